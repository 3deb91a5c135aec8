"""Evaluator objects with the reference's session duck type (MCTS.py:224-235):
    session.run(output_names=["policy", "value"], input_feed={"inputs": f32 (B,H,W,C)}[, depth=int])
        -> [policy (B, P), value (B, 1)]

`GazSession` is the CUDA network behind that call (what onnxruntime.InferenceSession is in Self_Play.py:226-236);
`Cache_Wrapper` mirrors Session_Cache.Cache_Wrapper (Session_Cache.py:13-26: outputs cached by the raw input bytes for
depth < max_cache_depth, look-ups stop after the first miss) with an in-memory dict instead of diskcache.
Checkpoints: the reference stores Keras `model.weights.h5` / `model.onnx`; neither h5py nor onnx exist in this image,
so the bridge format is a plain `.npz` of the Keras-layout arrays (conv (kh,kw,cin,cout), dense (in,out), BN
gamma/beta/mean/var) plus a JSON header, written by `save_checkpoint` - on a TensorFlow-equipped machine
`{w.path: w.numpy() for w in model.weights}` fills the same names.
"""
import json

import numpy as np

from . import netspec


def save_checkpoint(path, spec, weights):
    meta = dict(game=spec["game"], policy_head=spec["policy_head"], cfg=spec["cfg"])
    np.savez(path, __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **weights)


def load_checkpoint(path):
    z = np.load(path)
    meta = json.loads(bytes(z["__meta__"]).decode())
    spec = netspec.build_spec(meta["game"], meta["policy_head"], **meta["cfg"])
    weights = {k: z[k] for k in z.files if k != "__meta__"}
    return spec, weights


class GazSession:
    """CUDA policy/value network as a reference-style session."""

    def __init__(self, spec=None, weights=None, max_batch=64, device=0, checkpoint=None):
        from .net import Net
        if checkpoint is not None:
            spec, weights = load_checkpoint(checkpoint)
        self.spec = spec
        self.net = Net(spec, weights, max_batch=max_batch, device=device)
        self.max_batch = max_batch

    def get_inputs(self):   # onnxruntime-style introspection used by some reference scripts
        return [type("I", (), dict(name="inputs", shape=[None, self.spec["H"], self.spec["W"], self.spec["Cin"]]))()]

    def run(self, output_names=None, input_feed=None, **kw):
        x = np.asarray(input_feed["inputs"])
        x = x.reshape((-1,) + x.shape[-3:])
        st = np.ascontiguousarray(np.rint(x), dtype=np.int8)   # inputs are planes of {-1, 0, 1}
        outs_p, outs_v = [], []
        for i in range(0, len(st), self.max_batch):
            p, v = self.net.forward(st[i:i + self.max_batch])
            outs_p.append(p)
            outs_v.append(v.reshape(-1, 1))
        res = {"policy": np.concatenate(outs_p), "value": np.concatenate(outs_v)}
        names = output_names or ["policy", "value"]
        return [res[n] for n in names]

    def close(self):
        self.net.close()


class Cache_Wrapper:
    """Session_Cache.Cache_Wrapper (Session_Cache.py:4-26) with an in-memory store: look-ups happen until the first
    miss, outputs are stored for depth < max_cache_depth, same `run(output_names, input_feed, depth)` contract."""

    def __init__(self, session, path=None, max_cache_depth=2):
        self.session = session
        self.path = path
        self.finished_lookup = False if max_cache_depth > 0 else 0
        self.cache = {}
        self.max_cache_depth = max_cache_depth

    @staticmethod
    def _key(input_feed):
        return np.ascontiguousarray(np.asarray(input_feed["inputs"]).reshape(-1)).astype("<f4", copy=False).tobytes()

    def run(self, output_names, input_feed, depth=0):
        if self.max_cache_depth > 0 and not self.finished_lookup:
            outputs = self.cache.get(self._key(input_feed))
            if outputs is not None:
                return outputs
            self.finished_lookup = True
        outputs = self.session.run(output_names, input_feed)
        if depth < self.max_cache_depth:
            self.cache[self._key(input_feed)] = outputs
        return outputs
