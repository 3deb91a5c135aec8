"""Writes tests/golden/net_*.npz: outputs of the reference's OWN network builders on seeded inputs and weights
(TEST INFRASTRUCTURE ONLY; run in the build container, where /root/reference exists):

    python oracle/gen_net_golden.py            # rewrites the fixtures
    python oracle/gen_net_golden.py --check    # rebuilds them in memory and compares with the committed files

The builders (`Gomoku/Build_Model.py:10-88`, `Connect4/Build_Model.py:10-88`, `TicTacToe/Build_Model.py:8-69`, with
`Net/ResNet/ResNet_Block.py`, `Net/Stablemax.py`) are imported unmodified and executed under `oracle/keras_shim.py`
(TensorFlow is absent: the shim supplies the Keras layer definitions, the reference supplies the wiring).  Weights:
`netspec.init_weights(spec, seed)` pushed into the traced model through `keras_bridge.export_keras_weights`, i.e.
through the checkpoint names the bridge predicts - a wrong name or shape fails here.  `SE_Block`
(`Net/SE/SE_Block.py:4-23`, not wired into any reference model) gets its own fixture on a 5-D input.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("GAZ_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import keras_shim  # noqa: E402
from grok_alpha_zero_b200 import keras_bridge, netspec  # noqa: E402

SEED_W, SEED_X = 11, 23
# (fixture name, game, netspec head, netspec overrides, reference build_config / train_config switches, boards)
# RES_GAIN: the ten-block Gomoku trunk WITHOUT Squeeze-Excitation (the reference's builder has none) draws its conv2 kernels
# with residual gain 0.25 instead of init_weights' default 0.5 - what the SE gate (sigmoid ~ 0.5) does to the bench network -
# so that a random-init stack stays O(1) and the ABSOLUTE bf16 tolerance keeps its meaning (logits |max| 6.4 -> ~3).
RES_GAIN = {"gomoku_softmax": 0.25, "gomoku_stablemax": 0.25}
# "_g1" fixtures: the builders' own initialisation scale - he_normal at gain 1 everywhere (Gomoku/Build_Model.py:10-88,
# Connect4/Build_Model.py:22-75 pass kernel_initializer="he_normal" with no damping), kernels NOT pre-rounded to bf16.
# A random ten-block stack then reaches |logit| ~ 150, so these fixtures are judged with the scale-aware tolerance of
# tests/test_net_golden_gpu.py (the absolute north-star atol cannot hold for 8-bit-significand operands there).
UNDAMPED = dict(residual_gain=1.0, head_gain=1.0, bf16_kernels=False)
INIT = {"gomoku_softmax_g1": UNDAMPED, "connect4_softmax_g1": UNDAMPED}
CASES = [
    ("gomoku_softmax", "gomoku", "softmax", dict(num_blocks=10, use_se=False), dict(use_stablemax=False), dict(use_gumbel=False), 3),
    ("gomoku_stablemax", "gomoku", "stablemax", dict(num_blocks=10, use_se=False), dict(use_stablemax=True), dict(use_gumbel=False), 3),
    ("gomoku_linear", "gomoku", "linear", dict(num_blocks=4, use_se=False), dict(use_stablemax=False), dict(use_gumbel=True), 3),
    ("connect4_softmax", "connect4", "softmax", dict(num_blocks=5), dict(use_stablemax=False), dict(use_gumbel=False), 6),
    ("connect4_stablemax", "connect4", "stablemax", dict(num_blocks=3), dict(use_stablemax=True), dict(use_gumbel=False), 6),
    ("tictactoe_softmax", "tictactoe", "softmax", dict(num_blocks=2), dict(use_stablemax=False), dict(use_gumbel=False), 8),
    ("tictactoe_linear", "tictactoe", "linear", dict(num_blocks=2), dict(use_stablemax=False), dict(use_gumbel=True), 8),
    ("gomoku_softmax_g1", "gomoku", "softmax", dict(num_blocks=10, use_se=False), dict(use_stablemax=False), dict(use_gumbel=False), 3),
    ("connect4_softmax_g1", "connect4", "softmax", dict(num_blocks=5), dict(use_stablemax=False), dict(use_gumbel=False), 6),
]
_BUILDERS = {"gomoku": "Gomoku.Build_Model", "connect4": "Connect4.Build_Model", "tictactoe": "TicTacToe.Build_Model"}


def reference_model(game, spec, build_over, train_over):
    """the reference's build_model, traced under the shim"""
    import importlib
    keras_shim.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    mod = importlib.import_module(_BUILDERS[game])
    bc = dict(num_resnet_layers=spec["cfg"]["num_blocks"], num_filters=spec["cfg"]["filters"], use_grok_fast=False,
              use_orthograd=False, grok_fast_lambda=4.5, **build_over)
    tc = dict(train_over)
    shape, pshape = (spec["H"], spec["W"], spec["Cin"]), (spec["P"],)
    keras_shim.begin_trace()
    if game == "gomoku":          # Gomoku/Build_Model.py:10 takes the three config dicts as one tuple
        return mod.build_model(shape, pshape, (bc, tc, {}))
    return mod.build_model(shape, pshape, bc, tc)


def load_weights(model, spec, W):
    """netspec weights -> traced model, by the `<layer>/<variable>` names keras_bridge predicts"""
    exported = keras_bridge.export_keras_weights(spec, W)
    used = set()
    for v in model.weights:
        key = "/".join(v.path.split("/")[-2:])
        if key not in exported:
            raise KeyError("the reference model owns %r, which keras_bridge does not map" % v.path)
        v.assign(exported[key])
        used.add(key)
    if used != set(exported):
        raise KeyError("keras_bridge maps variables the reference model does not own: %r" % sorted(set(exported) - used))


def make_case(name, game, head, over, build_over, train_over, n):
    import net_util
    spec = netspec.build_spec(game, head, **over)
    init = dict(INIT.get(name, dict(residual_gain=RES_GAIN.get(name, 0.5))))
    W = netspec.init_weights(spec, seed=SEED_W, **init)
    states = net_util.random_states(game, n, seed=SEED_X)
    model = reference_model(game, spec, build_over, train_over)
    load_weights(model, spec, W)
    pol, val = (o.numpy() for o in model(states.astype(np.float32)))
    # the logits: the same weights behind the builder's use_gumbel=True head (Activation("linear"))
    lin = reference_model(game, spec, dict(build_over), dict(train_over, use_gumbel=True))
    load_weights(lin, spec, W)
    logits = lin(states.astype(np.float32))[0].numpy()
    return dict(states=states, policy=pol.astype(np.float32), value=val.astype(np.float32).reshape(-1),
                logits=logits.astype(np.float32), policy_dtype=np.array(str(pol.dtype)),
                meta=np.array(repr(dict(game=game, head=head, over=over, seed_w=SEED_W, seed_x=SEED_X,
                                         residual_gain=init["residual_gain"], init=init))),
                layer_names=np.array([l.name for l in model.layers if l._vars]))


def make_se_case():
    """SE_Block on a (B, 1, H, W, C) tensor with broadcast shape (1, 1, 1, C): GlobalAveragePooling3D squeezes the
    three middle axes, which for D = 1 is the mean over the board's cells."""
    import importlib
    import torch
    keras_shim.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    mod = importlib.import_module("Net.SE.SE_Block")
    C, Hh, Ww, B = 16, 5, 4, 3
    keras_shim.begin_trace()
    tf = sys.modules["tensorflow"]
    inp = tf.keras.layers.Input(batch_shape=(None, 1, Hh, Ww, C), name="inputs")
    out = mod.SE_Block(C, (1, 1, 1, C), ratio=2)(inp)
    model = tf.keras.Model(inputs=inp, outputs=[out])
    rng = np.random.default_rng(5)
    arrays = {}
    for v in model.weights:
        a = rng.normal(0, 0.5, size=v.shape).astype(np.float32)
        v.assign(a)
        arrays["/".join(v.path.split("/")[-2:])] = a
    x = rng.normal(0, 1, size=(B, 1, Hh, Ww, C)).astype(np.float32)
    y = model(torch.from_numpy(x))[0].numpy()
    return dict(x=x, y=y.astype(np.float32), w1=arrays["dense/kernel"], b1=arrays["dense/bias"],
                w2=arrays["dense_1/kernel"], b2=arrays["dense_1/bias"])


def build_all():
    import __graft_entry__ as ge
    ge.build_oracle()       # net_util.random_states plays random games on the C oracle
    out = {"net_" + c[0]: make_case(*c) for c in CASES}
    out["net_se_block"] = make_se_case()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    for name, d in build_all().items():
        path = os.path.join(GOLD, name + ".npz")
        if args.check:
            old = np.load(path)
            for k, v in d.items():
                if v.dtype.kind in "fc":
                    np.testing.assert_allclose(old[k], v, rtol=0, atol=1e-6, err_msg="%s:%s" % (name, k))
                else:
                    assert np.array_equal(old[k], v), (name, k)
            print("ok", name)
        else:
            np.savez_compressed(path, **d)
            print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
