"""Deterministic, tie-free hash evaluator (TEST INFRASTRUCTURE ONLY).

Bit-for-bit the same function as ``orc_hash_eval`` in oracle/mcts_oracle.c and
the CUDA engine's parity evaluator (csrc: ``hash_eval_kernel``).  It plays the
role of the network behind the reference's session duck type
``run(output_names, input_feed[, depth]) -> [policy (1,P), value (1,1)]``
(/root/reference/MCTS.py:224-235), so the unmodified reference, the C oracle and
the GPU engine can be driven by identical evaluator outputs.

Outputs are exactly representable float32 values built from an integer hash:
  policy_i = k_i            (probability mode, k_i in 1..2^20, distinct per i)
  logit_i  = k_i*2^-17 - 4  (logits mode, used for MCTS_Gumbel)
  value    = v*2^-23 - 1    (v a 24-bit integer)
"""
import numpy as np

_M64 = (1 << 64) - 1
GOLD = 0x9E3779B97F4A7C15


def _mix_int(z):
    z &= _M64
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & _M64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return z


def _mix_arr(z):
    z = z.astype(np.uint64)
    z ^= z >> np.uint64(30)
    z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return z


_K_CACHE = {}


def _weights(n):
    w = _K_CACHE.get(n)
    if w is None:
        w = _mix_arr(np.arange(1, n + 1, dtype=np.uint64) * np.uint64(GOLD))
        _K_CACHE[n] = w
    return w


def hash_eval(state, P, logits=False, salt=0):
    """state: integer-valued array (any shape, values in {-1,0,1}); returns (policy f32[P], value f32)."""
    s = (np.asarray(state).reshape(-1).astype(np.int64) + 2).astype(np.uint64)
    acc = int((s * _weights(s.size)).sum(dtype=np.uint64))
    h0 = _mix_int(acc + salt * 0xD1B54A32D192ED03)
    idx = np.arange(P, dtype=np.uint64)
    r = _mix_arr(np.uint64(h0) + (idx + np.uint64(1)) * np.uint64(GOLD))
    k = (((r >> np.uint64(52)) << np.uint64(8)) | idx).astype(np.int64) + 1
    if logits:
        policy = k.astype(np.float32) * np.float32(2.0 ** -17) - np.float32(4.0)
    else:
        policy = k.astype(np.float32)
    rv = _mix_int(h0 + 0x5851F42D4C957F2D)
    value = np.float32(np.float32(rv >> 40) * np.float32(2.0 ** -23) - np.float32(1.0))
    return policy.astype(np.float32), value


class HashSession:
    """Reference-compatible session (duck type of onnxruntime.InferenceSession.run)."""

    def __init__(self, P, logits=False, salt=0):
        self.P, self.logits, self.salt = P, logits, salt
        self.calls = 0

    def run(self, output_names=None, input_feed=None, **kw):
        x = input_feed["inputs"]
        self.calls += 1
        with np.errstate(all="ignore"):  # the reference sets np.seterr(all='raise') globally
            p, v = hash_eval(x, self.P, self.logits, self.salt)
        return [p.reshape(1, self.P), np.array([[v]], dtype=np.float32)]
