"""fp32 PyTorch restatement of the reference's Keras policy/value networks (TEST INFRASTRUCTURE ONLY).

Follows Gomoku/Build_Model.py:10-88, Connect4/Build_Model.py:10-88, TicTacToe/Build_Model.py:8-69,
Net/ResNet/ResNet_Block.py:27-41, Net/SE/SE_Block.py:15-23, Net/Stablemax.py:7-11 with Keras
semantics (NHWC, "same" padding, biases everywhere, BatchNorm eps=1e-3 in inference mode,
Dense-after-Reshape flattening in H,W,C order, exact-erf gelu, softmax in float64 for
Gomoku/Connect4 - MCTS.py:234 casts it back to float32).

Parity status: WIRING PINNED, ARITHMETIC UNPINNED.  TensorFlow / tf2onnx / onnxruntime (pinned in
requirements-training-*.txt: 2.18.0 / 1.16.1 / 1.20.1) are not installed here and the reference
holds no network golden vectors.  What is pinned: tests/golden/net_*.npz hold the
outputs of the reference's own, unmodified `build_model` functions executed under
oracle/keras_shim.py (a minimal stand-in for the Keras layers they call) on seeded weights and
positions; tests/test_net_golden.py holds this restatement to them at fp32 round-off (2e-5), so
the layer graph - skip connections, pre-activation order, flatten order, head activations and
dtypes - is the reference's.  What is not: the layer arithmetic on both sides is PyTorch's, not
TensorFlow's / onnxruntime's.  The CUDA network is judged against the fixtures and against this
restatement with BASELINE.json's tolerance: policy logits atol 2e-2, value atol 1e-2.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from grok_alpha_zero_b200.netspec import BN_EPS, is_tensor_core_conv  # noqa: E402


def _t(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


class NetOracle:
    """forward(states int8/float (B,H,W,C)) -> dict(logits, policy, value_pre, value) as float32/64 torch."""

    def __init__(self, spec, weights, dtype=torch.float32, bf16_sim=False, res_fp32=True):
        # bf16_sim: round every tensor-core operand to bf16 (what the CUDA trunk stores); res_fp32: the
        # residual stream itself stays float32 (only conv operands are rounded)
        self.spec, self.W, self.dtype, self.bf16_sim, self.res_fp32 = spec, weights, dtype, bf16_sim, res_fp32

    def _q(self, x):  # emulate a bf16 activation store
        return x.to(torch.bfloat16).to(self.dtype) if self.bf16_sim else x

    def _qr(self, x):  # residual-stream store
        return x if self.res_fp32 else self._q(x)

    def _wq(self, w):
        return w.to(torch.bfloat16).to(self.dtype) if self.bf16_sim else w

    def conv(self, x, name, quant_w=True):  # x NCHW
        k = _t(self.W[name + ".kernel"], self.dtype).permute(3, 2, 0, 1)  # HWIO -> OIHW
        if quant_w:
            k = self._wq(k)
        b = _t(self.W[name + ".bias"], self.dtype)
        return F.conv2d(x, k, b, padding=k.shape[-1] // 2)

    def bn(self, x, name, dim=1):
        g, b, m, v = (_t(self.W[name + "." + s], self.dtype) for s in ("gamma", "beta", "mean", "var"))
        scale = g / torch.sqrt(v + BN_EPS)
        shift = b - m * scale
        shape = [1] * x.dim()
        shape[dim] = -1
        return x * scale.view(shape) + shift.view(shape)

    def dense(self, x, name):
        return x @ _t(self.W[name + ".kernel"], self.dtype) + _t(self.W[name + ".bias"], self.dtype)

    def se(self, h, n):
        """Squeeze-Excitation on an NCHW tensor (Net/SE/SE_Block.py:15-23): mean over the cells -> Dense(C/ratio) -> relu
        -> Dense(C) -> sigmoid -> scale"""
        s = h.mean(dim=(2, 3))
        s = F.relu(self.dense(s, n + ".se1"))
        s = torch.sigmoid(self.dense(s, n + ".se2"))
        return h * s[:, :, None, None]

    @torch.no_grad()
    def forward(self, states):
        spec = self.spec
        x = torch.as_tensor(np.asarray(states)).to(self.dtype).permute(0, 3, 1, 2)  # NHWC -> NCHW
        out = {}
        for l in spec["layers"]:
            if l["op"] == "stem":
                x = self.bn(self.conv(x, l["name"], quant_w=False), l["bn"])
                x = F.relu(x) if l["act"] == "relu" else F.gelu(x)  # exact erf gelu
                x = self._qr(x)
            elif l["op"] == "block":
                n = l["name"]
                res = self._qr(self.conv(self._q(x), n + ".proj")) if l["proj"] else x
                h = self._q(F.relu(self.bn(x, n + ".bn1")))
                h = self.conv(h, n + ".conv1")
                h = self._q(F.relu(self.bn(h, n + ".bn2")))
                h = self.conv(h, n + ".conv2")
                if l["se"]:
                    h = self.se(self._qr(h), n)
                x = self._qr(h + res)
            else:
                h = x
                flat = False
                for hl in l["layers"]:
                    t = hl["t"]
                    if t == "bnrelu":
                        h = F.relu(self.bn(h, hl["name"], 1))
                        if not flat:
                            h = self._q(h)
                    elif t == "bn":
                        h = self.bn(h, hl["name"], 1)
                    elif t == "relu":
                        h = F.relu(h)
                    elif t == "conv":
                        big = is_tensor_core_conv(hl["cin"], hl["cout"])  # tensor-core path in the CUDA net
                        h = self.conv(h, hl["name"], quant_w=big)
                    elif t == "flatten":
                        h = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)  # H,W,C order
                        flat = True
                    elif t == "dense":
                        h = self.dense(h, hl["name"])
                if l["out"] == "policy":
                    out["logits"] = h.to(torch.float32)
                    if l["final"] == "softmax":
                        out["policy"] = torch.softmax(h.to(torch.float64), dim=-1).to(torch.float32)
                    elif l["final"] == "stablemax":
                        hf = h.to(torch.float32)
                        s = torch.where(hf >= 0, hf + 1.0, 1.0 / (1.0 - hf))  # Net/Stablemax.py:7-11
                        out["policy"] = s / s.sum(-1, keepdim=True)
                    else:
                        out["policy"] = h.to(torch.float32)
                else:
                    out["value_pre"] = h.to(torch.float32)
                    out["value"] = torch.tanh(h).to(torch.float32)
        return out


class OracleSession:
    """Reference session duck type (MCTS.py:224-235) backed by the fp32 oracle network."""

    def __init__(self, spec, weights, threads=None):
        self.net = NetOracle(spec, weights)
        if threads:
            torch.set_num_threads(threads)

    def run(self, output_names=None, input_feed=None, **kw):
        o = self.net.forward(np.asarray(input_feed["inputs"]))
        return [o["policy"].numpy(), o["value"].numpy().reshape(-1, 1)]
