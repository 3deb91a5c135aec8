"""A/B check of two builds of the library on the same weights and inputs (developer tool).
usage: python tools/ab_compare.py <variant.so under tests/_emul/> [game] [batch]   -- prints max |diff| of logits / policy / value"""
import os, sys
sys.path.insert(0, '.')
import numpy as np
from grok_alpha_zero_b200 import netspec, _lib, _net_symbols
from grok_alpha_zero_b200.net import Net

var = sys.argv[1]
game = sys.argv[2] if len(sys.argv) > 2 else "gomoku"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 512
spec = netspec.build_spec(game, "softmax")
W = netspec.init_weights(spec, seed=3)
rng = np.random.default_rng(5)
H, Wd, Cin = spec["H"], spec["W"], spec["Cin"]
stones = rng.integers(-1, 2, size=(B, H, Wd)).astype(np.int8)
states = np.zeros((B, H, Wd, Cin), np.int8)
states[..., 0] = rng.choice([-1, 1], size=(B, 1, 1))
states[..., Cin - 1] = stones
outs = []
syms = dict(_lib.SYMBOLS, **_net_symbols.SYMBOLS)
for path in (_lib.lib_path(), os.path.join("tests", "_emul", var)):
    lib = _lib.bind(path, syms)
    net = Net(spec, W, max_batch=B, lib=lib)
    outs.append(net.forward(states, want_logits=True))
    net.close()
if os.environ.get("AB_ORACLE"):      # both builds against the fp32 restatement (first 256 states)
    sys.path.insert(0, "oracle")
    from net_oracle import NetOracle
    m = min(B, 256)
    ref = NetOracle(spec, W).forward(states[:m])
    for tag, o in zip(("this build", var), outs):
        print("%-28s vs fp32 oracle: logits err %.4g value err %.4g" % (tag, np.abs(o[2][:m] - ref["logits"].numpy()).max(),
                                                                          np.abs(o[1][:m] - ref["value"].numpy().reshape(-1)).max()))
for name, a, b in zip(("policy", "value", "logits"), outs[0], outs[1]):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    print(name, "max|diff| %.3e" % np.abs(a - b).max(), "bit-identical" if np.array_equal(a, b) else "differs", "absmax %.3f" % np.abs(a).max())

d = np.abs(np.asarray(outs[0][2], np.float64) - np.asarray(outs[1][2], np.float64)).max(axis=1)
bad = np.nonzero(d > 0)[0]
print("boards whose logits differ: %d of %d; first %s; per-board max diff of those: %s" % (len(bad), B, bad[:24].tolist(), np.round(d[bad[:12]], 5).tolist()))
