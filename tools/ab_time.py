"""A/B timing of two builds of the library on ONE box (developer tool): the forward pass of both builds alternately.
usage: python tools/ab_time.py <variant.so under tests/_emul/> [game] [batch] [repeats]"""
import os, sys
sys.path.insert(0, '.')
import numpy as np
from grok_alpha_zero_b200 import netspec, _lib, _net_symbols
from grok_alpha_zero_b200.net import Net

var = sys.argv[1]
game = sys.argv[2] if len(sys.argv) > 2 else "gomoku"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
R = int(sys.argv[4]) if len(sys.argv) > 4 else 6
spec = netspec.build_spec(game, "softmax")
W = netspec.init_weights(spec, seed=0)
syms = dict(_lib.SYMBOLS, **_net_symbols.SYMBOLS)
nets = []
for path in (_lib.lib_path(), os.path.join("tests", "_emul", var)):
    nets.append(Net(spec, W, max_batch=B, lib=_lib.bind(path, syms)))
ms = [[], []]
for n in nets:
    n.time_forward(B, 3)      # warm-up
for r in range(R):
    for k, n in enumerate(nets):
        ms[k].append(n.time_forward(B, 4))
for k, tag in enumerate(("this build", var)):
    print("%-28s forward ms: median %.3f  min %.3f  all %s" % (tag, float(np.median(ms[k])), min(ms[k]), [round(x, 3) for x in ms[k]]))
