import sys, time
sys.path.insert(0, '.')
import numpy as np
from grok_alpha_zero_b200 import netspec
from grok_alpha_zero_b200.net import Net
game = sys.argv[1] if len(sys.argv) > 1 else "gomoku"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
spec = netspec.build_spec(game, "softmax")
W = netspec.init_weights(spec, seed=0)
net = Net(spec, W, max_batch=B)
fl = netspec.flops_per_eval(spec)
for n in sorted(set([B // 4, B])):
    ms = net.time_forward(n, 5)
    print(game, "batch", n, "forward %.3f ms" % ms, "evals/s %.3e" % (n / ms * 1e3), "TFLOP/s %.1f" % (n * fl / ms * 1e3 / 1e12))
net.profile(64)
ms = net.time_forward(B, 1)
import ctypes
tot, cnt, per = net.profile_read()
print("conv launches", cnt, "conv total ms (last passes)", tot)
print("per conv op ms:", [round(float(x), 3) for x in per if x > 0])
print("net bytes GB", net.bytes_allocated() / 1e9)
