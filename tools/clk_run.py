"""Per-role cycle accounting of the residual trunk kernel (developer tool): loads the -DGAZ_BLOCK_CLK build
(tools/build_variant.sh libgaz_clk.so -DGAZ_BLOCK_CLK) and runs one forward; run with GAZ_CONV_DBG=1024."""
import os, sys
sys.path.insert(0, '.')
import numpy as np
from grok_alpha_zero_b200 import netspec, _lib, _net_symbols
from grok_alpha_zero_b200.net import Net
game = sys.argv[1] if len(sys.argv) > 1 else "connect4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
spec = netspec.build_spec(game, "softmax")
W = netspec.init_weights(spec, seed=0)
lib = _lib.bind(os.path.join("tests", "_emul", "libgaz_clk.so"), dict(_lib.SYMBOLS, **_net_symbols.SYMBOLS))
net = Net(spec, W, max_batch=B, lib=lib)
print("forward ms", net.time_forward(B, 2))
