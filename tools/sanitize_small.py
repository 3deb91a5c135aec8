"""Small end-to-end run for compute-sanitizer (memcheck): both networks, both search modes, closed loop.
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
Not a pytest file: the sanitizer slows kernels by 10-100x, so the sizes are tiny."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import numpy as np

import engine_parity as ep
import net_util
from grok_alpha_zero_b200 import netspec
from grok_alpha_zero_b200.engine import Engine
from grok_alpha_zero_b200.net import Net

for game, over, n in (("gomoku", dict(num_blocks=3, use_se=True), 5), ("connect4", dict(num_blocks=2), 9),
                      ("tictactoe", {}, 7)):
    spec = netspec.build_spec(game, "softmax", **over)
    W = netspec.init_weights(spec, seed=1)
    net = Net(spec, W, max_batch=16)
    pol, val = net.forward(net_util.random_states(game, n, seed=4))
    assert np.isfinite(pol).all() and np.isfinite(val).all()
    eng = Engine(game, n_games=4, mode="puct", trees_per_game=2, c_puct_init=2.5, iters_hint=64)
    net.attach(eng)
    if eng.new_roots() > 0:
        eng.eval_net()
        eng.expand()
    eng.run_begin([24, 0] * 4)
    eng.rounds_net(24)
    vis, valsum, info = eng.root_dense()
    assert eng.status() == 0 and int(vis.sum()) > 0
    w = eng.apply_actions(info.reshape(4, 2, 4)[:, 0, 1].astype(np.int16))
    eng.prune(np.repeat(info.reshape(4, 2, 4)[:, 0, 1], 2))
    eng.close()
    net.close()
    print(game, "net + closed loop ok")

for mode, kw in (("puct", dict(c_puct_init=4.5)), ("gumbel", dict(m=8, c_visit=50.0, c_scale=1.0, activation_fn="stablemax"))):
    ep.batch_vs_oracle(None, "gomoku", 3, 40, seed=3, mode=mode, max_plies=8, **kw)
    print("gomoku", mode, "vs oracle ok")
print("done")
