import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np
from grok_alpha_zero_b200 import netspec
from grok_alpha_zero_b200.engine import Engine
from grok_alpha_zero_b200.net import Net
for game, n in (("tictactoe", 4096), ("connect4", 4096)):
    spec = netspec.build_spec(game, "softmax"); W = netspec.init_weights(spec, seed=0)
    eng = Engine(game, n_games=n, mode="puct", trees_per_game=1, c_puct_init=2.5, iters_hint=1200)
    net = Net(spec, W, max_batch=n); net.attach(eng)
    if eng.new_roots() > 0: eng.eval_net()
    eng.expand(); eng.run_begin([1200] * n)
    eng.rounds_net(8)
    eng.timer_begin(); eng.rounds_net(200, sync=False); ms = eng.timer_end()
    vis, _, info = eng.root_dense()
    print(game, "GAZ_GRAPH=%s" % os.environ.get("GAZ_GRAPH", "1"), "ms/round %.4f" % (ms / 200), "iters", int(info[:, 2].min()), int(info[:, 2].max()), "status", eng.status(), "checksum", int(vis.astype(np.int64).sum()))
