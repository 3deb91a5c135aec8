import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from grok_alpha_zero_b200.engine import Engine
for game, n_games, iters in [("connect4", 4096, 300), ("gomoku", 4096, 300)]:
    eng = Engine(game, n_games=n_games, mode="puct", trees_per_game=1, c_puct_init=2.5, iters_hint=iters)
    n = eng.new_roots(); eng.eval_hash(0, False); eng.expand()
    eng.run_begin([iters]*n_games)
    eng.rounds_hash(20)
    torch.cuda.synchronize()
    t0 = time.time(); eng.rounds_hash(iters - 20); dt = time.time() - t0
    print(game, "games", n_games, "rounds", iters-20, "time %.3fs" % dt, "sims/s %.3e" % (n_games*(iters-20)/dt), "status", eng.status(), "remaining", eng.remaining(), "MB", eng.bytes_allocated()/1e6)
    st = eng.root_stats(0); print(st["visits"][:10], st["n_nodes"], st["n_slots"])
    eng.close()
