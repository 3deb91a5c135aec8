"""One full self-play generation through the reference-named entry point `run_self_play` on one GPU (not a pytest file):
    python tools/selfplay_generation.py [connect4|tictactoe] [games] [OUT_DIR]
BASELINE configs[1]: Connect4, 800 sims/move (x1.5 = 1200 iterations, Self_Play.py:99), 5-block ResNet128, 4096 concurrent
games, Dirichlet noise and temperature schedule as Self_Play.py configures them, every game played to its end, trajectories
augmented and written by the replay writer.  Prints one JSON line: wall-clock positions/s and simulations/s of the whole call."""
import json
import os
import shutil
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from grok_alpha_zero_b200 import games, netspec  # noqa: E402
from grok_alpha_zero_b200.Self_Play import net_spec_from_configs, run_self_play  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
if world > 1:     # torchrun: games shard by id over the ranks, trajectories are gathered to rank 0 over NCCL
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
game = sys.argv[1] if len(sys.argv) > 1 else "connect4"
n_games = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
out = sys.argv[3] if len(sys.argv) > 3 else "/tmp/gaz_generation"
if rank == 0:
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
if world > 1:
    dist.barrier()
if game == "connect4":    # Connect4/Connect4.py:4-60 with BASELINE's depth and simulation count
    cls, sims, blocks, max_actions, alpha, cpuct = games.Connect4, 800, 5, 42, 0.5, 2.5
elif game in ("gomoku", "gumbel"):   # Gomoku/Gomoku.py:5-60 with BASELINE's depth; "gumbel" = configs[3] (m=16, n=64, StableMax)
    cls, sims, blocks, max_actions, alpha, cpuct = games.Gomoku, (800 if game == "gomoku" else 64), 10, 150, 0.05, 4.5
else:
    cls, sims, blocks, max_actions, alpha, cpuct = games.TicTacToe, 200, 2, 9, 1.0, 1.25
bc = {"num_resnet_layers": blocks, "num_filters": 128, "use_stablemax": False}
tc = dict(MCTS_iteration_limit=sims, use_gumbel=False, c_puct_init=cpuct, dirichlet_alpha=alpha, max_actions=max_actions,
          num_explore_actions_first=2, num_explore_actions_second=1, games_per_generation=n_games, games_per_gpu=n_games // world)
gname = "gomoku" if game == "gumbel" else game
if gname == "gomoku":
    bc["use_se"] = True
if game == "gumbel":
    bc["use_stablemax"] = True
    tc.update(use_gumbel=True, m=16, c_visit=50.0, c_scale=1.0)
spec = net_spec_from_configs(gname, bc, tc)
W = netspec.init_weights(spec, seed=0)
t0 = time.time()
merged = run_self_play(cls, (bc, tc, {}), out, weights=W, seed=3)
dt = time.time() - t0
if world > 1:
    dist.barrier()
    dt = time.time() - t0
    if rank != 0:
        dist.destroy_process_group()
        sys.exit(0)
positions = int(sum(g["length"] for g in merged))
winners = np.array([g["winner"] for g in merged])
files = {f: os.path.getsize(os.path.join(out, f)) for f in os.listdir(out)}
print(json.dumps(dict(
    what="run_self_play: one generation, every game to its end, replay file written", game=game, games=len(merged),
    n_gpus=world,
    sims_per_move=sims, iterations_per_move=int(sims * 1.5) if game != "gumbel" else sims, net="%d x ResNet%d bf16" % (blocks, spec["cfg"]["filters"]), positions=positions,
    seconds=round(dt, 2), positions_per_s=round(positions / dt, 1), nominal_sims_per_s=round(positions * (int(sims * 1.5) if game != "gumbel" else sims) / dt, 1),
    mean_game_length=round(positions / max(1, len(merged)), 2),
    winners={"-1": int((winners == -1).sum()), "0": int((winners == 0).sum()), "1": int((winners == 1).sum())},
    replay_files=files)), flush=True)
if world > 1:
    dist.destroy_process_group()
