"""Manual multi-GPU check (not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_selfplay_dist.py OUT_DIR
plays a small Connect4 generation with the CUDA network on every rank and gathers the trajectories to rank 0 over NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from grok_alpha_zero_b200 import games, netspec  # noqa: E402
from grok_alpha_zero_b200.Self_Play import ReplayWriter, run_self_play  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/gaz_selfplay"
bc = {"num_resnet_layers": 2, "num_filters": 128, "use_stablemax": False}
tc = dict(MCTS_iteration_limit=40, use_gumbel=False, c_puct_init=2.5, dirichlet_alpha=0.5, max_actions=42,
          num_explore_actions_first=3, num_explore_actions_second=2, games_per_generation=24, games_per_gpu=8)
spec = netspec.build_spec("connect4", "softmax", num_blocks=2)
W = netspec.init_weights(spec, seed=1)
merged = run_self_play(games.Connect4, (bc, tc, {}), out, weights=W, seed=9)
dist.barrier()
if dist.get_rank() == 0:
    w = ReplayWriter(out)
    st = w.data["game_stats"] if not w.use_h5 else None
    print("rank0 gathered", len(merged), "games; game_stats", None if st is None else st.tolist())
    assert len(merged) == 24 and sorted(g["game_id"] for g in merged) == list(range(24))
dist.destroy_process_group()
