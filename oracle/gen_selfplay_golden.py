"""Writes tests/golden/selfplay_*.npz: what the reference's OWN `Self_Play.play` (Self_Play.py:71-208) writes for seeded,
noise-free games (TEST INFRASTRUCTURE ONLY; run in the build container, where /root/reference exists):

    python oracle/gen_selfplay_golden.py            # rewrites the fixtures
    python oracle/gen_selfplay_golden.py --check    # regenerates in memory and compares with the committed files

`Self_Play.py`, `MCTS.py`, `MCTS_Gumbel.py` and the game classes are imported unmodified.  Stubs: `onnxruntime`
(annotations only) and `h5py` -> oracle/h5_stub.py (h5py is absent; the stub keeps the datasets in a dict so every
`create_dataset` / `game_stats` update of Self_Play.py:178-208 is captured as written).  Determinism (the reference draws
from the global numpy stream): exploration noise is switched off by replacing `MCTS._apply_dirichlet` with the identity and
`np.random.gumbel` with zeros while a game runs, tau = 0 through `num_explore_actions_* = 0`, the opening book has a
single entry of weight 1, `np.random.randint -> low` inside `_PUCT_select` (oracle/ref_shim.py), evaluator =
oracle/hash_eval.HashSession.  Each fixture holds TWO consecutive games appended to one file (dataset numbering and
`game_stats` accumulation), the second with another evaluator salt.
"""
import argparse
import contextlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import h5_stub  # noqa: E402
import ref_shim  # noqa: E402
from hash_eval import HashSession  # noqa: E402

BASE_TC = dict(MCTS_time_limit=None, use_njit=True, num_explore_actions_first=0, num_explore_actions_second=0,
               m=16, c_visit=50.0, c_scale=1.0, c_puct_init=2.5, dirichlet_alpha=0.3, use_gumbel=False)
# name, game, train_config overrides, build_config, salts of the two games
CASES = [
    ("tictactoe_puct", "tictactoe", dict(MCTS_iteration_limit=40, max_actions=9, c_puct_init=1.25), {}, (0, 1)),
    ("connect4_puct", "connect4", dict(MCTS_iteration_limit=100, max_actions=42), {}, (0, 3)),
    # max_actions cut-off while the game is still running (Self_Play.py:155-157)
    ("connect4_puct_cutoff", "connect4", dict(MCTS_iteration_limit=60, max_actions=7), {}, (1, 2)),
    # max_actions == the ply on which game 0 is WON: the reference still books a draw (winner = 0 overrides the win)
    ("connect4_puct_win_on_last_ply", "connect4", dict(MCTS_iteration_limit=100, max_actions="WIN_PLY"), {}, (0, 3)),
    # limit < n_legal -> 3 * n_legal iterations (MCTS.py:543-546); opening book entry never expanded by the idle tree
    ("gomoku_puct_book", "gomoku", dict(MCTS_iteration_limit=60, max_actions=6, c_puct_init=4.5,
                                        opening_actions=[[[7, 7], 1.0]]), {}, (0, 1)),
    ("gomoku_gumbel_stablemax", "gomoku", dict(MCTS_iteration_limit=32, max_actions=8, use_gumbel=True, m=8),
     dict(use_stablemax=True), (0, 1)),
    ("connect4_gumbel_softmax", "connect4", dict(MCTS_iteration_limit=24, max_actions=42, use_gumbel=True, m=4,
                                                 c_scale=0.1), {}, (0, 1)),
    ("tictactoe_puct_new_root", "tictactoe", dict(MCTS_iteration_limit=30, max_actions=9, c_puct_init=1.25,
                                                  create_new_root=True), {}, (2, 5)),
]


@contextlib.contextmanager
def quiet_noise(ref):
    """no Dirichlet / Gumbel noise while the reference plays (it hard-wires use_dirichlet / use_gumbel_noise = True)"""
    saved_d, saved_g = ref.MCTS._apply_dirichlet, np.random.gumbel
    ref.MCTS._apply_dirichlet = lambda self, legal_policy, epsilon: legal_policy
    np.random.gumbel = lambda loc=0.0, scale=1.0, size=None: np.zeros(size, dtype=np.float64)
    try:
        yield
    finally:
        ref.MCTS._apply_dirichlet = saved_d
        np.random.gumbel = saved_g


def load_reference():
    h5_stub.install()
    ref = ref_shim.load()
    import Self_Play as ref_sp       # /root/reference/Self_Play.py, unmodified
    return ref, ref_sp


def play_reference(ref, ref_sp, game, tc, bc, salts, folder):
    path = os.path.join(folder, "Self_Play_Data.h5")
    with h5_stub.File(path, "w") as f:
        f.create_dataset("game_stats", data=np.zeros(6, dtype=np.uint32))   # Gomoku/main.py:88-96 (Make_Dataset_File)
    histories = []
    for salt in salts:
        g = ref.games[game]()
        sess = HashSession(g.policy_shape[0], logits=bool(tc["use_gumbel"]), salt=salt)
        np.random.seed(0)
        with quiet_noise(ref):
            sp = ref_sp.Self_Play(g, sess, bc, tc, contextlib.nullcontext(), folder, generation=1)
            sp.play()
        histories.append(np.asarray(g.action_history, dtype=np.int16).reshape(len(g.action_history), -1))
    return dict(h5_stub.STORE[os.path.abspath(path)]), histories


def make_case(ref, ref_sp, name, game, over, bc, salts):
    tc = dict(BASE_TC, **over)
    folder = "/tmp/gaz_selfplay_golden/" + name
    if tc["max_actions"] == "WIN_PLY":      # find the ply on which game 0 ends with a win, then make it the cut-off
        probe, hist = play_reference(ref, ref_sp, game, dict(tc, max_actions=42), bc, salts[:1], folder + "_probe")
        assert probe["game_stats"][4] == 0, "probe game must end with a win"
        tc["max_actions"] = int(len(hist[0]))
    data, hist = play_reference(ref, ref_sp, game, tc, bc, salts, folder)
    out = {"ds/" + k: v for k, v in data.items()}
    out["dataset_order"] = np.array(list(data.keys()))
    for i, h in enumerate(hist):
        out["history_%d" % i] = h
    out["meta"] = np.array(repr(dict(game=game, train_config=tc, build_config=bc, salts=list(salts))))
    return out


def build_all():
    ref, ref_sp = load_reference()
    return {"selfplay_" + c[0]: make_case(ref, ref_sp, *c) for c in CASES}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    for name, d in build_all().items():
        path = os.path.join(GOLD, name + ".npz")
        if args.check:
            old = np.load(path)
            assert sorted(old.files) == sorted(d.keys()), name
            for k, v in d.items():
                assert np.array_equal(old[k], v), (name, k)
            print("ok", name)
        else:
            np.savez_compressed(path, **d)
            st = d["ds/game_stats"]
            print("wrote", path, os.path.getsize(path), "bytes; game_stats", st.tolist(),
                  "lengths", [len(d["history_%d" % i]) for i in range(2)])


if __name__ == "__main__":
    main()
