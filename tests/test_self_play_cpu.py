"""CPU suite for the batched self-play driver (emulated engine + the deterministic hash evaluator): trajectories equal
a `Self_Play.play`-style loop over the MCTS façade, results do not depend on slot count or rank count, the replay
file follows the reference schema, and the world_size-2 gloo gather reproduces the single-process file."""
import os
import subprocess
import sys

import numpy as np
import pytest

import emul_lib
from grok_alpha_zero_b200 import games
from grok_alpha_zero_b200.MCTS import MCTS
from grok_alpha_zero_b200.Self_Play import (BatchedSelfPlay, ReplayWriter, finalize_game, pack_games, run_self_play,
                                            unpack_games)
from hash_eval import HashSession

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return emul_lib.load()


def cfg(game, **kw):
    tc = dict(MCTS_iteration_limit=40, use_gumbel=False, c_puct_init=2.5, dirichlet_alpha=0.5, max_actions=42,
              num_explore_actions_first=0, num_explore_actions_second=0, games_per_generation=6, games_per_gpu=3)
    tc.update(kw)
    return {"use_stablemax": False}, tc


def facade_game(cls, tc, salt=0, lib=None):
    """Self_Play.play (Self_Play.py:71-157) over the façade: two trees, tau = 0, no noise"""
    g = cls()
    sess = HashSession(g.policy_shape[0], logits=False, salt=salt)
    kw = dict(use_dirichlet=False, tau=0.0, c_puct_init=tc["c_puct_init"], lib=lib)
    t1, t2 = MCTS(g, sess, **kw), MCTS(g, sess, **kw)
    states, pols, qs, zs = [], [], [], []
    w = -2
    while w == -2 and len(zs) < tc["max_actions"]:
        states.append(np.asarray(g.get_input_state(), dtype=np.int8).copy())
        tree = t1 if g.get_next_player() == -1 else t2
        a, rows = tree.run(iteration_limit=int(tc["MCTS_iteration_limit"] * 1.5), use_bar=False)
        pols.append(g.compute_policy_improvement(map(lambda r: r[:2], rows)))
        qs.append([r[2] for r in rows if np.array_equal(r[0], a)][0])
        zs.append(g.get_next_player())
        g.do_action(a)
        w = g.check_win()
        if w == -2:
            t1.prune_tree(a); t2.prune_tree(a)
    t1.close(); t2.close()
    return dict(states=np.array(states), policies=np.array(pols, np.float32), q=np.array(qs, np.float32),
                z=np.array(zs, np.float32), winner=w if w != -2 else 0)


@pytest.mark.parametrize("name", ["tictactoe", "connect4"])
def test_batched_games_equal_the_facade_loop(lib, name):
    cls = {"tictactoe": games.TicTacToe, "connect4": games.Connect4}[name]
    bc, tc = cfg(name)
    ref = facade_game(cls, tc, lib=lib)
    sp = BatchedSelfPlay(cls, bc, tc, list(range(5)), n_slots=2, evaluator="hash", lib=lib, use_noise=False)
    fin = sp.play()
    sp.close()
    assert sorted(g["game_id"] for g in fin) == list(range(5))          # slots were re-seated
    for g in fin:
        assert g["winner"] == ref["winner"] and g["length"] == len(ref["z"])
        assert np.array_equal(g["states"], ref["states"])
        assert np.array_equal(g["policies"].view(np.uint32), ref["policies"].view(np.uint32))
        assert np.allclose(g["q"], ref["q"], atol=1e-6) and np.array_equal(g["z"], ref["z"])


def test_results_do_not_depend_on_slots_or_ranks(lib):
    bc, tc = cfg("connect4", num_explore_actions_first=3, num_explore_actions_second=2, MCTS_iteration_limit=24,
                 opening_actions=[[3, 0.4]])
    runs = []
    for ids_list, slots in ([list(range(8))], 8), ([list(range(8))], 3), ([[0, 2, 4, 6], [1, 3, 5, 7]], 2):
        fin = []
        for ids in ids_list:
            sp = BatchedSelfPlay(games.Connect4, bc, tc, ids, n_slots=slots, evaluator="hash", lib=lib, seed=7,
                                 use_noise=True)   # device Dirichlet noise keyed by the global game id
            fin += sp.play()
            sp.close()
        runs.append({g["game_id"]: g for g in fin})
    assert len({tuple(runs[0][i]["states"][-1].reshape(-1)) for i in range(8)}) > 1   # explored games differ
    for other in runs[1:]:
        for i in range(8):
            a, b = runs[0][i], other[i]
            assert a["winner"] == b["winner"] and np.array_equal(a["states"], b["states"])
            assert np.array_equal(a["policies"], b["policies"]) and np.array_equal(a["q"], b["q"])


def test_gumbel_batched_self_play_runs_and_targets_are_distributions(lib):
    bc, tc = cfg("tictactoe", use_gumbel=True, MCTS_iteration_limit=16, m=4, c_visit=50.0, c_scale=1.0, max_actions=9)
    bc = {"use_stablemax": True}
    sp = BatchedSelfPlay(games.TicTacToe, bc, tc, list(range(4)), n_slots=4, evaluator="hash", lib=lib, seed=1)
    fin = sp.play()
    sp.close()
    assert len(fin) == 4
    for g in fin:
        assert np.allclose(g["policies"].sum(1), 1.0, atol=1e-4) and g["length"] <= 9 and g["winner"] in (-1, 0, 1)


def test_finalize_and_replay_schema(tmp_path):
    proto = games.TicTacToe()
    T = 5
    states = np.zeros((T, 3, 3, 2), np.int8)
    pols = np.full((T, 9), 1 / 9, np.float32)
    q = np.linspace(-0.5, 0.5, T).astype(np.float32)
    z = np.array([-1, 1, -1, 1, -1], np.float32)
    b, p, v = finalize_game(proto, states, pols, q, z, winner=-1)
    assert b.shape == (8, T, 3, 3, 2) and p.shape == (8, T, 9) and v.shape == (8, T, 1)
    assert np.allclose(v[0, :, 0], 0.5 * (-z + q))            # player -1 just won: z flipped (Self_Play.py:164-165)
    _, _, v0 = finalize_game(proto, states, pols, q, z, winner=0)
    assert np.allclose(v0[0, :, 0], 0.5 * q)                  # draw: z zeroed
    w = ReplayWriter(str(tmp_path))
    w.add_game(b, p, v, T, -1)
    w.add_game(b, p, v0, T, 0)
    w.flush()
    w2 = ReplayWriter(str(tmp_path))
    assert w2.games_done() == 2
    data = w2.data if not w2.use_h5 else None
    if data is not None:
        assert list(data["game_stats"]) == [T, 2 * T, 2, 1, 1, 0]
        assert {"boards_0", "policies_7", "values_15"} <= set(data.keys()) and len(data) == 1 + 3 * 16
        assert data["boards_8"].dtype == np.int8 and data["policies_8"].dtype == np.float32


def test_pack_unpack_round_trip():
    gs = [dict(game_id=3, winner=1, length=2, states=np.ones((2, 3, 3, 2), np.int8), policies=np.ones((2, 9), np.float32),
               q=np.zeros(2, np.float32), z=np.array([-1, 1], np.float32)),
          dict(game_id=9, winner=0, length=1, states=np.zeros((1, 3, 3, 2), np.int8), policies=np.zeros((1, 9), np.float32),
               q=np.ones(1, np.float32), z=np.array([-1], np.float32))]
    back = unpack_games(pack_games(gs))
    assert [g["game_id"] for g in back] == [3, 9]
    assert all(np.array_equal(a[k], b[k]) for a, b in zip(gs, back) for k in ("states", "policies", "q", "z"))
    assert unpack_games(pack_games([])) == []


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests")); sys.path.insert(0, os.path.join({root!r}, "oracle"))
import torch.distributed as dist
import emul_lib
from grok_alpha_zero_b200 import games
from grok_alpha_zero_b200.Self_Play import run_self_play
dist.init_process_group("gloo")
bc = {{"use_stablemax": False}}
tc = dict(MCTS_iteration_limit=24, use_gumbel=False, c_puct_init=2.5, dirichlet_alpha=0.5, max_actions=42,
          num_explore_actions_first=2, num_explore_actions_second=2, games_per_generation=6, games_per_gpu=2)
run_self_play(games.Connect4, (bc, tc, {{}}), {out!r}, evaluator="hash", lib=emul_lib.load(), seed=5)
dist.barrier()
dist.destroy_process_group()
'''


def test_world_size_2_gloo_gather_matches_single_process(tmp_path, lib):
    out2, out1 = str(tmp_path / "w2"), str(tmp_path / "w1")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=out2))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)], env=env, timeout=600)
    bc = {"use_stablemax": False}
    tc = dict(MCTS_iteration_limit=24, use_gumbel=False, c_puct_init=2.5, dirichlet_alpha=0.5, max_actions=42,
              num_explore_actions_first=2, num_explore_actions_second=2, games_per_generation=6, games_per_gpu=2)
    run_self_play(games.Connect4, (bc, tc, {}), out1, evaluator="hash", lib=lib, seed=5)
    a, b = ReplayWriter(out1), ReplayWriter(out2)
    if a.use_h5:
        pytest.skip("npz comparison only")
    assert set(a.data.keys()) == set(b.data.keys()) and int(a.data["game_stats"][2]) == 6
    for k in a.data:
        assert np.array_equal(a.data[k], b.data[k]), k


def test_tree_sizes_and_pool_relief(lib):
    """gaz_tree_sizes reports the pool occupancy root_stats reports per tree; with pools too small for the sub-tree
    reuse the driver gives the tree of the side to move a fresh root (MCTS.py:661-671 behaviour for an unseen action)
    instead of overflowing: the generation finishes with a clean status word."""
    from grok_alpha_zero_b200.Self_Play import BatchedSelfPlay
    bc = {"num_resnet_layers": 1, "num_filters": 128, "use_stablemax": False}
    tc = dict(MCTS_iteration_limit=40, use_gumbel=False, c_puct_init=2.5, dirichlet_alpha=0.5, max_actions=42,
              num_explore_actions_first=1, num_explore_actions_second=1)
    n = 5
    # limit = 60 iterations: one move needs 120 nodes of head-room, so a 150-node pool forces a rebuild as soon as more
    # than 30 nodes are kept
    sp = BatchedSelfPlay(games.Connect4, bc, tc, list(range(n)), n, evaluator="hash", lib=lib, seed=3,
                         node_cap=150, slot_cap=150 * 7)
    sp._seat(list(range(n)))
    sz = sp.eng.tree_sizes()
    assert sz.shape == (2 * n, 2)
    for t in range(2 * n):
        st = sp.eng.root_stats(t)
        assert (int(st["n_nodes"]), int(st["n_slots"])) == tuple(int(x) for x in sz[t])
    while sp.step():
        assert sp.eng.status() == 0
        assert (sp.eng.tree_sizes()[:, 0] <= 150).all()
    assert len(sp.finished) == n and sp.pool_rebuilds > 0
    sp.close()
    # roomy pools: same games, no rebuilds
    sp = BatchedSelfPlay(games.Connect4, bc, tc, list(range(n)), n, evaluator="hash", lib=lib, seed=3)
    assert len(sp.play()) == n and sp.pool_rebuilds == 0
    sp.close()


def test_bench_pools_follow_the_self_play_rule(lib):
    """bench.py sizes its tree pools with the rule of BatchedSelfPlay, so `hbm_bytes` is what a real generation needs."""
    import bench
    from grok_alpha_zero_b200.Self_Play import BatchedSelfPlay
    classes = {"gomoku": games.Gomoku, "connect4": games.Connect4, "tictactoe": games.TicTacToe}
    for name, cfg in bench.CONFIGS.items():
        gumbel = cfg["mode"] == "gumbel"
        tc = dict(MCTS_iteration_limit=cfg["sims"], use_gumbel=gumbel, c_puct_init=cfg["c_puct_init"], dirichlet_alpha=0.3,
                  max_actions=10, m=16, c_visit=50.0, c_scale=1.0)
        sp = BatchedSelfPlay(classes[cfg["game"]], {"use_stablemax": gumbel}, tc, [0], 1, evaluator="hash", lib=lib)
        assert sp.limit == cfg["limit"]
        assert (sp.eng.node_cap, sp.eng.slot_cap) == bench.tree_caps(cfg), name
        sp.close()


@pytest.mark.parametrize("name", ["tictactoe", "connect4", "gomoku"])
def test_augmentation_tables_reproduce_the_games_own_method(lib, name):
    """`gaz_augment` gathers through tables probed from the game class's own `augment_sample`; random trajectories must come
    out bit-identical to that method (8 dihedral copies for Gomoku / TicTacToe, the np.fliplr pair with its row-flip quirk
    for Connect4, Connect4.py:441-442)."""
    from grok_alpha_zero_b200.Self_Play import augmentation_tables, finalize_games
    cls = {"tictactoe": games.TicTacToe, "connect4": games.Connect4, "gomoku": games.Gomoku}[name]
    proto = cls()
    ps, pp = augmentation_tables(proto)
    S, P = np.asarray(proto.get_input_state()).size, proto.policy_shape[0]
    assert ps.shape == ({"connect4": 2}.get(name, 8), S) and pp.shape[1] == P
    assert all(sorted(r) == list(range(S)) for r in ps.tolist()) and all(sorted(r) == list(range(P)) for r in pp.tolist())
    rng = np.random.default_rng(3)
    fin = []
    for gid, T in enumerate((1, 5, 9)):
        fin.append(dict(game_id=gid, winner=int(rng.integers(-1, 2)), length=T,
                        states=rng.integers(-1, 2, size=(T,) + np.asarray(proto.get_input_state()).shape).astype(np.int8),
                        policies=rng.random((T, P), dtype=np.float32), q=rng.random(T, dtype=np.float32) * 2 - 1,
                        z=np.where(np.arange(T) % 2 == 0, -1.0, 1.0).astype(np.float32)))
    out = list(finalize_games(proto, fin, lib=lib, chunk_positions=6))     # chunking: 1 + 5 | 9
    assert [o[0]["game_id"] for o in out] == [0, 1, 2]
    for g, b, p, v in out:
        hb, hp, hv = finalize_game(proto, g["states"], g["policies"], g["q"], g["z"], g["winner"])
        assert b.shape == hb.shape and np.array_equal(b, hb)
        assert np.array_equal(p.view(np.uint32), np.asarray(hp, np.float32).view(np.uint32))
        assert np.array_equal(v, hv)


def test_streaming_writer_appends_and_resumes(tmp_path):
    """every add_game reaches the archive at once (nothing but game_stats stays in memory), and a finished archive can be
    re-opened to append (Self_Play.py:267-272: games_left = games_per_generation - game_stats[2])"""
    proto = games.TicTacToe()
    T = 3
    b = np.zeros((8, T, 3, 3, 2), np.int8)
    p = np.full((8, T, 9), 1 / 9, np.float32)
    v = np.zeros((8, T, 1), np.float32)
    w = ReplayWriter(str(tmp_path))
    w.add_game(b + 1, p, v, T, -1)
    assert os.path.getsize(os.path.join(str(tmp_path), "Self_Play_Data.npz")) > 8 * T * 18      # on disk before flush()
    w.add_game(b + 2, p, v + 0.5, T + 1, 1)
    w.flush()
    w2 = ReplayWriter(str(tmp_path))
    assert w2.games_done() == 2 and w2.n_datasets == 16
    w2.add_game(b + 3, p, v - 0.5, T, 0)
    w2.flush()
    with np.load(os.path.join(str(tmp_path), "Self_Play_Data.npz")) as z:
        assert len(z.files) == 1 + 3 * 24 and z["game_stats"].tolist() == [T + 1, 3 * T, 3, 1, 1, 1]
        assert int(z["boards_0"][0, 0, 0, 0]) == 1 and int(z["boards_8"][0, 0, 0, 0]) == 2 and int(z["boards_16"][0, 0, 0, 0]) == 3
        assert float(z["values_23"][0, 0]) == -0.5
    assert list(ReplayWriter(str(tmp_path)).data.keys())[:4] == ["game_stats", "boards_0", "policies_0", "values_0"]
