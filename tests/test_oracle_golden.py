"""Pins the C oracle (oracle/mcts_oracle.c) to the reference's own outputs (tests/golden)."""
import numpy as np
import pytest

import oracle as orc
from golden_util import case_id, f32bits, load
from hash_eval import hash_eval

PUCT = load("puct.json")
GUMBEL = load("gumbel.json")
KATS = load("kats.json")


def check_root(stats, rec, gumbel, where):
    ch = rec["children"]
    assert len(ch) == len(stats["action"]), where
    assert stats["root_visits"] == rec["root_visits"], where
    got = np.stack([stats["action"].astype(np.int64), stats["visits"].astype(np.int64),
                    f32bits(stats["values"]).astype(np.int64), f32bits(stats["prior"]).astype(np.int64),
                    stats["term"].astype(np.int64)], 1)
    want = np.array([r[:5] for r in ch], dtype=np.int64)
    np.testing.assert_array_equal(got, want, err_msg=where)
    if gumbel:
        np.testing.assert_array_equal(f32bits(stats["raw"]).astype(np.int64),
                                      np.array([r[5] for r in ch], dtype=np.int64), err_msg=where)


@pytest.mark.parametrize("case", PUCT, ids=case_id)
def test_puct_matches_reference(case):
    g = orc.OracleGame(case["game"])
    mk = lambda: orc.OracleTree(case["game"], False, salt=case["salt"], c_puct_init=case["c_puct_init"])
    t1, t2 = mk(), mk()
    t1.new_root(g)
    t2.new_root(g)
    for ply, mv in enumerate(case["moves"]):
        tree = t1 if g.next_player == -1 else t2
        assert (1 if tree is t1 else 2) == mv["tree"]
        a = tree.run(case["iters"])
        where = "%s ply %d" % (case_id(case), ply)
        check_root(tree.root_stats(), mv, False, where)
        assert a == mv["action"], where
        assert t1.n_evals + t2.n_evals == mv["evals"], where
        g.do_action(a)
        assert g.check_win() == mv["winner"], where
        if mv["winner"] == -2:
            t1.prune(a)
            t2.prune(a)
            check_root(t1.root_stats(), mv["after_prune"][0], False, where + " prune t1")
            check_root(t2.root_stats(), mv["after_prune"][1], False, where + " prune t2")


@pytest.mark.parametrize("case", GUMBEL, ids=case_id)
def test_gumbel_matches_reference(case):
    g = orc.OracleGame(case["game"])
    mk = lambda: orc.OracleTree(case["game"], True, salt=case["salt"], m=case["m"], c_visit=case["c_visit"],
                                c_scale=case["c_scale"], activation_fn=case["activation"])
    tree = mk()
    tree.new_root(g)
    evals = 0
    for ply, mv in enumerate(case["moves"]):
        a = tree.run(case["n"])
        where = "%s ply %d" % (case_id(case), ply)
        check_root(tree.root_stats(), mv, True, where)
        assert a == mv["action"], where
        assert evals + tree.n_evals == mv["evals"], where
        if case["activation"] == "stablemax" or True:
            # final pi' is always the softmax branch (MCTS_Gumbel.py:656-662): glibc exp on both sides
            np.testing.assert_array_equal(f32bits(tree.pi).astype(np.int64), np.array(mv["pi"], dtype=np.int64),
                                          err_msg=where)
        g.do_action(a)
        assert g.check_win() == mv["winner"], where
        if mv["winner"] == -2:
            if case["reuse"]:
                tree.prune(a)
            else:
                evals += tree.n_evals
                tree = mk()
                tree.new_root(g)


def test_hash_eval_kats():
    L = orc.lib()
    import ctypes as C
    for k in KATS["eval"]:
        st = np.array(k["state"], dtype=np.int8)
        p, v = hash_eval(st, k["P"], k["logits"], k["salt"])
        assert f32bits(p).tolist() == k["policy_bits"]
        assert int(f32bits(v)) == k["value_bits"]
        ctx = orc.HashEvalCtx(k["P"], int(k["logits"]), k["salt"])
        pol = np.zeros(k["P"], np.float32)
        val = C.c_float()
        L.orc_hash_eval(C.byref(ctx), st.ctypes.data_as(C.c_void_p), st.size, pol.ctypes.data_as(C.c_void_p),
                        C.byref(val))
        assert f32bits(pol).tolist() == k["policy_bits"]
        assert int(f32bits(val.value)) == k["value_bits"]


@pytest.mark.parametrize("game", ["tictactoe", "connect4", "gomoku"])
def test_game_kats(game):
    for plies in KATS["games"][game]:
        g = orc.OracleGame(game)
        for p in plies:
            assert len(g.legal()) == p["n_legal_before"]
            g.do_action(p["action"])
            assert g.check_win() == p["winner"]
            st = g.input_state().reshape(-1).astype(np.int64)
            assert int((st * np.arange(1, st.size + 1)).sum()) == p["state_sum"]
            if p["state"] is not None:
                assert st.tolist() == p["state"]
