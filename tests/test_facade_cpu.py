"""CPU suite: the reference-named façades (games, MCTS, MCTS_Gumbel) against the reference goldens and the C oracle.

The façades run on the test-only host emulation of the engine (tests/emul_lib.py); on the GPU box
tests/test_facade_gpu.py repeats the golden drive through the CUDA library."""
import numpy as np
import pytest

import emul_lib
import oracle as orc
from golden_util import case_id, f32bits, load
from grok_alpha_zero_b200 import games
from grok_alpha_zero_b200.MCTS import MCTS
from grok_alpha_zero_b200.MCTS_Gumbel import MCTS_Gumbel
from hash_eval import HashSession

PUCT = load("puct.json")
GUMBEL = load("gumbel.json")
CLS = {"gomoku": games.Gomoku, "connect4": games.Connect4, "tictactoe": games.TicTacToe}


@pytest.fixture(scope="module")
def lib():
    return emul_lib.load()


def drive_puct(case, lib, max_plies=None):
    """Self_Play.play pattern: two trees on one live game, both re-rooted after every ply."""
    name = case["game"]
    g = CLS[name]()
    sess = HashSession(g.policy_shape[0], logits=False, salt=case["salt"])
    kw = dict(use_dirichlet=False, tau=0.0, c_puct_init=case["c_puct_init"], lib=lib)
    t1, t2 = MCTS(g, sess, **kw), MCTS(g, sess, **kw)
    for ply, mv in enumerate(case["moves"][:max_plies]):
        tree = t1 if g.get_next_player() == -1 else t2
        move, rows = tree.run(iteration_limit=case["iters"], use_bar=False)
        where = "%s ply %d" % (case_id(case), ply)
        assert games.action_to_id(name, move) == mv["action"], where
        want = {c[0]: c for c in mv["children"] if c[1] > 0 or c[4] != 2}
        got = {games.action_to_id(name, r[0]): r for r in rows}
        for a, r in got.items():
            w = want[a]
            assert int(r[4]) == w[1], where                      # visits
            assert int(f32bits(r[3])) == w[2], where             # value sum, float32 bit pattern
            assert int(f32bits(r[5])) == w[3], where             # prior
            assert r[6] == mv["root_visits"], where
            assert (2 if r[7] is None else r[7]) == w[4], where  # is_terminal
        assert sorted(int(r[4]) for r in rows) == sorted(c[1] for c in mv["children"] if c[0] in got), where
        assert [int(r[4]) for r in rows] == sorted((int(r[4]) for r in rows), reverse=True), "rows sorted by visits"
        assert sess.calls == mv["evals"], where
        g.do_action(move)
        assert g.check_win() == mv["winner"], where
        if mv["winner"] != -2:
            break
        t1.prune_tree(move)
        t2.prune_tree(move)
    t1.close(); t2.close()


@pytest.mark.parametrize("case", [PUCT[0], PUCT[3], PUCT[7]], ids=case_id)
def test_mcts_facade_matches_reference_goldens(lib, case):
    drive_puct(case, lib, max_plies=12)


@pytest.mark.parametrize("case", [c for c in GUMBEL if not c["reuse"]][:6], ids=case_id)
def test_mcts_gumbel_facade_matches_reference_goldens(lib, case):
    drive_gumbel(case, lib)


def drive_gumbel(case, lib, check_pi=True):
    name = case["game"]
    g = CLS[name]()
    sess = HashSession(g.policy_shape[0], logits=True, salt=case["salt"])
    for ply, mv in enumerate(case["moves"][:8]):
        t = MCTS_Gumbel(g, sess, use_gumbel_noise=False, m=case["m"], c_visit=case["c_visit"], c_scale=case["c_scale"],
                        activation_fn=case["activation"], lib=lib)   # Self_Play.py:151-153: fresh tree per move
        move, rows = t.run(iteration_limit=case["n"], use_bar=False)
        where = "%s ply %d" % (case_id(case), ply)
        assert games.action_to_id(name, move) == mv["action"], where
        want = {c[0]: c for c in mv["children"]}
        assert len(rows) == len(want), where
        for r in rows:
            w = want[games.action_to_id(name, r[0])]
            assert int(r[4]) == w[1] and int(f32bits(r[3])) == w[2] and int(f32bits(r[5])) == w[3], where
            assert r[6] == mv["root_visits"], where
        pis = [float(r[1]) for r in rows]
        assert pis == sorted(pis, reverse=True), "rows sorted by pi'"
        if check_pi:  # final pi' is always the softmax branch: bit-exact with glibc exp (host emulation) only
            got_pi = {games.action_to_id(name, r[0]): int(f32bits(r[1])) for r in rows}
            want_pi = {c[0]: p for c, p in zip(mv["children"], mv["pi"])}
            assert got_pi == want_pi, where
        t.close()
        g.do_action(move)
        assert g.check_win() == mv["winner"], where
        if mv["winner"] != -2:
            break


def test_mcts_facade_api_rules(lib):
    g = games.TicTacToe()
    with pytest.warns(UserWarning):
        t = MCTS(g, None, use_dirichlet=True, tau=1e-4, lib=lib)        # tau < 5e-3 -> 0 (MCTS.py:116-120)
    assert t.tau == 0.0
    move, rows = t.run(iteration_limit=4, use_bar=False)                # limit < n_legal -> 3 * n_legal (MCTS.py:543-546)
    assert rows[0][6] >= 27 and len(rows) == 9
    assert abs(sum(r[1] for r in rows) - 1.0) < 1e-9
    with pytest.warns(UserWarning):
        t.update_hyperparams(c_puct_init=-1.0)
    t.update_hyperparams(c_puct_init=3.0, tau=1.0, dirichlet_alpha=0.3)
    assert t.c_puct_init == 3.0 and t.tau == 1.0 and t.dirichlet_alpha == 0.3
    g.do_action(move)
    t.prune_tree(move)
    move2, rows2 = t.run(iteration_limit=30, use_bar=False)
    assert g.board[move2[1]][move2[0]] == 0
    g.do_action(move2)
    t.prune_tree(move2, create_new_root=True)
    t.close()


def test_full_random_games_through_both_facades(lib):
    """Game_Tester.check_compatibility_with_MCTS (Game_Tester.py:480-513): one whole game, session=None."""
    for cls in (games.TicTacToe, games.Connect4):
        g = cls()
        t = MCTS(g, None, lib=lib)
        tg = MCTS_Gumbel(g, None, m=4, lib=lib)
        winner = -2
        while winner == -2:
            n = len(g.get_legal_actions())
            move, rows = t.run(iteration_limit=n, use_bar=False)
            mg, rg = tg.run(iteration_limit=n, use_bar=False)
            assert abs(sum(float(r[1]) for r in rows) - 1.0) < 1e-6
            assert abs(sum(float(r[1]) for r in rg) - 1.0) < 1e-4
            g.do_action(move)
            winner = g.check_win()
            if winner == -2:
                t.prune_tree(move)
                tg.prune_tree(move)
        t.close(); tg.close()


# ------------------------------------------------------------------ game classes vs the C oracle's int8 games --
@pytest.mark.parametrize("name", ["tictactoe", "connect4", "gomoku"])
def test_game_classes_match_oracle_games(name):
    rng = np.random.RandomState(11)
    for trial in range(12 if name != "gomoku" else 5):
        g, o = CLS[name](), orc.OracleGame(name)
        w = -2
        while w == -2:
            legal = g.get_legal_actions()
            ids = sorted(games.action_to_id(name, a) for a in legal)
            assert ids == sorted(int(a) for a in o.legal())
            if len(legal) == 0:
                break
            a = legal[rng.randint(len(legal))]
            g.do_action(a)
            o.do_action(games.action_to_id(name, a))
            assert np.array_equal(g.board.reshape(-1), o.board)
            assert np.array_equal(np.asarray(g.get_input_state(), dtype=np.int8), o.input_state())
            w = g.check_win()
            assert w == o.check_win()
            if name == "gomoku" and len(g.action_history) >= 225:
                break


def test_policy_masking_and_augmentation():
    rng = np.random.RandomState(3)
    g = games.Gomoku()
    for a in [(7, 7), (8, 7), (7, 8)]:
        g.do_action(a)
    pol = rng.rand(225).astype(np.float32)
    acts, p = g.get_legal_actions_policy_MCTS(g.board, -g.next_player, None, pol.copy())
    assert len(acts) == 222 and p.dtype == np.float32 and abs(float(p.sum()) - 1.0) < 1e-5
    s = np.float32(0)
    for v in pol[g.board.reshape(-1) == 0]:
        s = np.float32(s + v)
    assert np.array_equal(p.view(np.uint32), (pol[g.board.reshape(-1) == 0] / s).view(np.uint32))
    acts2, p2 = g.get_legal_actions_policy_MCTS(g.board, -g.next_player, None, pol.copy(), normalize=False)
    assert abs(float(p2.sum()) - 1.0) > 1e-3
    states = np.stack([g.get_input_state()] * 3)
    b8, p8 = g.augment_sample(states, np.stack([pol] * 3))
    assert b8.shape == (8, 3, 15, 15, 2) and b8.dtype == np.int8 and p8.shape == (8, 3, 225) and p8.dtype == np.float32
    assert np.array_equal(b8[1, 0], np.flipud(states[0])) and np.array_equal(p8[3, 0], np.rot90(pol.reshape(15, 15)).reshape(-1))
    c = games.Connect4()
    for a in [3, 3, 2, 4, 3]:
        c.do_action(a)
    st = c.get_input_state()
    assert st.shape == (6, 7, 4) and st[5, 3, 3] == -1 and np.all(st[..., 0] == st[..., 0])
    b2, pp = c.augment_sample(np.stack([st, st]), rng.rand(2, 7).astype(np.float32))
    assert b2.shape == (2, 2, 6, 7, 4) and pp.shape == (2, 2, 7)
    assert np.array_equal(c.compute_policy_improvement([[3, 0.75], [0, 0.25]]), np.array([0.25, 0, 0, 0.75, 0, 0, 0], np.float32))
