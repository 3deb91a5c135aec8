"""Shared parity drivers: engine (CUDA or emulation) vs the reference goldens / the C oracle."""
import numpy as np

import oracle as orc
from golden_util import case_id, f32bits
from grok_alpha_zero_b200.engine import Engine, TERM_DRAW, TERM_NONE, TERM_WIN
from hash_eval import hash_eval


def term_ref(term_code, action_player):
    """engine term code -> reference is_terminal (2 = None); winner = the player who moved"""
    if term_code == TERM_NONE:
        return 2
    if term_code == TERM_DRAW:
        return 0
    return action_player


def stats_rows(st, mover, gumbel):
    rows = np.stack([st["action"].astype(np.int64), st["visits"].astype(np.int64),
                     f32bits(st["values"]).astype(np.int64), f32bits(st["prior"]).astype(np.int64),
                     np.array([term_ref(t, mover) for t in st["term"]], dtype=np.int64)], 1)
    if gumbel:
        rows = np.concatenate([rows, f32bits(st["raw"]).astype(np.int64)[:, None]], 1)
    return rows


def check_root(st, rec, mover, gumbel, where):
    want = np.array(rec["children"], dtype=np.int64)
    got = stats_rows(st, mover, gumbel)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (where, got.shape, want.shape)
    np.testing.assert_array_equal(got, want, err_msg=where)
    assert st["root_visits"] == rec["root_visits"], where


def host_hash_evaluator(P, logits, salt):
    def ev(states):
        pol = np.zeros((len(states), P), np.float32)
        val = np.zeros(len(states), np.float32)
        for i, s in enumerate(states):
            pol[i], val[i] = hash_eval(s, P, logits, salt)
        return pol, val
    return ev


def run_all(eng, limits, salt, logits, host_eval=None):
    eng.run_begin(limits)
    guard = 0
    while eng.remaining() > 0:
        n = eng.select()
        if n > 0:
            if host_eval is not None:
                eng.eval_pending(host_eval)
            else:
                eng.eval_hash(salt, logits)
                eng.expand()
        guard += 1
        assert guard < 100000
    assert eng.status() == 0, "engine status %d" % eng.status()


def serve_roots(eng, n, salt, logits, host_eval=None):
    if n > 0:
        if host_eval is not None:
            eng.eval_pending(host_eval)
        else:
            eng.eval_hash(salt, logits)
            eng.expand()


def puct_golden_case(lib, case, use_host_eval=False):
    """Self_Play-style drive: one game, two trees, both re-rooted after each ply."""
    game = case["game"]
    eng = Engine(game, n_games=1, mode="puct", trees_per_game=2, c_puct_init=case["c_puct_init"],
                 iters_hint=case["iters"], lib=lib)
    P = eng.P
    he = host_hash_evaluator(P, False, case["salt"]) if use_host_eval else None
    serve_roots(eng, eng.new_roots(), case["salt"], False, he)
    next_player = -1
    for ply, mv in enumerate(case["moves"]):
        t = 0 if next_player == -1 else 1
        assert t + 1 == mv["tree"]
        lim = [0, 0]
        lim[t] = case["iters"]
        run_all(eng, lim, case["salt"], False, he)
        where = "%s ply %d" % (case_id(case), ply)
        st = eng.root_stats(t)
        check_root(st, mv, next_player, False, where)
        a = int(st["action"][int(np.argmax(st["visits"]))])
        assert a == mv["action"], where
        ev = eng.root_stats(0)["evals"] + eng.root_stats(1)["evals"]
        assert ev == mv["evals"], where
        w = eng.apply_actions([a])
        assert int(w[0]) == mv["winner"], where
        next_player = -next_player
        if mv["winner"] == -2:
            serve_roots(eng, eng.prune([a, a]), case["salt"], False, he)
            # after prune the side to move is next_player; children of the roots are its moves
            check_root(eng.root_stats(0), mv["after_prune"][0], next_player, False, where + " prune t1")
            check_root(eng.root_stats(1), mv["after_prune"][1], next_player, False, where + " prune t2")
    assert eng.status() == 0
    eng.close()


def gumbel_golden_case(lib, case, check_pi=True):
    game = case["game"]
    eng = Engine(game, n_games=1, mode="gumbel", trees_per_game=1, m=case["m"], c_visit=case["c_visit"],
                 c_scale=case["c_scale"], activation_fn=case["activation"], iters_hint=case["n"] * 4, lib=lib)
    salt = case["salt"]
    serve_roots(eng, eng.new_roots(), salt, True)
    next_player = -1
    evals_prev = 0
    for ply, mv in enumerate(case["moves"]):
        run_all(eng, [case["n"]], salt, True)
        where = "%s ply %d" % (case_id(case), ply)
        st = eng.root_stats(0)
        check_root(st, mv, next_player, True, where)
        a = int(st["action"][st["best_slot"]])
        assert a == mv["action"], where
        assert st["evals"] == mv["evals"], where
        pi = eng.gumbel_pi(0)[:st["L"]]
        want_pi = np.array(mv["pi"], dtype=np.uint32).view(np.float32)
        if check_pi:  # final pi' always uses exp(): bit-exact only where exp is glibc's (CPU emulation)
            np.testing.assert_array_equal(f32bits(pi), f32bits(want_pi), err_msg=where)
        else:         # CUDA exp: <= 2 ulp of float32 after the final rounding
            np.testing.assert_allclose(pi, want_pi, rtol=3e-7, atol=1e-30, err_msg=where)
        w = eng.apply_actions([a])
        assert int(w[0]) == mv["winner"], where
        next_player = -next_player
        if mv["winner"] == -2:
            if case["reuse"]:
                serve_roots(eng, eng.prune([a]), salt, True)
            else:
                # Self_Play.py:151-153: a fresh MCTS_Gumbel per move (m is reset with it)
                eng.set_gumbel_params(case["m"], case["c_visit"], case["c_scale"], case["activation"] == "softmax")
                serve_roots(eng, eng.new_roots(), salt, True)
    assert eng.status() == 0
    eng.close()


def random_position(game, rng, plies):
    """Random non-terminal position reached by `plies` random legal moves (oracle game rules)."""
    while True:
        g = orc.OracleGame(game)
        ok = True
        for _ in range(plies):
            legal = g.legal()
            g.do_action(int(legal[rng.randint(len(legal))]))
            if g.check_win() != -2:
                ok = False
                break
        if ok and len(g.legal()) > 0:
            return g


def batch_vs_oracle(lib, game, n_games, iters, seed, mode="puct", max_plies=12, salt=0, **kw):
    """n_games different positions searched concurrently on the engine; every tree's root statistics must
    equal the C oracle's (bit-exact), then one re-root per game is compared too."""
    rng = np.random.RandomState(seed)
    gumbel = mode == "gumbel"
    max_plies = min(max_plies, {"tictactoe": 5, "connect4": 30, "gomoku": 100}[game])
    eng = Engine(game, n_games=n_games, mode=mode, trees_per_game=1, iters_hint=iters * (4 if gumbel else 1),
                 lib=lib, **kw)
    games = [random_position(game, rng, rng.randint(0, max_plies + 1)) for _ in range(n_games)]
    for i, g in enumerate(games):
        eng.set_game(i, g.board, g.next_player, g.history)
    serve_roots(eng, eng.new_roots(), salt, gumbel)
    run_all(eng, [iters] * n_games, salt, gumbel)
    okw = dict(kw)
    if "c_puct_init" in okw:
        okw["c_puct_init"] = okw["c_puct_init"]
    actions = []
    trees = []
    for i, g in enumerate(games):
        t = orc.OracleTree(game, gumbel, salt=salt, **okw)
        t.new_root(g)
        a = t.run(iters)
        ref = t.root_stats()
        st = eng.root_stats(i)
        where = "%s game %d (hist %s)" % (game, i, g.history)
        np.testing.assert_array_equal(st["action"], ref["action"], err_msg=where)
        np.testing.assert_array_equal(st["visits"], ref["visits"], err_msg=where)
        np.testing.assert_array_equal(f32bits(st["values"]), f32bits(ref["values"]), err_msg=where)
        np.testing.assert_array_equal(f32bits(st["prior"]), f32bits(ref["prior"]), err_msg=where)
        assert st["root_visits"] == ref["root_visits"], where
        assert st["evals"] == t.n_evals, where
        if gumbel:
            np.testing.assert_array_equal(f32bits(st["raw"]), f32bits(ref["raw"]), err_msg=where)
            assert int(st["action"][st["best_slot"]]) == a, where
        else:
            assert int(st["action"][int(np.argmax(st["visits"]))]) == a, where
        actions.append(a)
        trees.append(t)
    # play the chosen action everywhere, re-root, search again
    w = eng.apply_actions(actions)
    acts2 = []
    for i, g in enumerate(games):
        g.do_action(actions[i])
        assert g.check_win() == int(w[i])
        acts2.append(actions[i] if w[i] == -2 else -1)
    serve_roots(eng, eng.prune(acts2), salt, gumbel)
    lim = [iters if a >= 0 else 0 for a in acts2]
    run_all(eng, lim, salt, gumbel)
    for i, g in enumerate(games):
        if acts2[i] < 0:
            continue
        t = trees[i]
        t.prune(actions[i])
        t.run(iters)
        ref = t.root_stats()
        st = eng.root_stats(i)
        where = "%s game %d after re-root" % (game, i)
        np.testing.assert_array_equal(st["action"], ref["action"], err_msg=where)
        np.testing.assert_array_equal(st["visits"], ref["visits"], err_msg=where)
        np.testing.assert_array_equal(f32bits(st["values"]), f32bits(ref["values"]), err_msg=where)
        assert st["root_visits"] == ref["root_visits"], where
    assert eng.status() == 0
    eng.close()
