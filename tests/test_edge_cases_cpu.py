"""Edge cases of the search path on the host emulation of the engine: terminal roots, forced single moves, full boards,
ragged batches (finished / idle games), against the C oracle where it applies."""
import numpy as np
import pytest

import emul_lib
import oracle as orc
from grok_alpha_zero_b200 import games
from grok_alpha_zero_b200.engine import Engine
from grok_alpha_zero_b200.MCTS import MCTS
from grok_alpha_zero_b200.MCTS_Gumbel import MCTS_Gumbel
from hash_eval import HashSession


@pytest.fixture(scope="module")
def lib():
    return emul_lib.load()


def test_terminal_root_takes_the_winning_move(lib):
    """create_expand_root with terminal replies (MCTS.py:315-344): children = terminal actions, no network call"""
    g = games.Connect4()
    for a in [3, 0, 3, 0, 3, 1]:        # player -1 has three in column 3 and is to move
        g.do_action(a)
    sess = HashSession(7)
    t = MCTS(g, sess, use_dirichlet=False, tau=0.0, lib=lib)
    assert sess.calls == 0              # terminal root: the evaluator is never asked
    move, rows = t.run(iteration_limit=50, use_bar=False)
    assert int(move) == 3 and rows[0][7] == -1   # is_terminal = the mover wins
    o = orc.OracleGame("connect4")
    for a in [3, 0, 3, 0, 3, 1]:
        o.do_action(a)
    ot = orc.OracleTree("connect4", False, salt=0)
    ot.new_root(o)
    ot.run(50)
    ref = ot.root_stats()
    st = t.engine.root_stats(0)
    assert np.array_equal(st["visits"], ref["visits"]) and st["root_visits"] == ref["root_visits"]
    t.close()


def test_single_legal_move_runs_one_iteration(lib):
    """MCTS.run: one legal move -> iteration_limit = 1 (MCTS.py:543-544)"""
    g = games.TicTacToe()
    for a in [(0, 0), (1, 1), (2, 2), (0, 2), (2, 0), (1, 0), (1, 2), (2, 1)]:   # one empty cell left: (0, 1), no winner yet
        g.do_action(a)
        assert g.check_win() == -2
    t = MCTS(g, HashSession(9), use_dirichlet=False, tau=0.0, lib=lib)
    move, rows = t.run(iteration_limit=500, use_bar=False)
    assert tuple(int(x) for x in move) == (0, 1) and len(rows) == 1
    assert t.engine.root_stats(0)["iter"] == 1
    tg = MCTS_Gumbel(g, HashSession(9, logits=True), m=8, lib=lib)
    mg, rg = tg.run(iteration_limit=16, use_bar=False)
    assert tuple(int(x) for x in mg) == (0, 1) and tg.m == 1                      # m clipped to the legal moves (MCTS_Gumbel.py:581)
    t.close(); tg.close()


def test_full_board_draw_and_gomoku_no_draw_detection(lib):
    eng = Engine("connect4", n_games=1, lib=lib)
    o = orc.OracleGame("connect4")
    order = [0, 1, 0, 1, 0, 1, 1, 0, 1, 0, 1, 0, 2, 3, 2, 3, 2, 3, 3, 2, 3, 2, 3, 2, 4, 5, 4, 5, 4, 5, 5, 4, 5, 4, 5, 4,
             6, 6, 6, 6, 6, 6]
    w = -2
    for a in order:
        w = int(eng.apply_actions([a])[0])
        o.do_action(a)
        assert w == o.check_win()
        if w != -2:
            break
    assert w == 0 and len(o.history) == 42                                       # full board, nobody connected four
    eng.close()
    # Gomoku: the reference has no draw detection (Gomoku.py:250-255 returns -2 on a full board); the device agrees
    g = games.Gomoku()
    eng = Engine("gomoku", n_games=1, lib=lib)
    rng = np.random.RandomState(0)
    w = -2
    while w == -2 and len(g.action_history) < 225:
        legal = g.get_legal_actions()
        a = legal[rng.randint(len(legal))]
        g.do_action(a)
        w = int(eng.apply_actions([games.action_to_id("gomoku", a)])[0])
        assert w == g.check_win()
    eng.close()


def test_ragged_batch_idle_and_finished_games(lib):
    """limits <= 0 = idle trees; finished games keep their winner and ignore further actions"""
    n = 5
    eng = Engine("tictactoe", n_games=n, trees_per_game=1, lib=lib)
    eng.new_roots(); eng.eval_hash(0, False); eng.expand()
    limits = [30, 0, 30, -1, 30]
    eng.run_begin(limits)
    while eng.remaining() > 0:
        if eng.select() > 0:
            eng.eval_hash(0, False)
        eng.expand()
    vis, _, info = eng.root_dense()
    assert [int(x) for x in info[:, 2]] == [30, 0, 30, 0, 30]
    assert int(vis[1].sum()) == 0 and int(vis[3].sum()) == 0 and int(vis[0].sum()) > 0
    # play game 0 to the end while the others idle (-1 = skip)
    o = orc.OracleGame("tictactoe")
    w = -2
    for a in [0, 3, 1, 4, 2]:
        acts = [-1] * n
        acts[0] = a
        ws = eng.apply_actions(acts)
        o.do_action(a)
        w = int(ws[0])
    assert w == -1 == o.check_win() and all(int(x) == -2 for x in eng.apply_actions([-1] * n)[1:])
    assert int(eng.apply_actions([5, -1, -1, -1, -1])[0]) == -1                  # finished game: action ignored, winner kept
    b, nxt, hl, win = eng.get_game(0)
    assert hl == 5 and win == -1
    assert eng.status() == 0
    eng.close()


def test_engine_rejects_bad_arguments(lib):
    from grok_alpha_zero_b200.engine import EngineError
    with pytest.raises(EngineError):
        Engine("gomoku", n_games=0, lib=lib)
    eng = Engine("gomoku", n_games=2, lib=lib)
    with pytest.raises(EngineError):
        eng.root_stats(7)                      # tree index out of range
    with pytest.raises(EngineError):
        eng.set_game(9, np.zeros((15, 15), np.int8), -1, [])
    with pytest.raises(EngineError):
        eng.root_stats(0)                      # no root yet
    with pytest.raises(ValueError):
        games.game_name_of(type("G", (), {"board": np.zeros((4, 4))})())
    eng.close()


def test_node_pool_overflow_is_reported_not_silent(lib):
    eng = Engine("gomoku", n_games=1, node_cap=16, slot_cap=4000, lib=lib)
    eng.new_roots(); eng.eval_hash(0, False); eng.expand()
    eng.run_begin([400])
    guard = 0
    while eng.remaining() > 0 and guard < 2000:
        if eng.select() > 0:
            eng.eval_hash(0, False)
        eng.expand()
        guard += 1
    assert eng.status() & 3                    # sticky overflow bits (1 = nodes, 2 = slots)
    eng.close()
