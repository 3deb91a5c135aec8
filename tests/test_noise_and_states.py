"""Device-side Dirichlet noise (Philox + Marsaglia-Tsang gamma) and the batched game-state read-back, on the host
emulation of the engine (CPU) - the same code the CUDA kernels run."""
import numpy as np
import pytest

import emul_lib
import oracle as orc
from grok_alpha_zero_b200.engine import Engine


@pytest.fixture(scope="module")
def lib():
    return emul_lib.load()


def root_priors(eng, n):
    out = np.zeros((n, eng.P), np.float64)
    for t in range(n):
        st = eng.root_stats(t)
        out[t, st["action"]] = st["prior"]
    return out


@pytest.mark.parametrize("alpha", [0.3, 1.11, 0.05])
def test_dirichlet_noise_statistics(lib, alpha):
    n, eps = 1500, 0.25
    clean = Engine("tictactoe", n_games=1, lib=lib)
    clean.new_roots(); clean.eval_hash(0, False); clean.expand()
    p = root_priors(clean, 1)[0]
    clean.close()
    eng = Engine("tictactoe", n_games=n, lib=lib)
    eng.set_noise(alpha, eps, seed=123)
    eng.new_roots(); eng.eval_hash(0, False); eng.expand()
    noisy = root_priors(eng, n)
    eng.close()
    d = (noisy - (1 - eps) * p) / eps                      # the Dirichlet(alpha) component (MCTS.py:243-245)
    assert np.all(d > -1e-5) and np.allclose(d.sum(1), 1.0, atol=1e-4)
    k = 9
    mean, var = 1.0 / k, (1.0 / k) * (1 - 1.0 / k) / (k * alpha + 1)
    assert np.allclose(d.mean(0), mean, atol=5 * np.sqrt(var / n) + 2e-3)
    assert np.allclose(d.var(0), var, rtol=0.25, atol=2e-3)
    assert len({tuple(np.round(r, 6)) for r in noisy}) > n * 0.99      # every tree has its own stream


def test_noise_is_reproducible_and_keyed(lib):
    def run(seed, keys):
        eng = Engine("connect4", n_games=4, lib=lib)
        eng.set_noise(0.5, 0.25, seed=seed)
        eng.set_tree_keys(keys)
        eng.new_roots(); eng.eval_hash(0, False); eng.expand()
        out = root_priors(eng, 4)
        eng.close()
        return out
    a = run(1, [10, 11, 12, 13])
    b = run(1, [13, 12, 11, 10])
    c = run(2, [10, 11, 12, 13])
    assert np.array_equal(a, b[::-1]) and not np.array_equal(a, c)
    with pytest.raises(Exception):
        Engine("connect4", n_games=1, lib=lib).set_noise(0.5, 1.5)


@pytest.mark.parametrize("name", ["tictactoe", "connect4", "gomoku"])
def test_get_states_and_set_games_round_trip(lib, name):
    rng = np.random.RandomState(2)
    n = 6
    eng = Engine(name, n_games=n, lib=lib)
    ogs = [orc.OracleGame(name) for _ in range(n)]
    for ply in range(7):
        acts = []
        for g in ogs:
            legal = g.legal()
            acts.append(int(legal[rng.randint(len(legal))]))
        w = eng.apply_actions(acts)
        for g, a in zip(ogs, acts):
            g.do_action(a)
        st, info = eng.get_states()
        for i, g in enumerate(ogs):
            assert np.array_equal(st[i], g.input_state()), (name, ply, i)
            assert info[i, 0] == g.next_player and info[i, 1] == len(g.history) and info[i, 3] == acts[i]
            assert info[i, 2] == w[i] == g.check_win()
        if (w != -2).any():
            break
    # batched upload of the same positions into a second engine reproduces the boards
    eng2 = Engine(name, n_games=n, lib=lib)
    boards = np.stack([g.board.reshape(eng.H, eng.W) for g in ogs])
    nxt = np.array([g.next_player for g in ogs])
    if name == "connect4":
        l3 = np.array([sum((a & 7) << (3 * k) for k, a in enumerate(reversed(g.history[-3:]))) for g in ogs])
    else:
        l3 = None
    eng2.set_games(boards, nxt, hist_lens=[len(g.history) for g in ogs], last3=l3,
                   last_actions=[g.history[-1] for g in ogs])
    st2, info2 = eng2.get_states()
    assert np.array_equal(st2, eng.get_states()[0]) and np.array_equal(info2[:, :2], eng.get_states()[1][:, :2])
    eng.close(); eng2.close()
