"""Device-side evaluation cache (gaz_eval_cache_enable; the role of Session_Cache.Cache_Wrapper, Session_Cache.py:4-26):
identical positions are evaluated once.  A hit returns exactly what the evaluator returned for that position (the full
input state is compared), so every search result must be bit-identical with the cache on - at any table size, in both
scopes - and the number of evaluator calls must go down.  CPU: host emulation + hash evaluator; GPU: CUDA, hash evaluator
and the CUDA network (packed miss batches: row independence of the network)."""
import numpy as np
import pytest

from grok_alpha_zero_b200.engine import Engine


def drive(lib, cache=None, game="connect4", n_games=5, moves=4, iters=150, net=None, distinct=False):
    eng = Engine(game, n_games=n_games, mode="puct", trees_per_game=2, c_puct_init=2.5, iters_hint=iters, lib=lib)
    if cache is not None:
        eng.enable_eval_cache(cache[0], shared=cache[1])
    if net is not None:
        net.attach(eng)
    if distinct:    # different openings per game
        for g in range(n_games):
            b = np.zeros((eng.H, eng.W), np.int8)
            if game == "connect4":
                b[5, g % 7] = -1
                eng.set_game(g, b, 1, [g % 7])
    serve = (lambda: eng.eval_net()) if net is not None else (lambda: eng.eval_hash(5, False))
    for k in range(2):
        m = np.zeros((n_games, 2), np.uint8)
        m[:, k] = 1
        if eng.new_roots(m.reshape(-1)) > 0:
            serve()
        eng.expand()
    out, mover = [], 1 if distinct else 0
    for mv in range(moves):
        lim = np.zeros((n_games, 2), np.int32)
        lim[:, mover] = iters
        eng.run_begin(lim.reshape(-1))
        while eng.remaining() > 0:
            if net is not None:
                eng.rounds_net(16)
            else:
                eng.rounds_hash(16, 5, False)
        vis, val, info = eng.root_dense()
        out.append((vis.copy(), val.copy(), info.copy()))
        act = info.reshape(n_games, 2, 4)[:, mover, 1].astype(np.int16)
        eng.apply_actions(act)
        if eng.prune(np.repeat(act, 2)) > 0:
            serve()
        eng.expand()
        mover ^= 1
    st, stats = eng.status(), eng.eval_cache_stats()
    eng.close()
    return out, st, stats


def same(a, b):
    return all(np.array_equal(x[0], y[0]) and np.array_equal(x[1].view(np.uint32), y[1].view(np.uint32)) and np.array_equal(x[2], y[2])
               for x, y in zip(a, b))


def check(lib, net_factory=None):
    mk = (lambda: net_factory()) if net_factory else (lambda: None)
    n0 = mk()
    ref, st, s0 = drive(lib, None, net=n0, distinct=True)
    assert st == 0 and s0["lookups"] == 0
    for entries, shared in ((1 << 16, False), (1 << 16, True), (1024, False)):     # 1024 entries: constant eviction
        nn = mk()
        got, st, stats = drive(lib, (entries, shared), net=nn, distinct=True)
        assert st == 0 and same(ref, got), (entries, shared)
        assert stats["lookups"] > 0 and stats["entries"] == entries
        if entries > 1024:
            assert stats["hits"] > 0, stats          # both trees of a game meet the same positions (and transpositions)
        if nn is not None:
            nn.close()
    # identical games advance in lock step: every copy of a position is requested in the SAME round, before any of them is
    # stored (look-ups precede the fill pass), so sharing the table between games buys nothing here - the hits are the
    # second tree of each game and transpositions, in both scopes
    a, _, sa = drive(lib, (1 << 16, True), n_games=4, moves=2, distinct=False)
    b, _, sb = drive(lib, (1 << 16, False), n_games=4, moves=2, distinct=False)
    c, _, _ = drive(lib, None, n_games=4, moves=2, distinct=False)
    assert same(a, c) and same(b, c)
    assert sa["hits"] >= sb["hits"] > 0 and sa["lookups"] == sb["lookups"]
    if n0 is not None:
        n0.close()


def test_cache_does_not_change_results_emulated():
    import emul_lib
    check(emul_lib.load())


@pytest.mark.gpu
def test_cache_does_not_change_results_cuda():
    check(None)


@pytest.mark.gpu
def test_cache_with_the_cuda_network_is_bit_identical():
    from grok_alpha_zero_b200 import netspec
    from grok_alpha_zero_b200.net import Net
    spec = netspec.build_spec("connect4", "softmax", num_blocks=2)
    W = netspec.init_weights(spec, seed=3)
    check(None, net_factory=lambda: Net(spec, W, max_batch=16))
