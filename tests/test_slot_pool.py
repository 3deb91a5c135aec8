"""The paged child-slot pool (gaz_core.cuh PAGE_SLOTS, gaz_config.slot_pool): trees take 4096-slot pages from one pool as they
grow and give them back when prune_tree compacts them.  Results must not depend on the pool size (bit for bit), pages must
come back, and an exhausted pool must stop the tree cleanly with the sticky status bit - never corrupt another tree.
CPU: the test-only host emulation; GPU (-m gpu): the CUDA library."""
import numpy as np
import pytest

from grok_alpha_zero_b200.engine import Engine


def play(lib, slot_pool, n_games=6, moves=3, iters=260, game="gomoku"):
    eng = Engine(game, n_games=n_games, mode="puct", trees_per_game=2, c_puct_init=4.5, node_cap=700, slot_cap=128 * 1024,
                 lib=lib, slot_pool=slot_pool)
    for k in range(2):
        m = np.zeros((n_games, 2), np.uint8)
        m[:, k] = 1
        if eng.new_roots(m.reshape(-1)) > 0:
            eng.eval_hash(3, False)
        eng.expand()
    out, free = [], [eng.pool_info()["free"]]
    mover, peak = 0, 0
    for mv in range(moves):
        lim = np.zeros((n_games, 2), np.int32)
        lim[:, mover] = iters
        eng.run_begin(lim.reshape(-1))
        while eng.remaining() > 0:
            eng.rounds_hash(32, 3, False)
        vis, val, info = eng.root_dense()
        out.append((vis.copy(), val.copy(), info.copy()))
        act = info.reshape(n_games, 2, 4)[:, mover, 1].astype(np.int16)
        free.append(eng.pool_info()["free"])
        peak = max(peak, int(eng.tree_sizes()[:, 1].max()))
        eng.apply_actions(act)
        if eng.prune(np.repeat(act, 2)) > 0:
            eng.eval_hash(3, False)
        eng.expand()
        free.append(eng.pool_info()["free"])
        mover ^= 1
    st, sizes, info = eng.status(), eng.tree_sizes(), eng.pool_info()
    info["peak_slots"] = peak
    eng.close()
    return out, free, st, sizes, info


def check_pool_independence(lib):
    full, free_full, st_full, sizes_full, info_full = play(lib, 0)
    assert st_full == 0 and info_full["pages"] == 12 * 32            # 12 trees x 128 K slots / 4096
    # a Gomoku node owns 225 slots: 18 blocks fill a page, the 19th must start a new one -> > 1 page per searching tree
    assert info_full["peak_slots"] > 2 * 4096
    used_peak = info_full["pages"] - min(free_full)
    small, free_small, st_small, sizes_small, info_small = play(lib, (used_peak + 1) * 4096)
    assert st_small == 0 and info_small["pages"] == used_peak + 1 < info_full["pages"] // 2
    for a, b in zip(full, small):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)) and np.array_equal(a[2], b[2])
    assert np.array_equal(sizes_full, sizes_small)
    # prune_tree gives pages back: after every re-rooting more pages are free than before it
    used = [info_small["pages"] - f for f in free_small]
    assert all(used[2 * k + 2] < used[2 * k + 1] for k in range(3)), used
    # a pool that is too small: the sticky SLOT overflow bit, no crash, and trees that got their pages are unharmed
    tiny, _, st_tiny, _, info_tiny = play(lib, 14 * 4096)
    assert st_tiny & 2 and info_tiny["pages"] == 14


def test_pool_size_does_not_change_results_emulated():
    import emul_lib
    check_pool_independence(emul_lib.load())


@pytest.mark.gpu
def test_pool_size_does_not_change_results_cuda():
    check_pool_independence(None)


@pytest.mark.gpu
def test_cuda_equals_the_width_1_emulation_through_prune_tree():
    """compute-sanitizer is not available on the pool, and the CPU suite runs the warp code with a
    cooperative width of 1, where lane races cannot exist.  This is the lane-race self-check for `compact_tree` and the
    scratch reuse: twelve trees searched, re-rooted (in-place compaction: the read - sync - write chunks) and searched again
    on the GPU, 32 lanes per tree and four trees per CTA, must equal the width-1 emulation bit for bit at every move."""
    import emul_lib
    cuda = play(None, 0, n_games=7, moves=4)
    emul = play(emul_lib.load(), 0, n_games=7, moves=4)
    assert cuda[2] == 0 and emul[2] == 0
    for a, b in zip(cuda[0], emul[0]):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)) and np.array_equal(a[2], b[2])
    assert np.array_equal(cuda[3], emul[3]) and cuda[1] == emul[1]        # tree sizes and the pool's free-page history
