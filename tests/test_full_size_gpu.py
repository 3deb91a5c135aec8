"""GPU suite at BASELINE.json's full sizes (16384 Gomoku games / 4096 Connect4 games per GPU, network batch
16384), checked through size-independent properties -- the C oracle cannot search 16384 trees in seconds:

* replication: position k is loaded into every game g with g % K == k; all replicas must produce bit-identical
  root statistics (a tree's result may not depend on which warp / slot / CTA owns it), and the K distinct
  results are compared with the C oracle bit for bit;
* bookkeeping: the iteration count follows MCTS.py:543-546, child visits sum to the root's visits (fresh root),
  |value sum| <= visits, the tau=0 move is the first arg-max of the visit counts in slot order (MCTS.py:603);
* network: a batch of 16384 states made of 64 distinct states repeated gives row-identical outputs that equal
  the small-batch forward bit for bit and stay within the bf16 tolerance of the fp32 oracle.
"""
import numpy as np
import pytest

import engine_parity as ep
import net_util
import oracle as orc
from golden_util import f32bits
from grok_alpha_zero_b200 import netspec
from grok_alpha_zero_b200.engine import Engine
from grok_alpha_zero_b200.net import Net
from net_oracle import NetOracle

pytestmark = pytest.mark.gpu


def _replicated_search(game, n_games, k_distinct, iters, mode, plies, lib=None, **kw):
    rng = np.random.RandomState(11)
    gumbel = mode == "gumbel"
    distinct = [ep.random_position(game, rng, rng.randint(0, plies + 1)) for _ in range(k_distinct)]
    eng = Engine(game, n_games=n_games, mode=mode, trees_per_game=1, iters_hint=iters * (4 if gumbel else 1), lib=lib, **kw)
    H, W = eng.H, eng.W
    boards = np.zeros((n_games, H, W), np.int8)
    nxt = np.zeros(n_games, np.int32)
    hl = np.zeros(n_games, np.int32)
    last = np.full(n_games, -1, np.int32)
    for g in range(n_games):
        p = distinct[g % k_distinct]
        boards[g] = np.asarray(p.board).reshape(H, W)
        nxt[g] = p.next_player
        hl[g] = len(p.history)
    if game == "connect4":
        # Connect4's win check looks at the last action (Connect4.py:380-411): use the per-game upload
        for g in range(n_games):
            p = distinct[g % k_distinct]
            eng.set_game(g, p.board, p.next_player, p.history)
    else:
        eng.set_games(boards, nxt, hist_lens=hl, last_actions=last)
    ep.serve_roots(eng, eng.new_roots(), 0, gumbel)
    ep.run_all(eng, [iters] * n_games, 0, gumbel)
    vis, val, info = eng.root_dense()
    assert eng.status() == 0
    # bookkeeping
    if not gumbel:
        # MCTS.py:543-546: one legal move -> 1 iteration, a limit below the number of legal moves -> 3 x legal moves
        n_legal = np.array([len(distinct[g % k_distinct].legal()) for g in range(n_games)])
        want = np.where(n_legal == 1, 1, np.where(iters < n_legal, 3 * n_legal, iters))
        assert (info[:, 2] == want).all()
        assert (info[:, 0] >= info[:, 2]).all()          # terminal expansions may back up more than one visit
        assert (vis.sum(1) == info[:, 0]).all()          # fresh root: every visit passed through one child
        slot_first_max = np.array([int(eng.root_stats(k)["action"][int(np.argmax(eng.root_stats(k)["visits"]))])
                                   for k in range(k_distinct)])
        assert (info[:k_distinct, 1] == slot_first_max).all()
    assert (np.abs(val) <= vis + 1e-6).all()
    # replication
    for k in range(k_distinct):
        rows = np.arange(k, n_games, k_distinct)
        assert (vis[rows] == vis[k]).all(), "replicas of position %d differ in visit counts" % k
        assert (f32bits(val[rows]) == f32bits(val[k])).all(), "replicas of position %d differ in value sums" % k
        assert (info[rows] == info[k]).all()
    # the distinct positions against the oracle
    okw = {a: b for a, b in kw.items()}
    for k, p in enumerate(distinct):
        t = orc.OracleTree(game, gumbel, salt=0, **okw)
        t.new_root(p)
        t.run(iters)
        ref = t.root_stats()
        dv = np.zeros(eng.P, np.uint32)
        dw = np.zeros(eng.P, np.float32)
        dv[ref["action"]] = ref["visits"]
        dw[ref["action"]] = ref["values"]
        np.testing.assert_array_equal(vis[k], dv, err_msg="position %d" % k)
        np.testing.assert_array_equal(f32bits(val[k]), f32bits(dw), err_msg="position %d" % k)
    eng.close()


def test_gomoku_16384_games_replicated_puct():
    _replicated_search("gomoku", 16384, 32, 300, "puct", 40, c_puct_init=4.5)


def test_connect4_4096_games_replicated_puct():
    _replicated_search("connect4", 4096, 32, 300, "puct", 20, c_puct_init=2.5)


def test_gomoku_16384_games_replicated_gumbel():
    _replicated_search("gomoku", 16384, 16, 64, "gumbel", 30, m=16, c_visit=50.0, c_scale=1.0,
                       activation_fn="stablemax")


def test_network_batch_16384_rows_are_independent():
    spec = netspec.build_spec("gomoku", "softmax")
    Wt = netspec.init_weights(spec, seed=1)
    base = net_util.random_states("gomoku", 64, seed=9)
    big = np.ascontiguousarray(np.tile(base, (256, 1, 1, 1)))
    net = Net(spec, Wt, max_batch=16384)
    pol, val, lg = net.forward(big, want_logits=True)
    assert pol.shape == (16384, 225)
    for r in range(1, 256):
        assert (f32bits(lg[r * 64:(r + 1) * 64]) == f32bits(lg[:64])).all(), "repeat %d differs" % r
        assert (f32bits(val[r * 64:(r + 1) * 64]) == f32bits(val[:64])).all()
    pol_s, val_s, lg_s = net.forward(base, want_logits=True)
    np.testing.assert_array_equal(f32bits(lg_s), f32bits(lg[:64]))
    np.testing.assert_array_equal(f32bits(val_s), f32bits(val[:64]))
    ref = NetOracle(spec, Wt).forward(base)
    assert np.abs(lg[:64] - ref["logits"].numpy()).max() <= 2e-2      # BASELINE.json: policy logits atol 2e-2
    assert np.abs(val[:64] - ref["value"].numpy().reshape(-1)).max() <= 1e-2   # value atol 1e-2
    np.testing.assert_allclose(pol.sum(1), 1.0, atol=1e-5)
    net.close()
