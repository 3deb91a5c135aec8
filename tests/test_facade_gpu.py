"""GPU suite for the reference-facing layer: façades, device noise, state read-back and the batched self-play driver
through the CUDA library (the CPU suite runs the same drivers on the host emulation)."""
import numpy as np
import pytest

import test_facade_cpu as F
import test_noise_and_states as N
from golden_util import case_id
from grok_alpha_zero_b200 import games, netspec
from grok_alpha_zero_b200.MCTS import MCTS
from grok_alpha_zero_b200.Self_Play import BatchedSelfPlay, ReplayWriter, run_self_play
from grok_alpha_zero_b200.session import GazSession

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [F.PUCT[0], F.PUCT[5], F.PUCT[7]], ids=case_id)
def test_mcts_facade_goldens_on_cuda(case):
    F.drive_puct(case, None, max_plies=10)


def test_gumbel_facade_goldens_on_cuda():
    for case in [c for c in F.GUMBEL if not c["reuse"] and c["activation"] == "stablemax"][:3]:
        F.drive_gumbel(case, None, check_pi=False)   # CUDA exp differs from glibc in the last bit


@pytest.mark.parametrize("alpha", [0.3, 0.05])
def test_dirichlet_noise_statistics_on_cuda(alpha):
    N.test_dirichlet_noise_statistics(None, alpha)


def test_noise_keys_and_state_round_trip_on_cuda():
    N.test_noise_is_reproducible_and_keyed(None)
    for name in ("tictactoe", "connect4", "gomoku"):
        N.test_get_states_and_set_games_round_trip(None, name)


def test_facade_with_attached_cuda_net_equals_host_session_path():
    """the same network through the device-resident path (leaves never leave HBM) and through session.run per leaf"""
    spec = netspec.build_spec("connect4", "softmax", num_blocks=2)
    W = netspec.init_weights(spec, seed=3)
    sess = GazSession(spec, W, max_batch=8)

    class HostOnly:   # hides .net so the façade uses the host evaluator boundary
        def run(self, output_names, input_feed, **kw):
            return sess.run(output_names, input_feed)

    res = []
    for s in (sess, HostOnly()):
        g = games.Connect4()
        t = MCTS(g, s, use_dirichlet=False, tau=0.0, c_puct_init=2.5)
        moves = []
        for _ in range(6):
            a, rows = t.run(iteration_limit=60, use_bar=False)
            moves.append((int(a), [(int(r[0]), int(r[4])) for r in rows]))
            g.do_action(a)
            if g.check_win() != -2:
                break
            t.prune_tree(a)
        t.close()
        res.append(moves)
    assert res[0] == res[1]
    sess.close()


def test_batched_self_play_with_cuda_net(tmp_path):
    bc = {"num_resnet_layers": 2, "num_filters": 128, "use_stablemax": False}
    tc = dict(MCTS_iteration_limit=40, use_gumbel=False, c_puct_init=2.5, dirichlet_alpha=0.5, max_actions=42,
              num_explore_actions_first=3, num_explore_actions_second=2, games_per_generation=12, games_per_gpu=8)
    spec = netspec.build_spec("connect4", "softmax", num_blocks=2)
    W = netspec.init_weights(spec, seed=1)
    runs = []
    for slots in (8, 5):
        sp = BatchedSelfPlay(games.Connect4, bc, tc, list(range(12)), n_slots=slots, evaluator="net", spec=spec, weights=W,
                             seed=9)
        runs.append({g["game_id"]: g for g in sp.play()})
        sp.close()
    for i in range(12):
        a, b = runs[0][i], runs[1][i]
        assert a["winner"] == b["winner"] and np.array_equal(a["states"], b["states"]) and \
            np.array_equal(a["policies"], b["policies"])
        assert np.allclose(a["policies"].sum(1), 1.0, atol=1e-5)
    out = str(tmp_path / "0")
    merged = run_self_play(games.Connect4, (bc, tc, {}), out, weights=W, seed=9)
    assert len(merged) == 12
    w = ReplayWriter(out)
    assert w.games_done() == 12
    run_self_play(games.Connect4, (bc, tc, {}), out, weights=W, seed=9)     # resume rule: nothing left to play
    assert ReplayWriter(out).games_done() == 12


def test_pool_relief_on_cuda_matches_the_emulated_engine():
    """Small tree pools force fresh roots (BatchedSelfPlay._relieve_full_pools, gaz_tree_sizes): the CUDA engine and the
    host emulation of the same warp code must play identical games, with the same number of rebuilds."""
    import emul_lib
    bc = {"num_resnet_layers": 1, "num_filters": 128, "use_stablemax": False}
    tc = dict(MCTS_iteration_limit=40, use_gumbel=False, c_puct_init=2.5, dirichlet_alpha=0.5, max_actions=42,
              num_explore_actions_first=1, num_explore_actions_second=1)
    runs = []
    for lib in (None, emul_lib.load()):
        sp = BatchedSelfPlay(games.Connect4, bc, tc, list(range(6)), 6, evaluator="hash", lib=lib, seed=3,
                             node_cap=150, slot_cap=150 * 7, use_noise=False)
        fin = {g["game_id"]: g for g in sp.play()}
        runs.append((fin, sp.pool_rebuilds))
        sp.close()
    (a, ra), (b, rb) = runs
    assert ra == rb and ra > 0
    for i in range(6):
        assert a[i]["winner"] == b[i]["winner"] and np.array_equal(a[i]["states"], b[i]["states"])
        assert np.array_equal(a[i]["policies"], b[i]["policies"])
