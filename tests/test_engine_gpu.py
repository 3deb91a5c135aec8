"""GPU suite: the CUDA engine through the C ABI against the reference goldens and the C oracle."""
import pytest

import engine_parity as ep
from golden_util import case_id, load

pytestmark = pytest.mark.gpu

PUCT = load("puct.json")
GUMBEL = load("gumbel.json")
# softmax-mode Gumbel depends on exp(): CUDA's exp is not bit-identical to glibc's, so those
# cases are checked separately with the documented "actions/visits equal in practice" bar.
GUMBEL_EXACT = [c for c in GUMBEL if c["activation"] == "stablemax"]
GUMBEL_SOFT = [c for c in GUMBEL if c["activation"] == "softmax"]


@pytest.mark.parametrize("case", PUCT, ids=case_id)
def test_puct_golden_gpu(case):
    ep.puct_golden_case(None, case)


def test_puct_golden_gpu_host_evaluator():
    ep.puct_golden_case(None, PUCT[0], use_host_eval=True)


@pytest.mark.parametrize("case", GUMBEL_EXACT, ids=case_id)
def test_gumbel_stablemax_golden_gpu(case):
    ep.gumbel_golden_case(None, case, check_pi=False)


@pytest.mark.parametrize("case", GUMBEL_SOFT, ids=case_id)
def test_gumbel_softmax_golden_gpu(case):
    ep.gumbel_golden_case(None, case, check_pi=False)


@pytest.mark.parametrize("game,n,iters,plies", [("tictactoe", 256, 60, 6), ("connect4", 256, 300, 30),
                                                ("gomoku", 96, 400, 60)])
def test_puct_batch_vs_oracle_gpu(game, n, iters, plies):
    ep.batch_vs_oracle(None, game, n, iters, seed=3, mode="puct", max_plies=plies, c_puct_init=2.5)


@pytest.mark.parametrize("game,n,iters", [("tictactoe", 64, 16), ("connect4", 64, 48), ("gomoku", 48, 64)])
def test_gumbel_batch_vs_oracle_gpu(game, n, iters):
    ep.batch_vs_oracle(None, game, n, iters, seed=5, mode="gumbel", max_plies=10, m=8, c_visit=50.0, c_scale=1.0,
                       activation_fn="stablemax")
