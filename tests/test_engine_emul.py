"""CPU suite: the engine's warp code (host emulation, width 1) against the reference goldens."""
import pytest

import emul_lib
import engine_parity as ep
from golden_util import case_id, load

PUCT = load("puct.json")
GUMBEL = load("gumbel.json")


@pytest.fixture(scope="module")
def lib():
    return emul_lib.load()


@pytest.mark.parametrize("case", PUCT, ids=case_id)
def test_puct_golden_emul(lib, case):
    ep.puct_golden_case(lib, case)


def test_puct_golden_emul_host_evaluator(lib):
    ep.puct_golden_case(lib, PUCT[0], use_host_eval=True)


@pytest.mark.parametrize("case", GUMBEL, ids=case_id)
def test_gumbel_golden_emul(lib, case):
    ep.gumbel_golden_case(lib, case)


@pytest.mark.parametrize("game,n,iters,plies", [("tictactoe", 24, 60, 6), ("connect4", 16, 150, 30),
                                                ("gomoku", 6, 260, 40)])
def test_puct_batch_vs_oracle_emul(lib, game, n, iters, plies):
    ep.batch_vs_oracle(lib, game, n, iters, seed=3, mode="puct", max_plies=plies, c_puct_init=2.5)


@pytest.mark.parametrize("game,n,iters,act", [("tictactoe", 12, 16, "stablemax"), ("connect4", 8, 48, "softmax"),
                                              ("gomoku", 4, 64, "stablemax")])
def test_gumbel_batch_vs_oracle_emul(lib, game, n, iters, act):
    ep.batch_vs_oracle(lib, game, n, iters, seed=5, mode="gumbel", max_plies=10, m=8, c_visit=50.0, c_scale=1.0,
                       activation_fn=act)
