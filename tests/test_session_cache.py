"""`Cache_Wrapper` (Session_Cache.py:4-26) and the `depth` the search hands to it (MCTS.py:224-235,346,468-472).

tests/golden/cache_depths.json holds the depth of every evaluator call the UNMODIFIED reference MCTS / MCTS_Gumbel made
(oracle/gen_cache_golden.py); the facades must make the same calls with the same depths in the same order."""
import json
import os

import numpy as np
import pytest

from grok_alpha_zero_b200 import games
from grok_alpha_zero_b200.MCTS import MCTS
from grok_alpha_zero_b200.MCTS_Gumbel import MCTS_Gumbel
from grok_alpha_zero_b200.session import Cache_Wrapper
from hash_eval import HashSession

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cache_depths.json")))
CLS = {"gomoku": games.Gomoku, "connect4": games.Connect4, "tictactoe": games.TicTacToe}


class Recorder(Cache_Wrapper):
    def __init__(self, session, **kw):
        super().__init__(session, **kw)
        self.depths = []

    def run(self, output_names, input_feed, depth=0):
        self.depths.append(int(depth))
        return super().run(output_names, input_feed, depth)


def drive(case, lib):
    g = CLS[case["game"]]()
    rec = Recorder(HashSession(g.policy_shape[0], logits=case["mode"] == "gumbel", salt=0), max_cache_depth=2)
    if case["mode"] == "puct":
        t = MCTS(g, rec, use_dirichlet=False, tau=0.0, c_puct_init=2.5, lib=lib)
    else:
        t = MCTS_Gumbel(g, rec, use_gumbel_noise=False, m=4, c_visit=50.0, c_scale=1.0, activation_fn="stablemax", lib=lib)
    moves = []
    for _ in range(case["plies"]):
        a, _rows = t.run(iteration_limit=case["limit"], use_bar=False)
        g.do_action(a)
        moves.append(np.asarray(a).reshape(-1).tolist())
        if g.check_win() != -2:
            break
        t.prune_tree(a)
    t.close()
    return rec, moves


def check(case, lib):
    rec, moves = drive(case, lib)
    assert moves == case["moves"]
    assert rec.depths == case["depths"]
    # a fresh cache misses on its first look-up, so every call reached the evaluator, and only depth < 2 was stored
    assert rec.session.calls == len(case["depths"]) and rec.finished_lookup is True
    assert 0 < len(rec.cache) <= sum(d < 2 for d in case["depths"])


@pytest.mark.parametrize("case", GOLD, ids=lambda c: "%s-%s" % (c["game"], c["mode"]))
def test_depths_match_the_reference_emulated(case):
    import emul_lib
    check(case, emul_lib.load())


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD, ids=lambda c: "%s-%s" % (c["game"], c["mode"]))
def test_depths_match_the_reference_cuda(case):
    check(case, None)


class Counting:
    def __init__(self):
        self.calls = 0

    def run(self, output_names, input_feed):
        self.calls += 1
        x = np.asarray(input_feed["inputs"], np.float32)
        return [np.full((1, 3), float(x.sum()), np.float32), np.array([[float(x.max())]], np.float32)]


def feed(v):
    return {"inputs": np.full((1, 2, 2, 1), v, np.float32)}


def test_cache_wrapper_semantics():
    """Session_Cache.py:13-26: look-ups until the first miss, stores only while depth < max_cache_depth"""
    inner = Counting()
    w = Cache_Wrapper(inner, max_cache_depth=2)
    w.run(["policy", "value"], feed(1), depth=0)        # miss -> finished_lookup, stored (depth 0 < 2)
    w.run(["policy", "value"], feed(2), depth=1)        # stored
    w.run(["policy", "value"], feed(3), depth=2)        # NOT stored
    assert inner.calls == 3 and len(w.cache) == 2 and w.finished_lookup is True
    w.run(["policy", "value"], feed(1), depth=0)        # no look-ups after the first miss: evaluated again
    assert inner.calls == 4
    # a second wrapper over the same store (the reference shares a diskcache directory between workers): hits until a miss
    w2 = Cache_Wrapper(inner, max_cache_depth=2)
    w2.cache = w.cache
    o = w2.run(["policy", "value"], feed(2), depth=1)
    assert inner.calls == 4 and float(o[0][0, 0]) == 8.0 and w2.finished_lookup is False
    w2.run(["policy", "value"], feed(3), depth=5)       # miss: evaluated, not stored, look-ups end
    assert inner.calls == 5 and w2.finished_lookup is True and len(w.cache) == 2
    w2.run(["policy", "value"], feed(2), depth=1)
    assert inner.calls == 6
    # max_cache_depth = 0: never looks up, never stores
    w0 = Cache_Wrapper(inner, max_cache_depth=0)
    w0.cache = w.cache
    w0.run(["policy", "value"], feed(1), depth=0)
    assert inner.calls == 7 and len(w.cache) == 2
