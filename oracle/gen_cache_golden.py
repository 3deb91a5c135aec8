"""Writes tests/golden/cache_depths.json: the `depth` argument the reference's MCTS / MCTS_Gumbel hand to a caching session
(MCTS.py:224-235,346,468-472; MCTS_Gumbel.py:262-272) call by call (TEST INFRASTRUCTURE ONLY, build container).

The session is a subclass of the reference's own `Session_Cache.Cache_Wrapper` (so `isinstance` at MCTS.py:102 holds) whose
`run` only records `depth` and forwards to the hash evaluator: `Cache_Wrapper.run` itself calls `ndarray.newbyteorder`,
which numpy 2 removed, so the reference's caching body cannot execute in this image - its semantics (Session_Cache.py:
13-26) are restated in grok_alpha_zero_b200/session.py and unit-tested in tests/test_session_cache.py.
    python oracle/gen_cache_golden.py [--check]
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden", "cache_depths.json")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from hash_eval import HashSession  # noqa: E402

CASES = [("connect4", "puct", 60, 2), ("tictactoe", "puct", 30, 3), ("gomoku", "puct", 40, 1), ("connect4", "gumbel", 16, 2)]


def build():
    ref = ref_shim.load()
    from Session_Cache import Cache_Wrapper

    class Recorder(Cache_Wrapper):
        def __init__(self, session):
            super().__init__(session, tempfile.mkdtemp(prefix="gaz_cache_"), max_cache_depth=2)
            self.depths = []

        def run(self, output_names, input_feed, depth):
            self.depths.append(int(depth))
            return self.session.run(output_names, input_feed)

    out = []
    for game, mode, limit, plies in CASES:
        g = ref.games[game]()
        rec = Recorder(HashSession(g.policy_shape[0], logits=mode == "gumbel", salt=0))
        if mode == "puct":
            t = ref.MCTS(g, rec, use_dirichlet=False, tau=0.0, c_puct_init=2.5)
        else:
            t = ref.MCTS_Gumbel(g, rec, use_gumbel_noise=False, m=4, c_visit=50.0, c_scale=1.0, activation_fn="stablemax")
        moves = []
        for _ in range(plies):
            np.random.seed(0)
            a, _rows = t.run(iteration_limit=limit, use_bar=False)
            g.do_action(a)
            moves.append(np.asarray(a).reshape(-1).tolist())
            if g.check_win() != -2:
                break
            t.prune_tree(a)
        out.append(dict(game=game, mode=mode, limit=limit, plies=plies, moves=moves, depths=rec.depths))
    return out


if __name__ == "__main__":
    data = build()
    if "--check" in sys.argv:
        assert json.load(open(OUT)) == data
        print("ok cache_depths")
    else:
        json.dump(data, open(OUT, "w"))
        print("wrote", OUT, [len(c["depths"]) for c in data])
