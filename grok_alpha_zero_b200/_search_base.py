"""Shared plumbing of the `MCTS` / `MCTS_Gumbel` façades: one engine tree behind the reference's class API.

The façade keeps the reference's calling conventions (it holds a REFERENCE to the caller's live `game` object and
reads `game.board`, `game.action_history`, `game.get_next_player()` where the reference does: MCTS.py:297-312,542,
661-671) and turns every call into C-ABI calls on a one-game engine (include/gaz_b200.h).  The evaluator is the
reference's session duck type `run(output_names, input_feed[, depth]) -> [policy (1,P), value (1,1)]`
(MCTS.py:224-235); a `GazSession` (CUDA network) is recognised and attached so leaves never leave HBM.
"""
import numpy as np

from . import games
from .engine import DIMS, Engine, EngineError


class SearchBase:
    MODE = "puct"

    def _make_engine(self, game, lib, max_nodes, **engine_kw):
        self.game_name = games.game_name_of(game)
        H, W, C, P = DIMS[self.game_name]
        L = 7 if self.game_name == "connect4" else P
        node_cap = int(max_nodes)
        slot_cap = node_cap * L + 256
        self.engine = Engine(self.game_name, n_games=1, mode=self.MODE, trees_per_game=1, node_cap=node_cap,
                             slot_cap=slot_cap, lib=lib, **engine_kw)
        self.P = P
        self._net_attached = False
        net = getattr(self.session, "net", None)
        if net is not None and hasattr(net, "attach") and lib is None:
            net.attach(self.engine)
            self._net_attached = True

    # ---- live game -> engine -------------------------------------------------------------------------------
    def _push_game(self):
        g = self.game
        hist = [games.action_to_id(self.game_name, a) for a in g.action_history]
        self.engine.set_game(0, np.asarray(g.board, dtype=np.int8), int(g.get_next_player()), hist)

    # ---- evaluator boundary (MCTS.py:224-240) -----------------------------------------------------------
    def _legal_mask(self, state):
        if self.game_name == "connect4":
            return state[0, :, 3] == 0                 # top row of the current-board plane
        return state[..., 1].reshape(-1) == 0

    def _host_outputs(self, state, depth):
        if self.session is None:   # MCTS._get_dummy_outputs: uniform random policy / value
            pol = np.random.uniform(low=0, high=1, size=(self.P,)).astype(np.float32, copy=False)
            val = np.random.uniform(low=-1, high=1, size=(1,))[0]
            return pol, np.float32(val)
        kwargs = {"output_names": ["policy", "value"],
                  "input_feed": {"inputs": np.expand_dims(state.astype(np.float32, copy=False), 0)}}
        if getattr(self, "cache_session", False):
            kwargs["depth"] = depth
        policy, value = self.session.run(**kwargs)
        return np.asarray(policy[0], dtype=np.float32), np.float32(np.asarray(value).reshape(-1)[0])

    def _post_policy(self, state, policy):
        """hook: PUCT applies Dirichlet noise to the legal priors here (MCTS.py:243-245,481-482)"""
        return policy

    def _serve(self, n_leaves):
        """answer the outstanding leaf requests and finish their expansion"""
        e = self.engine
        if n_leaves > 0:
            if self._net_attached and not self._needs_host_policy():
                e.eval_net()
            else:
                states, _ = e.get_leaves()
                # depth as the reference passes it to Session_Cache.Cache_Wrapper (MCTS.py:346,468-472)
                depths = e.get_leaf_depths(len(states)) if getattr(self, "cache_session", False) else np.zeros(len(states), np.int32)
                pol = np.zeros((len(states), self.P), np.float32)
                val = np.zeros(len(states), np.float32)
                for i, st in enumerate(states):
                    p, v = self._host_outputs(st, depth=int(depths[i]))
                    pol[i] = self._post_policy(st, p)
                    val[i] = v
                e.put_evals(pol, val)
        e.expand()

    def _needs_host_policy(self):
        return False

    def _check(self):
        st = self.engine.status()
        if st != 0:
            raise EngineError("search tree overflow / bad state (engine status %d); raise max_nodes" % st)

    def _rows_common(self):
        return self.engine.root_stats(0)

    def close(self):
        self.engine.close()
