"""`MCTS` - drop-in for the reference's PUCT search class (MCTS.py:75-671) on the B200 engine.

    from grok_alpha_zero_b200.MCTS import MCTS          # instead of `from MCTS import MCTS`
    mcts = MCTS(game, session, c_puct_init=2.5, tau=1.0)
    move, rows = mcts.run(iteration_limit=1200, use_bar=False)
    game.do_action(move); mcts.prune_tree(move)

Same constructor arguments, `run` / `prune_tree` / `update_hyperparams` semantics, row format
`[action, prob, winrate, value_sum, visits, prior, root_visits, is_terminal]` sorted by visits, tau rules
(tau < 5e-3 -> 0, `np.random.choice` consumed on every run, MCTS.py:601-616) and budget rules (one legal move -> 1
iteration, limit < n_legal -> 3*n_legal, MCTS.py:543-546).  With identical evaluator outputs, no Dirichlet noise
and tau = 0 the visit counts, value sums and chosen moves are bit-identical to the reference
(tests/test_facade_*.py).  Differences, all in the random parts: Dirichlet noise is drawn on the host per evaluated
leaf from numpy's global stream (same distribution, different stream position than the reference), and a terminal
parent's winning child is the lowest-index one instead of `np.random.randint` (MCTS.py:208).
"""
import time
from warnings import warn

import numpy as np

from . import games
from ._search_base import SearchBase
from .engine import TERM_DRAW, TERM_NONE


class MCTS(SearchBase):
    MODE = "puct"

    def __init__(self, game, session=None, use_njit=None, c_puct_init: float = 2.5, c_puct_base: float = 19_652,
                 use_dirichlet=True, dirichlet_alpha=1.11, dirichlet_epsilon=0.25, tau=1.0, fast_find_win=False,
                 max_nodes=16384, lib=None):
        self.game = game
        self.session = session
        self.cache_session = any(c.__name__ == "Cache_Wrapper" for c in type(session).__mro__)   # MCTS.py:102 isinstance
        self.fast_find_win = fast_find_win   # the device look-ahead always scans every reply (same results)
        self.use_njit = use_njit             # accepted for signature compatibility; there is no numba path
        self.c_puct_init = c_puct_init
        self.c_puct_base = c_puct_base
        self.use_dirichlet = use_dirichlet
        self.dirichlet_alpha = dirichlet_alpha
        self.dirichlet_epsilon = dirichlet_epsilon
        if tau != 0.0 and tau < 5e-3:
            warn("Tau cannot be smaller than 5e-3 as it will cause floating point errors")
            warn("If you want the most visited move, set tau = 0.0, defaulting tau to 0.0")
            tau = 0.0
        self.tau = tau
        self._make_engine(game, lib, max_nodes, c_puct_init=float(c_puct_init), c_puct_base=float(c_puct_base))
        self.create_expand_root()

    # ---- hyper-parameters (MCTS.py:134-168) -----------------------------------------------------------------
    def update_hyperparams(self, **kwargs) -> None:
        c_puct_init = kwargs.get("c_puct_init")
        if c_puct_init is not None:
            if c_puct_init < 0.0:
                warn(f"c_puct_init value is invalid, {c_puct_init} cannot be negative.")
            else:
                self.c_puct_init = c_puct_init
        c_puct_base = kwargs.get("c_puct_base")
        if c_puct_base is not None:
            if c_puct_base <= 0:
                warn("c_puct_base cannot be negative")
            else:
                self.c_puct_base = c_puct_base
        if c_puct_init is not None or c_puct_base is not None:
            self.engine.set_puct_params(float(self.c_puct_init), float(self.c_puct_base))
        dirichlet_alpha = kwargs.get("dirichlet_alpha")
        if dirichlet_alpha is not None:
            if dirichlet_alpha <= 0.0:
                warn("dirichlet_alpha cannot be less than or equal to 0")
            else:
                self.dirichlet_alpha = dirichlet_alpha
        dirichlet_epsilon = kwargs.get("dirichlet_epsilon")
        if dirichlet_epsilon is not None:
            if dirichlet_epsilon < 0.0 or dirichlet_epsilon >= 1.0:
                warn("dirichlet_epsilon cannot be negative nor bigger than 1")
            else:
                self.dirichlet_epsilon = dirichlet_epsilon
        tau = kwargs.get("tau")
        if tau is not None:
            if tau != 0.0 and tau <= 5e-3:
                warn("Tau can't be less than 5e-3. Changing tau = 0.0")
                tau = 0.0
            self.tau = tau

    # ---- evaluator hooks ------------------------------------------------------------------------------------
    def _needs_host_policy(self):
        return bool(self.use_dirichlet)

    def _post_policy(self, state, policy):
        if not self.use_dirichlet:
            return policy
        legal = self._legal_mask(state)
        p = policy[legal]
        p = p / games._seq_sum(p)
        noise = np.random.dirichlet(self.dirichlet_alpha * np.ones_like(p))
        p = ((1 - self.dirichlet_epsilon) * p + self.dirichlet_epsilon * noise).astype(np.float32, copy=False)
        out = np.zeros_like(policy)
        out[legal] = p
        return out

    # ---- root (MCTS.py:296-365) --------------------------------------------------------------------------
    def create_expand_root(self):
        self._push_game()
        self._serve(self.engine.new_roots())
        self._check()

    # ---- search (MCTS.py:528-618) ------------------------------------------------------------------------
    def run(self, iteration_limit=None, time_limit=None, use_bar=True):
        len_legal_actions = len(self.game.get_legal_actions())
        if len_legal_actions == 1:
            iteration_limit = 1
        elif (iteration_limit is not None and (iteration_limit is True or iteration_limit < len_legal_actions)) \
                and time_limit is None:
            iteration_limit = len_legal_actions * 3
        elif iteration_limit is None and time_limit is True:
            time_limit = 30.0
        bar = None
        if use_bar:
            from tqdm import tqdm
            bar = tqdm(total=iteration_limit if (iteration_limit and time_limit is None) else time_limit)
        e = self.engine
        e.run_begin([int(iteration_limit) if iteration_limit is not None else (1 << 30)])
        start = time.time()
        done = 0
        while e.remaining() > 0:
            if time_limit is not None and time.time() - start >= time_limit:
                break
            self._serve(e.select())
            done += 1
            if bar is not None:
                bar.update(1 if (iteration_limit and time_limit is None) else 0)
        if bar is not None:
            bar.close()
        self._check()

        st = e.root_stats(0)
        name = self.game_name
        mover = self.game.get_next_player()
        k = st["n_expanded"]
        visits = st["visits"][:k]
        values = st["values"][:k]
        with np.errstate(divide="ignore", invalid="ignore"):
            probs = visits / np.sum(visits)
            winrates = values / visits
        move_probs = []
        for i in range(k):
            term = int(st["term"][i])
            is_terminal = None if term == TERM_NONE else (0 if term == TERM_DRAW else mover)
            move_probs.append([games.id_to_action(name, int(st["action"][i])), probs[i], winrates[i], values[i],
                               visits[i], st["prior"][i], st["root_visits"], is_terminal])
        if self.tau == 0.0:
            prob_weights = np.zeros_like(visits)
            prob_weights[np.argmax(visits)] = 1.0
        else:
            exp = np.array(1.0 / self.tau, dtype=np.float64)
            prob_weights = (visits.astype(np.float64) ** exp) / (np.array(st["root_visits"], np.float64) ** exp)
            prob_weights /= np.sum(prob_weights)
            prob_weights = np.array(prob_weights, np.float64)
        chosen_index = np.random.choice(np.arange(len(move_probs)), size=1, replace=False, p=prob_weights)[0]
        move = move_probs[chosen_index][0]
        move_probs = sorted(move_probs, key=lambda x: x[4], reverse=True)
        return move, move_probs

    # ---- re-rooting (MCTS.py:657-671) --------------------------------------------------------------------
    def prune_tree(self, action, create_new_root=False):
        self._push_game()   # the caller has already played `action` on the live game
        self._serve(self.engine.prune([games.action_to_id(self.game_name, action)], create_new_root=create_new_root))
        self._check()
