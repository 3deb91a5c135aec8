"""`MCTS_Gumbel` - drop-in for the reference's Gumbel-AlphaZero search (MCTS_Gumbel.py:151-733) on the B200 engine.

Same constructor (`use_gumbel_noise`, `m`, `c_visit`, `c_scale`, `activation_fn`), `run` / `prune_tree` /
`update_hyperparams`, row format `[action, pi', winrate, value_sum, visits, logit_prior, root_visits, is_terminal]`
sorted by pi', and the reference's quirks: `time_limit` is ignored with a warning (:576-578), `m` is permanently
clipped to the number of legal moves (:581-582), root-child expansions are not counted in the budget (:626-628),
the final pi' always uses softmax (:655-662).  In StableMax mode and without Gumbel noise the visits, value sums
and moves are bit-identical to the reference under identical evaluator outputs; with noise the Gumbel(0,1) draw
comes from numpy's global stream like the reference's (`np.random.gumbel`, :592-594).
One deliberate difference: `prune_tree` on a move that was never expanded builds a fresh root (the reference's
`fill_empty_children` leaves evaluation-less stub children behind that `_set_root` cannot search from).
"""
from warnings import warn

import numpy as np

from . import games
from ._search_base import SearchBase
from .engine import TERM_DRAW, TERM_NONE


class MCTS_Gumbel(SearchBase):
    MODE = "gumbel"

    def __init__(self, game, session, use_gumbel_noise=False, use_njit=None, m=16, c_visit=50.0, c_scale=0.1,
                 activation_fn="softmax", fast_find_win=False, max_nodes=8192, lib=None):
        self.game = game
        self.session = session
        self.cache_session = any(c.__name__ == "Cache_Wrapper" for c in type(session).__mro__)   # MCTS.py:102 isinstance
        self.fast_find_win = fast_find_win
        self.use_gumbel_noise = use_gumbel_noise
        self.use_njit = use_njit
        self.m = m
        self.c_visit = c_visit
        self.c_scale = c_scale
        self.use_softmax = activation_fn == "softmax"
        self._make_engine(game, lib, max_nodes, m=int(m), c_visit=float(c_visit), c_scale=float(c_scale),
                          activation_fn=activation_fn)
        self.create_expand_root()

    def update_hyperparams(self, *args, **kwargs) -> None:
        m = kwargs.get("m")
        if m is not None:
            self.m = m
        c_scale = kwargs.get("c_scale")
        if c_scale is not None:
            self.c_scale = c_scale
        c_visit = kwargs.get("c_visit")
        if c_visit is not None:
            self.c_visit = c_visit
        self.engine.set_gumbel_params(int(self.m), float(self.c_visit), float(self.c_scale), self.use_softmax)

    def create_expand_root(self):
        self._push_game()
        self._serve(self.engine.new_roots())
        self._check()

    def run(self, iteration_limit=None, time_limit=None, use_bar=True):
        n_legal = len(self.game.get_legal_actions())
        if (iteration_limit is not None and iteration_limit is True) and time_limit is None:
            iteration_limit = n_legal * 3
        if time_limit is not None:
            iteration_limit = n_legal * 3
            warn("Time limit isn't allowed for gumbel MCTS defaulting to use 3 * len_legal_actions")
        if self.m > n_legal:
            self.m = n_legal
        assert iteration_limit > 0
        e = self.engine
        st0 = e.root_stats(0)
        if self.use_gumbel_noise:
            e.set_gumbel_noise(np.random.gumbel(loc=0.0, scale=1.0, size=(st0["L"],)).reshape(1, -1))
        else:
            e.set_gumbel_noise(None)
        bar = None
        if use_bar:
            from tqdm import tqdm
            bar = tqdm(total=iteration_limit)
        e.run_begin([int(iteration_limit)])
        while e.remaining() > 0:
            self._serve(e.select())
            if bar is not None:
                bar.update(1)
        if bar is not None:
            bar.close()
        self._check()

        st = e.root_stats(0)
        L = st["L"]
        pi = e.gumbel_pi(0)[:L]
        name = self.game_name
        mover = self.game.get_next_player()
        visits, values = st["visits"], st["values"]
        with np.errstate(divide="ignore", invalid="ignore"):
            mean_values = np.where(visits > 0, values / visits, -1.0).astype(np.float32, copy=False)
        mask = visits == 0
        mean_values[mask] = pi[mask]
        rows = []
        for i in range(L):
            term = int(st["term"][i])
            is_terminal = None if term == TERM_NONE else (0 if term == TERM_DRAW else mover)
            rows.append([games.id_to_action(name, int(st["action"][i])), pi[i], mean_values[i], values[i], visits[i],
                         st["prior"][i], st["root_visits"], is_terminal])
        best = st["best_slot"] if st["best_slot"] >= 0 else int(np.argmax(pi))
        move = rows[best][0]
        rows = sorted(rows, key=lambda x: x[1], reverse=True)
        return move, rows

    def prune_tree(self, action, create_new_root=False):
        self._push_game()
        self._serve(self.engine.prune([games.action_to_id(self.game_name, action)], create_new_root=create_new_root))
        self._check()
