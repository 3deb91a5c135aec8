"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GAME_IDS = {"tictactoe": 0, "connect4": 1, "gomoku": 2}
GAME_DIMS = {"tictactoe": (3, 3, 2, 9), "connect4": (6, 7, 4, 7), "gomoku": (15, 15, 2, 225)}  # H, W, C, P
TERM_NONE = 2

EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_int8), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float))


class HashEvalCtx(C.Structure):
    _fields_ = [("P", C.c_int), ("logits", C.c_int), ("salt", C.c_uint64)]


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "mcts_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B", "liboracle.so"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_tree_new.restype = C.c_void_p
        L.orc_tree_new.argtypes = [C.c_int, C.c_int]
        L.orc_tree_free.argtypes = [C.c_void_p]
        L.orc_set_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_puct.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.orc_set_gumbel.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int]
        L.orc_new_root.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_puct_run.argtypes = [C.c_void_p, C.c_int]
        L.orc_puct_run.restype = C.c_int
        L.orc_puct_iterate.argtypes = [C.c_void_p, C.c_int]
        L.orc_puct_iterate.restype = C.c_int
        L.orc_gumbel_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_gumbel_run.restype = C.c_int
        L.orc_prune.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for f in ("orc_root_L", "orc_root_n_exp", "orc_root_term", "orc_puct_best_action", "orc_n_nodes",
                  "orc_root_player"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int
        for f in ("orc_root_visits", "orc_n_evals", "orc_n_sims"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int64
        L.orc_root_stats.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.orc_root_board.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_game_legal.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.orc_game_legal.restype = C.c_int
        L.orc_game_do_action.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_game_check_win.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_game_check_win.restype = C.c_int
        L.orc_game_input_state.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleGame:
    """Minimal live-game mirror (board int8 flat, next_player, history of action ids)."""

    def __init__(self, game):
        self.name = game
        self.gid = GAME_IDS[game]
        self.H, self.W, self.C, self.P = GAME_DIMS[game]
        self.board = np.zeros(self.H * self.W, dtype=np.int8)
        self.next_player = -1
        self.history = []

    def legal(self):
        out = np.zeros(225, dtype=np.int16)
        n = lib().orc_game_legal(self.gid, _p(self.board), _p(out))
        return out[:n].copy()

    def do_action(self, a):
        lib().orc_game_do_action(self.gid, _p(self.board), int(a), self.next_player)
        self.history.append(int(a))
        self.next_player = -self.next_player

    def check_win(self):
        return lib().orc_game_check_win(self.gid, _p(self.board), -self.next_player, self.history[-1])

    def input_state(self):
        out = np.zeros(self.H * self.W * self.C, dtype=np.int8)
        h = np.array(self.history, dtype=np.int16)
        lib().orc_game_input_state(self.gid, _p(self.board), -self.next_player, _p(h), len(self.history), _p(out))
        return out.reshape(self.H, self.W, self.C)


class OracleTree:
    """One search tree of the C oracle (PUCT or Gumbel)."""

    def __init__(self, game, gumbel=False, evaluator=None, salt=0, c_puct_init=2.5, c_puct_base=19652.0, m=16,
                 c_visit=50.0, c_scale=0.1, activation_fn="softmax"):
        self.L = lib()
        self.game = game
        self.gid = GAME_IDS[game]
        self.H, self.W, self.Cc, self.P = GAME_DIMS[game]
        self.gumbel = gumbel
        self.t = self.L.orc_tree_new(self.gid, int(gumbel))
        self.L.orc_set_puct(self.t, c_puct_init, c_puct_base)
        self.L.orc_set_gumbel(self.t, m, c_visit, c_scale, int(activation_fn == "softmax"))
        if evaluator is None:  # built-in C hash evaluator
            self._ctx = HashEvalCtx(self.P, int(gumbel), salt)
            self._cb = None
            self.L.orc_set_eval(self.t, C.cast(self.L.orc_hash_eval, C.c_void_p), C.addressof(self._ctx))
        else:  # python callable (state int8 (H,W,C)) -> (policy f32[P], value)
            def cb(ctx, state, n, policy, value):
                st = np.ctypeslib.as_array(state, shape=(n,)).reshape(self.H, self.W, self.Cc)
                p, v = evaluator(st)
                np.ctypeslib.as_array(policy, shape=(self.P,))[:] = np.asarray(p, dtype=np.float32)
                value[0] = float(np.float32(v))
            self._cb = EVAL_FN(cb)
            self.L.orc_set_eval(self.t, C.cast(self._cb, C.c_void_p), None)

    def __del__(self):
        try:
            self.L.orc_tree_free(self.t)
        except Exception:
            pass

    def new_root(self, g: OracleGame):
        h = np.array(g.history if g.history else [0], dtype=np.int16)
        self.L.orc_new_root(self.t, _p(g.board), g.next_player, _p(h), len(g.history))

    def run(self, iteration_limit, gumbel_noise=None):
        if not self.gumbel:
            self.L.orc_puct_run(self.t, iteration_limit)
            return self.L.orc_puct_best_action(self.t)
        pi = np.zeros(225, dtype=np.float32)
        noise = None if gumbel_noise is None else _p(np.ascontiguousarray(gumbel_noise, dtype=np.float64))
        a = self.L.orc_gumbel_run(self.t, iteration_limit, noise, _p(pi))
        self.pi = pi[: self.L.orc_root_L(self.t)].copy()
        return a

    def iterate(self, n):
        """n more iterations of the PUCT run loop (a slice of a longer run; bench.py's CPU sample)."""
        assert not self.gumbel
        return self.L.orc_puct_iterate(self.t, int(n))

    def prune(self, action, create_new_root=False):
        self.L.orc_prune(self.t, int(action), int(create_new_root))

    def root_stats(self):
        n = self.L.orc_root_L(self.t)
        act = np.zeros(n, np.int16); vis = np.zeros(n, np.uint32); val = np.zeros(n, np.float32)
        pri = np.zeros(n, np.float32); raw = np.zeros(n, np.float32)
        term = np.zeros(n, np.int8); exp = np.zeros(n, np.int8)
        self.L.orc_root_stats(self.t, _p(act), _p(vis), _p(val), _p(pri), _p(raw), _p(term), _p(exp))
        return dict(action=act, visits=vis, values=val, prior=pri, raw=raw, term=term, expanded=exp,
                    root_visits=int(self.L.orc_root_visits(self.t)))

    @property
    def n_evals(self):
        return int(self.L.orc_n_evals(self.t))

    @property
    def n_sims(self):
        return int(self.L.orc_n_sims(self.t))
