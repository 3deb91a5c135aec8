"""Generate tests/golden/*.json by running the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container only:   python oracle/gen_golden.py
The reference (/root/reference) is imported through oracle/ref_shim.py and driven
by the deterministic hash evaluator (oracle/hash_eval.py).  Recorded per root
decision: searching tree, played action, every root child's (action, visits,
float32 value-sum bits, float32 prior bits), root.visits, evaluator calls.

Drive pattern = Self_Play.play (Self_Play.py:71-156): PUCT uses two trees per
game, both re-rooted after every ply (prune_tree); Gumbel rebuilds the tree each
move ("fresh") or re-roots it ("prune").
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from hash_eval import HashSession, hash_eval  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
DIMS = {"tictactoe": (3, 3), "connect4": (6, 7), "gomoku": (15, 15)}


def f32bits(x):
    return int(np.array([x], dtype=np.float32).view(np.uint32)[0])


def act_id(game, a):
    if game == "connect4":
        return int(a)
    H, W = DIMS[game]
    return int(a[1]) * W + int(a[0])


def root_record(game, tree, gumbel):
    r = tree.root
    rec = {"root_visits": int(r.visits), "children": []}
    n = len(r.child_visits)
    for i in range(n):
        ch = r.children[i] if i < len(r.children) else None
        if ch is None:
            if gumbel:
                a = act_id(game, r.child_legal_actions[i])
            else:  # unexpanded PUCT tail: the deque holds the remaining actions in prior order
                a = act_id(game, list(r.child_legal_actions)[i - len(r.children)])
            term = 2
        else:
            a = act_id(game, ch.action_history[-1])
            term = 2 if ch.is_terminal is None else int(ch.is_terminal)
        pri = r.child_logit_priors[i] if gumbel else r.child_prob_priors[i]
        row = [a, int(r.child_visits[i]), f32bits(r.child_values[i]), f32bits(pri), term]
        if gumbel:
            row.append(f32bits(r.child_raw_values[i]))
        rec["children"].append(row)
    return rec


def puct_game(ref, game, sims, salt, c_puct, max_plies):
    Game = ref.games[game]
    g = Game()
    P = g.policy_shape[0]
    sess = HashSession(P, logits=False, salt=salt)
    kw = dict(use_dirichlet=False, tau=0.0, c_puct_init=c_puct, fast_find_win=False)
    t1 = ref.MCTS(g, sess, **kw)
    t2 = ref.MCTS(g, sess, **kw)
    moves = []
    winner = -2
    while winner == -2 and len(moves) < max_plies:
        tree = t1 if g.get_next_player() == -1 else t2
        which = 1 if tree is t1 else 2
        action, rows = tree.run(iteration_limit=int(sims * 1.5), time_limit=None, use_bar=False)
        rec = root_record(game, tree, False)
        rec["tree"] = which
        rec["action"] = act_id(game, action)
        rec["evals"] = sess.calls
        g.do_action(action)
        winner = int(g.check_win())
        rec["winner"] = winner
        moves.append(rec)
        if winner == -2:
            t1.prune_tree(action, False)
            t2.prune_tree(action, False)
            rec["after_prune"] = [root_record(game, t1, False), root_record(game, t2, False)]
    return {"game": game, "mode": "puct", "sims": sims, "iters": int(sims * 1.5), "salt": salt,
            "c_puct_init": c_puct, "moves": moves}


def gumbel_game(ref, game, n, m, salt, activation, c_visit, c_scale, reuse, max_plies):
    Game = ref.games[game]
    g = Game()
    P = g.policy_shape[0]
    sess = HashSession(P, logits=True, salt=salt)

    def mk():
        return ref.MCTS_Gumbel(g, sess, use_gumbel_noise=False, m=m, c_visit=c_visit, c_scale=c_scale,
                               activation_fn=activation)

    tree = mk()
    moves = []
    winner = -2
    while winner == -2 and len(moves) < max_plies:
        action, rows = tree.run(iteration_limit=n, time_limit=None, use_bar=False)
        rec = root_record(game, tree, True)
        rec["action"] = act_id(game, action)
        rec["evals"] = sess.calls
        # final pi' in slot order (rows are sorted by pi'; rebuild from actions)
        pi_by_action = {act_id(game, r[0]): f32bits(r[1]) for r in rows}
        rec["pi"] = [pi_by_action[c[0]] for c in rec["children"]]
        g.do_action(action)
        winner = int(g.check_win())
        rec["winner"] = winner
        moves.append(rec)
        if winner == -2:
            if reuse:
                tree.prune_tree(action, False)
            else:
                tree = mk()
    return {"game": game, "mode": "gumbel", "n": n, "m": m, "salt": salt, "activation": activation,
            "c_visit": c_visit, "c_scale": c_scale, "reuse": reuse, "moves": moves}


def game_kats(ref, game, seed, n_games):
    """Random play-outs through the reference game classes: legal actions, check_win, input states."""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n_games):
        g = ref.games[game]()
        plies = []
        while True:
            legal = g.get_legal_actions()
            a = legal[rng.randint(len(legal))]
            g.do_action(a)
            w = int(g.check_win())
            st = np.asarray(g.get_input_state()).astype(np.int8)
            plies.append({"action": act_id(game, a), "winner": w, "n_legal_before": int(len(legal)),
                          "state_sum": int((st.astype(np.int64).reshape(-1) * np.arange(1, st.size + 1)).sum()),
                          "state": st.reshape(-1).tolist() if len(plies) < 6 else None})
            if w != -2 or len(g.get_legal_actions()) == 0:
                break
        out.append(plies)
    return out


def eval_kats():
    rng = np.random.RandomState(7)
    out = []
    for (n, P, logits) in [(18, 9, False), (168, 7, False), (450, 225, False), (450, 225, True), (168, 7, True)]:
        for salt in (0, 3):
            st = rng.randint(-1, 2, size=n).astype(np.int8)
            p, v = hash_eval(st, P, logits, salt)
            out.append({"state": st.tolist(), "P": P, "logits": logits, "salt": salt,
                        "policy_bits": [f32bits(x) for x in p], "value_bits": f32bits(v)})
    return out


def main():
    ref = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    cases = []
    # PUCT: whole TicTacToe / Connect4 games, Gomoku openings (BASELINE configs 1-3 sims)
    for salt in (0, 1, 2):
        cases.append(puct_game(ref, "tictactoe", 200, salt, 1.25, 9))
    for salt in (0, 1):
        cases.append(puct_game(ref, "connect4", 800, salt, 2.5, 42))
    cases.append(puct_game(ref, "connect4", 100, 5, 2.5, 42))
    cases.append(puct_game(ref, "gomoku", 800, 0, 4.5, 8))
    cases.append(puct_game(ref, "gomoku", 300, 1, 4.5, 30))
    with gzip.open(os.path.join(OUT, "puct.json.gz"), "wt") as f:
        json.dump(cases, f, separators=(",", ":"))
    print("puct cases:", len(cases), "root decisions:", sum(len(c["moves"]) for c in cases))

    gcases = []
    for act in ("stablemax", "softmax"):
        for reuse in (False, True):
            gcases.append(gumbel_game(ref, "gomoku", 64, 16, 0, act, 50.0, 1.0, reuse, 14))
            gcases.append(gumbel_game(ref, "connect4", 64, 7, 1, act, 50.0, 1.0, reuse, 42))
            gcases.append(gumbel_game(ref, "tictactoe", 16, 4, 2, act, 50.0, 2.0, reuse, 9))
    gcases.append(gumbel_game(ref, "gomoku", 200, 32, 3, "stablemax", 50.0, 0.1, False, 6))
    with gzip.open(os.path.join(OUT, "gumbel.json.gz"), "wt") as f:
        json.dump(gcases, f, separators=(",", ":"))
    print("gumbel cases:", len(gcases), "root decisions:", sum(len(c["moves"]) for c in gcases))

    kats = {"eval": eval_kats(),
            "games": {g: game_kats(ref, g, 11, 12 if g != "gomoku" else 6) for g in ("tictactoe", "connect4", "gomoku")}}
    with gzip.open(os.path.join(OUT, "kats.json.gz"), "wt") as f:
        json.dump(kats, f, separators=(",", ":"))
    print("kats written")


if __name__ == "__main__":
    main()
