import numpy as np

import oracle as orc


def random_states(game, n, seed):
    """n input states (int8 HWC) of random reachable positions, via the oracle's game rules."""
    rng = np.random.RandomState(seed)
    maxp = {"tictactoe": 5, "connect4": 30, "gomoku": 80}[game]
    out = []
    while len(out) < n:
        g = orc.OracleGame(game)
        plies = rng.randint(0, maxp + 1)
        ok = True
        for _ in range(plies):
            legal = g.legal()
            g.do_action(int(legal[rng.randint(len(legal))]))
            if g.check_win() != -2:
                ok = False
                break
        if ok:
            out.append(g.input_state())
    return np.stack(out).astype(np.int8)
