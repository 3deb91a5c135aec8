"""At which weight scale does the north-star tolerance (policy logits atol 2e-2, value atol 1e-2) stop holding?

Sweeps the residual / head gains of netspec.init_weights from the softened defaults to the builders' own he_normal
scale (gain 1, fp32 kernels) at BASELINE depth and prints, per case, the CUDA network's error against the fp32
restatement next to the error of the CPU bf16-operand simulation (oracle/net_oracle.py, bf16_sim=True) - the floor any
bf16-operand implementation has.  One JSON line per case; DESIGN 1.1 quotes them.   python tools/net_tolerance_sweep.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import net_util  # noqa: E402
from grok_alpha_zero_b200 import netspec  # noqa: E402
from net_oracle import NetOracle  # noqa: E402

CASES = [("gomoku", dict(use_se=False)), ("gomoku", dict(use_se=True)), ("connect4", {})]
GAINS = [(0.25, 0.5, True), (0.5, 0.5, True), (0.5, 0.5, False), (0.7, 0.7, False), (1.0, 1.0, False)]


def main():
    cuda = "--cpu" not in sys.argv
    if cuda:
        from grok_alpha_zero_b200.net import Net
    for game, over in CASES:
        for rg, hg, bk in GAINS:
            spec = netspec.build_spec(game, "linear", **over)
            W = netspec.init_weights(spec, seed=1, residual_gain=rg, head_gain=hg, bf16_kernels=bk)
            st = net_util.random_states(game, 24, seed=4)
            ref = NetOracle(spec, W).forward(st)
            sim = NetOracle(spec, W, bf16_sim=True).forward(st)
            rl, rv = ref["logits"].numpy(), ref["value"].numpy().reshape(-1)
            rec = dict(game=game, over=over, residual_gain=rg, head_gain=hg, bf16_kernels=bk,
                       logit_absmax=float(np.abs(rl).max()),
                       sim_logit_err=float(np.abs(sim["logits"].numpy() - rl).max()),
                       sim_value_err=float(np.abs(sim["value"].numpy().reshape(-1) - rv).max()))
            if cuda:
                net = Net(spec, W, max_batch=32)
                pol, val, lg = net.forward(st, want_logits=True)
                net.close()
                rec.update(cuda_logit_err=float(np.abs(lg - rl).max()), cuda_value_err=float(np.abs(val - rv).max()),
                           cuda_argmax_equal=bool((lg.argmax(-1) == rl.argmax(-1)).all()))
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
