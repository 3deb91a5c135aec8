"""smoke(): closed loop on cuda:0 - a few search rounds with the CUDA network as the evaluator, and the
network's outputs against the fp32 oracle on the same leaves (BASELINE tolerance)."""
import os
import sys

import numpy as np


def run_flagship():
    """The flagship kernels first, so that they sit at the head of any launch trace of smoke(): one forward pass of a
    Gomoku 2-block + Squeeze-Excitation network (the layer shapes of BASELINE configs[2]: tcgen05 stem + shortcut projection
    (stem_proj_kernel), both residual blocks with SE in one trunk launch (res_trunk_kernel), dx-merged C128->C32 head
    convolutions (conv_board_kernel<96>), mma.sync head convolutions and dense layers, tensor-core dense) on 6 boards,
    checked against the fp32 restatement with the north-star tolerance."""
    from . import netspec
    from .net import Net
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "oracle"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from net_oracle import NetOracle
    import net_util
    spec = netspec.build_spec("gomoku", "softmax", num_blocks=2, use_se=True)
    W = netspec.init_weights(spec, seed=1)
    st = net_util.random_states("gomoku", 6, seed=4)
    net = Net(spec, W, max_batch=8)
    tags = [t[4] for t in net.op_shapes()]
    assert any(t.startswith("block") for t in tags), "the fused residual-block kernel is not selected: %r" % (tags,)
    pol, val, lg = net.forward(st, want_logits=True)
    net.close()
    ref = NetOracle(spec, W).forward(st)
    el = float(np.abs(lg - ref["logits"].numpy()).max())
    ev = float(np.abs(val - ref["value"].numpy().reshape(-1)).max())
    assert el < 2e-2 and ev < 1e-2, (el, ev)
    print("flagship net smoke ok: gomoku 2 blocks + SE, max |dlogit| %.4f |dvalue| %.4f" % (el, ev))


def run():
    from . import netspec
    from .engine import Engine
    from .net import Net
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle"))
    from net_oracle import NetOracle  # the checker (smoke is one of the places allowed to use it)

    spec = netspec.build_spec("connect4", "softmax", num_blocks=2)
    W = netspec.init_weights(spec, seed=1)
    eng = Engine("connect4", n_games=64, mode="puct", trees_per_game=1, c_puct_init=2.5, iters_hint=64)
    net = Net(spec, W, max_batch=64)
    net.attach(eng)
    if eng.new_roots() > 0:
        eng.eval_net()
    eng.expand()
    eng.run_begin([24] * 64)
    eng.rounds_net(8)
    n = eng.select()
    st, _ = eng.get_leaves()
    assert n == len(st) and n > 0
    pol, val, lg = net.forward(st, want_logits=True)
    ref = NetOracle(spec, W).forward(st)
    assert np.abs(lg - ref["logits"].numpy()).max() < 2e-2, "policy logits outside atol 2e-2"
    assert np.abs(val - ref["value"].numpy().reshape(-1)).max() < 1e-2, "value outside atol 1e-2"
    eng.eval_net()
    eng.expand()
    vis, _, info = eng.root_dense()
    assert int(info[:, 2].min()) == 9 and eng.status() == 0
    eng.close()
    net.close()
    print("net smoke ok: %d leaves, max |dlogit| %.4f" % (n, float(np.abs(lg - ref["logits"].numpy()).max())))
