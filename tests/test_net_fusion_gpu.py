"""Round-2 network kernels at their edges (GPU, through the C ABI): trunk launches (several residual blocks per launch of
gaz_block::res_trunk_kernel), the fused stem + shortcut projection, the tile stem, the dx-merged head convolutions.
The launch counts pin the fusion rules of gaz_net_create; the numerical bar is BASELINE's (logits 2e-2, value 1e-2) against
the fp32 restatement, and rows must not depend on what else is in the batch (bit for bit)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import net_util  # noqa: E402
from grok_alpha_zero_b200 import netspec  # noqa: E402
from grok_alpha_zero_b200.net import Net  # noqa: E402
from net_oracle import NetOracle  # noqa: E402

pytestmark = pytest.mark.gpu


def _tags(net):
    return [t[4] for t in net.op_shapes() if t[4]]


def test_launch_counts_pin_the_fusion_rules():
    # Gomoku 10 x 128 + SE: stem+projection | 6 + 4 blocks | per head: dx-merged C128->C32, small head conv, tensor-core dense
    # layer (all its 128-wide slices in one launch) | policy_2 as a one-layer mma chain | policy_out | value dense chain
    # = 12 kernels per forward
    spec = netspec.build_spec("gomoku", "softmax")
    net = Net(spec, netspec.init_weights(spec, seed=0), max_batch=64)
    tags = _tags(net)
    assert tags.count("intrunk") == 8 and "block+se*6" in tags and "block+se*4" in tags, tags
    assert net.n_launches == 12, net.n_launches
    net.close()
    # Connect4 5 x 128: tile stem | one trunk launch | both head convolutions | both dense stacks | policy_out
    spec = netspec.build_spec("connect4", "softmax")
    net = Net(spec, netspec.init_weights(spec, seed=0), max_batch=64)
    assert "block+*5" in _tags(net), _tags(net)
    assert net.n_launches == 5, net.n_launches
    net.close()


CASES = [
    ("gomoku", dict(num_blocks=7, use_se=True), 9),        # two trunk launches: 6 blocks + 1 block
    ("gomoku", dict(num_blocks=6, use_se=False), 5),       # exactly one full launch, no SE
    ("connect4", dict(num_blocks=7), 13),                  # 6 + 1 blocks, partial last tile (13 boards = 3 tiles + 1 board)
    ("connect4", {}, 1),                                   # a single board in a single tile
]


@pytest.mark.parametrize("game,over,n", CASES, ids=lambda c: str(c).replace(" ", ""))
def test_edge_batches_hold_the_tolerance_and_rows_are_independent(game, over, n):
    spec = netspec.build_spec(game, "softmax", **over)
    W = netspec.init_weights(spec, seed=2)
    st = net_util.random_states(game, n, seed=11)
    ref = NetOracle(spec, W).forward(st)
    # max_batch far above n: most CTAs own no tile (n_my = 0), the rest one
    net = Net(spec, W, max_batch=700)
    pol, val, lg = net.forward(st, want_logits=True)
    assert np.isfinite(lg).all() and np.isfinite(val).all()
    assert np.abs(lg - ref["logits"].numpy()).max() <= 2e-2
    assert np.abs(val - ref["value"].numpy().reshape(-1)).max() <= 1e-2
    # the same positions in another order, behind 301 other boards: identical bits per position
    other = net_util.random_states(game, 301, seed=12)
    perm = np.random.RandomState(0).permutation(n)
    pol2, val2, lg2 = net.forward(np.concatenate([other, st[perm]]), want_logits=True)
    np.testing.assert_array_equal(lg2[301:], lg[perm])
    np.testing.assert_array_equal(val2[301:], val[perm])
    np.testing.assert_array_equal(pol2[301:], pol[perm])
    net.close()
