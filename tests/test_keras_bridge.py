"""CPU suite: Keras checkpoint bridge -- automatic layer names replayed from the reference builders'
creation order, import/export round trip, loud failures on missing or mis-shaped variables."""
import numpy as np
import pytest

from grok_alpha_zero_b200 import keras_bridge as kb
from grok_alpha_zero_b200 import netspec


def test_gomoku_names_follow_creation_order():
    spec = netspec.build_spec("gomoku", "softmax", use_se=False)       # Gomoku/Build_Model.py:21-86, 10 blocks
    n = kb.keras_layer_names(spec)
    assert n["eyes"] == "conv2d" and n["eyes_bn"] == "batch_normalization"
    # ResNet_Block.__init__ creates bn1, conv1, bn2, conv2, residual_conv (Net/ResNet/ResNet_Block.py:11-20)
    assert (n["block0.bn1"], n["block0.conv1"], n["block0.bn2"], n["block0.conv2"], n["block0.proj"]) == \
        ("batch_normalization_1", "conv2d_1", "batch_normalization_2", "conv2d_2", "conv2d_3")
    assert "block1.proj" not in n                                       # created (conv2d_6) but never built
    assert n["block1.conv1"] == "conv2d_4" and n["block9.conv2"] == "conv2d_29"
    assert n["policy_bn0"] == "batch_normalization_21" and n["policy_conv0"] == "conv2d_31"
    assert n["policy_conv1"] == "conv2d_32" and n["value_conv0"] == "conv2d_33" and n["value_conv1"] == "conv2d_34"
    assert n["policy_1"] == "policy_1" and n["value_3"] == "value_3"    # explicit names in the Gomoku builder
    assert n["value_bn4"] == "batch_normalization_29"


def test_connect4_and_tictactoe_dense_layers_are_auto_named():
    n = kb.keras_layer_names(netspec.build_spec("connect4", "softmax"))
    assert [n["policy_1"], n["policy_2"], n["policy_3"], n["value_1"], n["value_2"], n["value_3"]] == \
        ["dense", "dense_1", "dense_2", "dense_3", "dense_4", "dense_5"]
    assert n["policy_conv0"] == "conv2d_16" and n["value_conv0"] == "conv2d_17"     # 1 + 5 * 3 convolutions before
    n = kb.keras_layer_names(netspec.build_spec("tictactoe", "softmax"))
    assert n["block0.proj"] == "conv2d_3" and n["policy_conv0"] == "conv2d_7" and n["value_bn0"] == "batch_normalization_6"


@pytest.mark.parametrize("game,over,prefix", [("gomoku", dict(use_se=False, num_blocks=3), "res_net__block/"),
                                              ("connect4", {}, ""), ("tictactoe", {}, "functional/")])
def test_round_trip(game, over, prefix):
    spec = netspec.build_spec(game, "softmax", **over)
    W = netspec.init_weights(spec, seed=5)
    exported = kb.export_keras_weights(spec, W, prefix=prefix)
    exported = {k + ":0": v for k, v in exported.items()}              # Keras-2 spelling of a variable name
    back = kb.import_keras_weights(spec, exported)
    assert set(back) == set(W)
    for k in W:
        np.testing.assert_array_equal(back[k], W[k])
    assert set(kb.expected_shapes(spec)) == set(W)


def test_failures_are_loud():
    spec = netspec.build_spec("connect4", "softmax")
    W = netspec.init_weights(spec, seed=5)
    ex = kb.export_keras_weights(spec, W)
    missing = dict(ex)
    del missing["dense_3/bias"]
    with pytest.raises(KeyError):
        kb.import_keras_weights(spec, missing)
    wrong = dict(ex)
    wrong["conv2d_1/kernel"] = np.zeros((3, 3, 128, 64), np.float32)
    with pytest.raises(ValueError):
        kb.import_keras_weights(spec, wrong)
    with pytest.raises(ValueError):
        kb.keras_layer_names(netspec.build_spec("gomoku", "softmax"))   # use_se=True: no Keras counterpart


def test_net_rejects_weights_of_another_build_config():
    """Net() validates every array before anything is laid out for the device (no GPU needed to fail)."""
    from grok_alpha_zero_b200.net import check_weights
    spec64 = netspec.build_spec("tictactoe", "softmax")
    spec128 = netspec.build_spec("tictactoe", "softmax", filters=128)
    W = netspec.init_weights(spec64, seed=1)
    check_weights(spec64, W)
    with pytest.raises(ValueError, match="expects"):
        check_weights(spec128, W)
    se = netspec.build_spec("gomoku", "softmax", num_blocks=1, use_se=True)
    Wse = netspec.init_weights(se, seed=1)
    check_weights(se, Wse)
    del Wse["block0.se2.bias"]
    with pytest.raises(ValueError, match="missing"):
        check_weights(se, Wse)


def test_spec_from_reference_configs():
    """build_config -> netspec as the reference's builders read it: TicTacToe hard-codes ResNet_Block(64)
    (TicTacToe/Build_Model.py:22) and ignores num_filters; Connect4 / Gomoku use it."""
    from grok_alpha_zero_b200.Self_Play import net_spec_from_configs
    bc = {"num_resnet_layers": 2, "num_filters": 128, "use_stablemax": False}
    assert net_spec_from_configs("tictactoe", bc, {"use_gumbel": False})["cfg"]["filters"] == 64
    assert net_spec_from_configs("connect4", dict(bc, num_filters=96), {"use_gumbel": False})["cfg"]["filters"] == 96
    assert net_spec_from_configs("gomoku", bc, {"use_gumbel": True})["policy_head"] == "linear"


def test_reference_configs_build_the_reference_architecture():
    """`run_self_play(Gomoku, <reference configs>, folder)`: the reference's build_config has no `use_se` key and its
    builders never wire SE_Block in (Net/ResNet/ResNet_Block.py:27-41), so the spec must be SE-free, must map onto the
    Keras variable names, and a checkpoint exported through the bridge must load back."""
    from grok_alpha_zero_b200.Self_Play import net_spec_from_configs
    ref_build_config = {"num_resnet_layers": 4, "num_filters": 128, "rr_alpha": 0.05, "mixed_precision": None,
                        "use_stablemax": False}                                    # Gomoku/Gomoku.py:5-12
    spec = net_spec_from_configs("gomoku", ref_build_config, {"use_gumbel": False})
    assert spec["cfg"]["use_se"] is False and spec["cfg"]["num_blocks"] == 4
    vm = kb.keras_variable_map(spec)
    W = netspec.init_weights(spec, seed=3)
    assert set(vm) == set(W)                                                       # every array has a Keras name; no SE arrays
    back = kb.import_keras_weights(spec, kb.export_keras_weights(spec, W))
    assert all(np.array_equal(back[k], W[k]) for k in W)
    # opt-in stays possible for the benchmark configuration (BASELINE configs[2])
    assert net_spec_from_configs("gomoku", dict(ref_build_config, use_se=True), {})["cfg"]["use_se"] is True
