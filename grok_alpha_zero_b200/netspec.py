"""Architecture description + synthetic weights of the reference's policy/value networks.

Restates the three Keras builders as a flat layer list that both the CUDA loader
(grok_alpha_zero_b200/net.py) and the fp32 oracle (oracle/net_oracle.py) consume:
  Gomoku/Build_Model.py:10-88, Connect4/Build_Model.py:10-88, TicTacToe/Build_Model.py:8-69,
  Net/ResNet/ResNet_Block.py:5-41 (pre-activation block, 1x1 projection when C_in != C_out),
  Net/SE/SE_Block.py:4-23 (optional, ratio 2, applied to conv2's output before the skip add).
Keras semantics kept: NHWC, padding="same", every Conv2D/Dense has a bias, BatchNormalization
with epsilon 1e-3 in inference mode, Dense after Reshape flattens in (H, W, C) order.

Weights are plain numpy arrays in Keras layouts: conv kernel (kh, kw, cin, cout), dense kernel
(in, out), BN gamma/beta/mean/var (C,).  There are no trained checkpoints anywhere:
`init_weights` draws he_normal kernels everywhere (the reference zero-inits a few last layers,
which would make every prior tie), small random biases and non-trivial BN statistics so that
every fused affine is exercised.
"""
import numpy as np

BN_EPS = 1e-3
GAME_SHAPES = {"tictactoe": (3, 3, 2, 9), "connect4": (6, 7, 4, 7), "gomoku": (15, 15, 2, 225)}


def default_config(game):
    if game == "gomoku":      # BASELINE config 3: 10 blocks x 128 (+SE)
        return dict(num_blocks=10, filters=128, stem_filters=256, stem_kernel=3, stem_act="relu", use_se=True)
    if game == "connect4":    # BASELINE config 2: 5 blocks x 128
        return dict(num_blocks=5, filters=128, stem_filters=128, stem_kernel=3, stem_act="gelu", use_se=False)
    return dict(num_blocks=2, filters=64, stem_filters=128, stem_kernel=5, stem_act="gelu", use_se=False)


def build_spec(game, policy_head="softmax", **over):
    """Returns dict(game, H, W, Cin, P, cfg, layers=[...]) ; policy_head in softmax|stablemax|linear."""
    H, W, Cin, P = GAME_SHAPES[game]
    cfg = default_config(game)
    cfg.update(over)
    L = []
    F, SF = cfg["filters"], cfg["stem_filters"]
    L.append(dict(op="stem", name="eyes", k=cfg["stem_kernel"], cin=Cin, cout=SF, bn="eyes_bn", act=cfg["stem_act"]))
    c = SF
    for b in range(cfg["num_blocks"]):
        n = "block%d" % b
        L.append(dict(op="block", name=n, cin=c, cout=F, proj=(c != F), se=cfg["use_se"]))
        c = F
    if game == "gomoku":
        L.append(dict(op="head", name="policy", out="policy", final=policy_head, layers=[
            dict(t="bnrelu", name="policy_bn0", c=F), dict(t="conv", name="policy_conv0", k=3, cin=F, cout=32),
            dict(t="bnrelu", name="policy_bn1", c=32), dict(t="conv", name="policy_conv1", k=3, cin=32, cout=8),
            dict(t="flatten"),
            dict(t="bnrelu", name="policy_bn2", c=H * W * 8), dict(t="dense", name="policy_1", cin=H * W * 8, cout=512),
            dict(t="bnrelu", name="policy_bn3", c=512), dict(t="dense", name="policy_2", cin=512, cout=P)]))
        L.append(dict(op="head", name="value", out="value", final="tanh", layers=[
            dict(t="bnrelu", name="value_bn0", c=F), dict(t="conv", name="value_conv0", k=3, cin=F, cout=32),
            dict(t="bnrelu", name="value_bn1", c=32), dict(t="conv", name="value_conv1", k=1, cin=32, cout=4),
            dict(t="flatten"),
            dict(t="bnrelu", name="value_bn2", c=H * W * 4), dict(t="dense", name="value_1", cin=H * W * 4, cout=256),
            dict(t="bnrelu", name="value_bn3", c=256), dict(t="dense", name="value_2", cin=256, cout=128),
            dict(t="bnrelu", name="value_bn4", c=128), dict(t="dense", name="value_3", cin=128, cout=1)]))
    elif game == "connect4":
        for hn, outc in (("policy", P), ("value", 1)):
            L.append(dict(op="head", name=hn, out=hn, final=policy_head if hn == "policy" else "tanh", layers=[
                dict(t="conv", name=hn + "_conv0", k=3, cin=F, cout=8), dict(t="flatten"),
                dict(t="bnrelu", name=hn + "_bn0", c=H * W * 8), dict(t="dense", name=hn + "_1", cin=H * W * 8, cout=128),
                dict(t="bnrelu", name=hn + "_bn1", c=128), dict(t="dense", name=hn + "_2", cin=128, cout=64),
                dict(t="dense", name=hn + "_3", cin=64, cout=outc)]))
    else:
        L.append(dict(op="head", name="policy", out="policy", final=policy_head, layers=[
            dict(t="conv", name="policy_conv0", k=1, cin=F, cout=8), dict(t="bn", name="policy_bn0", c=8),
            dict(t="flatten"),
            dict(t="dense", name="policy_1", cin=H * W * 8, cout=128), dict(t="relu"),
            dict(t="dense", name="policy_2", cin=128, cout=64), dict(t="dense", name="policy_3", cin=64, cout=P)]))
        L.append(dict(op="head", name="value", out="value", final="tanh", layers=[
            dict(t="conv", name="value_conv0", k=1, cin=F, cout=4), dict(t="bn", name="value_bn0", c=4),
            dict(t="flatten"),
            dict(t="dense", name="value_1", cin=H * W * 4, cout=128), dict(t="dense", name="value_2", cin=128, cout=64),
            dict(t="relu"), dict(t="dense", name="value_3", cin=64, cout=1)]))
    return dict(game=game, H=H, W=W, Cin=Cin, P=P, cfg=cfg, layers=L, policy_head=policy_head)


def _he(rng, shape, fan_in):
    # Keras he_normal = truncated normal (|z| <= 2) with stddev sqrt(2/fan_in)/0.87962566
    std = np.sqrt(2.0 / fan_in) / 0.87962566103423978
    z = rng.standard_normal(size=shape)
    bad = np.abs(z) > 2
    while bad.any():
        z[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(z) > 2
    return (z * std).astype(np.float32)


def _bn(rng, c, w):
    w["gamma"] = rng.uniform(0.8, 1.2, size=c).astype(np.float32)
    w["beta"] = rng.uniform(-0.1, 0.1, size=c).astype(np.float32)
    w["mean"] = rng.uniform(-0.1, 0.1, size=c).astype(np.float32)
    w["var"] = rng.uniform(0.8, 1.25, size=c).astype(np.float32)


def round_bf16(a):
    """Round float32 values to the nearest bfloat16-representable float32 (ties to even)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(a))


def is_tensor_core_conv(cin, cout):
    """Convolutions that run on the tcgen05 path (operands in bf16)."""
    return cin >= 64 and cin % 64 == 0 and cout >= 16


def init_weights(spec, seed=0, residual_gain=0.5, head_gain=0.5, bf16_kernels=True):
    """Synthetic random-init weights (there are no trained checkpoints).
    he_normal kernels everywhere.  Two documented scalings keep a 10-block random stack O(1), which is
    what BN-calibrated trained networks look like and what makes an ABSOLUTE output tolerance
    meaningful: each block's conv2 kernel is scaled by `residual_gain`, the last dense layer of each
    head by `head_gain`.  With `bf16_kernels` the kernels of tensor-core convolutions are rounded to
    bf16-representable values, so the fp32 oracle and the bf16 CUDA trunk use the SAME weights."""
    rng = np.random.default_rng(seed)
    W = {}

    def conv(name, k, cin, cout, gain=1.0):
        kern = _he(rng, (k, k, cin, cout), k * k * cin) * np.float32(gain)
        if bf16_kernels and is_tensor_core_conv(cin, cout):
            kern = round_bf16(kern)
        W[name + ".kernel"] = kern
        W[name + ".bias"] = rng.uniform(-0.05, 0.05, size=cout).astype(np.float32)

    def dense(name, cin, cout, gain=1.0):
        W[name + ".kernel"] = _he(rng, (cin, cout), cin) * np.float32(gain)
        W[name + ".bias"] = rng.uniform(-0.05, 0.05, size=cout).astype(np.float32)

    def bn(name, c):
        d = {}
        _bn(rng, c, d)
        for k2, v in d.items():
            W[name + "." + k2] = v

    for l in spec["layers"]:
        if l["op"] == "stem":
            conv(l["name"], l["k"], l["cin"], l["cout"])
            bn(l["bn"], l["cout"])
        elif l["op"] == "block":
            n = l["name"]
            bn(n + ".bn1", l["cin"])
            conv(n + ".conv1", 3, l["cin"], l["cout"])
            bn(n + ".bn2", l["cout"])
            conv(n + ".conv2", 3, l["cout"], l["cout"], gain=residual_gain)
            if l["proj"]:
                conv(n + ".proj", 1, l["cin"], l["cout"])
            if l["se"]:
                r = l["cout"] // 2
                dense(n + ".se1", l["cout"], r)
                dense(n + ".se2", r, l["cout"])
        else:
            last = [h["name"] for h in l["layers"] if h["t"] == "dense"][-1]
            for h in l["layers"]:
                if h["t"] in ("bnrelu", "bn"):
                    bn(h["name"], h["c"])
                elif h["t"] == "conv":
                    conv(h["name"], h["k"], h["cin"], h["cout"])
                elif h["t"] == "dense":
                    dense(h["name"], h["cin"], h["cout"], head_gain if h["name"] == last else 1.0)
    return W


def bn_affine(W, name):
    """Inference BatchNorm as y = scale * x + shift (float32)."""
    g, b, m, v = (W[name + "." + k].astype(np.float64) for k in ("gamma", "beta", "mean", "var"))
    scale = g / np.sqrt(v + BN_EPS)
    shift = b - m * scale
    return scale.astype(np.float32), shift.astype(np.float32)


def flops_per_eval(spec):
    """Algorithmic FLOPs (2*MACs) of one forward pass."""
    H, W = spec["H"], spec["W"]
    hw = H * W
    f = 0
    for l in spec["layers"]:
        if l["op"] == "stem":
            f += 2 * hw * l["k"] ** 2 * l["cin"] * l["cout"]
        elif l["op"] == "block":
            f += 2 * hw * 9 * l["cin"] * l["cout"] + 2 * hw * 9 * l["cout"] * l["cout"]
            if l["proj"]:
                f += 2 * hw * l["cin"] * l["cout"]
            if l["se"]:
                f += 2 * 2 * l["cout"] * (l["cout"] // 2)
        else:
            for h in l["layers"]:
                if h["t"] == "conv":
                    f += 2 * hw * h["k"] ** 2 * h["cin"] * h["cout"]
                elif h["t"] == "dense":
                    f += 2 * h["cin"] * h["cout"]
    return f
