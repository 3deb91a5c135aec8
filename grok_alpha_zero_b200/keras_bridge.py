"""Weight bridge from the reference's Keras checkpoints to the CUDA network.

The reference keeps `model.weights.h5` (Keras) next to `model.onnx` (`Gomoku/main.py:139,166`).  Neither h5py,
TensorFlow nor onnx exist in this image, so the bridge is a flat `{keras path: array}` dictionary that a
TensorFlow-equipped machine writes with three lines:

    model.load_weights("Grok_Zero_Train/<gen>/model.weights.h5")
    np.savez("keras_weights.npz", **{w.path: w.numpy() for w in model.weights})

and `import_keras_weights(spec, dict(np.load("keras_weights.npz")))` turns into the `netspec` weight dictionary
(`Net(spec, W)`, `save_checkpoint`).  Arrays stay in Keras layouts (conv (kh, kw, cin, cout), dense (in, out), BN
gamma / beta / moving_mean / moving_variance); `net.build_ops` does the re-layout for the tcgen05 operands.

Layer names.  Only the Dense layers of the Gomoku builder carry explicit names (`Gomoku/Build_Model.py:43-83`);
everything else gets Keras' automatic names (`conv2d`, `conv2d_1`, ..., `batch_normalization_7`, `dense_2`), which
count layer OBJECTS in creation order per class.  `keras_layer_names` replays the creation order of the three
builders: stem conv + BN, then per `ResNet_Block.__init__` (`Net/ResNet/ResNet_Block.py:11-20`) bn1, conv1, bn2,
conv2 and `residual_conv` (created for every block, built only when C_in != C_out), then the policy head and the value
head in source order.  `main.py` builds the model in a fresh process (`Gomoku/main.py:99-139`), so the counters start
at zero.  A path is matched by its last two components (`<layer>/<variable>`), which is what stays stable between the
Keras 2 (`res_net__block/conv2d_1/kernel:0`) and Keras 3 (`res_net__block/conv2d_1/kernel`) spellings.

The reference's block has no Squeeze-Excitation (`SE_Block` is defined but not wired into `ResNet_Block.call`), so a
spec with `use_se=True` has no Keras counterpart for the SE dense layers and is rejected.
"""
import numpy as np

_BN_VARS = (("gamma", "gamma"), ("beta", "beta"), ("mean", "moving_mean"), ("var", "moving_variance"))


class _Counter:
    def __init__(self):
        self.n = {}

    def next(self, prefix):
        k = self.n.get(prefix, 0)
        self.n[prefix] = k + 1
        return prefix if k == 0 else "%s_%d" % (prefix, k)


def keras_layer_names(spec):
    """{netspec layer name: Keras layer name} for the model the reference's build_model creates for `spec`."""
    if spec["cfg"].get("use_se"):
        raise ValueError("the reference's ResNet_Block has no SE layers: a use_se spec has no Keras checkpoint")
    c = _Counter()
    names = {}
    explicit_dense = spec["game"] == "gomoku"      # Gomoku/Build_Model.py names its Dense layers
    for l in spec["layers"]:
        if l["op"] == "stem":
            names[l["name"]] = c.next("conv2d")
            names[l["bn"]] = c.next("batch_normalization")
        elif l["op"] == "block":
            n = l["name"]
            names[n + ".bn1"] = c.next("batch_normalization")
            names[n + ".conv1"] = c.next("conv2d")
            names[n + ".bn2"] = c.next("batch_normalization")
            names[n + ".conv2"] = c.next("conv2d")
            proj = c.next("conv2d")                  # residual_conv exists in every block
            if l["proj"]:
                names[n + ".proj"] = proj
        else:
            for h in l["layers"]:
                if h["t"] == "conv":
                    names[h["name"]] = c.next("conv2d")
                elif h["t"] in ("bn", "bnrelu"):
                    names[h["name"]] = c.next("batch_normalization")
                elif h["t"] == "dense":
                    names[h["name"]] = h["name"] if explicit_dense else c.next("dense")
    return names


def keras_variable_map(spec):
    """{netspec weight key: '<keras layer>/<keras variable>'} for every array `Net(spec, W)` consumes."""
    out = {}
    for ours, keras in keras_layer_names(spec).items():
        if keras.startswith("batch_normalization"):
            for a, b in _BN_VARS:
                out["%s.%s" % (ours, a)] = "%s/%s" % (keras, b)
        else:
            out[ours + ".kernel"] = keras + "/kernel"
            out[ours + ".bias"] = keras + "/bias"
    return out


def _suffix(path):
    p = path.split(":")[0].strip("/").split("/")
    return "/".join(p[-2:])


def expected_shapes(spec):
    """{netspec weight key: shape} in Keras layouts."""
    sh = {}

    def bn(n, c):
        for a, _ in _BN_VARS:
            sh["%s.%s" % (n, a)] = (c,)

    for l in spec["layers"]:
        if l["op"] == "stem":
            sh[l["name"] + ".kernel"] = (l["k"], l["k"], l["cin"], l["cout"])
            sh[l["name"] + ".bias"] = (l["cout"],)
            bn(l["bn"], l["cout"])
        elif l["op"] == "block":
            n = l["name"]
            bn(n + ".bn1", l["cin"])
            sh[n + ".conv1.kernel"] = (3, 3, l["cin"], l["cout"])
            sh[n + ".conv1.bias"] = (l["cout"],)
            bn(n + ".bn2", l["cout"])
            sh[n + ".conv2.kernel"] = (3, 3, l["cout"], l["cout"])
            sh[n + ".conv2.bias"] = (l["cout"],)
            if l["proj"]:
                sh[n + ".proj.kernel"] = (1, 1, l["cin"], l["cout"])
                sh[n + ".proj.bias"] = (l["cout"],)
        else:
            for h in l["layers"]:
                if h["t"] == "conv":
                    sh[h["name"] + ".kernel"] = (h["k"], h["k"], h["cin"], h["cout"])
                    sh[h["name"] + ".bias"] = (h["cout"],)
                elif h["t"] in ("bn", "bnrelu"):
                    bn(h["name"], h["c"])
                elif h["t"] == "dense":
                    sh[h["name"] + ".kernel"] = (h["cin"], h["cout"])
                    sh[h["name"] + ".bias"] = (h["cout"],)
    return sh


def import_keras_weights(spec, keras_arrays):
    """`keras_arrays`: {keras weight path: array} (e.g. dict(np.load('keras_weights.npz'))).
    Returns the netspec weight dictionary; raises KeyError / ValueError on a missing variable or a shape mismatch
    (a mismatch means the checkpoint was built with a different build_config than `spec`)."""
    by_suffix = {}
    for path, arr in keras_arrays.items():
        s = _suffix(path)
        if s in by_suffix:
            raise ValueError("two Keras variables end in %r: cannot match by layer/variable" % s)
        by_suffix[s] = np.asarray(arr)
    shapes = expected_shapes(spec)
    W = {}
    for ours, keras in keras_variable_map(spec).items():
        if keras not in by_suffix:
            raise KeyError("Keras checkpoint has no variable %r (needed for %s)" % (keras, ours))
        a = by_suffix[keras]
        if tuple(a.shape) != shapes[ours]:
            raise ValueError("%s: Keras variable %r has shape %s, the spec expects %s"
                             % (ours, keras, tuple(a.shape), shapes[ours]))
        W[ours] = np.ascontiguousarray(a, dtype=np.float32)
    return W


def export_keras_weights(spec, W, prefix=""):
    """Inverse of import_keras_weights: {'<prefix><layer>/<variable>': array} (for round-trip tests and for pushing
    weights back into a Keras model with `variable.assign`)."""
    return {prefix + keras: np.asarray(W[ours]) for ours, keras in keras_variable_map(spec).items()}
