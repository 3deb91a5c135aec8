"""Minimal stand-in for the `tensorflow.keras` API surface that the reference's model builders touch
(TEST INFRASTRUCTURE ONLY - never imported by the product).

TensorFlow 2.18 (requirements-training-*.txt) is absent from this image, so the reference's networks cannot be run
as they are.  This module lets the UNMODIFIED builder code - `Gomoku/Build_Model.py:10-88`,
`Connect4/Build_Model.py:10-88`, `TicTacToe/Build_Model.py:8-69`, `Net/ResNet/ResNet_Block.py:5-41`,
`Net/SE/SE_Block.py:4-23`, `Net/Stablemax.py:3-11` - execute: `install()` registers a module named `tensorflow`
whose layers are small float32 PyTorch-CPU functions with the published Keras semantics

    Conv2D        NHWC, kernel (kh, kw, cin, cout), bias, stride 1, padding "same"
    Dense         x @ kernel (in, out) + bias
    BatchNormalization   inference mode: gamma * (x - moving_mean) / sqrt(moving_variance + 1e-3) + beta
    Activation    relu | gelu (exact erf, Keras' default approximate=False) | softmax | tanh | sigmoid | linear
    Reshape       row-major over the non-batch axes        GlobalAveragePooling3D   mean over axes 1..3
    layer dtype   a layer built with dtype="float64" casts its input to float64 (Keras autocast)

and whose automatic layer names follow Keras' rule (snake_case class name + per-name counter, assigned in
`Layer.__init__`, e.g. conv2d, conv2d_1, batch_normalization_7, res_net__block).  What this buys: the WIRING of the
networks - which normalisation feeds which convolution, where the skip connections join, the flatten order, the
head activations and their dtypes, the layer creation order behind the checkpoint names - comes from running the
reference's own source, not from our restatement of it.  What it cannot give is TensorFlow's arithmetic: the layer
numerics are the documented Keras definitions evaluated by PyTorch.  `oracle/gen_net_golden.py` uses it to write
`tests/golden/net_*.npz`.

The functional API is traced eagerly: `Input` returns a symbolic tensor carrying a one-row dummy value (so
`.shape` works while the builder runs); every primitive operation appends `(function, input ids, output id)` to a
tape, and `Model.__call__` replays the tape on real data.
"""
import re
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

_TAPE = None          # list of (fn, [arg ids or constants], out id) while a builder runs
_UIDS = {}            # Keras' per-prefix name counters
_DT = {"float32": torch.float32, "float64": torch.float64, None: None}


def reset_names():
    """Keras counts layer names per process; main.py builds each model in a fresh process (Gomoku/main.py:99-139)."""
    _UIDS.clear()


def _snake(name):
    s = re.sub(r"(.)([A-Z][a-z]+)", r"\1_\2", name)
    return re.sub(r"([a-z])([A-Z])", r"\1_\2", s).lower()


def _auto_name(prefix):
    k = _UIDS.get(prefix, 0)
    _UIDS[prefix] = k + 1
    return prefix if k == 0 else "%s_%d" % (prefix, k)


class Sym:
    """symbolic tensor: id on the tape + the dummy value that gives it a shape"""
    _next = 0

    def __init__(self, value):
        self.value = value
        self.id = Sym._next
        Sym._next += 1

    @property
    def shape(self):
        return (None,) + tuple(self.value.shape[1:])

    def _bin(self, other, fn, swap=False):
        return _op((lambda a, b: fn(b, a)) if swap else fn, self, other)

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __iadd__(self, o): return self._bin(o, torch.add)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, torch.div, True)
    def __ge__(self, o): return self._bin(o, torch.ge)


def _val(a):
    return a.value if isinstance(a, Sym) else a


def _op(fn, *args):
    """run `fn` on the dummy values now and remember it for the replay"""
    def run(*vals):
        fdt = next((v.dtype for v in vals if torch.is_tensor(v) and v.is_floating_point()), torch.float32)
        vals = [v if torch.is_tensor(v) else torch.tensor(v, dtype=fdt) for v in vals]
        return fn(*vals)
    out = Sym(run(*[_val(a) for a in args]))
    _TAPE.append((run, [("s", a.id) if isinstance(a, Sym) else ("c", a) for a in args], out.id))
    return out


class Variable:
    def __init__(self, path, shape, init):
        self.path = path
        self.name = path.rsplit("/", 1)[-1]
        self.t = torch.full(shape, float(init), dtype=torch.float32)

    @property
    def shape(self):
        return tuple(self.t.shape)

    def numpy(self):
        return self.t.numpy().copy()

    def assign(self, a):
        a = torch.as_tensor(np.asarray(a), dtype=torch.float32)
        assert tuple(a.shape) == tuple(self.t.shape), (self.path, tuple(a.shape), tuple(self.t.shape))
        self.t.copy_(a)


_SCOPE = []            # names of the custom layers whose call() is running (variable path prefix)
_REGISTRY = None       # every primitive layer object created while a builder runs, in creation order


class Layer:
    def __init__(self, name=None, dtype=None, **kwargs):
        self.name = name or _auto_name(_snake(type(self).__name__))
        self.dtype = dtype
        self.built = False
        self._vars = []
        if _REGISTRY is not None:
            _REGISTRY.append(self)

    # -- Keras hooks ---------------------------------------------------------------------------------
    def build(self, input_shape):
        self.built = True

    def call(self, inputs):
        raise NotImplementedError

    def add_var(self, name, shape, init):
        v = Variable("/".join(_SCOPE + [self.name, name]), shape, init)
        self._vars.append(v)
        return v

    def __call__(self, inputs, *a, **k):
        if not self.built:
            self.build(inputs.shape)
            self.built = True
        if type(self).forward is not Layer.forward:      # primitive layer: one tape entry
            dt = _DT[self.dtype]
            return _op(lambda x: self.forward(x if dt is None else x.to(dt)), inputs)
        _SCOPE.append(self.name)                         # custom layer (reference code): its call() records itself
        try:
            return self.call(inputs, *a, **k)
        finally:
            _SCOPE.pop()

    def forward(self, x):  # overridden by primitive layers
        raise NotImplementedError


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", kernel_initializer=None, **kw):
        super().__init__(**kw)
        ks = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        assert tuple(strides) == (1, 1) and padding == "same" and ks[0] == ks[1] and ks[0] % 2 == 1
        self.filters, self.k = int(filters), ks[0]

    def build(self, input_shape):
        self.kernel = self.add_var("kernel", (self.k, self.k, input_shape[-1], self.filters), 0.0)
        self.bias = self.add_var("bias", (self.filters,), 0.0)

    def forward(self, x):
        w = self.kernel.t.to(x.dtype).permute(3, 2, 0, 1)
        y = F.conv2d(x.permute(0, 3, 1, 2), w, self.bias.t.to(x.dtype), padding=self.k // 2)
        return y.permute(0, 2, 3, 1)


class Dense(Layer):
    def __init__(self, units, kernel_initializer=None, bias_initializer=None, **kw):
        super().__init__(**kw)
        self.units = int(units)

    def build(self, input_shape):
        self.kernel = self.add_var("kernel", (input_shape[-1], self.units), 0.0)
        self.bias = self.add_var("bias", (self.units,), 0.0)

    def forward(self, x):
        return x @ self.kernel.t.to(x.dtype) + self.bias.t.to(x.dtype)


class BatchNormalization(Layer):
    def __init__(self, epsilon=1e-3, **kw):
        super().__init__(**kw)
        self.epsilon = epsilon

    def build(self, input_shape):
        c = input_shape[-1]
        self.gamma = self.add_var("gamma", (c,), 1.0)
        self.beta = self.add_var("beta", (c,), 0.0)
        self.moving_mean = self.add_var("moving_mean", (c,), 0.0)
        self.moving_variance = self.add_var("moving_variance", (c,), 1.0)

    def forward(self, x):
        g, b, m, v = (t.t.to(x.dtype) for t in (self.gamma, self.beta, self.moving_mean, self.moving_variance))
        return (x - m) * (g / torch.sqrt(v + self.epsilon)) + b


class Activation(Layer):
    def __init__(self, activation, **kw):
        super().__init__(**kw)
        self.fn = {"relu": F.relu, "gelu": F.gelu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "linear": lambda x: x,
                   "softmax": lambda x: torch.softmax(x, dim=-1)}[activation]

    def forward(self, x):
        return self.fn(x)


class Reshape(Layer):
    def __init__(self, target_shape, **kw):
        super().__init__(**kw)
        self.target = tuple(int(t) for t in target_shape)

    def forward(self, x):
        return x.reshape((x.shape[0],) + self.target)


class GlobalAveragePooling3D(Layer):
    def forward(self, x):
        return x.mean(dim=(1, 2, 3))


def Input(batch_shape=None, shape=None, name=None, **kw):
    dims = tuple(batch_shape[1:]) if batch_shape is not None else tuple(shape)
    s = Sym(torch.zeros((1,) + dims, dtype=torch.float32))
    s.input_name = name
    return s


class Model:
    def __init__(self, inputs=None, outputs=None, **kw):
        self.inputs, self.outputs = inputs, list(outputs)
        self.tape = list(_TAPE)
        self.layers = list(_REGISTRY)

    @property
    def weights(self):
        return [v for l in self.layers for v in l._vars]

    @torch.no_grad()
    def __call__(self, x, training=False):
        env = {self.inputs.id: torch.as_tensor(np.asarray(x), dtype=torch.float32)}
        for fn, args, out in self.tape:
            env[out] = fn(*[env[a] if k == "s" else a for k, a in args])
        return [env[o.id] for o in self.outputs]

    def predict(self, x, **kw):
        return [o.numpy() for o in self(x)]


def begin_trace():
    """start a fresh tape / layer registry / name counters (one model build = one process in the reference)"""
    global _TAPE, _REGISTRY
    _TAPE, _REGISTRY = [], []
    del _SCOPE[:]
    reset_names()


def install():
    """register the stand-in as `tensorflow` (and stub `Net.Grok_Model`, whose custom train_step classes the builders
    import but do not need for inference)"""
    tf = types.ModuleType("tensorflow")
    keras = types.ModuleType("tensorflow.keras")
    layers = types.ModuleType("tensorflow.keras.layers")
    for c in (Layer, Conv2D, Dense, BatchNormalization, Activation, Reshape, GlobalAveragePooling3D):
        setattr(layers, c.__name__, c)
    layers.Input = Input
    keras.layers, keras.Model = layers, Model
    tf.keras = keras
    tf.where = lambda c, a, b: _op(torch.where, c, a, b)
    tf.reduce_sum = lambda x, axis=None, keepdims=False: _op(lambda t: t.sum(dim=axis, keepdim=keepdims), x)
    tf.math = types.ModuleType("tensorflow.math")

    def divide_no_nan(a, b):
        return _op(lambda x, y: torch.where(y == 0, torch.zeros_like(x / y), x / y), a, b)
    tf.math.divide_no_nan = divide_no_nan
    gm = types.ModuleType("Net.Grok_Model")
    for n in ("Grok_Fast_EMA_Model", "Ortho_Model", "Ortho_Grok_Fast_EMA_Model"):
        setattr(gm, n, Model)
    sys.modules.update({"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.layers": layers,
                        "tensorflow.math": tf.math, "Net.Grok_Model": gm})
    return tf
