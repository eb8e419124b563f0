"""Minimal emulation of the TensorFlow 2.0 eager API surface that pierremtb/PINNs-TF2.0 touches.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Purpose: let the reference's OWN, UNMODIFIED source files
(utils/neuralnetwork.py, utils/custom_lbfgs.py, utils/logger.py and the *InformedNN classes of the example scripts) execute
in this container, where TensorFlow 2.0.0-rc0 cannot be installed, so that the restated oracle can be pinned against what
the reference's code computes (tests/golden/make_reference_fixtures.py).  The arithmetic back end is torch CPU float64
autograd; what is emulated is TensorFlow's *semantics*:

* ``GradientTape`` records only while it is active: a gradient is ``None`` unless the target is connected to the source
  through operations executed inside the tape's context (``_reach`` bookkeeping below), sources must be watched tensors or
  trainable variables, a non-persistent tape serves one ``gradient`` call, gradient computations are themselves recorded by
  whatever tapes are active at that moment (nested / persistent tapes give higher-order derivatives), ``output_gradients``
  is supported (the dummy-gradient trick of 1d-burgers/inf_disc_burgers.py:66-75).
* ``tf.keras``: ``Sequential`` (``layers`` hides the ``InputLayer``), ``InputLayer``, ``Lambda``, ``Dense`` with
  ``glorot_normal`` (TF 2.0 VarianceScaling: truncated normal, stddev sqrt(2/(fan_in+fan_out))/0.87962566), zero biases,
  ``get_weights/set_weights``, ``trainable_variables`` (a fresh list per access, kernel then bias per layer);
  ``optimizers.Adam`` with the OptimizerV2 update (epsilon=None -> 1e-7, epsilon outside the bias correction).
* eager tensors: immutable values, ``x += y`` rebinds, ``.numpy()``, ``__format__`` of scalars, truthiness of scalar
  comparisons, numpy operands accepted on either side.

Known limits (documented, not hidden): TensorFlow's random stream cannot be reproduced (parity runs load weights);
float32 numpy operands mixed with float64 tensors are promoted exactly (real TF may refuse the mixed-dtype MatMul of
inf_disc_burgers.py:83); only the symbols the reference uses exist.
"""
import itertools
import sys
import types

import numpy as np
import torch

__version__ = "2.0.0-rc0 (API emulation on torch %s, oracle/tf_emulation)" % torch.__version__
float64, float32 = "float64", "float32"
_TORCH_DTYPE = {"float64": torch.float64, "float32": torch.float32, None: None, np.float64: torch.float64,
                np.float32: torch.float32, torch.float64: torch.float64, torch.float32: torch.float32}
_uid = itertools.count(1)
_active_tapes = []                 # innermost last
_floatx = ["float32"]


# ----------------------------------------------------------------------------------------------------------- tensors
def _payload(x, like=None):
    """torch payload of an operand (Tensor, numpy array, python scalar, list of those)."""
    if isinstance(x, Tensor):
        return x.t
    if isinstance(x, (list, tuple)) and any(isinstance(v, Tensor) for v in x):
        return torch.stack([_payload(v) for v in x])
    a = np.asarray(x)
    if a.dtype == np.float32 or a.dtype.kind in "iu" and like is not None and like.dtype.is_floating_point:
        a = a.astype(np.float64)           # exact promotion (see module docstring)
    t = torch.from_numpy(np.ascontiguousarray(a).reshape(-1).copy()).reshape(a.shape)      # keeps the numpy dtype, also for 0-d
    if like is not None and t.dtype != like.dtype and t.dtype.is_floating_point:
        t = t.to(like.dtype)
    return t


def _reach_of(inputs):
    """For every active tape: the watched sources the new value is connected to through recorded operations."""
    reach = {}
    for tape in _active_tapes:
        r = set()
        for inp in inputs:
            if isinstance(inp, Tensor):
                r |= inp._reach.get(tape.uid, frozenset())
                if tape._watches(inp):
                    r.add(inp.uid)
        if r:
            reach[tape.uid] = frozenset(r)
    return reach


def _op(fn, *args):
    """Apply a torch function to operands (tensors, numpy arrays, python scalars) and record tape connectivity."""
    ts = [a for a in args if isinstance(a, Tensor)]
    like = ts[0].t if ts else None
    return Tensor(fn(*[_payload(a, like) for a in args]), _reach_of(ts))


def _op_list(fn, values):
    """Same for operations over a list of operands (concat, stack)."""
    ts = [v for v in values if isinstance(v, Tensor)]
    like = ts[0].t if ts else None
    return Tensor(fn([_payload(v, like) for v in values]), _reach_of(ts))


class Tensor(object):
    """Eager tensor: an immutable value (``t``) plus the tape connectivity it was produced with."""
    __array_ufunc__ = None          # numpy defers to our reflected operators
    __array_priority__ = 1000

    def __init__(self, t, reach=None):
        self.t = t
        self.uid = next(_uid)
        self._reach = reach or {}

    # value access
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def dtype(self):
        return str(self.t.dtype).replace("torch.", "")

    def numpy(self):
        return self.t.detach().numpy().copy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def __float__(self):
        return float(self.t.detach())

    def __bool__(self):
        return bool(self.t.detach())

    def __format__(self, spec):
        if self.t.numel() == 1 and spec:
            return format(float(self), spec)
        return str(self)

    def __repr__(self):
        return "tf.Tensor(%s, shape=%s, dtype=%s)" % (self.t.detach().numpy(), self.shape, self.dtype)

    __str__ = __repr__
    __hash__ = object.__hash__

    # structure
    def __getitem__(self, idx):
        return _op(lambda a: a[idx], self)

    # arithmetic (no in-place variants on purpose: `x += d` must rebind, custom_lbfgs.py:175)
    def __add__(self, o): return _op(lambda a, b: a + b, self, o)
    def __radd__(self, o): return _op(lambda a, b: b + a, self, o)
    def __sub__(self, o): return _op(lambda a, b: a - b, self, o)
    def __rsub__(self, o): return _op(lambda a, b: b - a, self, o)
    def __mul__(self, o): return _op(lambda a, b: a * b, self, o)
    def __rmul__(self, o): return _op(lambda a, b: b * a, self, o)
    def __truediv__(self, o): return _op(lambda a, b: a / b, self, o)
    def __rtruediv__(self, o): return _op(lambda a, b: b / a, self, o)
    def __pow__(self, o): return _op(lambda a, b: a ** b, self, o)
    def __neg__(self): return _op(lambda a: -a, self)
    def __abs__(self): return _op(torch.abs, self)

    def _cmp(self, o, fn):
        return Tensor(fn(self.t.detach(), _payload(o, self.t).detach()))

    def __lt__(self, o): return self._cmp(o, torch.lt)
    def __le__(self, o): return self._cmp(o, torch.le)
    def __gt__(self, o): return self._cmp(o, torch.gt)
    def __ge__(self, o): return self._cmp(o, torch.ge)


class Variable(Tensor):
    def __init__(self, initial_value, dtype=None, trainable=True, name=None):
        t = _payload(initial_value).detach().clone()
        if _TORCH_DTYPE.get(dtype) is not None:
            t = t.to(_TORCH_DTYPE[dtype])
        Tensor.__init__(self, t.requires_grad_(bool(trainable)))
        self.trainable = trainable
        self.name = name

    def assign(self, value):
        with torch.no_grad():
            self.t.copy_(_payload(value, self.t).detach().reshape(self.t.shape))
        return self

    def assign_sub(self, value):
        with torch.no_grad():
            self.t.sub_(_payload(value, self.t).detach().reshape(self.t.shape))
        return self


# ------------------------------------------------------------------------------------------------------------- tapes
class GradientTape(object):
    def __init__(self, persistent=False, watch_accessed_variables=True):
        self.uid = next(_uid)
        self.persistent = persistent
        self.auto = watch_accessed_variables
        self._watched = {}
        self._used = False

    def __enter__(self):
        _active_tapes.append(self)
        return self

    def __exit__(self, *exc):
        assert _active_tapes and _active_tapes[-1] is self, "tapes must nest"
        _active_tapes.pop()
        return False

    def _watches(self, x):
        return x.uid in self._watched or (self.auto and isinstance(x, Variable) and x.trainable)

    def watch(self, x):
        for v in (x if isinstance(x, (list, tuple)) else [x]):
            if not isinstance(v, Tensor):
                raise ValueError("Passed in object of type %s, not tf.Tensor" % type(v).__name__)
            if not v.t.requires_grad:           # a constant becomes a differentiation source from here on
                v.t = v.t.detach().clone().requires_grad_(True)
            self._watched[v.uid] = v            # keeps the source alive, so uids stay unique

    def gradient(self, target, sources, output_gradients=None):
        if self._used and not self.persistent:
            raise RuntimeError("GradientTape.gradient can only be called once on non-persistent tapes.")
        self._used = True
        single = not isinstance(sources, (list, tuple))
        srcs = [sources] if single else list(sources)
        if not isinstance(target, Tensor):
            raise ValueError("target must be a tf.Tensor")
        reach = target._reach.get(self.uid, frozenset())
        live = [s for s in srcs if isinstance(s, Tensor) and s.uid in reach and s.t.requires_grad]
        out = {}
        if live and target.t.requires_grad:
            og = _payload(output_gradients, target.t) if output_gradients is not None else torch.ones_like(target.t)
            recording = bool(_active_tapes)          # the backward computation is recorded iff some tape is active now
            gs = torch.autograd.grad(target.t, [s.t for s in live], grad_outputs=og, create_graph=recording,
                                     retain_graph=True, allow_unused=True)
            extra = [output_gradients] if isinstance(output_gradients, Tensor) else []
            for s, g in zip(live, gs):
                if g is not None:
                    out[s.uid] = Tensor(g if recording else g.detach(), _reach_of([target] + extra + live))
        res = [out.get(s.uid) if isinstance(s, Tensor) else None for s in srcs]
        return res[0] if single else res


# --------------------------------------------------------------------------------------------------------------- ops
def convert_to_tensor(value, dtype=None, name=None):
    if isinstance(value, Tensor):
        return value
    t = _payload(value)
    if _TORCH_DTYPE.get(dtype) is not None:
        t = t.to(_TORCH_DTYPE[dtype])
    ins = [v for v in value if isinstance(v, Tensor)] if isinstance(value, (list, tuple)) else []
    return Tensor(t, _reach_of(ins))


def _none_guard(x, what):
    if x is None:
        raise ValueError("None values not supported (%s)." % what)


def reduce_mean(x, axis=None): return _op(lambda a: a.mean() if axis is None else a.mean(axis), x)
def reduce_sum(x, axis=None): return _op(lambda a: a.sum() if axis is None else a.sum(axis), x)
def square(x): return _op(lambda a: a * a, x)
def exp(x): return _op(torch.exp, x)
def matmul(a, b): return _op(lambda p, q: p @ q, a, b)


def abs(x):  # noqa: A001  (mirrors tf.abs)
    _none_guard(x, "abs")
    return _op(torch.abs, x)


def reshape(x, shape):
    _none_guard(x, "reshape")
    shp = [int(s) for s in shape]
    return _op(lambda a: a.reshape(shp), x)


def concat(values, axis):
    for v in values:
        _none_guard(v, "concat")
    return _op_list(lambda vs: torch.cat(vs, dim=axis), list(values))


def stack(values, axis=0):
    return _op_list(lambda vs: torch.stack(vs, dim=axis), list(values))


def ones(shape, dtype=None):
    return Tensor(torch.ones([int(s) for s in shape], dtype=_TORCH_DTYPE.get(dtype) or torch.float32))


def zeros(shape, dtype=None):
    return Tensor(torch.zeros([int(s) for s in shape], dtype=_TORCH_DTYPE.get(dtype) or torch.float32))


def function(fn=None, **kw):
    return fn if fn is not None else (lambda f: f)


def executing_eagerly():
    return True


def print(*args, **kw):  # noqa: A001  (mirrors tf.print; writes to stderr like TF does)
    import builtins
    builtins.print(*args, file=sys.stderr, **kw)


nn = types.SimpleNamespace(tanh=lambda x: _op(torch.tanh, x))
math = types.SimpleNamespace(tanh=nn.tanh)
test = types.SimpleNamespace(is_gpu_available=lambda *a, **k: False)
_rng = [np.random.default_rng(0)]
random = types.SimpleNamespace(set_seed=lambda s: _rng.__setitem__(0, np.random.default_rng(s)))


# ------------------------------------------------------------------------------------------------------------- keras
class _InputLayer(object):
    def __init__(self, input_shape=None, **kw):
        self.input_shape = tuple(input_shape)
        self.trainable_variables = []


class _Lambda(object):
    def __init__(self, function, **kw):
        self.fn = function
        self.trainable_variables = []

    def __call__(self, x):
        return self.fn(x)

    def get_weights(self):
        return []


def _glorot_normal(fan_in, fan_out):
    std = np.sqrt(2.0 / (fan_in + fan_out)) / 0.87962566103423978
    out = _rng[0].standard_normal((fan_in, fan_out))
    bad = np.abs(out) > 2.0
    while bad.any():                                     # truncated normal: resample outside two standard deviations
        out[bad] = _rng[0].standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return out * std


class _Dense(object):
    def __init__(self, units, activation=None, kernel_initializer="glorot_uniform", **kw):
        self.units, self.activation, self.init = int(units), activation, kernel_initializer
        self.kernel = self.bias = None

    def build(self, fan_in):
        if self.init != "glorot_normal":
            raise NotImplementedError("only glorot_normal is used by the reference (utils/neuralnetwork.py:34,37)")
        self.kernel = Variable(_glorot_normal(fan_in, self.units), dtype=_floatx[0])
        self.bias = Variable(np.zeros(self.units), dtype=_floatx[0])

    @property
    def trainable_variables(self):
        return [self.kernel, self.bias]

    def __call__(self, x):
        z = matmul(x, self.kernel) + self.bias
        return self.activation(z) if self.activation is not None else z

    def get_weights(self):
        return [self.kernel.numpy(), self.bias.numpy()]

    def set_weights(self, weights):
        k, b = (np.asarray(w) for w in weights)
        if k.shape != self.kernel.shape or b.shape != self.bias.shape:
            raise ValueError("Layer weight shape %s not compatible with provided weight shape %s" % (self.kernel.shape, k.shape))
        self.kernel.assign(k)
        self.bias.assign(b)


class _Sequential(object):
    def __init__(self):
        self._layers = []
        self._width = None

    def add(self, layer):
        if isinstance(layer, _InputLayer):
            self._width = layer.input_shape[-1]
        elif isinstance(layer, _Dense):
            layer.build(self._width)
            self._width = layer.units
        self._layers.append(layer)

    @property
    def layers(self):
        return [l for l in self._layers if not isinstance(l, _InputLayer)]     # Keras hides the InputLayer

    @property
    def trainable_variables(self):
        return [v for l in self.layers for v in l.trainable_variables]         # a new list on every access

    def __call__(self, x):
        if not isinstance(x, Tensor):
            x = convert_to_tensor(x, dtype=_floatx[0])
        for l in self.layers:
            x = l(x)
        return x

    def summary(self):
        import builtins
        for i, l in enumerate(self.layers):
            builtins.print("layer %d: %s %s" % (i, type(l).__name__.strip("_"), getattr(l, "units", "")))


class _Adam(object):
    """OptimizerV2 Adam of TF 2.0 (resource_apply_dense): t = iterations + 1, lr_t = lr sqrt(1-b2^t)/(1-b1^t),
    m += (g-m)(1-b1), v += (g^2-v)(1-b2), var -= lr_t m / (sqrt(v) + eps); epsilon=None -> backend epsilon 1e-7."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False, **kw):
        self.lr, self.b1, self.b2 = float(learning_rate), float(beta_1), float(beta_2)
        self.eps = 1e-7 if epsilon is None else float(epsilon)
        if amsgrad:
            raise NotImplementedError
        self.iterations = 0
        self._slots = {}

    def apply_gradients(self, grads_and_vars):
        gv = [(g, v) for g, v in grads_and_vars if g is not None]
        if not gv:
            raise ValueError("No gradients provided for any variable.")
        t = self.iterations + 1
        lr_t = self.lr * np.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        with torch.no_grad():
            for g, var in gv:
                if var.uid not in self._slots:
                    self._slots[var.uid] = (torch.zeros_like(var.t), torch.zeros_like(var.t))
                m, v = self._slots[var.uid]
                gt = g.t.detach()
                m += (gt - m) * (1.0 - self.b1)
                v += (gt * gt - v) * (1.0 - self.b2)
                var.t -= lr_t * m / (torch.sqrt(v) + self.eps)
        self.iterations = t


keras = types.SimpleNamespace(
    Sequential=_Sequential,
    layers=types.SimpleNamespace(InputLayer=_InputLayer, Lambda=_Lambda, Dense=_Dense),
    optimizers=types.SimpleNamespace(Adam=_Adam),
    backend=types.SimpleNamespace(set_floatx=lambda d: _floatx.__setitem__(0, d), floatx=lambda: _floatx[0],
                                  epsilon=lambda: 1e-7),
)
