"""Minimal stand-in for the handful of top-level TensorFlow symbols the reference *scripts* touch outside the class
bodies: tf.random.set_seed, tf.concat, tf.convert_to_tensor, tf.Variable, tf.exp, version/banner helpers.
It performs NO training arithmetic: the tape bodies inside the reference subclasses are replaced by the fused CUDA
kernels (neuralnetwork.NeuralNetwork.__init_subclass__).  Only on sys.path when a reference script is run through
run_reference_script.py."""
import numpy as np

__version__ = "2.0.0-shim (pinn_b200; no TensorFlow arithmetic)"


class _Arr(np.ndarray):
    def numpy(self):
        return np.asarray(self)


def _t(a, dtype=None):
    return np.ascontiguousarray(np.asarray(a, dtype=dtype or np.float64)).view(_Arr)


def convert_to_tensor(a, dtype=None):
    return _t(a, np.dtype(dtype) if isinstance(dtype, str) else dtype)


def concat(values, axis=0):
    return _t(np.concatenate([np.asarray(v) for v in values], axis=axis))


def stack(values, axis=0):
    return _t(np.stack([np.asarray(v) for v in values], axis=axis))


def ones(shape, dtype=None):
    return _t(np.ones([int(v) for v in shape]))


def exp(a):
    return _t(np.exp(np.asarray(a)))


def executing_eagerly():
    return True


class Variable(object):
    def __init__(self, value, dtype=None):
        self._v = np.array(value, dtype=np.float64)

    def numpy(self):
        return self._v

    def assign(self, value):
        self._v = np.array(value, dtype=np.float64)
        return self


class _Random(object):
    seed = 1234

    def set_seed(self, s):
        _Random.seed = int(s)
        try:
            import neuralnetwork
            neuralnetwork.NeuralNetwork.weight_seed = int(s)
        except Exception:
            pass


random = _Random()


class _Test(object):
    @staticmethod
    def is_gpu_available():
        return True


test = _Test()


class GradientTape(object):
    def __init__(self, *a, **k):
        raise RuntimeError("tf.GradientTape is not available: PDE residuals are evaluated by the fused CUDA kernels")
