"""TensorFlow-2.0 API emulation used to EXECUTE the reference's own Python sources in this container (test infrastructure,
see tensorflow/__init__.py).  `path()` is the directory to put first on sys.path so that `import tensorflow` resolves here."""
import os


def path():
    return os.path.dirname(os.path.abspath(__file__))
