# Merge with a DSAC-v2 checkout further down sys.path (its networks/cnn.py stays importable).
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
