"""Shadow of the reference's `optimizers` module: put this directory first on sys.path (or PYTHONPATH)
and the reference's train.py picks up the native multi-tensor optimizers."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from convasr_b200.optimizers import *  # noqa: F401,F403,E402
from convasr_b200 import optimizers as _impl  # noqa: E402

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
