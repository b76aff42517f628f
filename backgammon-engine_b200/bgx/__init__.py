"""bgx — host-side mirror of the libbgx C-ABI (include/bgx.h).

`bgx.host`    single-position functions behind the compat Game object (CPU code in libbgx)
`bgx.engine`  BatchEngine: the batched GPU entry points (sm_100a kernels; no CPU fallback)
`bgx.synth`   seeded synthetic positions for the enumeration sweep
`bgx.model`   TDLGammonModel with the reference's contract (model.py:31-222) + GPU fast paths
"""
import os as _os
import sys as _sys

_LIBDIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "lib")
if _LIBDIR not in _sys.path:
    _sys.path.append(_LIBDIR)      # where the compat module backgammon_env*.so is built (appended: explicit paths win)

from . import lib  # noqa: E402,F401
from .lib import BgxError, FIRST_PARITY, FIRST_ROLLOFF  # noqa: E402,F401
