"""bgx — host-side mirror of the libbgx C-ABI (include/bgx.h).

`bgx.host`    single-position functions behind the compat Game object (CPU code in libbgx)
`bgx.engine`  BatchEngine: the batched GPU entry points (sm_100a kernels; no CPU fallback)
`bgx.synth`   seeded synthetic positions for the enumeration sweep
`bgx.model`   TDLGammonModel with the reference's contract (model.py:31-222) + GPU fast paths
"""
from . import lib  # noqa: F401
from .lib import BgxError, FIRST_PARITY, FIRST_ROLLOFF  # noqa: F401
