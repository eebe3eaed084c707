"""sprintz_b200 -- B200-native (sm_100a) Sprintz encode/decode path.

Drop-in for ONE path of dblalock/sprintz: cpp/Compress/sprintz.h:16-32.  See DESIGN.md / INTEGRATION.md.
The compute is hand-written CUDA behind a C ABI (include/sprintz_b200.h, libsprintz_b200.so); this package
is only the host-side mirror of the reference interface.  There is no CPU fallback.
"""
from ._lib import DELTA, XFF, SprintzError, build, lib  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import FrameCodec  # noqa: F401
