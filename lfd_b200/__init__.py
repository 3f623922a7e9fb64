"""lfd_b200 - B200-native drop-in for ``lfd.detecttrails``' per-frame detection path.

Same public names as /root/reference/lfd/detecttrails/__init__.py:36-71 (``setup``, and the star
re-exports of removestars / processfield / detecttrails).  The pixel work runs in liblfd_b200.so
(hand-written sm_100a kernels behind the C ABI of include/lfd_b200.h); importing this package does not
need a GPU, creating a handle does, and there is no CPU fallback."""
import os as _os

DEBUG_PATH = None


def setup(bosspath, photoobjpath, photoreduxpath, debugpath):
    """Set BOSS, BOSS_PHOTOOBJ, PHOTO_REDUX and DEBUG_PATH (lfd/detecttrails/__init__.py:36-66)."""
    global DEBUG_PATH
    _os.environ["BOSS"] = bosspath
    _os.environ["BOSS_PHOTOOBJ"] = photoobjpath
    _os.environ["PHOTO_REDUX"] = photoreduxpath
    _os.environ["DEBUG_PATH"] = debugpath
    DEBUG_PATH = debugpath


from .removestars import *      # noqa: E402,F401,F403
from .processfield import *     # noqa: E402,F401,F403
from .detecttrails import *     # noqa: E402,F401,F403
