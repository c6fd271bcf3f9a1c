"""chess_b200: B200-native implementation of the `chess sim` comparison path.

Flat re-export like chess/__init__.py of the reference, restricted to the
modules on that path (oe, helpers, sim).
"""
from .version import __version__
from .oe import *  # noqa: F401,F403
from .helpers import *  # noqa: F401,F403
from .sim import *  # noqa: F401,F403
from .engine import CompareEngine, DeviceContacts  # noqa: F401
