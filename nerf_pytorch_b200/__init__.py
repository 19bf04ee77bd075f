"""Import alias: the package lives in the directory `nerf-pytorch_b200/` (the name the build spec
fixes); a dash cannot appear in a Python identifier, so this shim puts that directory on the
package path and re-exports its public surface."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "nerf-pytorch_b200")
__path__.insert(0, _real)

from .api import *          # noqa: E402,F401,F403
from .api import __all__    # noqa: E402,F401
