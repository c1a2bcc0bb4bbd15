"""Import shim: the package lives in ``crust-render_b200/`` (not a valid Python
identifier), so ``import crust_render_b200`` resolves its submodules there."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "crust-render_b200"))

from .rt import *  # noqa: E402,F401,F403
from .rt import __all__ as _rt_all  # noqa: E402

from .render import *  # noqa: E402,F401,F403
from .render import __all__ as _render_all  # noqa: E402

__all__ = list(_rt_all) + list(_render_all)
