"""Import shim: the package directory is ``montreal-forced-aligner_b200/`` (not a valid Python
identifier), so ``import mfa_b200`` maps onto it."""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "montreal-forced-aligner_b200")
__path__ = [_pkg_dir]
with open(_os.path.join(_pkg_dir, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg_dir, "__init__.py"), "exec"))
