"""Import shim: the package directory is named ``spin-ed_b200`` (not an identifier), so this
module re-exports it under the importable name ``spin_ed_b200``."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "spin-ed_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
