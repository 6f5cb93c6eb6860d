"""Import shim: the package directory is named `gnss-sdr-rs_b200/` (not an importable identifier),
so this module makes it importable as `gnss_sdr_rs_b200` by pointing __path__ at it."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "gnss-sdr-rs_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
