"""Importable alias of the package that lives in ``vip-cup-2022_b200/`` (a directory name Python cannot import
directly).  ``import vipcup_b200`` resolves every submodule from that directory."""
import os as _os

_real = _os.path.normpath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "vip-cup-2022_b200"))
__path__.insert(0, _real)  # type: ignore[name-defined]

with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
