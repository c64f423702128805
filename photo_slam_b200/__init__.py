"""Import shim: the package directory is ``photo-slam_b200/`` (the name the project layout prescribes),
which is not a valid Python identifier. ``import photo_slam_b200`` resolves to it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "photo-slam_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
