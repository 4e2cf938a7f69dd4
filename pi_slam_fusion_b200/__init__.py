"""Import shim: the package lives in `pi-slam-fusion_b200/` (the reference's name, not a Python identifier).

`import pi_slam_fusion_b200` resolves sub-modules from that directory.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pi-slam-fusion_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
