"""Import shim: the package directory is `pomcp.jl_b200/` (not an importable
name because of the dot); `import pomcp_jl_b200` loads it from there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pomcp.jl_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
