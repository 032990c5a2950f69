"""Import shim: the product package lives in the directory `yolo-u_b200/` (the repo's layout contract), which is
not a valid Python identifier.  `import yolo_u_b200` resolves to that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "yolo-u_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
