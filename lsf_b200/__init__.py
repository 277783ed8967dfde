"""Import alias: the product package lives in ``levelsetfusion-python_b200/`` (a directory name Python cannot
import); this stub makes it importable as ``lsf_b200`` by pointing the package path at that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "levelsetfusion-python_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
