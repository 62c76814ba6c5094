"""Import alias: the product package lives in the directory
``industrial-quality-control-computer-vision-system-_b200/`` (not a valid Python identifier);
``import qcgraft`` resolves its sub-modules from there."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "industrial-quality-control-computer-vision-system-_b200")
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
