"""Import alias: the package directory is ``genomic-resistance-mapping-grm-_b200/`` (not a
valid Python identifier), so ``import grm_b200`` maps onto it."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "genomic-resistance-mapping-grm-_b200")
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
del _f
