"""Importable alias of the package directory `instanceseg-without-voxelwise-labeling_b200/`
(its name is not a valid Python identifier).  `import b200seg` executes that package's
__init__ with this module's __path__ pointing at it, so `b200seg.boxes_3d` etc. resolve there."""
import os as _os

_real = _os.path.normpath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..",
                                        "instanceseg-without-voxelwise-labeling_b200"))
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
