"""Importable alias of the hyphen-named package directory ``livecell-instance-segmentation_b200``.

``import livecell_instance_segmentation_b200 as lcr`` and ``import livecell_instance_segmentation_b200.ops``
both resolve to the one real package / sub-module object (no duplicate module instances).
"""
import importlib as _importlib
import importlib.abc as _abc
import importlib.util as _util
import os as _os
import sys as _sys

_ALIAS = __name__
_REAL = "livecell-instance-segmentation_b200"

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)


class _AliasFinder(_abc.MetaPathFinder, _abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(_ALIAS + "."):
            return _util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        return _importlib.import_module(_REAL + spec.name[len(_ALIAS):])

    def exec_module(self, module):
        return None


if not any(isinstance(f, _AliasFinder) for f in _sys.meta_path):
    _sys.meta_path.insert(0, _AliasFinder())
_sys.modules[__name__] = _importlib.import_module(_REAL)
