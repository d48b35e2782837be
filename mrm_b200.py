"""Import alias: the package directory is named after the reference repository
(``music-recommendation-multimodal_b200``), which is not a valid Python identifier.
``import mrm_b200`` resolves to that package, and ``mrm_b200.<sub>`` to the SAME module object as
``music-recommendation-multimodal_b200.<sub>`` whichever spelling is imported first (a finder on
``sys.meta_path`` maps the short dotted names onto the real modules) — two copies of a module would mean two
sets of ctypes structure classes and two library handles."""
import importlib
import importlib.abc
import importlib.machinery
import sys

_REAL = "music-recommendation-multimodal_b200"
_pkg = importlib.import_module(_REAL)


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(__name__ + "."):
            return importlib.machinery.ModuleSpec(fullname, self, origin=_REAL + fullname[len(__name__):])
        return None

    def create_module(self, spec):
        return importlib.import_module(spec.origin)

    def exec_module(self, module):
        return None


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
sys.modules[__name__] = _pkg
