"""Import alias: the package directory is named after the reference repository
(``music-recommendation-multimodal_b200``), which is not a valid Python identifier.
``import mrm_b200`` resolves to that package."""
import importlib
import sys

_pkg = importlib.import_module("music-recommendation-multimodal_b200")
sys.modules[__name__] = _pkg
