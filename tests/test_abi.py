"""The C-ABI shared library loads and exports every symbol include/tt_b200.h declares
(no compute calls: this runs without a GPU)."""
import os
import re

import pytest


def test_library_exports_every_declared_symbol():
    import mrm_b200
    from mrm_b200 import _lib
    l = mrm_b200.lib()
    header = open(os.path.join(os.path.dirname(__file__), "..", "include", "tt_b200.h")).read()
    declared = set(re.findall(r"\b(tt_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(l, name), f"{name} declared in tt_b200.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    assert l.tt_version() == 1
    assert l.tt_last_error() is not None


def test_no_cpu_fallback_without_library(monkeypatch, tmp_path):
    from mrm_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_LIB_PATH", tmp_path / "missing.so")
    with pytest.raises(_lib.TTError):
        _lib.lib()


def test_product_package_never_imports_the_oracle():
    root = os.path.join(os.path.dirname(__file__), "..", "music-recommendation-multimodal_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"
