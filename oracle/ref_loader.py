"""Loader of the UNMODIFIED reference (test / benchmark infrastructure only, like everything under oracle/).

Two places can hold the reference's own Python modules:
  * ``/root/reference/src``            — the read-only checkout (dev container only), used by
                                          tests/golden/make_golden.py to generate the golden vectors;
  * ``baseline/_ref``                  — its offline pip install (``__graft_entry__.build()`` runs
        python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>
    ), git-ignored but shipped to the GPU box, used by ``bench.py --impl reference``. setuptools treats the
    checkout's ``src/`` as a src-layout, so the install holds ``models/``, ``train.py``, ``evaluate_metrics.py``
    at its top level; the reference imports itself as ``src.*``, so ``src`` is registered as a package whose
    search path is that directory. No reference file is edited or copied into the tracked tree.

Three harness-side stubs make the hot path importable: ``peft`` (not installed; only used by
the out-of-scope text encoder), ``src.data.dataset`` (absent from the reference repository itself), and the four
modality encoders replaced by identity modules, so that ``batch['target_audio'|'target_image'|
'target_input_ids'|'target_tabular']`` carry the precomputed (B, 128) embeddings (north_star).
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from typing import Optional

import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED = os.path.join(ROOT, "baseline", "_ref")
CHECKOUT = "/root/reference"


class _Identity(nn.Module):
    def __init__(self, *a, **kw):
        super().__init__()

    def forward(self, x, *rest):
        return x


def reference_location() -> Optional[str]:
    """'checkout', 'installed' or None."""
    if os.path.isdir(os.path.join(CHECKOUT, "src", "models")):
        return "checkout"
    if os.path.isdir(os.path.join(INSTALLED, "models")):
        return "installed"
    return None


def import_reference(prefer: str = "checkout", dataset_cls=None):
    """Returns the reference's own modules (two_tower, evaluate_metrics, train) — train is None if it cannot be
    imported. ``prefer='installed'`` takes baseline/_ref even when the checkout exists (what the GPU box sees)."""
    where = reference_location()
    if where is None:
        raise ImportError("the reference is neither at /root/reference nor installed under baseline/_ref "
                          "(run `python __graft_entry__.py` in the dev container)")
    if prefer == "installed" and os.path.isdir(os.path.join(INSTALLED, "models")):
        where = "installed"
    for name in [m for m in sys.modules if m == "src" or m.startswith("src.")]:
        del sys.modules[name]
    peft = types.ModuleType("peft")
    peft.get_peft_model = lambda m, c: m
    peft.LoraConfig = lambda **kw: None
    peft.TaskType = types.SimpleNamespace(FEATURE_EXTRACTION=0)
    sys.modules["peft"] = peft
    if where == "checkout":
        if CHECKOUT not in sys.path:
            sys.path.insert(0, CHECKOUT)
        importlib.import_module("src")
    else:
        pkg = types.ModuleType("src")
        pkg.__path__ = [INSTALLED]
        pkg.__package__ = "src"
        sys.modules["src"] = pkg
    data = types.ModuleType("src.data")
    data.__path__ = []
    ds = types.ModuleType("src.data.dataset")
    ds.MultimodalDataset = dataset_cls if dataset_cls is not None else type("MultimodalDataset", (), {})
    sys.modules["src.data"] = data
    sys.modules["src.data.dataset"] = ds
    item_tower = importlib.import_module("src.models.item_tower")
    for name in ("AudioEncoder", "VisualEncoder", "TextEncoder", "TabularEncoder"):
        setattr(item_tower, name, _Identity)
    two_tower = importlib.import_module("src.models.two_tower")
    evalm = importlib.import_module("src.evaluate_metrics")
    try:
        train = importlib.import_module("src.train")
    except Exception:  # pragma: no cover - tqdm / joblib missing
        train = None
    return two_tower, evalm, train, where


def build_reference_model(two_tower, cfg, sd, dtype, dropout: float = 0.0):
    """The reference's TwoTowerModel with the synthetic state dict loaded (strict)."""
    m = two_tower.TwoTowerModel(
        vocab_size=cfg.vocab_size, tabular_input_dim=cfg.modality_dim, num_genders=cfg.num_genders,
        num_countries=cfg.num_countries, max_seq_len=cfg.max_seq_len,
        user_embedding_dim=cfg.embedding_dim, user_num_heads=cfg.num_heads,
        user_num_layers=cfg.num_layers, user_dropout=dropout, item_embedding_dim=cfg.embedding_dim,
        audio_dim=cfg.modality_dim, visual_dim=cfg.modality_dim, text_dim=cfg.modality_dim,
        tabular_dim=cfg.modality_dim, temperature=cfg.temperature)
    m.item_tower.fusion_layer[3].p = dropout   # hard-coded Dropout(0.1), item_tower.py:126
    m.load_state_dict(sd, strict=True)
    return m.to(dtype)
