"""Global-retrieval evaluation with the reference's entry points (src/evaluate_metrics.py:24-192).

  compute_all_item_embeddings(model, item_features, item_ids, batch_size, device) -> (dense table, vocab_size)
  calculate_metrics_global(model, val_loader, item_embeddings, device, k_list=[10, 20]) -> dict

``item_embeddings`` is the reference's cache tensor: fp32 (V, 256), row i = normalised embedding of item
id i, row 0 = padding zeros (also accepted: the saved dict {'item_embeddings', 'vocab_size'}).
"""
from __future__ import annotations

import logging
from typing import Dict, Sequence

import torch

from .retrieval import CatalogIndex, metrics_from_embeddings

logger = logging.getLogger(__name__)


def compute_all_item_embeddings(model, item_features: Dict[str, torch.Tensor], item_ids: torch.Tensor,
                                batch_size: int, device, vocab_size: int, shard=None):
    """Catalog indexing (src/evaluate_metrics.py:24-104) on precomputed modality embeddings:
    item tower in eval mode, NaN -> 0, re-normalise with eps 1e-8, scatter into the dense table.

    ``shard=(rank, world)`` indexes only the rank-th contiguous slice of the item list (eval-mode BatchNorm has
    no cross-row dependency, so the catalog splits freely over GPUs); rows of other slices stay
    zero, so the full table is the SUM over ranks (`dist.all_reduce`) or each rank keeps its slice."""
    model.eval()
    D = model.engine.cfg.embedding_dim
    dense = torch.zeros(vocab_size, D)
    n_all = item_ids.shape[0]
    lo, n = 0, n_all
    if shard is not None:
        r, w = shard
        per = (n_all + w - 1) // w
        lo, n = min(r * per, n_all), min(n_all, (r + 1) * per)
    with torch.no_grad():
        for s in range(lo, n, batch_size):
            sl = slice(s, min(n, s + batch_size))
            emb = model.get_item_embedding(images=item_features["target_image"][sl], audio=item_features["target_audio"][sl],
                                           input_ids=item_features["target_input_ids"][sl], attention_mask=None,
                                           tabular=item_features["target_tabular"][sl])
            if torch.isnan(emb).any():
                logger.warning(f"NaNs detected in model output for batch {s // batch_size}")
                emb = torch.nan_to_num(emb, nan=0.0)
            emb = torch.nn.functional.normalize(emb, p=2, dim=1, eps=1e-8)
            dense[item_ids[sl].long()] = emb.cpu()
    return dense, vocab_size


def calculate_metrics_global(model, val_loader, item_embeddings, device, k_list: Sequence[int] = (10, 20),
                             index: CatalogIndex = None) -> Dict[str, float]:
    """Reference signature (src/evaluate_metrics.py:106). User embeddings come from the CUDA user
    tower batch by batch; scoring / top-K / metrics run once over all validation rows with the
    fused retrieval kernels (the per-row results do not depend on the batching)."""
    logger.info(f"Iniciando evaluación global con K={list(k_list)}...")
    model.eval()
    if isinstance(item_embeddings, dict):
        item_embeddings = item_embeddings["item_embeddings"]
    if index is None:
        index = CatalogIndex(item_embeddings, device=device)
    users, targets = [], []
    with torch.no_grad():
        for batch in val_loader:
            users.append(model.get_user_embedding(history_ids=batch["history_ids"].to(device),
                                                  history_mask=batch["history_mask"].to(device),
                                                  user_gender=batch["user_gender"].to(device),
                                                  user_country=batch["user_country"].to(device)))
            targets.append(batch["target_id"].to(device))
    return metrics_from_embeddings(torch.cat(users), torch.cat(targets), index, list(k_list))
