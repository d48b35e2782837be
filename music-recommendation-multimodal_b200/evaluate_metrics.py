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


def _shard_slice(n_all: int, shard):
    if shard is None:
        return 0, n_all
    r, w = shard
    per = (n_all + w - 1) // w
    return min(r * per, n_all), min(n_all, (r + 1) * per)


def index_catalog_device(model, item_features: Dict[str, torch.Tensor], item_ids: torch.Tensor, vocab_size: int,
                         batch_size: int = 131072, shard=None, out=None):
    """Catalog indexing (src/evaluate_metrics.py:24-104) entirely on the device: returns (table fp32 (V, 256),
    table bf16 (V, 256)) on the model's GPU — the dense cache in the reference's layout (row i = embedding of item
    id i, row 0 and unlisted ids zero) and the copy the scoring kernel reads. Host features are moved batch by
    batch (pinned or not); device features are used in place. ``shard=(rank, world)``: only the rank-th contiguous
    slice of the ITEM LIST is indexed (eval-mode BatchNorm has no cross-row dependency); the other
    rows stay zero, so the full table is the sum over ranks. ``out=(table, table_bf16)`` reuses caller buffers."""
    eng = model.engine
    model.eval()
    if hasattr(model, "_sync_shadow"):
        model._sync_shadow()
    dev, D = eng.device, eng.cfg.embedding_dim
    if out is None:
        table = torch.zeros(vocab_size, D, device=dev)
        table_bf16 = torch.zeros(vocab_size, D, device=dev, dtype=torch.bfloat16)
    else:
        table, table_bf16 = out
    lo, hi = _shard_slice(item_ids.shape[0], shard)
    keys = ("target_audio", "target_image", "target_input_ids", "target_tabular")
    on_dev = all(item_features[k].is_cuda for k in keys) and item_ids.is_cuda
    # host inputs: stream them in chunks so the device never holds more than one chunk of features
    chunk = (hi - lo) if on_dev else max(batch_size, 1 << 18)
    for s0 in range(lo, hi, max(chunk, 1)):
        s1 = min(hi, s0 + chunk)
        feats = {k: item_features[k][s0:s1].to(dev, dtype=torch.float32, non_blocking=True).contiguous() for k in keys}
        ids = item_ids[s0:s1].to(dev, dtype=torch.long, non_blocking=True).contiguous()
        eng.index_items(feats, ids, table, table_bf16, batch_size=batch_size)
    return table, table_bf16


def build_catalog_index(model, item_features, item_ids, vocab_size: int, batch_size: int = 131072) -> CatalogIndex:
    """Index the catalog and hand the device tables straight to retrieval (no host round trip, no re-cast)."""
    table, table_bf16 = index_catalog_device(model, item_features, item_ids, vocab_size, batch_size)
    return CatalogIndex.from_device_tables(table, table_bf16)


def compute_all_item_embeddings(model, item_features: Dict[str, torch.Tensor], item_ids: torch.Tensor,
                                batch_size: int, device, vocab_size: int, shard=None):
    """The reference's return contract (src/evaluate_metrics.py:24-104): (dense CPU FloatTensor (V, 256),
    vocab_size) — what `main` saves as {'item_embeddings', 'vocab_size'} (:323-326). The work is
    `index_catalog_device`; the table crosses to the host once."""
    # the reference's batch_size is its DataLoader's; eval-mode rows are independent, so the device path batches
    # by what fills the GPU
    table, _ = index_catalog_device(model, item_features, item_ids, vocab_size, max(batch_size, 131072), shard)
    return table.cpu(), vocab_size


def calculate_metrics_global(model, val_loader, item_embeddings, device, k_list: Sequence[int] = (10, 20),
                             index: CatalogIndex = None, group=None) -> Dict[str, float]:
    """Reference signature (src/evaluate_metrics.py:106). User embeddings come from the CUDA user
    tower batch by batch; scoring / top-K / metrics run once over all validation rows with the
    fused retrieval kernels (the per-row results do not depend on the batching).

    With a catalog index built over one shard per rank (``index.is_sharded``) every rank iterates the SAME loader;
    the user tower then runs on 1 / world of the batches per rank (batch i on rank i % world) and the embeddings are
    all-gathered (1 KB per user) before the sharded scoring — the tower is not replicated work."""
    logger.info(f"Iniciando evaluación global con K={list(k_list)}...")
    model.eval()
    if isinstance(item_embeddings, dict):
        item_embeddings = item_embeddings["item_embeddings"]
    if index is None:
        index = CatalogIndex(item_embeddings, device=device)
    import torch.distributed as dist
    world = dist.get_world_size(group) if index.is_sharded else 1
    rank = dist.get_rank(group) if index.is_sharded else 0
    users, targets, sizes = [], [], []
    with torch.no_grad():
        for i, batch in enumerate(val_loader):
            n = batch["history_ids"].shape[0]
            sizes.append(n)
            targets.append(batch["target_id"].to(device))
            if i % world != rank:
                users.append(None)
                continue
            users.append(model.get_user_embedding(history_ids=batch["history_ids"].to(device),
                                                  history_mask=batch["history_mask"].to(device),
                                                  user_gender=batch["user_gender"].to(device),
                                                  user_country=batch["user_country"].to(device)))
    D = model.engine.cfg.embedding_dim
    if world > 1:
        # one sum-all-reduce of the (U, 256) matrix in which every rank filled its own batches
        full = torch.zeros(sum(sizes), D, device=device)
        off = 0
        for n, u in zip(sizes, users):
            if u is not None:
                full[off:off + n] = u
            off += n
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
        all_users = full
    else:
        all_users = torch.cat(users)
    return metrics_from_embeddings(all_users, torch.cat(targets), index, list(k_list), group=group)
