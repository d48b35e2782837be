"""Synthetic stand-in for the reference's private data (its ``src/data/dataset.py`` is
absent from the repository): seeded batches with the inferred batch
contract, seeded parameters with the reference's key names / shapes / init scales, and
seeded retrieval catalogs.

Everything is generated on the CPU with an explicit ``torch.Generator`` so the same
(config, seed) gives the same tensors here, on the GPU box, in the golden-vector script
and in bench.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, Optional

import torch


@dataclass
class TwoTowerConfig:
    """Hyper-parameters of the hot path (defaults = reference defaults,
    src/models/two_tower.py:9-31 and src/train.py:289-297)."""
    vocab_size: int = 10_001
    num_genders: int = 3
    num_countries: int = 50
    max_seq_len: int = 50
    embedding_dim: int = 256
    num_heads: int = 4
    num_layers: int = 2
    dropout: float = 0.1
    modality_dim: int = 128          # audio / visual / text / tabular widths (all 128)
    fusion_hidden: int = 512
    temperature: float = 0.07

    @property
    def ff_dim(self) -> int:
        return 4 * self.embedding_dim

    def as_dict(self):
        return asdict(self)


def _gen(seed: int) -> torch.Generator:
    return torch.Generator(device="cpu").manual_seed(int(seed))


def _xavier_normal(shape, g):
    fan_out, fan_in = shape[0], shape[1]
    std = math.sqrt(2.0 / (fan_in + fan_out))
    return torch.randn(shape, generator=g) * std


def _xavier_uniform(shape, g):
    fan_out, fan_in = shape[0], shape[1]
    a = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=g) * 2 - 1) * a


def _kaiming_linear(out_f, in_f, g):
    """torch's nn.Linear default init (kaiming_uniform a=sqrt(5)) -> U(-1/sqrt(in), 1/sqrt(in))."""
    b = 1.0 / math.sqrt(in_f)
    w = (torch.rand((out_f, in_f), generator=g) * 2 - 1) * b
    bias = (torch.rand((out_f,), generator=g) * 2 - 1) * b
    return w, bias


def make_state_dict(cfg: TwoTowerConfig, seed: int = 0, perturb: float = 0.05) -> Dict[str, torch.Tensor]:
    """Seeded fp32 parameters under the reference's state-dict keys.

    Scales follow the reference initialisers (Xavier-normal for the user tower's
    Linear/Embedding, torch defaults for the item fusion). ``perturb`` adds small noise to
    the biases / LayerNorm / BatchNorm affine terms (zeros / ones at reference init) so that
    parity tests exercise them.
    """
    g = _gen(seed)
    D, FF = cfg.embedding_dim, cfg.ff_dim
    sd: Dict[str, torch.Tensor] = {}

    def noise(n, base):
        return base + perturb * torch.randn((n,), generator=g)

    ut = "user_tower."
    sd[ut + "item_embedding.weight"] = _xavier_normal((cfg.vocab_size, D), g)
    sd[ut + "gender_embedding.weight"] = _xavier_normal((cfg.num_genders, 16), g)
    sd[ut + "country_embedding.weight"] = _xavier_normal((cfg.num_countries, 32), g)
    sd[ut + "position_embedding.weight"] = _xavier_normal((cfg.max_seq_len, D), g)
    for l in range(cfg.num_layers):
        p = f"{ut}transformer_encoder.layers.{l}."
        sd[p + "self_attn.in_proj_weight"] = _xavier_uniform((3 * D, D), g)
        sd[p + "self_attn.in_proj_bias"] = noise(3 * D, 0.0)
        sd[p + "self_attn.out_proj.weight"] = _xavier_normal((D, D), g)
        sd[p + "self_attn.out_proj.bias"] = noise(D, 0.0)
        sd[p + "linear1.weight"] = _xavier_normal((FF, D), g)
        sd[p + "linear1.bias"] = noise(FF, 0.0)
        sd[p + "linear2.weight"] = _xavier_normal((D, FF), g)
        sd[p + "linear2.bias"] = noise(D, 0.0)
        sd[p + "norm1.weight"] = noise(D, 1.0)
        sd[p + "norm1.bias"] = noise(D, 0.0)
        sd[p + "norm2.weight"] = noise(D, 1.0)
        sd[p + "norm2.bias"] = noise(D, 0.0)
    sd[ut + "layer_norm.weight"] = noise(D, 1.0)
    sd[ut + "layer_norm.bias"] = noise(D, 0.0)
    sd[ut + "fusion_layer.0.weight"] = _xavier_normal((D, D + 48), g)
    sd[ut + "fusion_layer.0.bias"] = noise(D, 0.0)
    sd[ut + "fusion_layer.1.weight"] = noise(D, 1.0)
    sd[ut + "fusion_layer.1.bias"] = noise(D, 0.0)
    sd[ut + "fusion_layer.3.weight"] = _xavier_normal((D, D), g)
    sd[ut + "fusion_layer.3.bias"] = noise(D, 0.0)

    it = "item_tower.fusion_layer."
    H, Fin = cfg.fusion_hidden, 4 * cfg.modality_dim
    w, b = _kaiming_linear(H, Fin, g)
    sd[it + "0.weight"], sd[it + "0.bias"] = w, b
    sd[it + "1.weight"] = noise(H, 1.0)
    sd[it + "1.bias"] = noise(H, 0.0)
    sd[it + "1.running_mean"] = torch.zeros(H)
    sd[it + "1.running_var"] = torch.ones(H)
    sd[it + "1.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    w, b = _kaiming_linear(D, H, g)
    sd[it + "4.weight"], sd[it + "4.bias"] = w, b
    sd[it + "5.weight"] = noise(D, 1.0)
    sd[it + "5.bias"] = noise(D, 0.0)
    return sd


def make_batch(cfg: TwoTowerConfig, batch_size: int, seq_len: Optional[int] = None, seed: int = 1,
               full_length: bool = False, num_users: int = 850, zipf: bool = True) -> Dict[str, torch.Tensor]:
    """One training/validation batch with the contract of SURVEY.md §8a-0.

    history_ids are right-padded with 0, every history has length >= 1; the four modality
    inputs carry precomputed (B, 128) embeddings (north_star: the encoders are out of scope).
    The key names are the reference's (src/models/two_tower.py:82-95).
    """
    g = _gen(seed)
    L = cfg.max_seq_len if seq_len is None else seq_len
    B, V = batch_size, cfg.vocab_size
    if full_length:
        lens = torch.full((B,), L, dtype=torch.long)
    else:
        lens = torch.randint(1, L + 1, (B,), generator=g)
    if zipf:
        # Zipf(alpha ~ 1) over [1, V): inverse-CDF of 1/x on [1, V)
        u = torch.rand((B, L), generator=g, dtype=torch.float64)
        ids = torch.exp(u * math.log(V - 1)).floor().long().clamp_(1, V - 1)
    else:
        ids = torch.randint(1, V, (B, L), generator=g)
    pos = torch.arange(L).unsqueeze(0)
    valid = pos < lens.unsqueeze(1)
    ids = ids * valid
    m = cfg.modality_dim
    batch = {
        "history_ids": ids,
        "history_mask": valid.long(),
        "user_gender": torch.randint(0, cfg.num_genders, (B,), generator=g),
        "user_country": torch.randint(0, cfg.num_countries, (B,), generator=g),
        "user_idx": torch.randint(0, num_users, (B,), generator=g),
        "target_id": torch.randint(1, V, (B,), generator=g),
        "target_audio": torch.randn((B, m), generator=g),
        "target_image": torch.randn((B, m), generator=g),
        "target_input_ids": torch.randn((B, m), generator=g),      # precomputed text embedding
        "target_attention_mask": torch.ones((B, 1), dtype=torch.long),  # ignored in scope
        "target_tabular": torch.randn((B, m), generator=g),
    }
    return batch


def make_c2_parity_batch(cfg: TwoTowerConfig, batch_size: int = 256, seed: int = 100) -> Dict[str, torch.Tensor]:
    """The batch of the c2-shape parity pins (tests/test_bench_shapes.py, tests/golden/anchor_c2.pt): bench-style
    batch (all histories full, ~unique users) with eight same-user collisions planted so the -1e4 mask is live."""
    batch = make_batch(cfg, batch_size, seed=seed, full_length=True, num_users=1_000_000)
    batch["user_idx"][:8] = batch["user_idx"][8:16]
    return batch


def _quantize(t: torch.Tensor, step: float) -> torch.Tensor:
    return torch.round(t.clamp(-1, 1) / step) * step


def make_catalog(num_items: int, dim: int = 256, seed: int = 2, grid: float = 0.0) -> torch.Tensor:
    """Item-embedding table in the reference's cache layout (evaluate_metrics.py:92-104):
    (num_items + 1, dim) fp32, row i = unit-norm embedding of item id i, row 0 = zeros.

    grid > 0 (a power of two >= 2^-7) quantises every entry to a multiple of ``grid`` in
    [-1, 1] instead of normalising: all 256-term dot products are then exact in fp32 and in
    bf16-input tensor-core MMA whatever the summation order, and score ties are frequent —
    the bit-exact top-K fixture.
    """
    g = _gen(seed)
    t = torch.randn((num_items + 1, dim), generator=g)
    if grid > 0:
        t = _quantize(t * 0.25, grid)
    else:
        t = torch.nn.functional.normalize(t, dim=1)
    t[0] = 0
    return t


def make_queries(table: torch.Tensor, num_users: int, seed: int = 3, noise: float = 3.3,
                 grid: float = 0.0):
    """User embeddings with a planted target so Recall/NDCG are non-degenerate:
    user ~ E[target] + noise * (a random vector of the same scale as a table row).
    noise ~ 3.3 gives Recall@10 ~ 0.6 at 1M items, ~5 at 5k items.
    Returns (users (U, dim) fp32, target ids (U,) int64)."""
    g = _gen(seed)
    N1, D = table.shape
    targets = torch.randint(1, N1, (num_users,), generator=g)
    n = torch.randn((num_users, D), generator=g)
    if grid > 0:
        u = _quantize((table[targets] + noise * 0.25 * n) / (1.0 + noise), grid)
    else:
        u = torch.nn.functional.normalize(table[targets] + noise * n / math.sqrt(D), dim=1)
    return u, targets


def make_item_features(cfg: TwoTowerConfig, n_items: int, vocab_size: int, seed: int = 71, nan_rows=(),
                       state_dict: Optional[Dict[str, torch.Tensor]] = None):
    """Catalog-indexing input (src/evaluate_metrics.py:24-104 on precomputed modality embeddings): four (n_items,
    128) feature tensors and the item id of every row — a random subset of [1, vocab_size) in random order, so ids
    that are not listed keep a zero row. ``nan_rows`` get a NaN feature (exercises the NaN -> 0 branch).
    With ``state_dict`` the BatchNorm running statistics in it are replaced by non-trivial seeded values.
    Returns (features dict, item ids int64 (n_items,))."""
    g = _gen(seed)
    feats = {k: torch.randn(n_items, cfg.modality_dim, generator=g)
             for k in ("target_audio", "target_image", "target_input_ids", "target_tabular")}
    for r in nan_rows:
        feats["target_audio"][r, 3] = float("nan")
    ids = torch.randperm(vocab_size - 1, generator=g)[:n_items] + 1
    if state_dict is not None:
        state_dict["item_tower.fusion_layer.1.running_mean"] = 0.1 * torch.randn(cfg.fusion_hidden, generator=g)
        state_dict["item_tower.fusion_layer.1.running_var"] = 0.5 + torch.rand(cfg.fusion_hidden, generator=g)
    return feats, ids
