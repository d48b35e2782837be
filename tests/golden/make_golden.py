"""Generate golden vectors from the REFERENCE's own modules (dev container only).

Imports /root/reference unmodified, with the three harness-side stubs of SURVEY.md §8c
(`peft`, `src.data.dataset`, identity modality encoders), loads the seeded synthetic
parameters of `synthetic.make_state_dict`, runs the reference's TwoTowerModel.forward /
backward, get_user_embedding and calculate_metrics_global on seeded synthetic batches and
stores the OUTPUTS (not the inputs: those are regenerated from the seed) as small .pt files.

    python tests/golden/make_golden.py          # writes tests/golden/*.pt (small fixtures, ~20 s)
    python tests/golden/make_golden.py --all    # + anchor_c2.pt: the reference at the c2 shape in fp64 and under
                                                #   bf16 autocast (minutes of CPU); --anchor-only for just that

/root/reference does not exist on the GPU box; tests only read the committed fixtures.
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

import mrm_b200  # noqa: E402
from mrm_b200 import synthetic  # noqa: E402


from oracle import ref_loader  # noqa: E402


def import_reference(dataset_cls=None):
    two_tower, evalm, train, _ = ref_loader.import_reference("checkout", dataset_cls=dataset_cls)
    return two_tower, evalm, train


def build_reference_model(two_tower, cfg, sd, dtype):
    return ref_loader.build_reference_model(two_tower, cfg, sd, dtype, dropout=0.0)


def cast_batch(batch, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in batch.items()}


def golden_train(two_tower, name, cfg, B, seed_w, seed_b, full_length=False):
    sd = synthetic.make_state_dict(cfg, seed=seed_w)
    batch = synthetic.make_batch(cfg, B, seed=seed_b, full_length=full_length)
    out = {"config": cfg.as_dict(), "batch_size": B, "seed_w": seed_w, "seed_b": seed_b,
           "full_length": full_length}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        m = build_reference_model(two_tower, cfg, sd, dtype)
        m.train()
        loss, logits, u, i = m(cast_batch(batch, dtype))
        loss.backward()
        grads = {k: p.grad.detach() for k, p in m.named_parameters() if p.grad is not None}
        out[tag] = {
            "loss": loss.detach(), "logits": logits.detach(), "user_emb": u.detach(),
            "item_emb": i.detach(),
            "bn_running_mean": m.item_tower.fusion_layer[1].running_mean.detach().clone(),
            "bn_running_var": m.item_tower.fusion_layer[1].running_var.detach().clone(),
        }
        # Parameter gradients: full tensors for the small ones, norm + a few rows for the table.
        gsel = {}
        for k, g in grads.items():
            if k == "user_tower.item_embedding.weight":
                ids = torch.unique(batch["history_ids"])[:8]
                gsel[k + "#rows_ids"] = ids
                gsel[k + "#rows"] = g[ids].clone()
                gsel[k + "#norm"] = g.norm()
                gsel[k + "#row0_absmax"] = g[0].abs().max()
            elif g.numel() <= 1024:
                gsel[k] = g.clone()
            else:
                gsel[k + "#norm"] = g.norm()
                gsel[k + "#head"] = g.flatten()[:64].clone()
        out[tag]["grads"] = gsel
        # eval-mode user embedding (fast-path encoder) and item embedding
        m.eval()
        with torch.no_grad():
            b = cast_batch(batch, dtype)
            out[tag]["eval_user_emb"] = m.get_user_embedding(
                b["history_ids"], b["history_mask"], b["user_gender"], b["user_country"])
            out[tag]["eval_user_emb_nomask"] = m.get_user_embedding(b["history_ids"])
            out[tag]["eval_item_emb"] = m.get_item_embedding(
                b["target_image"], b["target_audio"], b["target_input_ids"],
                b["target_attention_mask"], b["target_tabular"])
    # Tolerance anchor: the reference's own modules under bf16 autocast,
    # measured against its fp64 run. Per-tensor relative gradient error and output errors.
    m64 = build_reference_model(two_tower, cfg, sd, torch.float64)
    m64.train()
    l64, lg64, u64, i64 = m64(cast_batch(batch, torch.float64))
    l64.backward()
    g64 = {k: p.grad.detach() for k, p in m64.named_parameters() if p.grad is not None}
    mb = build_reference_model(two_tower, cfg, sd, torch.float32)
    mb.train()
    with torch.autocast(device_type="cpu", dtype=torch.bfloat16):
        lb, lgb, ub, ib = mb(cast_batch(batch, torch.float32))
    lb.backward()
    gb = {k: p.grad.detach() for k, p in mb.named_parameters() if p.grad is not None}
    auto = {
        "loss_abs": abs(lb.item() - l64.item()),
        "logits_abs": (lgb.double() - lg64).abs().max().item(),
        "user_emb_abs": (ub.double() - u64).abs().max().item(),
        "item_emb_abs": (ib.double() - i64).abs().max().item(),
        "grad_rel": {k: ((gb[k].double() - g64[k]).norm() / g64[k].norm().clamp_min(1e-300)).item() for k in g64},
        "grad_abs": {k: (gb[k].double() - g64[k]).norm().item() for k in g64},
        "grad_norm64": {k: g64[k].norm().item() for k in g64},
    }
    out["bf16_autocast_err"] = auto
    torch.save(out, os.path.join(HERE, name))
    worst = sorted(((v, k) for k, v in auto["grad_rel"].items() if auto["grad_norm64"][k] > 1e-6), reverse=True)[:4]
    print(name, "loss f64", out["f64"]["loss"].item(), "f32", out["f32"]["loss"].item(),
          "| reference under bf16 autocast: loss", auto["loss_abs"], "logits", auto["logits_abs"], "user",
          auto["user_emb_abs"], "item", auto["item_emb_abs"], "worst grad rel", worst)


class _FakeModel(nn.Module):
    """calculate_metrics_global only needs .eval() and .get_user_embedding(); feeding it
    precomputed user embeddings isolates the scoring/top-K/metric part."""

    def __init__(self, users):
        super().__init__()
        self.users = users
        self.pos = 0

    def get_user_embedding(self, history_ids, history_mask=None, user_gender=None, user_country=None):
        n = history_ids.shape[0]
        u = self.users[self.pos:self.pos + n]
        self.pos += n
        return u


def golden_retrieval(evalm, name, num_items, num_users, k_list, grid, seed, noise):
    table = synthetic.make_catalog(num_items, 256, seed=seed, grid=grid)
    users, targets = synthetic.make_queries(table, num_users, seed=seed + 1, noise=noise, grid=grid)
    loader = []
    for s in range(0, num_users, 64):
        n = min(64, num_users - s)
        loader.append({"history_ids": torch.ones((n, 4), dtype=torch.long),
                       "history_mask": torch.ones((n, 4), dtype=torch.long),
                       "user_gender": torch.zeros(n, dtype=torch.long),
                       "user_country": torch.zeros(n, dtype=torch.long),
                       "target_id": targets[s:s + n]})
    metrics = evalm.calculate_metrics_global(_FakeModel(users), loader, table, torch.device("cpu"),
                                             k_list=list(k_list))
    # the reference's own scores; top-K under the canonical order (stable sort) for index pins
    scores = users @ table.t()
    scores[:, 0] = -float("inf")
    vals, idx = torch.sort(scores, dim=1, descending=True, stable=True)
    kmax = max(k_list)
    # cross-check: torch.topk's value multiset must agree with the canonical one
    tv, ti = torch.topk(scores, kmax, dim=1)
    assert torch.equal(tv, vals[:, :kmax])
    ties = (vals[:, :kmax][:, 1:] == vals[:, :kmax][:, :-1]).sum().item()
    print(name, "adjacent ties inside the top-K lists:", ties)
    out = {"num_items": num_items, "num_users": num_users, "k_list": list(k_list), "grid": grid,
           "seed": seed, "noise": noise, "metrics": metrics, "topk_idx": idx[:, :kmax].to(torch.int32),
           "topk_val": vals[:, :kmax].clone()}
    torch.save(out, os.path.join(HERE, name))
    print(name, metrics)


def golden_anchor(two_tower, name, cfg, batch, seed_w):
    """Tolerance anchor only at a shape too large to store outputs for: the error of the
    REFERENCE's own modules under bf16 autocast against its fp64 run — loss, embeddings, logits and every
    parameter gradient (relative / absolute L2, fp64 norm). A few hundred numbers."""
    sd = synthetic.make_state_dict(cfg, seed=seed_w)
    m64 = build_reference_model(two_tower, cfg, sd, torch.float64)
    m64.train()
    l64, lg64, u64, i64 = m64(cast_batch(batch, torch.float64))
    l64.backward()
    g64 = {k: p.grad.detach() for k, p in m64.named_parameters() if p.grad is not None}
    mb = build_reference_model(two_tower, cfg, sd, torch.float32)
    mb.train()
    with torch.autocast(device_type="cpu", dtype=torch.bfloat16):
        lb, lgb, ub, ib = mb(cast_batch(batch, torch.float32))
    lb.backward()
    gb = {k: p.grad.detach() for k, p in mb.named_parameters() if p.grad is not None}
    auto = {
        "loss64": l64.item(),
        "loss_abs": abs(lb.item() - l64.item()),
        "logits_abs": (lgb.double() - lg64).abs().max().item(),
        "user_emb_abs": (ub.double() - u64).abs().max().item(),
        "item_emb_abs": (ib.double() - i64).abs().max().item(),
        "grad_rel": {k: ((gb[k].double() - g64[k]).norm() / g64[k].norm().clamp_min(1e-300)).item() for k in g64},
        "grad_abs": {k: (gb[k].double() - g64[k]).norm().item() for k in g64},
        "grad_norm64": {k: g64[k].norm().item() for k in g64},
    }
    torch.save({"config": cfg.as_dict(), "seed_w": seed_w, "bf16_autocast_err": auto}, os.path.join(HERE, name))
    worst = sorted(((v, k) for k, v in auto["grad_rel"].items() if auto["grad_norm64"][k] > 1e-6), reverse=True)[:5]
    print(name, "loss64", auto["loss64"], "autocast: loss", auto["loss_abs"], "logits", auto["logits_abs"], "user",
          auto["user_emb_abs"], "item", auto["item_emb_abs"], "worst grad rel", worst)


class _CatalogDataset(torch.utils.data.Dataset):
    """Harness stand-in for the absent src/data/dataset.py in compute_all_item_embeddings
    (evaluate_metrics.py:40-48): serves the precomputed modality embeddings of `FEATURES` row by row."""
    FEATURES = None

    def __init__(self, interactions_df, item_id_mapper, img_dir=None, audio_dir=None, text_data=None,
                 tokenizer=None, encoders=None, **kw):
        self.df, self.mapper = interactions_df.reset_index(drop=True), item_id_mapper

    def __len__(self):
        return len(self.df)

    def __getitem__(self, i):
        row = int(self.df.loc[i, "row"])
        f = type(self).FEATURES
        return {"target_id": torch.tensor(self.mapper[self.df.loc[i, "track_id"]]),
                "target_image": f["target_image"][row], "target_audio": f["target_audio"][row],
                "target_input_ids": f["target_input_ids"][row], "target_attention_mask": torch.ones(1, dtype=torch.long),
                "target_tabular": f["target_tabular"][row]}


def golden_evaluate(two_tower, train, name, cfg, B, n_batches, seed_w, seed_b, k):
    """The reference's own evaluate() (src/train.py:78-111) on seeded batches, eval mode, fp64 and fp32."""
    import torch.distributed as dist
    if not dist.is_initialized():      # evaluate() calls dist.get_rank() unguarded
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29541")
        dist.init_process_group("gloo", rank=0, world_size=1)
    sd = synthetic.make_state_dict(cfg, seed=seed_w)
    out = {"config": cfg.as_dict(), "batch_size": B, "n_batches": n_batches, "seed_w": seed_w, "seed_b": seed_b, "k": k}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        m = build_reference_model(two_tower, cfg, sd, dtype)
        batches = [cast_batch(synthetic.make_batch(cfg, B, seed=seed_b + j), dtype) for j in range(n_batches)]
        out[tag] = {"recall": train.evaluate(m, batches, torch.device("cpu"), k=k)}
        m.eval()
        with torch.no_grad():
            out[tag]["hits_per_batch"] = []
            for b in batches:
                _, logits, _, _ = m(b)
                _, topi = torch.topk(logits, k=k, dim=1)
                out[tag]["hits_per_batch"].append(int((topi == torch.arange(B).unsqueeze(1)).any(dim=1).sum()))
    torch.save(out, os.path.join(HERE, name))
    print(name, {t: out[t]["recall"] for t in ("f64", "f32")}, out["f64"]["hits_per_batch"])


def golden_index(two_tower, evalm, name, cfg, n_items, vocab_size, seed_w, seed_f, batch_size, nan_rows):
    """The reference's own compute_all_item_embeddings (src/evaluate_metrics.py:24-104) over a synthetic item
    list: a permuted subset of the ids (rows of ids not listed stay zero) with a few NaN feature rows (NaN -> 0
    branch, :79-81). Stores the dense (V, 256) fp32 table it returns."""
    import pandas as pd
    from types import SimpleNamespace
    sd = synthetic.make_state_dict(cfg, seed=seed_w)
    # features, ids and non-trivial BatchNorm running statistics (as after some training)
    feats, ids = synthetic.make_item_features(cfg, n_items, vocab_size, seed=seed_f, nan_rows=nan_rows, state_dict=sd)
    _CatalogDataset.FEATURES = feats
    mapper = {f"t{int(i)}": int(i) for i in range(1, vocab_size)}
    ds = SimpleNamespace(item_id_mapper=mapper, img_dir=None, audio_dir=None, text_data=None, tokenizer=None, encoders=None)
    df = pd.DataFrame({"track_id": [f"t{int(i)}" for i in ids], "row": list(range(n_items))})
    m = build_reference_model(two_tower, cfg, sd, torch.float32)
    dense, V = evalm.compute_all_item_embeddings(m, ds, df, batch_size, torch.device("cpu"))
    assert V == vocab_size and dense.shape == (vocab_size, cfg.embedding_dim)
    out = {"config": cfg.as_dict(), "n_items": n_items, "vocab_size": vocab_size, "seed_w": seed_w, "seed_f": seed_f,
           "batch_size": batch_size, "nan_rows": list(nan_rows), "dense": dense}
    torch.save(out, os.path.join(HERE, name))
    print(name, "rows set", int((dense.abs().sum(1) > 0).sum()), "of", n_items, "| NaNs", int(torch.isnan(dense).sum()))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    two_tower, evalm, train = import_reference(dataset_cls=_CatalogDataset)
    if "--anchor-only" in sys.argv or "--all" in sys.argv:
        c2 = synthetic.TwoTowerConfig(vocab_size=100_001, max_seq_len=200, dropout=0.0)
        golden_anchor(two_tower, "anchor_c2.pt", c2, synthetic.make_c2_parity_batch(c2), seed_w=0)
        if "--anchor-only" in sys.argv:
            return
    small = synthetic.TwoTowerConfig(vocab_size=501, num_genders=3, num_countries=7, max_seq_len=12,
                                     embedding_dim=64, num_heads=4, num_layers=2, modality_dim=16,
                                     fusion_hidden=512)
    # the reference hard-codes the 512-wide fusion hidden layer and 16/32-d demographic embeddings
    golden_train(two_tower, "train_small.pt", small, B=8, seed_w=10, seed_b=11)
    c1 = synthetic.TwoTowerConfig(vocab_size=10_001, max_seq_len=50)
    golden_train(two_tower, "train_c1.pt", c1, B=32, seed_w=0, seed_b=1)
    c1s = synthetic.TwoTowerConfig(vocab_size=2_001, max_seq_len=200)
    golden_train(two_tower, "train_l200.pt", c1s, B=16, seed_w=20, seed_b=21)
    golden_retrieval(evalm, "retrieval_grid.pt", 5_000, 192, (10, 20, 50, 100), 2.0 ** -7, 30, 5.0)
    golden_retrieval(evalm, "retrieval_grid_coarse.pt", 5_000, 192, (10, 20, 50, 100), 2.0 ** -3, 50, 5.0)
    golden_retrieval(evalm, "retrieval_float.pt", 5_000, 192, (10, 20, 50, 100), 0.0, 40, 5.0)
    ev = synthetic.TwoTowerConfig(vocab_size=2_001, max_seq_len=20)
    golden_evaluate(two_tower, train, "evaluate_inbatch.pt", ev, B=32, n_batches=3, seed_w=60, seed_b=61, k=10)
    golden_index(two_tower, evalm, "index_catalog.pt", synthetic.TwoTowerConfig(vocab_size=401), n_items=300,
                 vocab_size=401, seed_w=70, seed_f=71, batch_size=64, nan_rows=(5, 130))


if __name__ == "__main__":
    main()
