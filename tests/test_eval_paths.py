"""Device paths of SURVEY.md §8(f) rows 2 and 3 against the oracle and the reference's own outputs:
catalog indexing (src/evaluate_metrics.py:24-104) and in-batch evaluate (src/train.py:78-111)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _model(cfg, sd):
    from mrm_b200.models import TwoTowerModel
    m = TwoTowerModel(vocab_size=cfg.vocab_size, num_genders=cfg.num_genders, num_countries=cfg.num_countries,
                      max_seq_len=cfg.max_seq_len, user_dropout=0.0)
    m.load_state_dict(sd)
    return m


@pytest.mark.parametrize("R,C,pos0", [(64, 64, 0), (37, 120, 50), (256, 2048, 1024)])
def test_inbatch_recall_kernel_equals_the_oracle_rule(R, C, pos0):
    """Counts, not approximations: logits on a coarse grid (many exact ties, incl. with the positive) and -1e4
    collision fills; the device (hits, rows) pair must equal the oracle's canonical-order rule for every k."""
    from mrm_b200 import ops
    from oracle import two_tower_oracle as oracle
    g = torch.Generator().manual_seed(R * 1000 + C)
    S = torch.round(torch.randn(R, C, generator=g) * 2) / 2
    S[torch.rand(R, C, generator=g) < 0.02] = -1e4
    Sd = torch.zeros(R, (C + 7) // 8 * 8 + 8).cuda()[:, :C]      # padded leading dimension
    Sd.copy_(S)
    sq = S[:, pos0:pos0 + R] if C != R else S
    for k in (1, 5, 10, R):
        acc = torch.zeros(2, device="cuda")
        ops.inbatch_recall(Sd, pos0, k, acc)
        ops.inbatch_recall(Sd, pos0, k, acc)                      # accumulates across batches
        # oracle on the square problem: shift the positive to column i by rolling each row's view
        diag = S[torch.arange(R), pos0 + torch.arange(R)].unsqueeze(1)
        col = torch.arange(C).unsqueeze(0)
        better = (S > diag) | ((S == diag) & (col < (pos0 + torch.arange(R)).unsqueeze(1)))
        hits = int((better.sum(1) < k).sum())
        if C == R:
            assert hits == int(oracle.inbatch_hits(S, k).sum())
        assert acc.tolist() == [2.0 * hits, 2.0 * R], (k, acc.tolist(), hits)


def test_evaluate_matches_the_reference_golden():
    """train.evaluate on the CUDA path vs the reference's evaluate() on the same seeded batches. The reference ran
    in fp64/fp32; bf16 operands can move a logit across the k-th rank, so the recall may differ by a few rows of
    the 96 — and it must equal the oracle rule applied to the CUDA path's own logits exactly."""
    from mrm_b200 import synthetic
    from mrm_b200.train import evaluate
    from oracle import two_tower_oracle as oracle
    gold = torch.load(os.path.join(GOLDEN, "evaluate_inbatch.pt"), weights_only=False)
    cfg = synthetic.TwoTowerConfig(**gold["config"])
    sd = synthetic.make_state_dict(cfg, seed=gold["seed_w"])
    m = _model(cfg, sd)
    batches = [synthetic.make_batch(cfg, gold["batch_size"], seed=gold["seed_b"] + j) for j in range(gold["n_batches"])]
    r = evaluate(m, batches, torch.device("cuda"), k=gold["k"])
    n = gold["batch_size"] * gold["n_batches"]
    assert abs(r - gold["f64"]["recall"]) <= 3.0 / n + 1e-7, (r, gold["f64"]["recall"])
    m.eval()
    with torch.no_grad():
        logits = [m({k: v.cuda() for k, v in b.items()})[1].cpu() for b in batches]
    assert r == pytest.approx(oracle.evaluate_inbatch(logits, gold["k"]), abs=1e-7)


def test_catalog_indexing_matches_the_reference_golden():
    from mrm_b200 import synthetic
    from mrm_b200.evaluate_metrics import build_catalog_index, compute_all_item_embeddings, index_catalog_device
    gold = torch.load(os.path.join(GOLDEN, "index_catalog.pt"), weights_only=False)
    cfg = synthetic.TwoTowerConfig(**gold["config"])
    sd = synthetic.make_state_dict(cfg, seed=gold["seed_w"])
    feats, ids = synthetic.make_item_features(cfg, gold["n_items"], gold["vocab_size"], seed=gold["seed_f"],
                                              nan_rows=gold["nan_rows"], state_dict=sd)
    m = _model(cfg, sd)
    ref = gold["dense"]
    dense, V = compute_all_item_embeddings(m, feats, ids, gold["batch_size"], torch.device("cuda"), gold["vocab_size"])
    assert V == gold["vocab_size"] and dense.device.type == "cpu" and dense.shape == ref.shape
    assert not torch.isnan(dense).any()
    assert torch.equal(dense.abs().sum(1) > 0, ref.abs().sum(1) > 0)          # same rows populated, NaN rows zero
    err = (dense - ref).abs().max().item()
    assert err <= 5e-3, err                                                   # unit vectors, bf16 tensor-core operands
    norms = dense.norm(dim=1)
    assert ((norms - 1).abs()[norms > 0]).max().item() <= 1e-6
    # device-resident route: same table, bf16 copy = rounding of the fp32 one, features already on the GPU
    dfeats = {k: v.cuda() for k, v in feats.items()}
    t32, t16 = index_catalog_device(m, dfeats, ids.cuda(), gold["vocab_size"], batch_size=128)
    assert torch.equal(t32.cpu(), dense) or (t32.cpu() - dense).abs().max().item() <= 1e-6
    assert torch.equal(t16, t32.to(torch.bfloat16))
    idx = build_catalog_index(m, dfeats, ids.cuda(), gold["vocab_size"])
    assert idx.table.shape == (gold["vocab_size"], 256) and not idx.is_sharded
    # two item-list shards (what two ranks index) add up to the full table
    parts = [index_catalog_device(m, feats, ids, gold["vocab_size"], batch_size=64, shard=(r, 2))[0] for r in range(2)]
    assert torch.equal(parts[0] + parts[1], t32) or (parts[0] + parts[1] - t32).abs().max().item() <= 1e-6
