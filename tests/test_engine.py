"""End-to-end parity of the CUDA path (forward, backward, AdamW) with the reference.

Forward outputs are compared with the golden vectors produced by the REFERENCE's own modules
(tests/golden/*.pt, fp64 run); full parameter gradients with the CPU oracle (itself pinned to
those vectors in test_oracle_golden.py).

Tolerance rule. The kernels compute with bf16 tensor-core operands, fp32
accumulation and an fp32 residual stream; the reference trains under 16-bit autocast. The
goldens therefore also store the error of the REFERENCE's own modules under bf16 autocast
against its fp64 run ("bf16_autocast_err"), and every quantity must satisfy
    err(ours, fp64) <= max(2 * err(reference under bf16 autocast, fp64), floor)
with floors: normalised embeddings 5e-3 abs, logits (x 1/0.07) 8e-2 abs, loss 5e-3 abs,
parameter gradients 3e-2 * ||g_ref|| + 1e-5 * sqrt(numel) in L2 norm per tensor.
Measured values are appended to gpurun_out/parity_report.jsonl.
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _report(**kw):
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_report.jsonl", "a") as f:
        f.write(json.dumps(kw) + "\n")


def _setup(name):
    from mrm_b200 import synthetic
    from mrm_b200.engine import TwoTowerEngine
    gold = torch.load(os.path.join(GOLDEN, name), weights_only=False)
    cfg = synthetic.TwoTowerConfig(**gold["config"])
    cfg.dropout = 0.0
    sd = synthetic.make_state_dict(cfg, seed=gold["seed_w"])
    batch = synthetic.make_batch(cfg, gold["batch_size"], seed=gold["seed_b"], full_length=gold["full_length"])
    eng = TwoTowerEngine(cfg)
    eng.load_state_dict(sd)
    dbatch = {k: v.cuda() for k, v in batch.items()}
    return gold, cfg, sd, batch, eng, dbatch


@pytest.mark.parametrize("name", ["train_c1.pt", "train_l200.pt"])
def test_forward_matches_reference_golden(name):
    gold, cfg, sd, batch, eng, dbatch = _setup(name)
    loss, logits, u, i = eng.forward(dbatch, training=True)
    torch.cuda.synchronize()
    g, auto = gold["f64"], gold["bf16_autocast_err"]
    eu = (u.cpu().double() - g["user_emb"]).abs().max().item()
    ei = (i.cpu().double() - g["item_emb"]).abs().max().item()
    el = (logits.cpu().double() - g["logits"]).abs().max().item()
    eloss = abs(loss.item() - g["loss"].item())
    _report(test="forward", fixture=name, user_emb_abs=eu, item_emb_abs=ei, logits_abs=el, loss_abs=eloss,
            reference_bf16_autocast={k: auto[k] for k in ("user_emb_abs", "item_emb_abs", "logits_abs", "loss_abs")})
    assert eu <= max(2 * auto["user_emb_abs"], 5e-3), eu
    assert ei <= max(2 * auto["item_emb_abs"], 5e-3), ei
    assert el <= max(2 * auto["logits_abs"], 8e-2), el
    assert eloss <= max(2 * auto["loss_abs"], 5e-3), eloss
    assert (eng.bn_running_mean.cpu().double() - g["bn_running_mean"]).abs().max().item() <= 2e-3
    assert (eng.bn_running_var.cpu().double() - g["bn_running_var"]).abs().max().item() <= 2e-3
    assert eng.bn_num_batches.item() == 1


@pytest.mark.parametrize("name", ["train_c1.pt", "train_l200.pt"])
def test_backward_matches_oracle(name):
    from oracle import two_tower_oracle as oracle
    gold, cfg, sd, batch, eng, dbatch = _setup(name)
    eng.forward(dbatch, training=True)
    eng.backward()
    torch.cuda.synchronize()
    b64 = {k: (v.double() if v.is_floating_point() else v) for k, v in batch.items()}
    _, _, _, _, grads, _ = oracle.loss_and_grads(sd, b64, cfg.temperature, cfg.num_heads, dtype=torch.float64)
    auto = gold["bf16_autocast_err"]
    rows, bad = [], []
    for k, ref in grads.items():
        got = eng.g[k].cpu().double()
        err = (got - ref).norm().item()
        bound = max(2 * auto["grad_abs"].get(k, 0.0), 3e-2 * ref.norm().item() + 1e-5 * ref.numel() ** 0.5)
        rows.append((err / max(ref.norm().item(), 1e-12), k, err, bound))
        if err > bound:
            bad.append((k, err, bound))
    rows.sort(reverse=True)
    _report(test="backward", fixture=name,
            worst_rel=[(r, k) for r, k, _, _ in rows if grads[k].norm().item() > 1e-6][:5],
            reference_bf16_autocast_worst_rel=sorted(
                ((v, k) for k, v in auto["grad_rel"].items() if auto["grad_norm64"][k] > 1e-6), reverse=True)[:5])
    assert not bad, bad[:6]
    assert eng.g["user_tower.item_embedding.weight"][0].abs().max().item() == 0.0


def test_eval_embeddings_match_reference_golden():
    gold, cfg, sd, batch, eng, dbatch = _setup("train_c1.pt")
    g = gold["f64"]
    eng.bn_running_mean.copy_(g["bn_running_mean"].float())
    eng.bn_running_var.copy_(g["bn_running_var"].float())
    B, L = dbatch["history_ids"].shape
    ws = eng.workspace(B, L)
    eng.refresh_shadow()
    auto = gold["bf16_autocast_err"]
    tol_u, tol_i = max(2 * auto["user_emb_abs"], 5e-3), max(2 * auto["item_emb_abs"], 5e-3)
    u = eng.user_forward(ws, dbatch["history_ids"], dbatch["history_mask"], dbatch["user_gender"],
                         dbatch["user_country"], training=False)
    eu = (u.cpu().double() - g["eval_user_emb"]).abs().max().item()
    i = eng.item_forward(ws, dbatch["target_audio"], dbatch["target_image"], dbatch["target_input_ids"],
                         dbatch["target_tabular"], training=False)
    ei = (i.cpu().double() - g["eval_item_emb"]).abs().max().item()
    z = torch.zeros_like(dbatch["user_gender"])
    u2 = eng.user_forward(ws, dbatch["history_ids"], None, z, z, training=False)
    eu2 = (u2.cpu().double() - g["eval_user_emb_nomask"]).abs().max().item()
    _report(test="eval_embeddings", user_emb_abs=eu, item_emb_abs=ei, user_emb_nomask_abs=eu2, tol_user=tol_u, tol_item=tol_i)
    assert eu <= tol_u and eu2 <= tol_u, (eu, eu2, tol_u)
    assert ei <= tol_i, (ei, tol_i)


def test_train_steps_follow_oracle_adamw():
    """Three fused steps (fwd + bwd + AdamW) track the oracle's fp32 training loop."""
    from oracle import two_tower_oracle as oracle
    gold, cfg, sd, batch, eng, dbatch = _setup("train_c1.pt")
    p = {k: v.clone() for k, v in sd.items()}
    m = {k: torch.zeros_like(v) for k, v in p.items() if v.is_floating_point()}
    v2 = {k: torch.zeros_like(v) for k, v in p.items() if v.is_floating_point()}
    losses_ref, losses = [], []
    for t in range(1, 4):
        loss, _, _, _, grads, _ = oracle.loss_and_grads(p, batch, cfg.temperature, cfg.num_heads)
        losses_ref.append(loss.item())
        for k, gk in grads.items():
            p[k], m[k], v2[k] = oracle.adamw_step(p[k], gk, m[k], v2[k], t, lr=1e-3)
        losses.append(eng.train_step(dbatch, lr=1e-3).item())
    torch.cuda.synchronize()
    for a, b in zip(losses, losses_ref):
        assert abs(a - b) <= 5e-2, (losses, losses_ref)
    assert losses[-1] < losses[0]
    # parameters moved the same way, element-wise. Adam normalises every element's step to ~lr, so elements whose
    # gradient is rounding noise may move by up to 3 * lr in either direction; everywhere else the update must
    # match: per dense tensor, the relative L2 difference of the update vectors stays small and no element is
    # further from the oracle than the total distance an element can travel in three steps.
    worst = []
    for k in ("user_tower.fusion_layer.3.weight", "user_tower.fusion_layer.0.weight", "item_tower.fusion_layer.4.weight",
              "user_tower.transformer_encoder.layers.1.linear2.weight", "user_tower.transformer_encoder.layers.0.self_attn.in_proj_weight"):
        d_ref = p[k] - sd[k]
        d_got = eng.p[k].cpu() - sd[k]
        rel = ((d_ref - d_got).norm() / d_ref.norm()).item()
        worst.append((rel, k))
        assert rel <= 0.5, (k, rel)
        assert (d_ref - d_got).abs().max().item() <= 2 * 3 * 1e-3 + 1e-6, k
    _report(test="adamw_3step_updates", rel_update_diff=sorted(worst, reverse=True))
    assert eng.grad.abs().max().item() == 0.0     # zeroed by the fused step


def test_dropout_training_runs_and_is_seeded():
    gold, cfg, sd, batch, eng, dbatch = _setup("train_c1.pt")
    eng.cfg.dropout = 0.1
    l1 = eng.forward(dbatch, training=True)[0].item()
    eng.backward()
    l2 = eng.forward(dbatch, training=True)[0].item()
    assert l1 == l2                       # same seed counter -> same masks
    eng.grad.zero_()
    eng.adamw_step()                      # advances the seed counter
    l3 = eng.forward(dbatch, training=True)[0].item()
    assert l3 != l1
    assert abs(l1 - gold["f64"]["loss"].item()) < 0.5
    assert torch.isfinite(eng.grad).all()


def test_gathered_negatives_two_virtual_ranks_match_global_oracle():
    """Data parallel with all-gathered negatives (BASELINE config 4), emulated on one GPU with two
    engines holding the same weights: each 'rank' runs its towers on its own half batch (per-rank
    BatchNorm statistics, reference DDP semantics), the embeddings / ids / row log-sum-exps are
    exchanged by hand exactly where the NCCL all-gathers sit, and the averaged gradients must equal
    the gradient of the GLOBAL symmetric InfoNCE computed by the oracle on the same embeddings."""
    from mrm_b200 import synthetic
    from mrm_b200.engine import TwoTowerEngine
    from oracle import two_tower_oracle as oracle
    cfg = synthetic.TwoTowerConfig(vocab_size=2001, max_seq_len=50, dropout=0.0)
    sd = synthetic.make_state_dict(cfg, seed=3)
    G, B = 2, 16
    batches = [synthetic.make_batch(cfg, B, seed=40 + r, num_users=12) for r in range(G)]   # collisions across ranks
    engs = []
    for r in range(G):
        e = TwoTowerEngine(cfg)
        e.load_state_dict(sd)
        engs.append(e)
    wss = [e.forward_towers({k: v.cuda() for k, v in b.items()}, training=True) for e, b in zip(engs, batches)]
    U_all = torch.cat([ws["un_bf"] for ws in wss])
    I_all = torch.cat([ws["in_bf"] for ws in wss])
    uid_all = torch.cat([b["user_idx"] for b in batches]).cuda()
    gs = []
    for r, (e, ws) in enumerate(zip(engs, wss)):
        g = e.gathered_workspace(ws, G)
        g["U_all"].copy_(U_all); g["I_all"].copy_(I_all); g["uid_all"].copy_(uid_all)
        e.loss_forward(ws, batches[r]["user_idx"].cuda(), gathered=g, rank=r)
        gs.append(g)
    lse_r_all = torch.cat([ws["lse_r"] for ws in wss])
    lse_c_all = torch.cat([ws["lse_c"] for ws in wss])
    loss = 0.0
    for r, (e, ws, g) in enumerate(zip(engs, wss, gs)):
        g["lse_r_all"].copy_(lse_r_all); g["lse_c_all"].copy_(lse_c_all)
        loss += e.loss_value(ws, G * B).item()
        e.backward()
    torch.cuda.synchronize()
    avg_grad = {k: sum(e.g[k] for e in engs).cpu().double() / G for k in engs[0].g}

    # oracle: per-rank towers (per-rank BN statistics), global loss on the concatenated embeddings
    p = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(oracle.TRAINABLE_SKIP)
             else v.clone()) for k, v in sd.items()}
    us, its = [], []
    for b in batches:
        b64 = {k: (v.double() if v.is_floating_point() else v) for k, v in b.items()}
        us.append(oracle.l2_normalize(oracle.user_tower(p, b64["history_ids"], b64["user_gender"], b64["user_country"],
                                                        b64["history_mask"], cfg.num_heads)))
        it, _ = oracle.item_fusion(p, b64["target_audio"], b64["target_image"], b64["target_input_ids"],
                                   b64["target_tabular"], training=True)
        its.append(oracle.l2_normalize(it))
    ref_loss, _ = oracle.infonce(torch.cat(us), torch.cat(its), cfg.temperature, uid_all.cpu())
    ref_loss.backward()
    assert abs(loss - ref_loss.item()) <= 2e-2, (loss, ref_loss.item())
    bad = []
    for k, got in avg_grad.items():
        ref = p[k].grad
        if k == "user_tower.item_embedding.weight":
            ref = ref.clone(); ref[0] = 0
        err = (got - ref).norm().item()
        # bf16 operand noise on a 2 x 16 batch: the reference itself under bf16 autocast shows 6-9 %
        # per-tensor error on the comparable fixtures; a wrong exchange would show O(1) errors
        # (the Linear bias in front of BatchNorm has an exactly-zero true gradient: absolute floor)
        if err > 0.15 * ref.norm().item() + 5e-5 * ref.numel() ** 0.5:
            bad.append((k, err, ref.norm().item()))
    assert not bad, bad[:5]


def test_pruned_last_layer_equals_full_computation():
    """The single-row last layer is exact: same loss / embeddings / gradients as running
    the last layer on every position (differences are bf16 rounding of different but equivalent paths)."""
    gold, cfg, sd, batch, eng, dbatch = _setup("train_l200.pt")
    from mrm_b200.engine import TwoTowerEngine
    full = TwoTowerEngine(cfg)
    full.load_state_dict(sd)
    full.prune_last_layer = False
    assert eng.prune_last_layer
    out = {}
    for name, e in (("pruned", eng), ("full", full)):
        loss, logits, u, i = e.forward(dbatch, training=True)
        e.backward()
        torch.cuda.synchronize()
        out[name] = (loss.item(), u.clone(), {k: v.clone() for k, v in e.g.items()})
    assert abs(out["pruned"][0] - out["full"][0]) <= 5e-3
    assert (out["pruned"][1] - out["full"][1]).abs().max().item() <= 3e-3
    for k, gp in out["pruned"][2].items():
        gf = out["full"][2][k]
        err = (gp - gf).norm().item()
        # two different bf16 paths: each is within ~6 % of the fp64 oracle on its own
        assert err <= 0.12 * gf.norm().item() + 5e-5 * gf.numel() ** 0.5, (k, err, gf.norm().item())
    # and the full path still matches the reference golden on its own
    g = gold["f64"]
    assert abs(out["full"][0] - g["loss"].item()) <= 5e-3


def test_lastq_attention_matches_torch():
    from mrm_b200 import ops
    B, L, H = 5, 77, 4
    gen = torch.Generator().manual_seed(9)
    qkv = torch.randn(B * L, 3 * H * 64, generator=gen).cuda().bfloat16()
    last = torch.tensor([0, 76, 30, 5, 63], dtype=torch.int32).cuda()
    q = torch.stack([qkv.view(B, L, -1)[b, last[b], :H * 64] for b in range(B)]).contiguous()
    ctx = torch.empty(B, H * 64, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, device="cuda")
    ops.attn_lastq_fwd(q, qkv, last, ctx, lse, B, L, H)
    dctx = torch.randn(B, H * 64, generator=gen).cuda().bfloat16()
    dq = torch.empty_like(q)
    dqkv = torch.full_like(qkv, float("nan"))
    ops.attn_lastq_bwd(q, qkv, last, ctx, dctx, lse, dq, dqkv, B, L, H)
    torch.cuda.synchronize()
    x = qkv.float().view(B, L, 3, H, 64)
    for b in range(B):
        n = int(last[b]) + 1
        qq = q[b].float().view(H, 64).clone().requires_grad_(True)
        k = x[b, :n, 1].clone().requires_grad_(True)          # (n, H, 64)
        v = x[b, :n, 2].clone().requires_grad_(True)
        s = torch.einsum("hd,nhd->hn", qq, k) / 8.0
        o = torch.einsum("hn,nhd->hd", torch.softmax(s, dim=-1), v)
        assert (ctx[b].float().view(H, 64) - o).abs().max().item() < 2e-2
        assert (lse[b] - torch.logsumexp(s, dim=-1)).abs().max().item() < 2e-3
        o.backward(dctx[b].float().view(H, 64))
        assert (dq[b].float().view(H, 64) - qq.grad).abs().max().item() < 3e-2 * max(1.0, qq.grad.abs().max().item())
        dk = dqkv.view(B, L, 3, H, 64)[b, :, 1].float()
        dv = dqkv.view(B, L, 3, H, 64)[b, :, 2].float()
        assert (dk[:n] - k.grad).abs().max().item() < 3e-2 * max(1.0, k.grad.abs().max().item())
        assert (dv[:n] - v.grad).abs().max().item() < 3e-2 * max(1.0, v.grad.abs().max().item())
        assert dk[n:].abs().max().item() == 0.0 if n < L else True
        assert dv[n:].abs().max().item() == 0.0 if n < L else True


def test_seq512_training_step_matches_oracle():
    """BASELINE config 5's sequence length (L = 512, four attention tiles, LONG backward layout)."""
    from mrm_b200 import synthetic
    from mrm_b200.engine import TwoTowerEngine
    from oracle import two_tower_oracle as oracle
    cfg = synthetic.TwoTowerConfig(vocab_size=3001, max_seq_len=512, dropout=0.0)
    sd = synthetic.make_state_dict(cfg, seed=12)
    batch = synthetic.make_batch(cfg, 8, seed=13)
    batch["history_ids"][0, :] = torch.randint(1, 3001, (512,))      # one full-length history
    batch["history_mask"][0, :] = 1
    eng = TwoTowerEngine(cfg)
    eng.load_state_dict(sd)
    loss, _, u, i = eng.forward({k: v.cuda() for k, v in batch.items()}, training=True)
    eng.backward()
    torch.cuda.synchronize()
    ref_loss, _, ref_u, ref_i, grads, _ = oracle.loss_and_grads(sd, batch, cfg.temperature, cfg.num_heads)
    assert abs(loss.item() - ref_loss.item()) <= 2e-2
    assert (u.cpu() - ref_u).abs().max().item() <= 1e-2
    for k in ("user_tower.transformer_encoder.layers.0.self_attn.in_proj_weight",
              "user_tower.transformer_encoder.layers.0.linear1.weight",
              "user_tower.transformer_encoder.layers.1.self_attn.in_proj_weight",
              "user_tower.position_embedding.weight", "user_tower.item_embedding.weight"):
        ref = grads[k]
        err = (eng.g[k].cpu() - ref).norm().item()
        assert err <= 0.15 * ref.norm().item() + 5e-5 * ref.numel() ** 0.5, (k, err, ref.norm().item())


def test_row_sharded_table_step_equals_replicated_table():
    """Config-5 layout on one rank (the all-to-all degenerates to a copy; the gloo tests cover the exchange):
    fetched-rows buffer + per-token gradient rows + owner-side scatter-add + local-row AdamW reproduce the
    replicated table's training steps."""
    import dataclasses
    from mrm_b200.engine import TwoTowerEngine
    from mrm_b200.sharding import RowShardedTable
    from mrm_b200.train import TrainStepRunner
    gold, cfg, sd, batch, eng, dbatch = _setup("train_c1.pt")
    B, L = dbatch["history_ids"].shape
    name = "user_tower.item_embedding.weight"
    ref_runner = TrainStepRunner(eng, B, L, lr=1e-3, use_graph=False)
    cfg2 = dataclasses.replace(cfg, vocab_size=2)
    eng2 = TwoTowerEngine(cfg2)
    table = RowShardedTable(cfg.vocab_size, 256, 0, 1, eng2.device)
    runner = TrainStepRunner(eng2, B, L, lr=1e-3, use_graph=True, sharded_table=table)
    eng2.load_state_dict(sd)
    table.load_full(sd[name].cuda())
    for step in range(3):
        ref_runner.load_batch(dbatch)
        runner.load_batch(dbatch)
        a = ref_runner.step_resident().item()
        b = runner.step_resident().item()
        assert abs(a - b) < 2e-3, (step, a, b)
    torch.cuda.synchronize()
    d = (table.weight - eng.p[name]).abs()
    assert d.max().item() < 3e-3 and d.mean().item() < 2e-5, (d.max().item(), d.mean().item())
    moved = (eng.p[name].cpu() - sd[name]).abs().max().item()
    assert moved > 1e-4                       # the table really trained
    assert table.grad.abs().max().item() == 0.0 and eng2.table_rows_grad.abs().max().item() == 0.0


def test_prune_toggle_after_a_step_restores_the_reference_schedule():
    """prune_last_layer switched off on an engine whose (B, L) workspace already exists (built without the
    full-layer backward buffers): the next step must run every layer on every position and agree with the
    pruned step it replaces."""
    gold, cfg, sd, batch, eng, dbatch = _setup("train_c1.pt")
    l_pruned = eng.forward(dbatch, training=True)[0].item()
    eng.backward()
    torch.cuda.synchronize()
    g_pruned = eng.grad.clone()
    eng.grad.zero_()
    eng.prune_last_layer = False
    l_full = eng.forward(dbatch, training=True)[0].item()
    eng.backward()
    torch.cuda.synchronize()
    assert abs(l_full - l_pruned) <= 2e-3
    rel = ((eng.grad - g_pruned).norm() / g_pruned.norm()).item()
    assert rel <= 3e-2, rel


def test_deterministic_table_gradient_mode_of_the_engine():
    """engine.deterministic_table_grad: the ID-table gradient of a whole backward is bit-identical from run to run and
    agrees with the default (floating-point atomics) mode to fp32 rounding; every other gradient is untouched."""
    gold, cfg, sd, batch, eng, dbatch = _setup("train_l200.pt")
    key = "user_tower.item_embedding.weight"

    def grads(det):
        eng.deterministic_table_grad = det
        eng.grad.zero_()
        eng.forward(dbatch, training=True)
        eng.backward()
        torch.cuda.synchronize()
        return eng.grad.clone(), eng.g[key].clone()

    flat_a, tab_a = grads(True)
    flat_b, tab_b = grads(True)
    assert torch.equal(tab_a, tab_b)
    flat_c, tab_c = grads(False)
    scale = float(tab_c.abs().max())
    assert scale > 0 and float((tab_a - tab_c).abs().max()) <= 2e-6 * scale
    assert float(tab_a[0].abs().max()) == 0.0           # padding_idx row
