"""Test helper: one process, one GPU, G "virtual ranks" of the data-parallel step with all-gathered negatives
(train.TrainStepRunner's sequence with the NCCL exchanges replaced by torch.cat / copies between G engines)."""
import torch


def virtual_dp_step(engs, dbatches, backward=True):
    """engs[r] holds rank r's replica, dbatches[r] its device batch. Returns (global loss, per-rank workspaces).
    After the call engs[r].grad holds rank r's gradient (scaled so that the MEAN over ranks is the gradient of the
    global loss, as the reduce-scatter(AVG) of the real step makes it) and ws['dun'] / ws['din'] the gradient
    w.r.t. the rank's normalised embeddings times G."""
    G = len(engs)
    B = dbatches[0]["history_ids"].shape[0]
    wss = [e.forward_towers(b, training=True) for e, b in zip(engs, dbatches)]
    U_all = torch.cat([ws["un_bf"] for ws in wss])
    I_all = torch.cat([ws["in_bf"] for ws in wss])
    has_uid = "user_idx" in dbatches[0]
    uid_all = torch.cat([b["user_idx"] for b in dbatches]) if has_uid else None
    gs = []
    for r, (e, ws) in enumerate(zip(engs, wss)):
        g = e.gathered_workspace(ws, G)
        g["U_all"].copy_(U_all)
        g["I_all"].copy_(I_all)
        if has_uid:
            g["uid_all"].copy_(uid_all)
        e.loss_forward(ws, dbatches[r].get("user_idx"), gathered=g, rank=r)
        gs.append(g)
    lr_all = torch.cat([ws["lse_r"] for ws in wss])
    lc_all = torch.cat([ws["lse_c"] for ws in wss])
    total = 0.0
    for e, ws, g in zip(engs, wss, gs):
        g["lse_r_all"].copy_(lr_all)
        g["lse_c_all"].copy_(lc_all)
        total += e.loss_value(ws, G * B).item()
        if backward:
            e.backward()
    return total, wss


def virtual_dp_step_one_engine(eng, dbatches):
    """The same step with ONE engine replaying the ranks in turn (dropout must be off: towers are recomputed
    instead of kept, so memory stays that of one rank even at 8 x 512 x 200). Returns (global loss, mean over
    ranks of the flat gradient = gradient of the global loss, dU_all, dI_all = d loss / d normalised embeddings)."""
    G = len(dbatches)
    B = dbatches[0]["history_ids"].shape[0]
    has_uid = "user_idx" in dbatches[0]
    uid_all = torch.cat([b["user_idx"] for b in dbatches]) if has_uid else None
    U, I = [], []
    for b in dbatches:
        ws = eng.forward_towers(b, training=True)
        U.append(ws["un_bf"].clone())
        I.append(ws["in_bf"].clone())
    U_all, I_all = torch.cat(U), torch.cat(I)

    def rows(r):
        ws = eng.forward_towers(dbatches[r], training=True)
        g = eng.gathered_workspace(ws, G)
        g["U_all"].copy_(U_all)
        g["I_all"].copy_(I_all)
        if has_uid:
            g["uid_all"].copy_(uid_all)
        eng.loss_forward(ws, dbatches[r].get("user_idx"), gathered=g, rank=r)
        return ws, g

    lse_r, lse_c = [], []
    for r in range(G):
        ws, _ = rows(r)
        lse_r.append(ws["lse_r"].clone())
        lse_c.append(ws["lse_c"].clone())
    lr_all, lc_all = torch.cat(lse_r), torch.cat(lse_c)
    total, gsum, dU, dI = 0.0, torch.zeros_like(eng.grad), [], []
    for r in range(G):
        ws, g = rows(r)
        g["lse_r_all"].copy_(lr_all)
        g["lse_c_all"].copy_(lc_all)
        total += eng.loss_value(ws, G * B).item()
        eng.grad.zero_()
        eng.backward()
        gsum += eng.grad
        # the engine scales per-rank gradients by 0.5 / B_local (the all-reduce averages): / G = global-loss scale
        dU.append(ws["dun"].clone() / G)
        dI.append(ws["din"].clone() / G)
    eng.grad.zero_()
    return total, gsum / G, torch.cat(dU), torch.cat(dI)
