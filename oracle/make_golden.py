"""Generates tests/golden/*.npz by RUNNING THE REFERENCE (imported, unmodified, from
/root/reference) on seeded inputs.  Only runs in the build container; the GPU box has no
/root/reference, which is why the outputs are committed.

    python oracle/make_golden.py            # rewrites tests/golden/

Every fixture stores the inputs, the initial weights and the reference's outputs, so both
the oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_parity.py,
tests/test_gpu_lthm_step.py) can be checked against what the reference itself computed.
torch version is recorded.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def ref_modules():
    if not REF.exists():
        raise SystemExit("/root/reference is not mounted: golden vectors can only be regenerated "
                         "in the build container")
    sys.path.insert(0, str(REF))
    import commons.feature_utils as fu  # noqa: E402
    import commons.layers as cl  # noqa: E402
    import commons.transformers.layers as tl  # noqa: E402
    return cl, tl, fu


EDGE_IDS = [0, 1, -1, 2 ** 62, -2 ** 63, 2 ** 63 - 1, 12345678901234, -987654321,
            2, -2, 2 ** 32, -2 ** 32, 2 ** 31 - 1, -2 ** 31, 999, 1000, 1001, -999, -1000, -1001]


def seeded_ids(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-2 ** 63, 2 ** 63 - 1, (n,), generator=g, dtype=torch.int64)


def main():
    cl, tl, fu = ref_modules()
    OUT.mkdir(parents=True, exist_ok=True)
    meta = dict(torch_version=torch.__version__)

    # ---- 1. row-index known answers (KShiftEmbedding.get_row_idx) ----------------
    ids = torch.cat([torch.tensor(EDGE_IDS, dtype=torch.int64), seeded_ids(236, 11)])
    sizes = [1, 2, 7, 1000, 999983, 1 << 20, 1000000, (1 << 31) - 1, (1 << 34), 3 * (1 << 40) + 17]
    cols = [0, 1, 2, 3, 7, 8, 15, 16, 31, 32, 33, 62, 63]
    rows = np.zeros((len(sizes), len(cols), ids.numel()), dtype=np.int64)
    for si, n_rows in enumerate(sizes):
        m = cl.KShiftEmbedding(4, 2, num_shifts=2)  # tiny table; only the hashing is used
        m._num_embeddings = n_rows
        for ci, c in enumerate(cols):
            rows[si, ci] = m.get_row_idx(ids, c).numpy()
    np.savez_compressed(OUT / "row_index.npz", ids=ids.numpy(), sizes=np.array(sizes, dtype=np.int64),
                        cols=np.array(cols, dtype=np.int64), rows=rows, **meta)

    # ---- 2. FlatEmbedding forward (+ normalise, padding_idx) ----------------------
    torch.manual_seed(1234)
    fe = cl.FlatEmbedding(1000, 32, padding_idx=0)
    fe_n = cl.FlatEmbedding(1000, 32, normalize_output=True)
    fe_n._emb_table.weight.data.copy_(fe._emb_table.weight.data)
    fids = seeded_ids(6 * 50, 7).view(6, 50)
    fids[:, 40:] = 0  # right padding
    np.savez_compressed(OUT / "flat_embedding.npz", weight=fe._emb_table.weight.detach().numpy(),
                        ids=fids.numpy(), out=fe(fids).detach().numpy(),
                        out_norm=fe_n(fids).detach().numpy(), **meta)

    # ---- 3. KShiftEmbedding forward ----------------------------------------------
    torch.manual_seed(1235)
    kids = seeded_ids(4 * 33, 8).view(4, 33)
    ks = {}
    for k, norm in ((4, False), (8, False), (16, True), (16, False)):
        m = cl.KShiftEmbedding(1009, 32, num_shifts=k, normalize_output=norm)
        if "weight" not in ks:
            ks["weight"] = m.emb.weight.detach().numpy().copy()
        m.emb.weight.data.copy_(torch.from_numpy(ks["weight"]))
        ks[f"out_k{k}_{'norm' if norm else 'scale'}"] = m(kids).detach().numpy()
    np.savez_compressed(OUT / "kshift_embedding.npz", ids=kids.numpy(), **ks, **meta)

    # ---- 4. KShift fwd + bwd + Adagrad (embedding_module_gen.train_model body) ------
    torch.manual_seed(1236)
    n_prod, dim, k = 1024, 32, 16
    prod_ids = seeded_ids(n_prod, 9)
    target = F.normalize(torch.randn(n_prod, dim), p=2.0, dim=-1)
    model = cl.KShiftEmbedding(int(1.15 * n_prod), dim, num_shifts=k, normalize_output=True)
    w0 = model.emb.weight.detach().numpy().copy()
    optim = torch.optim.Adagrad(model.parameters(), lr=5e-1)
    crit = nn.MSELoss()
    losses = []
    for step in range(3):  # loop body of embedding_module_gen.py:148-153
        optim.zero_grad()
        y = model(prod_ids)
        loss = crit(y, target)
        loss.backward()
        optim.step()
        losses.append(loss.item())
    np.savez_compressed(OUT / "kshift_adagrad_train.npz", ids=prod_ids.numpy(), target=target.numpy(),
                        weight0=w0, weight3=model.emb.weight.detach().numpy(),
                        state_sum3=optim.state[model.emb.weight]["sum"].numpy(),
                        losses=np.array(losses, dtype=np.float64), k=k, lr=0.5, **meta)

    # ---- 5. FlatEmbedding fwd + bwd + Adagrad with an upstream gradient -------------
    torch.manual_seed(1237)
    fm = cl.FlatEmbedding(500, 32)
    w0 = fm._emb_table.weight.detach().numpy().copy()
    tids = seeded_ids(16 * 50, 10).view(16, 50)
    gout = torch.randn(16, 50, 32, generator=torch.Generator().manual_seed(4321))
    optim = torch.optim.Adagrad(fm.parameters(), lr=5e-1)
    for step in range(2):
        optim.zero_grad()
        fm(tids).backward(gout)
        optim.step()
    np.savez_compressed(OUT / "flat_adagrad_train.npz", ids=tids.numpy(), grad_out=gout.numpy(),
                        weight0=w0, weight2=fm._emb_table.weight.detach().numpy(),
                        state_sum2=optim.state[fm._emb_table.weight]["sum"].numpy(), lr=0.5, **meta)

    # ---- 6. QREmbedding (ctor repaired ONLY by running nn.Module.__init__ first) -----
    torch.manual_seed(1238)
    qr = cl.QREmbedding.__new__(cl.QREmbedding)
    nn.Module.__init__(qr)
    cl.QREmbedding.__init__(qr, 10007, 32, True)
    qids = torch.cat([torch.tensor(EDGE_IDS, dtype=torch.int64), seeded_ids(108, 12)])
    # second repair: forward passes `round_mode=` (commons/layers.py:117); torch spells the
    # keyword `rounding_mode`.  Translate the keyword for the duration of the call, nothing else.
    real_div = torch.div
    torch.div = lambda a, b, round_mode=None, **kw: real_div(a, b, rounding_mode=round_mode, **kw)
    try:
        out_n = qr(qids).detach().numpy()
        qr.normalize_output = False
        out_p = qr(qids).detach().numpy()
    finally:
        torch.div = real_div
    np.savez_compressed(OUT / "qr_embedding.npz", ids=qids.numpy(), num_embeddings=10007,
                        weight_q=qr.emb_q.weight.detach().numpy(),
                        weight_r=qr.emb_r.weight.detach().numpy(), out_norm=out_n, out_plain=out_p,
                        **meta)

    # ---- 7. CosineVectorEmbedding ---------------------------------------------------
    torch.manual_seed(1239)
    cv = tl.CosineVectorEmbedding(32, 64, n_proj=32, num_bins=12)
    x = torch.randn(3, 17, 32)
    out = cv(x)
    z = F.normalize(x, p=2.0, dim=-1) @ cv.projection_mat
    idxs = torch.bucketize(z, cv.grid).view(-1, 32) + cv.pos_offset.unsqueeze(0)
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    out.backward(go)
    np.savez_compressed(OUT / "cosine_vector_embedding.npz", x=x.numpy(),
                        projection_mat=cv.projection_mat.numpy(), grid=cv.grid.numpy(),
                        pos_offset=cv.pos_offset.numpy(), weight=cv.emb.weight.detach().numpy(),
                        idxs=idxs.numpy(), out=out.detach().numpy(), grad_out=go.numpy(),
                        grad_weight=cv.emb.weight.grad.numpy(), **meta)

    # ---- 8. id production: xxhash + pad_array -----------------------------------------
    seed = fu.hash_feature_name_to_int("product_id")
    strings = ["12345", "abc", "ABC", "", "product-42", "NA"]
    np.savez_compressed(OUT / "feature_utils.npz", seed=seed, strings=np.array(strings),
                        ids=np.array([fu.hash_string_to_long(s, seed, False) for s in strings], dtype=np.int64),
                        ids_lower=np.array([fu.hash_string_to_long(s, seed, True) for s in strings], dtype=np.int64),
                        pad_short=fu.pad_array([5, 6, 7], 5), pad_long=fu.pad_array(list(range(10)), 4),
                        **meta)
    # ---- 11. train_mask_model loop body (embedding_module_gen.py:70-118): KShift(D = 4, k = 16) + MLP + BCE ----
    torch.manual_seed(1240)
    n_prod, mdim, k = 2048, 4, 16
    prod_ids = seeded_ids(n_prod, 14)
    mask_model = nn.Sequential(cl.KShiftEmbedding(int(1.15 * n_prod), mdim, num_shifts=k, normalize_output=False),
                               cl.MLP(mdim, 1, [mdim * 16]))
    sd0 = {k_: v.detach().numpy().copy() for k_, v in mask_model.state_dict().items()}
    optim = torch.optim.Adagrad(mask_model.parameters(), lr=5e-1)
    crit = nn.BCEWithLogitsLoss()
    gneg = torch.Generator().manual_seed(15)
    losses, negs = [], []
    for step in range(3):  # loop body of embedding_module_gen.py:104-115
        pos = prod_ids[torch.randperm(n_prod, generator=gneg)[:1024]]
        neg = torch.randint(-2 ** 63, 2 ** 63 - 1, (pos.size(0),), dtype=torch.int64, generator=gneg)
        ids_this = torch.cat([pos, neg], dim=0)
        target_this = torch.cat([torch.ones_like(pos), torch.zeros_like(neg)], dim=0)
        prediction = mask_model(ids_this).squeeze(1)
        loss = crit(prediction, target_this.float())
        loss.backward()
        optim.step()
        optim.zero_grad()
        losses.append(loss.item())
        negs.append(ids_this.numpy().copy())
    np.savez_compressed(OUT / "mask_model_train.npz", ids=np.stack(negs), losses=np.array(losses, dtype=np.float64),
                        k=k, lr=0.5, **{f"sd0/{k_}": v for k_, v in sd0.items()},
                        **{f"sd3/{k_}": v.detach().numpy() for k_, v in mask_model.state_dict().items()}, **meta)

    # ---- 9. streaming logQ (commons/layers.py:189-237; train_step with the two evident repairs) ----
    sys.path.insert(0, str(OUT.parent))
    import harness_lthm as H  # tests/harness_lthm.py: the repaired LTHM-step harness
    L = H.reference_layers()
    lq = L.LogQ(num_buckets=257, hash_offsets=[0, 34144, 7465477], alpha=0.05, p_init=0.01)
    lq_ids = torch.cat([torch.tensor(EDGE_IDS, dtype=torch.int64), seeded_ids(400, 13) % 3000])
    fwd0 = lq(lq_ids).numpy().copy()
    steps_fwd = []
    for step in range(4):
        sub = lq_ids[torch.randperm(lq_ids.numel(), generator=torch.Generator().manual_seed(step))[:300]]
        lq.train_step(sub, step)
        steps_fwd.append(lq(lq_ids).numpy().copy())
    np.savez_compressed(OUT / "streaming_logq.npz", ids=lq_ids.numpy(), fwd0=fwd0, fwd_steps=np.stack(steps_fwd),
                        b=np.stack([m.b.numpy() for m in lq.models]), a=np.stack([m.a.numpy() for m in lq.models]),
                        offsets=np.array([0, 34144, 7465477], dtype=np.int64), num_buckets=257, alpha=0.05,
                        p_init=0.01, **meta)

    # ---- 10. the repaired-harness LTHM training step (BASELINE configs[0]), reference classes on CPU ----
    cfg = H.HarnessConfig()
    torch.manual_seed(cfg.seed)
    model = H.LTHMStep(cfg, L)
    model._model.product_emb_module.emb.weight.data.copy_(H.kshift_table(cfg))
    batch = H.make_batch(cfg)
    replaced = H.fix_margins(model, batch)
    margin = H.bucket_margin(model, batch)
    big = "_model.product_emb_module.emb.weight"
    sd0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items() if k != big}
    res = H.run_step(model, batch, steps=2)
    names = H.embedding_param_names(model)
    sd2 = model.state_dict()
    out = res["output"]
    fixture = {f"sd0/{k}": v for k, v in sd0.items()}
    fixture.update({f"grad/{n}": res["grads"][n].numpy() for n in names if n != big})
    fixture.update({f"sd2/{n}": sd2[n].numpy() for n in names if n != big})
    fixture.update({f"sd2/{k}": v.numpy() for k, v in sd2.items() if "_log_q_calc" in k})
    w0 = H.kshift_table(cfg)
    np.savez_compressed(
        OUT / "lthm_step.npz", product_ids=batch["product_ids"].numpy(), labels=batch["labels"].numpy(),
        timestamp=batch["timestamp"].numpy(), losses=np.array(res["losses"], dtype=np.float64),
        kshift_checksum=np.array([w0.double().sum().item(), (w0.double() ** 2).sum().item()]),
        kshift_grad_absmax=float(res["grads"][big].abs().max()),
        kshift_after=sd2[big][::97].numpy(),  # AdamW on a detached table is pure decay
        next_token_emb=out["next_token_emb"][::8].numpy(), next_token_rowsum=out["next_token_emb"].sum(-1).numpy(),
        current_token_emb=out["current_token_emb"][::8].numpy(),
        current_token_rowsum=out["current_token_emb"].sum(-1).numpy(),
        current_token_mask=out["current_token_mask"].numpy(), current_token_ids=out["current_token_ids"].numpy(),
        margin=margin, ids_replaced=replaced, **fixture, **meta)
    print(f"lthm_step: losses {res['losses']}, trim -> {out['current_token_ids'].shape}, margin {margin:.2e}, "
          f"{replaced} ids re-drawn")
    for f in sorted(OUT.glob("*.npz")):
        print(f"{f.name:36s} {f.stat().st_size / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()
