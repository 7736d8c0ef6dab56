"""Generates tests/golden/*.npz by RUNNING THE REFERENCE (imported, unmodified, from
/root/reference) on seeded inputs.  Only runs in the build container; the GPU box has no
/root/reference, which is why the outputs are committed.

    python oracle/make_golden.py            # rewrites tests/golden/

Every fixture stores the inputs, the initial weights and the reference's outputs, so both
the oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_golden.py) can
be checked against what the reference itself computed.  torch version is recorded.
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
    for f in sorted(OUT.glob("*.npz")):
        print(f"{f.name:36s} {f.stat().st_size / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()
