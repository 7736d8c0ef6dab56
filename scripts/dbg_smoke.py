import sys, torch
sys.path.insert(0, '.')
import recommendations_b200 as R
from oracle import embedding_oracle as O
dev='cuda:0'
g = torch.Generator().manual_seed(7)
ids = torch.randint(-2 ** 63, 2 ** 63 - 1, (64, 50), generator=g, dtype=torch.int64)
ids[:, 40:] = 0
grad = torch.randn(64, 50, 64, generator=g)
torch.manual_seed(1234)
w0 = torch.randn(5000, 64)
w, state = w0.clone(), torch.zeros_like(w0)
rows = O.row_index(ids, 5000, 0)
dg = O.dense_grad(rows, grad, 5000)
O.adagrad_step(w, dg, state, lr=0.5)
cnt = torch.bincount(rows.view(-1), minlength=5000)
nbad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    m = R.FlatEmbedding(5000, 64, device=dev, fused_optimizer=R.FusedOptimizerConfig(kind="adagrad", lr=0.5))
    m.load_state_dict({"_emb_table.weight": w0})
    out = m(ids.to(dev))
    if it % 2:
        _ = out.cpu()
    out.backward(grad.to(dev))
    got = m._emb_table.weight.cpu()
    bad = ((got - w).abs() > 1e-6 + 1e-5 * w.abs()).nonzero()
    if bad.shape[0]:
        nbad += 1
        print('iter', it, 'bad', bad.shape[0], 'rows', sorted(set(bad[:, 0].tolist()))[:10], 'cols', sorted(set(bad[:, 1].tolist()))[:16])
        for r, c in bad[:4].tolist():
            print('   ', r, c, 'count', int(cnt[r]), 'w0', float(w0[r, c]), 'g', float(dg[r, c]), 'want', float(w[r, c]), 'got', float(got[r, c]),
                  'state', float(m._emb_table._buffers['opt_state1'][r, c]), float(state[r, c]))
print('iterations with mismatches:', nbad)
