import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import recommendations_b200 as R
from oracle import embedding_oracle as O
from tolerances import *
DEV = "cuda:0"
b, l, n_rows, d_in, k = 256, 50, 10_000, 32, 8
g = torch.Generator().manual_seed(7)
ids = torch.randint(-2 ** 63, 2 ** 63 - 1, (b, l), generator=g, dtype=torch.int64)
valid = torch.randint(1, l + 1, (b,), generator=g)
ids[torch.arange(l).unsqueeze(0) >= valid.unsqueeze(1)] = 0
torch.manual_seed(1234)
w = torch.randn(n_rows, d_in)
for _ in range(6):
    torch.randn(1)
ks2 = R.KShiftEmbedding(n_rows, d_in, num_shifts=k, normalize_output=True, device=DEV,
                        fused_optimizer=R.FusedOptimizerConfig(kind="adagrad", lr=0.5))
ks2.load_state_dict({"emb.weight": w})
target = torch.nn.functional.normalize(torch.randn(b, l, d_in, generator=g), dim=-1)
loss = torch.nn.functional.mse_loss(ks2(ids.to(DEV)), target.to(DEV)); loss.backward()
wt = torch.nn.Parameter(w.to(DEV)); opt_t = torch.optim.Adagrad([wt], lr=0.5)
torch.nn.functional.mse_loss(O.kshift_embedding(wt, ids.to(DEV), k, normalize=True), target.to(DEV)).backward()
Gt = wt.grad.detach().cpu().double().clone()
opt_t.step()
# fp64 replay
w64 = w.double(); rows = [O.row_index(ids, n_rows, c) for c in range(k)]
x = sum(w64[r] for r in rows).requires_grad_(True)
lo = torch.nn.functional.mse_loss(torch.nn.functional.normalize(x, p=2.0, dim=-1), target.double())
(dx,) = torch.autograd.grad(lo, x)
G = torch.zeros_like(w64); A = torch.zeros_like(w64); Cn = torch.zeros(n_rows)
for r in rows:
    G.index_add_(0, r.reshape(-1), dx.reshape(-1, 32)); A.index_add_(0, r.reshape(-1), dx.reshape(-1, 32).abs()); Cn.index_add_(0, r.reshape(-1), torch.ones(r.numel()))
ours = ks2.emb.weight.cpu().double(); twin = wt.detach().cpu().double()
w64n = w64 - 0.5 * G / (G.abs() + 1e-10)
budget = torch.minimum(0.5 * 8 * EPS32 * A / (G.abs() + 1e-10), torch.full_like(A, 1.0))
err = (ours - twin).abs(); base = 1e-5 + 1e-5 * twin.abs()
ratio = err / (base + 4 * budget)
idx = torch.nonzero(ratio > 1.0)
print("violations", idx.shape[0])
for (r, c) in idx[:16].tolist():
    print(f"row {r} col {c} terms {int(Cn[r])} G64 {G[r,c]:.3e} Gtwin {Gt[r,c]:.3e} A {A[r,c]:.3e} budget {budget[r,c]:.3e} ours-w {ours[r,c]-w64[r,c]:.6f} twin-w {twin[r,c]-w64[r,c]:.6f} fp64-w {w64n[r,c]-w64[r,c]:.6f}")
