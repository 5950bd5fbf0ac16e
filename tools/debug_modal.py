import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import femol_b200 as F
from femol_b200 import lobpcg as lp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
inner = int(sys.argv[2]) if len(sys.argv) > 2 else 600
mesh = F.mesh.rectangle_Q4(10., 10., n, n)
p = F.FEM_Problem('modal', 'plate', mesh)
p.define_materials(F.materials.IsotropicMaterial(200e9, 0.3, 8000)); p.define_tensors(0.05)
p.add_fixed_domain(F.domains.inside_box([[0, 10.]], [[0, 10.]]), ddls=[0, 1, 5])
p.add_fixed_domain(F.domains.inside_box([0, 10.], [[0, 10.]]), ddls=[2, 4])
p.add_fixed_domain(F.domains.inside_box([[0, 10.]], [0, 10.]), ddls=[2, 3])
p.assemble('K'); p.assemble('M')
d = p.device
print('K finite', bool(torch.isfinite(p.K_values).all()), 'M finite', bool(torch.isfinite(p.M_values).all()), 'nnz', d.nnz, flush=True)
mask = p._fixed_mask() | torch.full((d.n_nodes,), 32, dtype=torch.uint8, device='cuda')
m = 8
g = torch.Generator(device='cuda'); g.manual_seed(1234)
X0 = torch.randn((d.n, m), dtype=torch.float64, device='cuda', generator=g)
bits = (mask.to(torch.int32)[:, None] >> torch.arange(6, device='cuda', dtype=torch.int32)[None]) & 1
X0[bits.reshape(-1).bool()] = 0
ops = lp.DeviceOps(d.plan, p.K_values, p.M_values, mask, m, 2, X0, inner_iters=inner, inner_tol=1e-3)
print('dinv finite', bool(torch.isfinite(ops.dinv).all()), flush=True)
orig_gram = ops.gram
def gram(a, b, ca, qa, cb, qb):
    G = orig_gram(a, b, ca, qa, cb, qb)
    if not np.isfinite(G).all():
        print('NON-FINITE gram', a, b, ca, qa, cb, qb, flush=True)
        for name in ('S', 'KS', 'MS'):
            t = ops.buf[ops.cur][name]
            print('  ', name, [bool(torch.isfinite(t[:, c]).all()) for c in range(3 * m)], flush=True)
        raise SystemExit(1)
    return G
ops.gram = gram
orig_prec = ops.precondition
def prec():
    orig_prec()
    W = ops.buf[ops.cur]['S'][:, m:2 * m]
    print('  precond: W finite', bool(torch.isfinite(W).all()), 'max', float(W.abs().max()), 'bp scalars',
          ops._bp_ws[-6 * 16:].reshape(6, 16)[:, :m].cpu().numpy().round(3).tolist() if ops._bp_ws is not None else None, flush=True)
ops.precondition = prec
t0 = time.time()
lam, rel, its = lp.lobpcg(ops, m, 6, tol=1e-8, max_iter=40, verbose=True)
print('done', its, time.time() - t0, np.sqrt(lam[:6]) / 2 / np.pi, rel)
