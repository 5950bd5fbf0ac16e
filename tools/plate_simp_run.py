"""SIMP_COMP on a Reissner-Mindlin plate (6 dofs per node): CG iterations per step with the multigrid preconditioner."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import femol_b200 as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mesh = F.mesh.rectangle_Q4(1.0, 1.0, n, n)
p = F.FEM_Problem('displacement', 'plate', mesh)
p.define_materials(F.materials.isotropic_bending_benchmark()); p.define_tensors(float(os.environ.get('T', '0.1')))
p.add_fixed_domain(F.domains.inside_box([0], [[0, 1.]]))
p.add_forces([0, 0, -1.0, 0, 0, 0], F.domains.inside_box([1.], [[0.45, 0.55]]))
p.solver_rtol = float(os.environ.get('RTOL', '1e-8'))
opt = F.SIMP_COMP(p, volfrac=0.4, penal=3, rmin=1.5)
opt.archive = False
its, ms = [], []
for k in range(steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    change, c = opt.iterate()
    torch.cuda.synchronize(); ms.append((time.perf_counter() - t0) * 1e3)
    its.append(p.solver_info['iterations'])
    print('step %d: %d CG iterations, %.1f ms, relres %.2e, status %d, c %.5g' % (k + 1, its[-1], ms[-1], p.solver_info['relres'], p.solver_info['status'], c), flush=True)
