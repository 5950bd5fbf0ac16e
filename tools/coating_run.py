"""SIMP_COMP with a coating material on the cfg-2 mesh (two element matrices: K = K_base + x^p K_coat)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import femol_b200 as F  # noqa: E402

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2000, 1000)
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
mesh = F.mesh.rectangle_Q4(float(nx), float(ny), nx, ny)
p = F.FEM_Problem('displacement', 'plane', mesh)
p.define_materials(F.materials.general_isotropic(), F.materials.IsotropicMaterial(5, 0.25, 2)); p.define_tensors(1, 0.5)
p.add_fixed_domain(F.domains.inside_box([0], [[0, float(ny)]]), ddls=[0, 1])
p.add_forces([0, -1.0], F.domains.inside_box([float(nx)], [[ny / 2 - 1, ny / 2 + 1]]))
p.solver_rtol = 1e-8
opt = F.SIMP_COMP(p, volfrac=0.4, penal=3, rmin=1.5)
opt.archive = False
its, ms = [], []
for k in range(steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    change, c = opt.iterate()
    torch.cuda.synchronize(); ms.append((time.perf_counter() - t0) * 1e3)
    its.append(p.solver_info['iterations'])
print('coating %dx%d: CG iterations %s' % (nx, ny, its))
print('ms/step (after 3 warm-up steps) %.2f -> %.1f it/s; relres %.2e, c %.4f' % (sum(ms[3:]) / len(ms[3:]), 1e3 * len(ms[3:]) / sum(ms[3:]), p.solver_info['relres'], c))
