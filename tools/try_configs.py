"""Exploratory timings of the plate / modal configs at reduced sizes (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import femol_b200 as F

def plate_static(n, precond):
    mesh = F.mesh.rectangle_Q4(1.0, 1.0, n, n)
    p = F.FEM_Problem('displacement', 'plate', mesh)
    p.define_materials(F.materials.isotropic_bending_benchmark()); p.define_tensors(0.1)
    p.add_fixed_domain(F.domains.inside_box([[0, 1.]], [[0, 1.]]), ddls=[0, 1, 5])
    p.add_fixed_domain(F.domains.inside_box([0, 1.], [[0, 1.]]), ddls=[2, 4])
    p.add_fixed_domain(F.domains.inside_box([[0, 1.]], [0, 1.]), ddls=[2, 3])
    p.add_forces([0, 0, -1.0, 0, 0, 0], F.domains.inside_box([[0.01, 0.99]], [[0.01, 0.99]]))
    p.preconditioner = precond
    p.solver_rtol = 1e-8
    p.solver_max_iter = 20000
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p.assemble('K')
    torch.cuda.synchronize(); t1 = time.perf_counter()
    U = p.solve_device()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    w = float(U[2::6].abs().max())
    print('plate %dx%d %s: assemble %.3fs solve %.3fs %s wmax=%.6e (thin-plate 0.00406*q*a^4/D=%.6e)' % (
        n, n, precond, t1 - t0, t2 - t1, p.solver_info, w, 0.00406 / (10920 * 0.1 ** 3 / (12 * (1 - 0.09)))), flush=True)

def modal(n, k=6, inner=200):
    mesh = F.mesh.rectangle_Q4(10., 10., n, n)
    p = F.FEM_Problem('modal', 'plate', mesh)
    p.define_materials(F.materials.IsotropicMaterial(200e9, 0.3, 8000)); p.define_tensors(0.05)
    p.add_fixed_domain(F.domains.inside_box([[0, 10.]], [[0, 10.]]), ddls=[0, 1, 5])
    p.add_fixed_domain(F.domains.inside_box([0, 10.], [[0, 10.]]), ddls=[2, 4])
    p.add_fixed_domain(F.domains.inside_box([[0, 10.]], [0, 10.]), ddls=[2, 3])
    p.n_modes = k; p.modal_inner_iters = inner; p.modal_tol = 1e-8; p.modal_max_iter = 60
    torch.cuda.synchronize(); t0 = time.perf_counter()
    w, v = p.solve(verbose=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print('modal %dx%d inner=%d: %.2fs its=%s relres=%s f=%s (NAFEMS 2.377 5.942 5.942 9.507 11.884 11.884)' % (
        n, n, inner, t1 - t0, p.solver_info['iterations'], np.array2string(p.solver_info['relres'], precision=1), np.array2string(w, precision=4)), flush=True)

if __name__ == '__main__':
    for n in (100, 300):
        for pc in ('gmg', 'block_jacobi'):
            try:
                plate_static(n, pc)
            except Exception as e:
                print('plate', n, pc, 'EXC', e, flush=True)
    for n, inner in ((50, 200), (100, 200), (100, 600), (200, 600)):
        try:
            modal(n, inner=inner)
        except Exception as e:
            print('modal', n, 'EXC', e, flush=True)
