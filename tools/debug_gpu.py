"""Ad-hoc GPU diagnostics (development aid; not part of the product or the tests)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import femol_b200 as F
import femol_oracle as o, cases


def big_pcg(nx, ny):
    print('==== big pcg', nx, ny)
    Lx, Ly = float(nx), float(ny)
    mesh = F.mesh.rectangle_Q4(Lx, Ly, nx, ny)
    p = F.FEM_Problem('displacement', 'plane', mesh)
    p.define_materials(F.materials.general_isotropic()); p.define_tensors(1)
    p.add_fixed_domain(F.domains.inside_box([0], [[0, Ly]]), ddls=[0, 1])
    p.add_forces([0, -1.0], F.domains.inside_box([Lx], [[Ly / 2 - 1, Ly / 2 + 1]]))
    x = torch.full((mesh.N_ele,), 0.4, dtype=torch.float64, device='cuda')
    p.assemble('K', X=x, p=3)
    d = p.device
    kv = p.K_values
    print('nnz', d.nnz, 'nnz_node', d.nnz_node, 'max_val', d.max_valence, 'max_row', d.max_row_nodes)
    print('vals finite', bool(torch.isfinite(kv).all()), 'min', float(kv.min()), 'max', float(kv.max()), 'sum', float(kv.sum()))
    print('F nonzeros', int((p._F_dev != 0).sum()), 'F sum', float(p._F_dev.sum()))
    indptr, indices = d.csr_pattern()
    print('indptr last', int(indptr[-1]), 'indices max', int(indices.max()), 'n', d.n)
    # reference product with torch sparse csr
    A = torch.sparse_csr_tensor(indptr.to(torch.int64), indices.to(torch.int64), kv, size=(d.n, d.n))
    v = torch.randn(d.n, dtype=torch.float64, device='cuda')
    y = torch.empty_like(v)
    F._lib.call('femol_spmv', d.plan, kv.data_ptr(), v.data_ptr(), y.data_ptr(), F._lib.stream())
    yr = A @ v
    print('spmv err', float((y - yr).abs().max() / yr.abs().max()), 'vAv', float(v @ yr))
    # diagonal via unit check: expected interior diagonal of 0.4^3 * Ke sum
    Ke = o.ke_batch('Q4', mesh.points[mesh.cells['quad'][:1]][:, :, :2], 2, o.iso_tensors(1, 0.3, 1))[0]
    print('expected interior diag', 0.4 ** 3 * (Ke[0, 0] + Ke[2, 2] + Ke[4, 4] + Ke[6, 6]))
    rows = torch.arange(d.n, device='cuda')
    diag = torch.zeros(d.n, dtype=torch.float64, device='cuda')
    rowidx = torch.repeat_interleave(rows, (indptr[1:] - indptr[:-1]).to(torch.int64))
    isd = rowidx == indices.to(torch.int64)
    diag[rowidx[isd]] = kv[isd]
    print('diag min', float(diag.min()), 'max', float(diag.max()), 'count', int(isd.sum()))
    for pc in ('jacobi', 'block_jacobi'):
        p.preconditioner = pc
        p._U_dev = None
        p.solver_max_iter = 300
        try:
            p.solve_device()
        except Exception as e:
            print(pc, 'EXC', e)
        print(pc, p.solver_info)


def lobpcg_ss():
    print('==== lobpcg ss')
    from femol_b200 import lobpcg as lp
    from numpy_ops import NumpyOps
    problem = cases.build_problem(F, cases.CASES['modal_q4_plate_ss_8x8'])
    problem.assemble('K'); problem.assemble('M')
    d = problem.device
    m, k = 8, 6
    X0 = np.random.default_rng(0).standard_normal((d.n, m))
    fixed = problem.fixed_dofs()
    X0[fixed] = 0
    for kind, nm in ((F._lib.PRECOND_JACOBI, 'jacobi'), (F._lib.PRECOND_BLOCK_JACOBI, 'block')):
        dops = lp.DeviceOps(d.plan, problem.K_values, problem.M_values, problem._fixed_mask(), m, kind, torch.from_numpy(X0).cuda())
        ld, rd, it = lp.lobpcg(dops, m, k, tol=1e-8, max_iter=2000)
        print(nm, it, np.sqrt(np.abs(ld[:6])) / (2 * np.pi), rd[:6])
    nops = NumpyOps(problem.K, problem.M, fixed, m, X0, precond='jacobi')
    ln, rn, it = lp.lobpcg(nops, m, k, tol=1e-8, max_iter=2000)
    print('numpy', it, np.sqrt(np.abs(ln[:6])) / (2 * np.pi), rn[:6])
    problem.n_modes = 6
    w, v = problem.solve(verbose=False)
    print('api', w, problem.solver_info)


if __name__ == '__main__':
    which = sys.argv[1:] or ['lobpcg', 'big']
    if 'lobpcg' in which:
        lobpcg_ss()
    if 'big' in which:
        big_pcg(800, 400)
        big_pcg(2000, 1000)


def big_simp(nx, ny, iters=3):
    print('==== big simp', nx, ny)
    sys.path.insert(0, ROOT)
    import bench
    problem, opt = bench.build_gpu_problem(F, nx, ny, 1e-8, 'block_jacobi')
    import time
    for it in range(iters):
        t0 = time.time()
        try:
            opt._U = opt.FEM_solver(None)
        except Exception as e:
            print('solve EXC', e, problem.solver_info)
            kv = problem.K_values
            print(' K finite', bool(torch.isfinite(kv).all()), float(kv.min()), float(kv.max()))
            x = opt.x_device
            print(' x finite', bool(torch.isfinite(x).all()), float(x.min()), float(x.max()), float(x.sum()))
            return
        torch.cuda.synchronize(); t1 = time.time()
        U = opt._U
        print(it, 'solve', problem.solver_info, 'U finite', bool(torch.isfinite(U).all()), 'Umax', float(U.abs().max()), 'time', t1 - t0)
        c, dc = opt.objective_function(None)
        print('   c', float(c), 'dc finite', bool(torch.isfinite(dc).all()), float(dc.min()), float(dc.max()))
        opt._filter_sensibility(None)
        print('   dcf finite', bool(torch.isfinite(opt._dc).all()), float(opt._dc.min()), float(opt._dc.max()))
        opt._get_new_x(None)
        opt._x, opt._xn = opt._xn, opt._x
        x = opt.x_device
        print('   x finite', bool(torch.isfinite(x).all()), float(x.min()), float(x.max()), 'mean', float(x.mean()), 'scal', opt._scal.cpu().numpy())
        torch.cuda.synchronize(); print('   iter time', time.time() - t0)


if __name__ == '__main__' and 'simp' in sys.argv[1:]:
    big_simp(2000, 1000)
