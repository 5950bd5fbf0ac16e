import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, ctypes
import femol_b200 as F, cases
for n in (8, 10, 16):
    for hg in (False, True):
        mesh = F.mesh.rectangle_Q4(10., 10., n, n)
        p = F.FEM_Problem('displacement', 'plate', mesh)
        p.define_materials(F.materials.IsotropicMaterial(200e9, 0.3, 8000)); p.define_tensors(0.05)
        p.add_fixed_domain(F.domains.inside_box([[0, 10.]], [[0, 10.]]), ddls=[0, 1])
        p.add_fixed_domain(F.domains.inside_box([0], [[0, 10.]]))
        p.add_forces([0, 0, -1.0, 0, 0, 0], F.domains.inside_box([10.], [[0, 10.]]))
        p.gmg_hourglass = hg
        p.preconditioner = 'gmg'
        p.solver_rtol = 1e-8
        U = p.solve_device().cpu().numpy()
        K = p.K
        print('n=%d hourglass %s: %s true relres %.2e' % (n, hg, {k: p.solver_info[k] for k in ('iterations', 'status')}, np.linalg.norm(K @ U - p.F) / np.linalg.norm(p.F)), flush=True)
        # 8 iterations only, random right-hand side (what the LOBPCG preconditioner does)
        g = p._gmg_ready()
        d = p.device
        rng = np.random.default_rng(1)
        r = rng.standard_normal(d.n); r[p.fixed_dofs()] = 0; r[5::6] = 0
        rd = torch.from_numpy(r).cuda(); z = torch.zeros_like(rd)
        ws = torch.empty(int(F._lib.load().femol_pcg_gmg_workspace_doubles(d.plan)), dtype=torch.float64, device='cuda')
        it, rr = ctypes.c_int(0), ctypes.c_double(0.0)
        rc = F._lib.call('femol_pcg_gmg', d.plan, g, p.K_values.data_ptr(), rd.data_ptr(), z.data_ptr(), 1e-3, 8, ws.data_ptr(), ctypes.byref(it), ctypes.byref(rr), F._lib.stream())
        zz = z.cpu().numpy()
        print('     8-iteration solve: rc %d its %d reported relres %.2e true %.2e' % (rc, it.value, rr.value, np.linalg.norm(K @ zz - r) / np.linalg.norm(r)), flush=True)
