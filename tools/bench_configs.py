"""Secondary measurements for the other BASELINE.json configs (not the headline bench line):

  cfg 3  Reissner-Mindlin Q4 plate static solve, 1000 x 1000 (6.0 M dof), assemble + PCG
  cfg 4  SIMP_VIBE on a Mindlin plate, 500 x 500 (1.5 M dof), lowest 6 modes by LOBPCG, a few iterations
  cfg 5  T6 plane-stress static solve (unstructured path, per-element matrices), ~1 M elements

    python tools/bench_configs.py [cfg3] [cfg4] [cfg5] [--scale 0.5]

One JSON line per config (written by the caller into profiles/).  Sizes can be scaled down with --scale.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import femol_b200 as F  # noqa: E402


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


def cfg3(scale):
    n = int(1000 * scale)
    mesh = F.mesh.rectangle_Q4(1.0, 1.0, n, n)
    p = F.FEM_Problem('displacement', 'plate', mesh)
    p.define_materials(F.materials.isotropic_bending_benchmark()); p.define_tensors(0.1)
    p.add_fixed_domain(F.domains.inside_box([[0, 1.]], [[0, 1.]]), ddls=[0, 1, 5])
    p.add_fixed_domain(F.domains.inside_box([0, 1.], [[0, 1.]]), ddls=[2, 4])
    p.add_fixed_domain(F.domains.inside_box([[0, 1.]], [0, 1.]), ddls=[2, 3])
    p.add_forces([0, 0, -1.0, 0, 0, 0], F.domains.inside_box([[0.001, 0.999]], [[0.001, 0.999]]))
    p.solver_rtol = 1e-8
    out = {'config': 'cfg3 plate static %dx%d' % (n, n), 'ndof': 6 * (n + 1) ** 2}
    for pc in ('gmg', 'block_jacobi'):
        p.preconditioner = pc
        p._U_dev = None
        t0 = sync(); p.assemble('K'); t1 = sync(); U = p.solve_device(); t2 = sync()
        out[pc] = {'assemble_s': t1 - t0, 'solve_s': t2 - t1, 'pcg': dict(p.solver_info),
                   'w_max': float(U[2::6].abs().max())}
        # again with the plans / hierarchy in place (what every later solve of the same mesh costs): re-assembly,
        # hierarchy refresh, solve from a zero start
        p._U_dev = None
        t0 = sync(); p.assemble('K'); t1 = sync(); U = p.solve_device(); t2 = sync()
        out[pc]['second_assemble_s'] = t1 - t0
        out[pc]['second_solve_s'] = t2 - t1
    out['nnz'] = p.device.nnz
    return out


def cfg4(scale, iters=3):
    n = int(500 * scale)
    mesh = F.mesh.rectangle_Q4(10., 10., n, n)
    p = F.FEM_Problem('modal', 'plate', mesh)
    p.define_materials(F.materials.IsotropicMaterial(200e9, 0.3, 8000)); p.define_tensors(0.05)
    p.add_fixed_domain(F.domains.inside_box([[0, 10.]], [[0, 10.]]), ddls=[0, 1, 5])
    p.add_fixed_domain(F.domains.inside_box([0, 10.], [[0, 10.]]), ddls=[2, 4])
    p.add_fixed_domain(F.domains.inside_box([[0, 10.]], [0, 10.]), ddls=[2, 3])
    p.n_modes = 6
    p.modal_inner_iters = 600
    p.modal_gmg_iters = int(os.environ.get('MODAL_GMG_ITERS', p.modal_gmg_iters))
    p.modal_inner_tol = float(os.environ.get('MODAL_INNER_TOL', p.modal_inner_tol))
    p.modal_tol = 1e-8
    t0 = sync(); w, v = p.solve(verbose=False); t1 = sync()
    out = {'config': 'cfg4 SIMP_VIBE plate %dx%d' % (n, n), 'ndof': 6 * (n + 1) ** 2, 'nnz': p.device.nnz,
           'solid_modes_hz': w.tolist(), 'nafems_fv13_hz': [2.377, 5.942, 5.942, 9.507, 11.884, 11.884],
           'modal_solve_s': t1 - t0, 'lobpcg': {'iterations': p.solver_info['iterations'],
                                                'max_relres': float(np.max(p.solver_info['relres']))}}
    opt = F.SIMP_VIBE(p, volfrac=0.5, p=3, q=1, rmin=1.5)
    opt.archive = False
    v_ref = torch.from_numpy(v[0]).cuda()
    ts, lam = [], []
    for _ in range(iters):
        t0 = sync(); opt.iterate(v_ref); t1 = sync()
        ts.append(t1 - t0); lam.append(float(opt.lmbd))
    out['vibe_iteration_s'] = ts
    out['tracked_hz'] = lam
    out['mac'] = opt.macs
    return out


def cfg5(scale):
    n = int(700 * scale)
    mesh = F.mesh.rectangle_T6(float(n), float(n), n, n)
    p = F.FEM_Problem('displacement', 'plane', mesh)
    p.define_materials(F.materials.general_isotropic()); p.define_tensors(1)
    p.add_fixed_domain(F.domains.inside_box([0], [[0, float(n)]]), ddls=[0, 1])
    p.add_forces([0, -1.0], F.domains.inside_box([float(n)], [[0, float(n)]]))
    p.solver_rtol = 1e-8
    x = {'triangle': np.random.default_rng(0).uniform(0.3, 1.0, mesh.N_ele)}
    t0 = sync(); p.assemble('K', X=x, p=3); t1 = sync(); p.assemble('K', X=x, p=3); t2 = sync()
    U = p.solve_device(); t3 = sync()
    return {'config': 'cfg5-like T6 plane static %dx%d (unstructured path)' % (n, n), 'n_elem': mesh.N_ele,
            'ndof': 2 * mesh.N_nodes, 'nnz': p.device.nnz, 'first_assemble_s (incl. plan + Ke batch)': t1 - t0,
            'reassemble_s': t2 - t1, 'solve_s': t3 - t2, 'pcg': dict(p.solver_info)}


def cfg5simp(scale, iters=5):
    """BASELINE cfg 5: T6 plane-stress SIMP_COMP on ~16 M dof (rectangle_T6 1414 x 1414), one GPU, lattice multigrid."""
    n = int(1414 * scale)
    t0 = time.perf_counter()
    mesh = F.mesh.rectangle_T6(float(n), float(n), n, n)
    p = F.FEM_Problem('displacement', 'plane', mesh)
    p.define_materials(F.materials.general_isotropic()); p.define_tensors(1)
    p.add_fixed_domain(F.domains.inside_box([0], [[0, float(n)]]), ddls=[0, 1])
    p.add_forces([0, -1.0], F.domains.inside_box([float(n)], [[n / 2 - 1, n / 2 + 1]]))
    p.solver_rtol = 1e-8
    opt = F.SIMP_COMP(p, volfrac=0.4, penal=3, rmin=1.5)
    opt.archive = False
    t_mesh = time.perf_counter() - t0
    ts, its, cs = [], [], []
    for k in range(iters):
        t0 = sync(); opt.iterate(); t1 = sync()
        ts.append(t1 - t0); its.append(p.solver_info['iterations']); cs.append(opt.c)
    return {'config': 'cfg5 T6 plane SIMP_COMP %dx%d (lattice multigrid, one GPU)' % (n, n), 'n_elem': mesh.N_ele,
            'ndof': 2 * mesh.N_nodes, 'nnz': p.device.nnz, 'host_mesh_and_setup_s': t_mesh,
            'iteration_s (first includes plan, element matrices, filter lists, hierarchy)': ts,
            'pcg_iterations': its, 'preconditioner': p.solver_info['preconditioner'], 'compliance': cs,
            'relres': p.solver_info['relres'], 'gpu_mem_gb': torch.cuda.max_memory_allocated() / 1e9}


if __name__ == '__main__':
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    scale = 1.0
    if '--scale' in sys.argv:
        scale = float(sys.argv[sys.argv.index('--scale') + 1])
        args = [a for a in args if a != str(scale) and a != sys.argv[sys.argv.index('--scale') + 1]]
    which = args or ['cfg3', 'cfg4', 'cfg5']
    for name in which:
        try:
            res = {'cfg3': cfg3, 'cfg4': cfg4, 'cfg5': cfg5, 'cfg5simp': cfg5simp}[name](scale)
        except Exception as e:  # keep going: these are exploratory
            res = {'config': name, 'error': repr(e)}
        print(json.dumps(res), flush=True)
        torch.cuda.empty_cache()
