"""Time the fine-level products at the cfg-2 size: CSR SpMV vs the matrix-free stencil (three epilogue modes).

    python tools/time_apply.py [nx ny] [plane|plate]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import femol_b200 as F  # noqa: E402
import bench  # noqa: E402


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    model = sys.argv[3] if len(sys.argv) > 3 else 'plane'
    L = F._lib.require_gpu()
    if model == 'plane':
        problem, opt = bench.build_gpu_problem(F, nx, ny, 1e-8, 'auto')
        x = opt.x_device
    else:
        mesh = F.mesh.rectangle_Q4(float(nx), float(ny), nx, ny)
        problem = F.FEM_Problem('displacement', 'plate', mesh)
        problem.define_materials(F.materials.IsotropicMaterial(10920, 0.3, 1)); problem.define_tensors(0.5)
        problem.add_fixed_domain(F.domains.inside_box([0], [[0, float(ny)]]))
        x = torch.rand(mesh.N_ele, dtype=torch.float64, device='cuda') * 0.9 + 0.1
    problem.assemble('K', X=x, p=3)
    d = problem.device
    kv = problem.K_values
    st = F._lib.stream()
    nd = d.ndof
    xv = torch.randn(d.n, dtype=torch.float64, device='cuda'); yv = torch.empty_like(xv); rv = torch.randn_like(xv)
    dinv = torch.empty(d.n_nodes * nd * nd, dtype=torch.float64, device='cuda')
    F._lib.call('femol_precond_setup', d.plan, kv.data_ptr(), 2, dinv.data_ptr(), st)
    t = bench.time_kernel(torch, lambda: L.femol_spmv(d.plan, kv.data_ptr(), xv.data_ptr(), yv.data_ptr(), st), 20)
    print('%s %dx%d ndof %d: CSR femol_spmv %.1f us' % (model, nx, ny, d.n, t * 1e6))
    for m in range(3):
        t = bench.time_kernel(torch, lambda: L.femol_apply_epilogue(d.plan, kv.data_ptr(), xv.data_ptr(), yv.data_ptr(), m,
                                                                    rv.data_ptr(), dinv.data_ptr(), 0.7, st), 50)
        print('   matrix-free mode %d: %.1f us' % (m, t * 1e6))


if __name__ == '__main__':
    main()
