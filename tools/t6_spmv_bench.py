"""T6 lattice mesh on one GPU: time the level-0 CSR product (FP64), one V-cycle and a SIMP iteration.

    [FEMOL_SPMV_WIDE=0|1] [FEMOL_GMG_F32_FINE=0|1] python tools/t6_spmv_bench.py [n]
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import femol_b200 as F  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, 'tools'))
from dist_simp_check import build_problem  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3   # us


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 707
    L = F._lib
    gp = build_problem(F, 't6', n, n)
    gp.solver_rtol = 1e-8
    opt = F.SIMP_COMP(gp, volfrac=0.4, penal=3, rmin=1.5)
    opt.archive = False
    for _ in range(3):
        opt.iterate()
    d = gp.device
    x = torch.randn(d.n, dtype=torch.float64, device=d.dev)
    y = torch.empty_like(x)
    t_spmv = timed(lambda: L.call('femol_spmv', d.plan, L.ptr(gp._Kv), L.ptr(x), L.ptr(y), L.stream()))
    g = gp._gmg_ready()
    t_vc = timed(lambda: L.call('femol_gmg_apply', g, L.ptr(x), L.ptr(y), L.stream()))
    t_it = timed(lambda: opt.iterate(), reps=4)
    nnz = d.nnz
    out = {'n': n, 'ndof': d.n, 'nnz': nnz, 'wide': os.environ.get('FEMOL_SPMV_WIDE'), 'f32_fine': os.environ.get('FEMOL_GMG_F32_FINE'),
           'spmv_f64_us': t_spmv, 'spmv_f64_GBps': (8.0 * nnz + 4.0 * nnz / 4 + 16.0 * d.n) / t_spmv / 1e3,
           'vcycle_us': t_vc, 'simp_iteration_ms': t_it / 1e3, 'pcg': gp.solver_info['iterations']}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
