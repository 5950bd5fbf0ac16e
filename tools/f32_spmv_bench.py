"""Single-precision operator product (the multigrid levels' kernel) against the HBM roofline.

    python tools/f32_spmv_bench.py            (T6 707 x 707, Q4 1000 x 500 = level 1 of config 2, Q4 2000 x 1000)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import torch  # noqa: E402
import femol_b200 as F  # noqa: E402
from dist_simp_check import build_problem  # noqa: E402
from t6_spmv_bench import timed  # noqa: E402


def main():
    L = F._lib
    peak = 6551.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    for kind, nx, ny in (('t6', 707, 707), ('q4', 1000, 500), ('q4', 2000, 1000)):
        gp = build_problem(F, kind, nx, ny)
        gp.assemble('K')
        d = gp.device
        v32 = torch.empty(d.nnz, dtype=torch.float32, device='cuda')
        L.call('femol_to_f32', d.plan, gp.K_values.data_ptr(), v32.data_ptr(), L.stream())
        x = torch.randn(d.n, dtype=torch.float64, device='cuda')
        r = torch.randn(d.n, dtype=torch.float64, device='cuda')
        y = torch.empty_like(x)
        dinv = torch.empty(d.n_nodes * 4, dtype=torch.float64, device='cuda')
        L.call('femol_precond_setup', d.plan, gp.K_values.data_ptr(), 2, dinv.data_ptr(), L.stream())
        for mode in (0, 1, 2):
            # algorithmic bytes: 16 B block + 4 B column per node pair, row pointers, x read once, y written (+ rhs, Dinv, x again)
            nbytes = 20 * d.nnz_node + 4 * d.n_nodes + 16 * d.n + (8 * d.n if mode >= 1 else 0) + (24 * d.n if mode == 2 else 0)
            row = {'mesh': '%s %dx%d' % (kind, nx, ny), 'nodes': d.n_nodes, 'node_blocks': d.nnz_node, 'mode': mode, 'MB': nbytes / 1e6}
            t = timed(lambda: L.call('femol_apply_f32', d.plan, v32.data_ptr(), x.data_ptr(), y.data_ptr(), mode, r.data_ptr(),
                                     dinv.data_ptr(), 0.7, L.stream()), reps=30)
            row['us'] = round(t, 2)
            row['frac_of_hbm_peak'] = round(nbytes / t / 1e3 / peak, 3)
            print(json.dumps(row))
        del gp, v32


if __name__ == '__main__':
    main()
