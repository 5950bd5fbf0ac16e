"""GMG parameter sweep on the cfg-2 problem (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import femol_b200 as F
import bench

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2000, 1000)
for nu in (1, 2, 3):
    for omega in (0.5, 0.6, 0.7, 0.8):
        problem, opt = bench.build_gpu_problem(F, nx, ny, 1e-8, 'gmg')
        problem.gmg_smoothing_steps, problem.gmg_omega = nu, omega
        its, ts = [], []
        try:
            for k in range(6):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                opt.iterate()
                torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
                its.append(problem.solver_info['iterations'])
            print('nu=%d omega=%.1f  pcg its %s  ms/step %s' % (nu, omega, its, ['%.1f' % (1e3 * t) for t in ts]), flush=True)
        except Exception as e:
            print('nu=%d omega=%.1f FAILED %s %s' % (nu, omega, e, problem.solver_info), flush=True)
        del problem, opt
        torch.cuda.empty_cache()
