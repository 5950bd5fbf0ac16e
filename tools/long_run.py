"""CG iterations and time per SIMP iteration over a long optimisation run on cfg 2 (the bench times steps 4-8 only)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import femol_b200 as F  # noqa: E402
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
problem, opt = bench.build_gpu_problem(F, 2000, 1000, 1e-8, 'auto')
problem.gmg_smoothing_steps = int(os.environ.get('NU', '1'))
problem.gmg_omega = float(os.environ.get('OMEGA', '0.7'))
its, ms = [], []
for k in range(steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    change, c = opt.iterate()
    torch.cuda.synchronize(); ms.append((time.perf_counter() - t0) * 1e3)
    its.append(problem.solver_info['iterations'])
    if k % 10 == 9:
        print('steps %3d-%3d: CG iterations %s, ms/step %.1f, change %.3f, c %.2f' % (k - 8, k + 1, its[-10:], sum(ms[-10:]) / 10, change, c), flush=True)
print('total %.2f s for %d steps: %.1f it/s' % (sum(ms) / 1e3, steps, steps / (sum(ms) / 1e3)))
