"""Per-kernel device-time breakdown of SIMP iterations at the cfg-2 size, from CUPTI activity records
(torch.profiler; sees every kernel of the process, including the ones replayed from the CUDA graph).

    python tools/kernel_breakdown.py [nx ny] [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
import femol_b200 as F  # noqa: E402
import bench  # noqa: E402


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    problem, opt = bench.build_gpu_problem(F, nx, ny, 1e-8, 'auto')
    for _ in range(3):
        opt.iterate()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(steps):
            opt.iterate()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, 'device_time_total', None)
        if t is None:
            t = getattr(e, 'cuda_time_total', 0)
        if t > 0 and e.device_type.name != 'CPU':
            rows.append((t, e.count, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print('total device time %.3f ms over %d steps (%.3f ms/step); pcg %s' % (tot / 1e3, steps, tot / 1e3 / steps, problem.solver_info))
    for t, c, k in rows[:40]:
        print('%9.1f us/step  %6.1f%%  %6.1f launches/step  %8.1f us/launch  %s' % (t / steps, 100 * t / tot, c / steps, t / c, k[:150]))


if __name__ == '__main__':
    main()
