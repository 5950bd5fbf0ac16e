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
    # idle time between consecutive device activities (launch latency inside the graphs, host round trips)
    ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                 if e.device_type.name != 'CPU'), key=lambda t: t[0])
    if ev:
        span = ev[-1][1] - ev[0][0]
        gaps = [(max(0.0, b[0] - a[1]), b[2]) for a, b in zip(ev[:-1], ev[1:])]
        small = [g for g, _ in gaps if g <= 5.0]
        mid = [g for g, _ in gaps if 5.0 < g <= 50.0]
        big = [(g, n) for g, n in gaps if g > 50.0]
        print('span %.3f ms/step; idle %.3f ms/step: %d gaps <= 5 us (%.3f ms/step), %d gaps 5-50 us (%.3f ms/step), %d gaps > 50 us (%.3f ms/step)'
              % (span / 1e3 / steps, sum(g for g, _ in gaps) / 1e3 / steps, len(small) // steps, sum(small) / 1e3 / steps,
                 len(mid) // steps, sum(mid) / 1e3 / steps, len(big) // steps, sum(g for g, _ in big) / 1e3 / steps))
        for g, n in sorted(big, reverse=True)[:12]:
            print('   gap %8.1f us before %s' % (g, n[:90]))
    # per-launch durations of the multilevel kernels (one cluster of values per multigrid level)
    per = {}
    for a, b, name in ev:
        per.setdefault(name[:60], []).append(b - a)
    for name, ds in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        if len(ds) >= 4 * steps and any(k in name for k in ('restrict', 'prolong', 'spmv_node', 'jacobi')):
            ds = sorted(ds)
            q = [ds[int(f * (len(ds) - 1))] for f in (0.1, 0.3, 0.5, 0.7, 0.9)]
            print('   %-60s us at 10/30/50/70/90 %%: %s' % (name, ' '.join('%.1f' % v for v in q)))
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
