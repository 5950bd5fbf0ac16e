"""Strip-distributed SIMP_COMP against the single-GPU SIMP_COMP on the same problem.

    python tools/dist_simp_check.py --mesh q4 --nx 96 --ny 48 --iters 3                      (one rank)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_simp_check.py --mesh t6 --nx 64 --ny 64 --iters 3

Prints one JSON line (rank 0) and exits non-zero when the distributed run departs from the single-GPU one."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def build_problem(F, kind, nx, ny, model='plane'):
    Lx, Ly = float(nx), float(ny)
    mesh = {'q4': F.mesh.rectangle_Q4, 't3': F.mesh.rectangle_T3, 't6': F.mesh.rectangle_T6}[kind](Lx, Ly, nx, ny)
    gp = F.FEM_Problem('displacement', model, mesh)
    if model == 'plane':
        gp.define_materials(F.materials.general_isotropic()); gp.define_tensors(1)
        gp.add_fixed_domain(F.domains.inside_box([0], [[0, Ly]]), ddls=[0, 1])
        gp.add_forces([0, -1.0], F.domains.inside_box([Lx], [[Ly / 2 - 1, Ly / 2 + 1]]))
    else:
        gp.define_materials(F.materials.IsotropicMaterial(10920, 0.3, 1)); gp.define_tensors(2.0)
        gp.add_fixed_domain(F.domains.inside_box([0], [[0, Ly]]))
        gp.add_forces([0, 0, -1e-2, 0, 0, 0], F.domains.inside_box([Lx], [[0, Ly]]))
    return gp


def kernel_profile(opt, steps, rank):
    from torch.profiler import profile, ProfilerActivity
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            opt.iterate()
        torch.cuda.synchronize()
    if rank != 0:
        return
    ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type.name != 'CPU'),
                key=lambda t: t[0])
    span = ev[-1][1] - ev[0][0]
    busy, cur_end = 0.0, ev[0][0]
    for a_, b_, _ in ev:
        if b_ > cur_end:
            busy += b_ - max(a_, cur_end)
            cur_end = b_
    print('PROFILE span %.3f ms/step, device busy %.3f ms/step, pcg %s' % (span / 1e3 / steps, busy / 1e3 / steps, opt.solver_info))
    gaps = sorted(((max(0.0, b[0] - a[1]), b[2]) for a, b in zip(ev[:-1], ev[1:])), reverse=True)
    for g, n in gaps[:10]:
        print('PROFILE    gap %8.1f us before %s' % (g, n[:100]))
    per = {}
    for a_, b_, name in ev:
        d = per.setdefault(name[:110], [0.0, 0])
        d[0] += b_ - a_; d[1] += 1
    tot = sum(v[0] for v in per.values())
    for name, (t, c) in sorted(per.items(), key=lambda kv: -kv[1][0])[:32]:
        print('PROFILE %9.1f us/step %5.1f%% %7.1f launches/step %8.1f us/launch  %s' % (t / steps, 100 * t / tot, c / steps, t / c, name))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mesh', default='q4')
    ap.add_argument('--model', default='plane')
    ap.add_argument('--nx', type=int, default=96)
    ap.add_argument('--ny', type=int, default=48)
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--ndist', type=int, default=None)
    ap.add_argument('--rtol', type=float, default=1e-9)
    ap.add_argument('--no-reference', action='store_true', help='skip the single-GPU run (large meshes)')
    ap.add_argument('--out', default=None)
    ap.add_argument('--latency', action='store_true', help='time back-to-back halo exchanges / scalar all-reduces (CUDA events)')
    ap.add_argument('--profile', type=int, default=0, help='profile this many extra iterations with CUPTI (rank 0 prints)')
    a = ap.parse_args()
    import femol_b200 as F
    from femol_b200 import dist as fd
    if os.environ.get('FEMOL_AB_LIB'):          # A/B runs against another build of the library
        F._lib.LIB_PATH = os.environ['FEMOL_AB_LIB']
    rank, world = fd.init_process_group()
    if world == 1:
        torch.cuda.set_device(0)
    t0 = time.perf_counter()
    gp = build_problem(F, a.mesh, a.nx, a.ny, a.model)
    gp.solver_rtol = a.rtol
    comm = fd.create_comm(F)
    opt = fd.DistributedSIMP_COMP(F, gp, comm, volfrac=0.4, penal=3, rmin=1.5, n_dist=a.ndist)
    t_setup = time.perf_counter() - t0
    times, log = [], []
    for it in range(a.iters):
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t1 = time.perf_counter()
        change, c = opt.iterate()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t1)
        log.append({'c': c, 'change': change, 'pcg': opt.solver_info['iterations'], 'relres': opt.solver_info['relres'],
                    'status': opt.solver_info['status']})
    X = opt.density()
    lat = None
    if a.latency:
        L = F._lib
        d = opt.local_problem.device
        v = torch.zeros(d.n, dtype=torch.float64, device=d.dev)
        sc = torch.zeros(16, dtype=torch.float64, device=d.dev)
        lat = {}
        for name, fn in (('halo_exchange_level0_us', lambda: L.call('femol_gmg_dist_exchange', opt._gmg, 0, L.ptr(v), gp.N_dof, L.stream())),
                         ('halo_exchange_level%d_us' % (opt.part.n_dist - 1),
                          lambda: L.call('femol_gmg_dist_exchange', opt._gmg, opt.part.n_dist - 1, L.ptr(v), gp.N_dof, L.stream())),
                         ('allreduce_2_doubles_us', lambda: L.call('femol_dist_allreduce_sum', comm, L.ptr(sc), 2, L.stream()))):
            for _ in range(20):
                fn()
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(500):
                fn()
            e1.record()
            torch.cuda.synchronize()
            lat[name] = e0.elapsed_time(e1) / 500 * 1e3
    if a.profile:
        kernel_profile(opt, a.profile, rank)
        for mode in (True, 'host'):
            opt.debug_timing, opt.timing = mode, {}
            for _ in range(a.profile):
                opt.iterate()
            opt.debug_timing = False
            if rank == 0:
                print('PROFILE host sections (ms/step, %s): %s' % ('synchronised' if mode is True else 'host time only',
                                                                  {k: round(1e3 * v / a.profile, 3) for k, v in opt.timing.items()}))
    res = {'mesh': a.mesh, 'model': a.model, 'nx': a.nx, 'ny': a.ny, 'world': world, 'n_dist': opt.part.n_dist, 'cuts': opt.part.cuts,
           'ndof': int(gp.mesh.N_nodes * gp.N_dof), 'setup_s': t_setup, 'iteration_s': times, 'log': log,
           'gpu_mem_gb': torch.cuda.max_memory_allocated() / 1e9, 'transport': getattr(opt, 'transport', None),
           'launches': int(F._lib.load().femol_launch_count(0)), 'latency': lat}
    ok = True
    if not a.no_reference and rank == 0:
        ref = build_problem(F, a.mesh, a.nx, a.ny, a.model)
        ref.solver_rtol = a.rtol
        so = F.SIMP_COMP(ref, volfrac=0.4, penal=3, rmin=1.5)
        so.archive = False
        ref_log, ref_times = [], []
        for it in range(a.iters):
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            change, c = so.iterate()
            torch.cuda.synchronize()
            ref_times.append(time.perf_counter() - t1)
            ref_log.append({'c': c, 'change': change, 'pcg': ref.solver_info['iterations']})
        res['ref_iteration_s'] = ref_times
        Xr = so.density()
        res['ref_log'] = ref_log
        res['density_err'] = float(np.abs(X - Xr).max())
        res['c_rel_err'] = [abs(l['c'] - r['c']) / abs(r['c']) for l, r in zip(log, ref_log)]
        ok = res['density_err'] < 1e-4 and max(res['c_rel_err']) < 1e-6
    res['ok'] = ok
    if rank == 0:
        line = json.dumps(res)
        print(line)
        if a.out:
            with open(a.out, 'a') as fh:
                fh.write(line + '\n')
    if world > 1:
        torch.distributed.barrier()
        F._lib.load().femol_dist_destroy(comm)
        torch.distributed.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
