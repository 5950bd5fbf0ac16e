"""Headline benchmark: SIMP_COMP iterations/s on the ~4 M DOF Q4 plane-stress cantilever (BASELINE.json
configs[1] mesh, 2000 x 1000), plus the roofline of the dominant kernel and a CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl femol|reference] [--nx 2000 --ny 1000]

A "step" is one SIMP_COMP iteration (TOM_SIMP.py:48-60): assemble K(x^p) -> PCG solve -> compliance
sensitivities -> filter -> OC update.  Steps continue one optimisation trajectory (densities evolve; the PCG
is warm-started from the previous displacement, as any SIMP driver would).

  value  device-resident throughput: densities stay in HBM between steps; timed with CUDA events.
  e2e    the same step through the public API with HOST buffers: pinned-host densities copied in, new
         densities + (change, objective) copied out, every step, inside the timed region.
  N > 1  independent designs sharded one per GPU (north-star "design batches"): no data-path collective,
         weak scaling; time = max over ranks.  A second leg (`partitioned` in the JSON line) runs ONE T6 problem
         strip-partitioned over all N GPUs (strong scaling) next to the same problem on one GPU.
  --impl reference   the CPU oracle port (vectorised numpy/scipy restatement of the reference algorithm; the
         reference itself is pure Python and cannot travel to the GPU box) on a bounded sample mesh.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'simp_iterations_per_s'
UNIT = 'SIMP iterations/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='femol', choices=['femol', 'reference'])
    ap.add_argument('--nx', type=int, default=2000)
    ap.add_argument('--ny', type=int, default=1000)
    ap.add_argument('--rtol', type=float, default=1e-8)
    ap.add_argument('--precond', default='auto')
    ap.add_argument('--sample-nx', type=int, default=400, help='CPU baseline sample mesh')
    ap.add_argument('--sample-ny', type=int, default=200)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-procs', type=int, default=None, help='CPU legs: parallel design replicas (default: every host core, at most 64)')
    ap.add_argument('--part-n', type=int, default=1414, help='N > 1: T6 mesh (n x n elements) of the strip-partitioned leg; 0 = skip')
    return ap.parse_args()


def workload_name(nx, ny):
    return 'Q4 plane-stress cantilever SIMP_COMP, %dx%d structured mesh (%.2fM DOF), volfrac 0.4, p=3, rmin 1.5' % (
        nx, ny, 2 * (nx + 1) * (ny + 1) / 1e6)


def workload_config(nx, ny, world):
    """The `config` object of the JSON line: identical for both arms (femol and --impl reference); everything that
    describes HOW an arm solved the workload (solver, iteration counts, set-up time) goes to the `solver` key."""
    n = 2 * (nx + 1) * (ny + 1)
    nnz = 4 * (3 * nx + 1) * (3 * ny + 1)          # nnz(K) of an nx x ny Q4 plane mesh
    return {'workload': workload_name(nx, ny), 'ndof': n, 'nnz': nnz,
            'l2_policy': 'inputs larger than L2: one step touches %.0f MB of CSR values (assembly, level-1 operator) and '
                         '%d MB of vectors per PCG iteration; 126 MB L2' % (8 * nnz / 1e6, 6 * 8 * n // 1000000),
            'parallelism': 'one independent design per GPU' if world > 1 else 'single GPU'}


# ---------------------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------------------------
def oracle_problem(nx, ny):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import femol_oracle as o
    Lx, Ly = float(nx), float(ny)
    pts, cells = o.rectangle_points_quads(Lx, Ly, nx, ny)
    fixed = [(o.inside_box([0], [[0, Ly]]), [0, 1])]
    forces = [([0, -1.0], o.inside_box([Lx], [[Ly / 2 - 1, Ly / 2 + 1]]))]
    return o.CompProblem('Q4', pts, cells, 2, o.iso_tensors(1, 0.3, 1), fixed, forces, structured=True,
                         volfrac=0.4, penal=3, rmin=1.5)


def _cpu_worker(job):
    """One process = one design: `warmup` untimed + `steps` timed oracle SIMP iterations on the sample mesh."""
    nx, ny, steps, warmup = job
    for var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[var] = '1'
    P = oracle_problem(nx, ny)
    for _ in range(warmup):
        P.iterate()
    t0 = time.perf_counter()
    for _ in range(steps):
        P.iterate()
    return (time.perf_counter() - t0) / max(steps, 1)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_sample(nx, ny, full_elems, steps=1, warmup=0, procs=None):
    """The reference's CPU path (numpy/scipy oracle port; SuperLU and the Python host code are single-threaded) on
    every host core the box offers: `procs` independent designs in parallel, one per core - the CPU counterpart of
    the north star's "design batches" - each on an nx x ny sample of the workload.  Seconds per iteration are
    scaled linearly in the element count to the full mesh (favourable to the CPU: its sparse direct solve grows
    faster than linearly).  Returns (seconds per sample step of the slowest process, aggregate full-size
    iterations/s over all processes, the same for one process alone, processes used)."""
    import multiprocessing as mp
    procs = procs or min(host_cores(), 64)
    scale = full_elems / float(nx * ny)
    if procs == 1:
        dts = [_cpu_worker((nx, ny, steps, warmup))]
    else:
        with mp.get_context('spawn').Pool(procs) as pool:
            dts = pool.map(_cpu_worker, [(nx, ny, steps, warmup)] * procs)
    dt = max(dts)
    return dt, procs / (dt * scale), 1.0 / (min(dts) * scale), procs


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    full = args.nx * args.ny
    t_all = time.perf_counter()
    dt, value, single, procs = cpu_sample(args.sample_nx, args.sample_ny, full, steps=args.steps, warmup=args.warmup,
                                          procs=args.cpu_procs)
    frac = args.sample_nx * args.sample_ny / float(full)
    sample = ('%d independent designs in parallel (one process per host core), each step = one SIMP iteration of the '
              'numpy/scipy oracle on a %dx%d mesh (%.3f of the %dx%d elements); value = processes / (seconds per step x '
              '%.0f), i.e. seconds scaled linearly in elements; one process alone: %.5f it/s'
              % (procs, args.sample_nx, args.sample_ny, frac, args.nx, args.ny, 1.0 / frac, single))
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup,
            # wall-clock per SAMPLE step as measured (what the driver's own clock sees); the full-size figure
            # the metric is quoted on is ms_per_step_full_size = ms_per_step / sample fraction
            'ms_per_step': dt * 1e3, 'ms_per_step_full_size': dt * 1e3 / frac,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args.nx, args.ny, args.gpus),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': procs, 'kind': 'port', 'sample': sample,
                             'single_process_value': single},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0, 'wall_s': time.perf_counter() - t_all}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# GPU leg
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def build_gpu_problem(F, nx, ny, rtol, precond):
    Lx, Ly = float(nx), float(ny)
    mesh = F.mesh.rectangle_Q4(Lx, Ly, nx, ny)
    problem = F.FEM_Problem('displacement', 'plane', mesh)
    problem.define_materials(F.materials.general_isotropic())
    problem.define_tensors(1)
    problem.add_fixed_domain(F.domains.inside_box([0], [[0, Ly]]), ddls=[0, 1])
    problem.add_forces([0, -1.0], F.domains.inside_box([Lx], [[Ly / 2 - 1, Ly / 2 + 1]]))
    problem.solver_rtol = rtol
    problem.preconditioner = precond
    opt = F.SIMP_COMP(problem, volfrac=0.4, penal=3, rmin=1.5)
    opt.archive = False
    return problem, opt


def run_partitioned(F, args, rank, world):
    """N > 1, second leg (north star: "for large meshes, work is partitioned across 2, 4 and 8 GPUs ... with halo and
    dot-product reductions over NVLink"; BASELINE config 5 at a size whose set-up fits the bench budget): ONE T6
    plane-stress SIMP_COMP problem on all N GPUs (femol_b200.dist.DistributedSIMP_COMP: strips of lattice rows,
    distributed multigrid PCG, halo exchanges / reductions through peer memory), K timed steps after W warm-up steps,
    CUDA events, max over ranks; then the same steps on rank 0 alone (the single-GPU SIMP_COMP) for the speed-up and
    the density parity.  Strong scaling: the problem is fixed, N grows."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from femol_b200 import dist as fd
    n = args.part_n

    def build():
        mesh = F.mesh.rectangle_T6(float(n), float(n), n, n)
        gp = F.FEM_Problem('displacement', 'plane', mesh)
        gp.define_materials(F.materials.general_isotropic())
        gp.define_tensors(1)
        gp.add_fixed_domain(F.domains.inside_box([0], [[0, float(n)]]), ddls=[0, 1])
        gp.add_forces([0, -1.0], F.domains.inside_box([float(n)], [[n / 2.0 - 1, n / 2.0 + 1]]))
        gp.solver_rtol = args.rtol
        return gp

    def timed(iterate, info):
        for _ in range(args.warmup):
            iterate()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        its = []
        e0.record()
        for _ in range(args.steps):
            iterate()
            its.append(info()['iterations'])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3, its

    t0 = time.perf_counter()
    gp = build()
    comm = fd.create_comm(F)
    opt = fd.DistributedSIMP_COMP(F, gp, comm, volfrac=0.4, penal=3, rmin=1.5)
    setup_s = time.perf_counter() - t0
    launches0 = int(F._lib.load().femol_launch_count(0))
    dt, its = timed(opt.iterate, lambda: opt.solver_info)
    launches = int(F._lib.load().femol_launch_count(0)) - launches0
    t = torch.tensor([dt], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t[0])
    stats = (F._lib.c_i64 * 3)()
    healthy = F._lib.load().femol_dist_p2p_status(comm, stats) == 0
    X = opt.density()
    out = {'workload': 'T6 plane-stress cantilever SIMP_COMP, %dx%d elements (%.2fM DOF), ONE problem on %d GPUs'
                       % (n, n, gp.mesh.N_nodes * 2 / 1e6, world),
           'scaling': 'strong', 'value': args.steps / dt, 'unit': UNIT, 'ms_per_step': dt / args.steps * 1e3,
           'pcg_iterations_per_step': its, 'distributed_levels': opt.part.n_dist, 'row_cuts': opt.part.cuts,
           'transport': opt.transport, 'peer_memory_healthy': healthy,
           'p2p_kernels': {'halo': int(stats[0]), 'allreduce': int(stats[1]), 'gather': int(stats[2])},
           'gpu_launches_this_rank': launches, 'setup_s': setup_s}
    transport = opt.transport
    del opt
    torch.cuda.synchronize()
    dist.barrier()
    F._lib.load().femol_dist_destroy(comm)
    if rank == 0:
        ref_p = build()
        ref = F.SIMP_COMP(ref_p, volfrac=0.4, penal=3, rmin=1.5)
        ref.archive = False
    else:
        ref_p = ref = None
    if rank == 0:
        for _ in range(args.warmup):
            ref.iterate()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rits = []
        e0.record()
        for _ in range(args.steps):
            ref.iterate()
            rits.append(ref_p.solver_info['iterations'])
        e1.record()
        torch.cuda.synchronize()
        rdt = e0.elapsed_time(e1) * 1e-3
        out['single_gpu'] = {'value': args.steps / rdt, 'ms_per_step': rdt / args.steps * 1e3, 'pcg_iterations_per_step': rits}
        out['speedup_vs_single_gpu'] = rdt / dt
        out['density_max_abs_diff_vs_single_gpu'] = float(np.abs(X - ref.density()).max())
        del ref, ref_p
    dist.barrier()
    return out if rank == 0 else None


def time_kernel(torch, fn, reps):
    """Average device time of `fn` (one kernel launch) with CUDA events on the current stream."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / reps


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import femol_b200 as F

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=torch.device('cuda', local), timeout=datetime.timedelta(seconds=300))
    L = F._lib.require_gpu()
    t_setup = time.perf_counter()
    problem, opt = build_gpu_problem(F, args.nx, args.ny, args.rtol, args.precond)
    d = problem.device
    n_elem = d.n_elem
    setup_s = time.perf_counter() - t_setup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps -------------------------------------------------------------------------
    for _ in range(args.warmup):
        opt.iterate()
    barrier()
    # both timed legs run the SAME K steps of the trajectory: state after the warm-up (densities + warm-start
    # displacement) is restored before the end-to-end leg (later SIMP steps need more CG iterations)
    x_start = opt.x_device.clone()
    u_start = problem._U_dev.clone()
    sampler = ClockSampler(local) if rank == 0 else None
    L.femol_launch_count(1)
    F._lib.graph_stats(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pcg_iters, log = [], []
    ev0.record()
    for _ in range(args.steps):
        opt.iterate()
        pcg_iters.append(problem.solver_info['iterations'])
        log.append((opt.change, opt.c, problem.solver_info['relres']))
    ev1.record()
    barrier()
    launches = int(L.femol_launch_count(0))
    graph_stats = F._lib.graph_stats()
    dt = ev0.elapsed_time(ev1) * 1e-3

    # ---- end-to-end steps with host buffers ---------------------------------------------------------------
    x_host = torch.empty(n_elem, dtype=torch.float64).pin_memory()
    x_host.copy_(x_start)
    problem._U_dev.copy_(u_start)
    out_host = torch.empty(n_elem, dtype=torch.float64).pin_memory()
    e2e_iters = []
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        opt.x_device.copy_(x_host, non_blocking=True)        # H2D: this step's input densities
        opt.iterate()                                         # includes the (change, objective) read-back
        out_host.copy_(opt.x_device, non_blocking=True)       # D2H: this step's result
        torch.cuda.current_stream().synchronize()
        x_host, out_host = out_host, x_host
        e2e_iters.append(problem.solver_info['iterations'])
    e1.record()
    barrier()
    dt_e2e = e0.elapsed_time(e1) * 1e-3
    clocks = sampler.stop() if sampler else None   # sampled over both timed regions (device-resident + e2e)

    if world > 1:
        t = torch.tensor([dt, dt_e2e], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e = float(t[0]), float(t[1])

    # ---- roofline of the dominant kernel (SpMV inside PCG) and of the assembly ----------------------------
    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        peak_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'
        n, nnz, nnzn = d.n, d.nnz, d.nnz_node
        xv = torch.randn(n, dtype=torch.float64, device='cuda')
        yv = torch.empty_like(xv)
        kv = problem.K_values
        st = F._lib.stream()
        t_spmv = time_kernel(torch, lambda: L.femol_spmv(d.plan, kv.data_ptr(), xv.data_ptr(), yv.data_ptr(), st), 20)
        # SURVEY 8(d) algorithmic bytes (the figure `achieved` is defined on) and the compulsory bytes of OUR
        # kernels, which move less: the SpMV reads the node adjacency (4 B per ndof^2 block) instead of one
        # int32 index per entry; the assembly walks a 4-byte source list per node block instead of a
        # 4*size^2-byte scatter map per element.  traffic = dram read + write of one ncu --set full capture.
        spmv_survey = 12 * nnz + 4 * (n + 1) + 16 * n
        spmv_kernel = 8 * nnz + 4 * nnzn + 4 * (d.n_nodes + 1) + 16 * n
        asm_survey = n_elem * (8 + 4 * (d.nn * d.ndof) ** 2) + 8 * nnz
        asm_kernel = 8 * nnz + 4 * (d.n_nodes + 1) + 4 * nnzn * 4 + 4 * nnzn + 8 * n_elem + d.n_nodes
        grid_asm = bool(problem.structured_grid_shape())
        if grid_asm:      # structured grid: scatter pattern derived from coordinates, a pure write stream
            asm_kernel = 8 * nnz + 4 * (d.n_nodes + 1) + 8 * n_elem + d.n_nodes
        Eb, pe = problem._element_matrices('K', False)
        vals = torch.empty(nnz, dtype=torch.float64, device='cuda')
        mask = problem._fixed_mask()
        xp = d.xpow
        # femol_assemble = density_power (x^p, 2 M elements) + the assembly kernel
        t_asm = time_kernel(torch, lambda: L.femol_assemble(d.plan, Eb.data_ptr(), pe, None, 0, opt.x_device.data_ptr(), 3.0,
                                                            mask.data_ptr(), xp.data_ptr(), vals.data_ptr(), st), 10)
        total_pcg = int(np.sum(pcg_iters))
        nu = problem.gmg_smoothing_steps
        gmg = problem.solver_info.get('preconditioner') == 'gmg'
        matrix_free = bool(gmg and problem.matrix_free and problem.structured_grid_shape())
        spmv_csr = {'kernel': 'spmv_node_kernel<2,4> (femol_spmv: assembled CSR values; unstructured meshes, coarse multigrid levels, block eigen-solver)',
                    'bound': 'hbm', 'achieved': spmv_survey / t_spmv / 1e9, 'peak': peak, 'unit': 'GB/s',
                    'frac': spmv_survey / t_spmv / 1e9 / peak, 'traffic': 828.0e6, 'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': spmv_survey, 'avg_launch_s': t_spmv,
                    'kernel_compulsory_bytes_per_launch': spmv_kernel,
                    'achieved_on_kernel_bytes': spmv_kernel / t_spmv / 1e9,
                    'frac_on_kernel_bytes': spmv_kernel / t_spmv / 1e9 / peak,
                    'traffic_source': 'profiles/r01_ncu_full_summary.md (dram__bytes_read.sum + dram__bytes_write.sum)'}
        if matrix_free:
            # The fine-level products of the solve run matrix-free (csrc/stencil.cu): per PCG iteration one plain
            # product (q = K p, fused dot), one with the residual epilogue and one with the Jacobi epilogue.
            # Algorithmic bytes of one launch = what the operator needs from HBM: u read + y written (16 B / dof),
            # one density per element, the Dirichlet mask; + rhs (modes 1, 2) and the block-diagonal inverse (mode 2).
            nd = d.ndof
            rv = torch.randn(n, dtype=torch.float64, device='cuda')
            dinv = torch.empty(d.n_nodes * nd * nd, dtype=torch.float64, device='cuda')
            F._lib.call('femol_precond_setup', d.plan, kv.data_ptr(), 2, dinv.data_ptr(), st)
            base = 16 * n + 8 * n_elem + d.n_nodes
            mf_bytes = [base, base + 8 * n, base + 8 * n + 8 * nd * n]
            mf_t = [time_kernel(torch, lambda m=m: L.femol_apply_epilogue(d.plan, kv.data_ptr(), xv.data_ptr(), yv.data_ptr(), m,
                                                                       rv.data_ptr(), dinv.data_ptr(), 0.7, st), 20)
                    for m in range(3)]
            per_it = [1, 1, 1] if nu == 1 else [1 + 2 * nu, 0, 0]
            n_launch = [total_pcg * c for c in per_it]
            n_launch[0] += 2 * args.steps                    # initial residual (+ restart check) of every solve
            tot_b = sum(b * c for b, c in zip(mf_bytes, n_launch))
            tot_t = sum(t * c for t, c in zip(mf_t, n_launch))
            roofline = {'kernel': 'stencil_tile_kernel<2,false,32,8,2,2,mode 0|1|2>: matrix-free fine-level K p of PCG and of the V-cycle '
                                  '(launch-weighted over the three epilogue modes)',
                        'bound': 'hbm', 'achieved': tot_b / tot_t / 1e9, 'peak': peak, 'unit': 'GB/s',
                        'frac': tot_b / tot_t / 1e9 / peak, 'traffic': 55.0e6, 'peak_source': peak_src,
                        'traffic_source': 'profiles/r02_ncu_summary.md (mode 0 launch: 50.1 MB read + 4.9 MB written; round 1: 54.9 MB; the vectors are partly L2-resident between kernels)',
                        'algorithmic_bytes_per_launch': tot_b / max(1, sum(n_launch)), 'avg_launch_s': tot_t / max(1, sum(n_launch)),
                        'per_mode': [{'mode': m, 'bytes': mf_bytes[m], 'avg_launch_s': mf_t[m], 'GBps': mf_bytes[m] / mf_t[m] / 1e9,
                                      'launches_in_timed_region': n_launch[m]} for m in range(3)],
                        'equivalent_csr_bytes_per_launch': spmv_survey,
                        'speedup_vs_csr_product': t_spmv / mf_t[0],
                        'launches_in_timed_region': sum(n_launch),
                        'share_of_step': tot_t / dt if dt > 0 else None,
                        'note': 'issue-bound, not HBM-bound: 64 DFMA + staging per node, see DESIGN.md section 5'}
        else:
            fine_spmv = total_pcg * ((1 + 2 * nu) if gmg else 1) + 2 * args.steps
            roofline = dict(spmv_csr)
            roofline['launches_in_timed_region'] = fine_spmv
            roofline['share_of_step'] = fine_spmv * t_spmv / dt if dt > 0 else None
        roofline_asm = {'kernel': 'femol_assemble = density_power + ' + ('assemble_grid_kernel<2,false> (thread per node, smem-staged coalesced stores)' if grid_asm else 'assemble_slots_kernel<2,4,9,4>'), 'bound': 'hbm',
                        'achieved': asm_survey / t_asm / 1e9, 'peak': peak, 'unit': 'GB/s',
                        'frac': asm_survey / t_asm / 1e9 / peak, 'traffic': None if grid_asm else 938.0e6,
                        'algorithmic_bytes_per_launch': asm_survey, 'avg_launch_s': t_asm,
                        'kernel_compulsory_bytes_per_launch': asm_kernel,
                        'frac_on_kernel_bytes': asm_kernel / t_asm / 1e9 / peak,
                        'launches_in_timed_region': args.steps,
                        'share_of_step': args.steps * t_asm / dt if dt > 0 else None}
        # stress recovery: N_e (16 conn + 8 x + 32 out) + 16 n (xy) + 8 n (U) algorithmic bytes
        Tb = torch.from_numpy(np.ascontiguousarray(problem.tensors[0], dtype=np.float64)).cuda()
        s_out = torch.empty((4, n_elem), dtype=torch.float64, device='cuda')
        t_str = time_kernel(torch, lambda: L.femol_stress_q4(d.plan, d.xy.data_ptr(), xv.data_ptr(), Tb.data_ptr(), None,
                                                             opt.x_device.data_ptr(), s_out.data_ptr(), st), 20)
        stress_bytes = n_elem * 56 + 24 * d.n_nodes
        roofline_stress = {'kernel': 'stress_q4_kernel (Mesh.stress_from_displacement)', 'bound': 'hbm',
                           'achieved': stress_bytes / t_str / 1e9, 'peak': peak, 'unit': 'GB/s',
                           'frac': stress_bytes / t_str / 1e9 / peak, 'traffic': None,
                           'algorithmic_bytes_per_launch': stress_bytes, 'avg_launch_s': t_str,
                           'launches_in_timed_region': 0}
        cpu = None
        if not args.no_cpu_baseline:
            cdt, cval, csingle, cprocs = cpu_sample(args.sample_nx, args.sample_ny, args.nx * args.ny, steps=2, warmup=0,
                                                    procs=args.cpu_procs)
            cpu = {'value': cval, 'unit': UNIT, 'cores': cprocs, 'kind': 'port', 'single_process_value': csingle,
                   'sample': '%d independent designs in parallel (one process per host core), 2 SIMP iterations each of the '
                             'numpy/scipy oracle on a %dx%d mesh (%.1f s per iteration), seconds scaled linearly to %dx%d elements'
                             % (cprocs, args.sample_nx, args.sample_ny, cdt, args.nx, args.ny)}
        value = world * args.steps / dt
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': workload_config(args.nx, args.ny, world),
                'solver': {'precond': problem.solver_info.get('preconditioner', args.precond), 'rtol': args.rtol,
                           'pcg_iterations_per_step': pcg_iters, 'warm_start': True, 'setup_s': setup_s,
                           'cuda_graph': graph_stats},
                'e2e': {'value': world * args.steps / dt_e2e, 'unit': UNIT, 'h2d_bytes_per_step': 8 * n_elem,
                        'd2h_bytes_per_step': 8 * n_elem + 64, 'pcg_iterations_per_step': e2e_iters},
                'gpu_launches': launches, 'roofline': roofline, 'roofline_spmv_csr': spmv_csr,
                'roofline_assembly': roofline_asm, 'roofline_stress': roofline_stress,
                'cpu_baseline': cpu, 'clocks': clocks,
                'trajectory': {'change': [l[0] for l in log], 'compliance': [l[1] for l in log],
                               'pcg_relres': [l[2] for l in log]}}
    if world > 1 and args.part_n > 0:
        del opt, problem
        torch.cuda.empty_cache()
        try:
            part = run_partitioned(F, args, rank, world)
        except Exception as exc:      # the headline line must survive a failure of the secondary leg
            part = {'error': '%s: %s' % (type(exc).__name__, exc)} if rank == 0 else None
        if line is not None:
            line['partitioned'] = part
    if line is not None:          # printed before the final collectives: a failed secondary leg must not cost the headline
        emit(line)
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    global _REAL_STDOUT
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        # NCCL prints a version banner on stdout (C stdio) when a communicator is created: for the duration of the run
        # file descriptor 1 points to stderr, and the JSON line goes to a duplicate of the original stdout
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)
        run_gpu(args)


if __name__ == '__main__':
    main()
