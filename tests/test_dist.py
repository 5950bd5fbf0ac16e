"""Host logic of the multi-GPU decomposition (femol_b200/dist.py) on CPU: gloo, world size 2 (and 3).
numpy/scipy stand in for the device kernels: every rank assembles its LOCAL mesh with the oracle, and a CG
whose halo exchange / reductions go through torch.distributed must reproduce the global solution."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case_name, out_q):
    for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import cases
    import femol_oracle as o
    from femol_b200.dist import Partition, shard_items
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        case = cases.CASES[case_name]
        arr = o.case_arrays(case, cases.jitter_points)
        nd = arr['ndof']
        pts, cells = arr['points'], arr['cells']
        part = Partition(pts.shape[0], cells, rank, world)
        part.exchange_halo_lists(part.gather_all_ghosts())
        # --- invariants of the partition ---
        assert part.local_to_global[:part.n_owned].tolist() == list(range(part.node_begin, part.node_end))
        assert np.all(np.diff(part.ghost_global) > 0)
        assert np.array_equal(part.local_to_global[part.local_cells], cells[part.local_elements])
        owners = torch.zeros(cells.shape[0], dtype=torch.int64)
        owners[torch.from_numpy(part.local_elements[part.element_owner_is_me])] = 1
        dist.all_reduce(owners)
        assert bool((owners == 1).all()), 'every element has exactly one owner'
        # --- local assembly (oracle on the local mesh), rows of owned nodes only ---
        larr = dict(arr)
        larr['points'] = pts[part.local_to_global]
        larr['cells'] = part.local_cells
        larr['fixed_idx'] = o.fixed_dofs(larr['points'], nd, arr['fixed'])
        X = cases.density_field(cells.shape[0])
        K_loc = o.global_matrices(larr, 'K', X[part.local_elements], 3)
        n_own = part.n_owned * nd
        K_rows = K_loc[:n_own]
        K_glob = o.global_matrices(arr, 'K', X, 3)
        gidx = (part.local_to_global[:, None] * nd + np.arange(nd)[None]).reshape(-1)
        # the owned rows of the local matrix ARE the global rows (halo elements assembled redundantly)
        ref_rows = K_glob[gidx[:n_own]][:, gidx]
        assert abs(K_rows - ref_rows).max() < 1e-12 * abs(K_glob).max()
        # --- distributed Jacobi-CG with halo exchange over gloo ---
        F_loc = arr['F'][gidx][:n_own]
        dinv = 1.0 / K_rows.diagonal(0) if False else 1.0 / np.asarray(K_loc.diagonal())[:n_own]
        send = dict(part.send)
        recv = {p_: (s_, c_) for p_, s_, c_ in part.recv}

        def halo(v):                      # v: local vector (owned + ghost dofs)
            reqs = []
            bufs = {}
            for q in part.neighbours():
                sb = torch.from_numpy(np.ascontiguousarray(v.reshape(-1, nd)[send[q]].reshape(-1)))
                reqs.append(dist.isend(sb, q))
                rb = torch.empty(recv[q][1] * nd, dtype=torch.float64)
                bufs[q] = rb
                reqs.append(dist.irecv(rb, q))
            for r_ in reqs:
                r_.wait()
            for q, rb in bufs.items():
                s_, c_ = recv[q]
                v[s_ * nd:(s_ + c_) * nd] = rb.numpy()

        def gsum(*vals):
            t = torch.tensor(vals, dtype=torch.float64)
            dist.all_reduce(t)
            return t.tolist()

        n_loc = part.local_to_global.shape[0] * nd
        u = np.zeros(n_loc); p = np.zeros(n_loc)
        r = F_loc.copy()
        z = dinv * r
        p[:n_own] = z
        rz, ff = gsum(r @ z, F_loc @ F_loc)
        for it in range(5000):
            halo(p)
            q = K_rows @ p
            (pq,) = gsum(p[:n_own] @ q)
            a = rz / pq
            u[:n_own] += a * p[:n_own]
            r -= a * q
            z = dinv * r
            rz_new, rr = gsum(r @ z, r @ r)
            if rr <= 1e-22 * ff:
                break
            p[:n_own] = z + (rz_new / rz) * p[:n_own]
            rz = rz_new
        U_ref = o.solve_static(K_glob, arr['F'])
        err = np.abs(u[:n_own] - U_ref[gidx[:n_own]]).max() / np.abs(U_ref).max()
        assert err < 1e-7, err
        # --- design-batch sharding ---
        items = list(range(7))
        mine = shard_items(items, rank, world)
        t = torch.zeros(7, dtype=torch.int64); t[mine] = 1
        dist.all_reduce(t)
        assert t.tolist() == [1] * 7
        out_q.put((rank, 'ok', it))
    except Exception as e:  # pragma: no cover
        import traceback
        out_q.put((rank, 'fail: %s\n%s' % (e, traceback.format_exc()), -1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,case_name', [(2, 'comp_q4_30x15'), (2, 'comp_t6_8x4'), (3, 'plate_q4_jitter_6x5')])
def test_row_block_partition_gloo(world, case_name):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, it in sorted(results):
        assert status == 'ok', 'rank %d: %s' % (rank, status)


def test_partition_single_rank():
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import femol_oracle as o
    from femol_b200.dist import Partition
    pts, cells = o.rectangle_points_quads(4., 3., 4, 3)
    part = Partition(pts.shape[0], cells, 0, 1)
    assert part.n_owned == pts.shape[0] and part.ghost_global.size == 0 and part.recv == []
    assert np.array_equal(part.local_cells, cells)
    # strips of a structured grid: two ranks, the halo of each is exactly one node row
    p0 = Partition(pts.shape[0], cells, 0, 2, node_ranges=[0, 10, 20])
    p1 = Partition(pts.shape[0], cells, 1, 2, node_ranges=[0, 10, 20])
    assert p0.ghost_global.tolist() == [10, 11, 12, 13, 14] and p1.ghost_global.tolist() == [5, 6, 7, 8, 9]
    p0.exchange_halo_lists([p0.ghost_global, p1.ghost_global])
    p1.exchange_halo_lists([p0.ghost_global, p1.ghost_global])
    assert p0.send[0][0] == 1 and p0.send[0][1].tolist() == [5, 6, 7, 8, 9]
    assert p1.send[0][0] == 0 and p1.send[0][1].tolist() == [0, 1, 2, 3, 4]


# ---- the real thing: femol_pcg_dist over NCCL on 2 GPUs (skipped on boxes with fewer) -------------------------------
def _gpu_worker(rank, world, port, nx, ny, model, out_q):
    for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import femol_b200 as F
        from femol_b200 import dist as fd
        import cases
        Lx, Ly = float(nx), float(ny)
        mesh = F.mesh.rectangle_Q4(Lx, Ly, nx, ny)
        gp = F.FEM_Problem('displacement', model, mesh)
        if model == 'plane':
            gp.define_materials(F.materials.general_isotropic()); gp.define_tensors(1)
            gp.add_fixed_domain(F.domains.inside_box([0], [[0, Ly]]), ddls=[0, 1])
            gp.add_forces([0, -1.0], F.domains.inside_box([Lx], [[0, Ly]]))
        else:
            gp.define_materials(F.materials.IsotropicMaterial(10920, 0.3, 1)); gp.define_tensors(0.5)
            gp.add_fixed_domain(F.domains.inside_box([0], [[0, Ly]]))
            gp.add_forces([0, 0, -1e-2, 0, 0, 0], F.domains.inside_box([Lx], [[0, Ly]]))
        X = cases.density_field(mesh.N_ele)
        comm = fd.create_comm(F)
        ds = fd.DistributedStatic(F, gp, rank, world, comm)
        ds.assemble(X, 3)
        rtol = 1e-9 if model == 'plane' else 1e-7
        U_own = ds.solve(rtol=rtol)
        info = dict(ds.solver_info)
        assert info['converged'], info
        # gather the owned pieces on every rank and compare with the single-GPU solve of the whole problem
        nd = gp.N_dof
        sizes = [int((ds.part.bounds[r + 1] - ds.part.bounds[r]) * nd) for r in range(world)]
        parts = [torch.empty(s, dtype=torch.float64, device='cuda') for s in sizes]
        dist.all_gather(parts, U_own.contiguous())
        U = torch.cat(parts).cpu().numpy()
        gp.preconditioner = 'block_jacobi'
        gp.solver_rtol = rtol
        gp.assemble('K', X={'quad': X}, p=3)
        U_ref = gp.solve_device().cpu().numpy()
        err = np.abs(U - U_ref).max() / np.abs(U_ref).max()
        cerr = abs(gp.F @ U - gp.F @ U_ref) / abs(gp.F @ U_ref)
        assert cerr < 1e-6 and err < 1e-4, (err, cerr, info, gp.solver_info)
        F._lib.load().femol_dist_destroy(comm)
        out_q.put((rank, 'ok', (info['iterations'], gp.solver_info['iterations'], err)))
    except Exception as e:  # pragma: no cover
        import traceback
        out_q.put((rank, 'fail: %s\n%s' % (e, traceback.format_exc()), -1))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize('nx,ny,model', [(96, 48, 'plane'), (24, 13, 'plate')])
def test_pcg_dist_two_gpus(nx, ny, model):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, nx, ny, model, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = []
    try:
        for _ in procs:
            results.append(q.get(timeout=150))
            if results[-1][1] != 'ok':      # a failed rank leaves its peer waiting in a collective: stop now
                break
    finally:
        for p in procs:
            p.join(timeout=5 if (results and results[-1][1] != 'ok') else 60)
            if p.is_alive():
                p.kill()
    assert len(results) == len(procs) and all(r[1] == 'ok' for r in results), results


# =====================================================================================================================
# Strip partition with multigrid (femol_b200/dist.py StripPartition / DistributedSIMP_COMP)
# =====================================================================================================================
@pytest.mark.parametrize('Ny,world,n_dist', [(2828, 8, 4), (48, 2, 2), (1000, 4, 5), (129, 3, 3), (17, 1, 2), (64, 2, 1)])
def test_strip_partition_levels_coarsen_like_the_global_grid(Ny, world, n_dist):
    from femol_b200.dist import StripPartition
    parts = [StripPartition(Ny, r, world, n_dist) for r in range(world)]
    A = 1 << n_dist
    assert all(c % A == 0 for c in parts[0].cuts[1:-1]) and parts[0].cuts[0] == 0 and parts[0].cuts[-1] == Ny + 1
    ny_g = Ny
    for l in range(n_dist + 1):
        ob = [p.levels[l]['own_begin'] for p in parts]
        oe = [p.levels[l]['own_end'] for p in parts]
        assert ob[0] == 0 and oe[-1] == ny_g + 1 and all(oe[i] == ob[i + 1] for i in range(world - 1))   # owned rows tile the level
        for p in parts:
            lev = p.levels[l]
            assert lev['Ny'] == ny_g and lev['row0'] <= lev['own_begin'] < lev['own_end'] <= lev['top'] + 1
            if p.rank > 0:
                assert lev['own_begin'] - lev['row0'] >= lev['h_up']              # the exchanged rows exist locally
            if p.rank < world - 1:
                assert lev['top'] - (lev['own_end'] - 1) >= lev['h_down']
            if l < n_dist:
                nyl = lev['top'] - lev['row0']
                nxt = p.levels[l + 1]
                # ceil(n / 2) coarsening of the local grid lands on the local grid of the next level
                assert lev['row0'] % 2 == 0 and lev['row0'] // 2 == nxt['row0']
                assert (nyl + 1) // 2 == nxt['top'] - nxt['row0']
                if p.rank < world - 1:
                    assert nyl % 2 == 0                                           # no spurious trailing single cell
        ny_g = (ny_g + 1) // 2


def test_strip_partition_rejects_thin_strips():
    from femol_b200.dist import StripPartition
    with pytest.raises(ValueError):
        StripPartition(30, 0, 4, 4)


def test_coarsen_mask_host_matches_the_device_rule():
    from femol_b200.dist import coarsen_mask_host
    rng = np.random.default_rng(0)
    for nx, ny in [(5, 6), (8, 4), (7, 7), (1, 3)]:
        m = rng.integers(0, 4, size=(ny + 1) * (nx + 1)).astype(np.uint8)
        out, nxc, nyc = coarsen_mask_host(m, nx, ny)
        assert (nxc, nyc) == ((nx + 1) // 2, (ny + 1) // 2)
        for J in range(nyc + 1):
            for I in range(nxc + 1):            # csrc/gmg.cu coarsen_mask: coarse (I, J) sits on fine (min(2I, nx), min(2J, ny))
                assert out[J * (nxc + 1) + I] == m[min(2 * J, ny) * (nx + 1) + min(2 * I, nx)]


def _strip_worker(rank, world, port, kind, nx, ny, n_dist, out_q):
    for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import femol_oracle as o
    import cases
    from femol_b200 import dist as fd
    from femol_b200 import mesh as fmesh
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        gm = {'q4': fmesh.rectangle_Q4, 't6': fmesh.rectangle_T6, 't3': fmesh.rectangle_T3}[kind](float(nx), float(ny), nx, ny)
        ctype = gm.contains[0]
        cells = np.asarray(gm.cells[ctype])
        pts = gm.points[:, :2]
        # lattice of the mesh (host logic of FEM_Problem.node_lattice, without a device)
        hx = np.unique(np.round(pts[:, 0], 9)); hy = np.unique(np.round(pts[:, 1], 9))
        Nx, Ny = hx.shape[0] - 1, hy.shape[0] - 1
        li = np.rint(pts[:, 0] / (hx[1] - hx[0])).astype(np.int64); lj = np.rint(pts[:, 1] / (hy[1] - hy[0])).astype(np.int64)
        lat = lj * (Nx + 1) + li
        lat_node = np.empty(gm.N_nodes, dtype=np.int64); lat_node[lat] = np.arange(gm.N_nodes)
        if kind == 'q4':
            assert np.array_equal(lat_node, np.arange(gm.N_nodes))
            lat_node = None
        h_up0 = fd.strip_halo_rows(cells, Nx, lat_node, gm.N_nodes)
        assert h_up0 == (2 if kind == 't6' else 1)
        part = fd.StripPartition(Ny, rank, world, n_dist, h_up0=h_up0)
        cut = fd.strip_local_mesh(cells, Nx, lat_node, gm.N_nodes, part)
        g_nodes = cut['nodes']
        # local connectivity = global connectivity in local numbering; local nodes in lattice order
        assert np.array_equal(g_nodes[cut['cells']], cells[cut['elements']])
        lp = pts[g_nodes]
        assert np.all(np.diff(lp[:, 1].reshape(-1, Nx + 1)[:, 0]) > 0) and np.all(np.diff(lp[:Nx + 1, 0]) > 0)
        # every element has exactly one owner; the all-gather index puts every owned value in its global slot
        owned_global = cut['elements'][cut['owned']]
        m, index = fd.owned_gather_index(owned_global, cells.shape[0], world)
        send = torch.full((m,), -1.0, dtype=torch.float64)
        send[:owned_global.shape[0]] = torch.from_numpy(owned_global.astype(np.float64))
        recv = [torch.empty(m, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(recv, send)
        full = torch.full((cells.shape[0] + 1,), -7.0, dtype=torch.float64)
        full[torch.from_numpy(index)] = torch.cat(recv)
        assert np.array_equal(full[:-1].numpy(), np.arange(cells.shape[0], dtype=np.float64))
        # owned rows of the LOCAL matrix are the global rows (local assembly needs no communication)
        nd = 2
        etype = {'q4': 'Q4', 't6': 'T6', 't3': 'T3'}[kind]
        arr = {'etype': etype, 'points': gm.points, 'cells': cells, 'ndof': nd, 'tensors': o.iso_tensors(1.0, 0.3, 1.0),
               'coat_tensors': None, 'structured': False, 'fixed_idx': np.zeros(0, dtype=np.int64)}
        X = cases.density_field(cells.shape[0])
        Kg = o.global_matrices(arr, 'K', X, 3)
        larr = dict(arr); larr['points'] = gm.points[g_nodes]; larr['cells'] = cut['cells']
        Kl = o.global_matrices(larr, 'K', X[cut['elements']], 3)
        npx = Nx + 1
        lo, hi = (part.own_begin - part.row0) * npx, (part.own_end - part.row0) * npx
        rows = (np.arange(lo, hi)[:, None] * nd + np.arange(nd)[None]).reshape(-1)
        gidx = (g_nodes[:, None] * nd + np.arange(nd)[None]).reshape(-1)
        ref = Kg[gidx[rows]][:, gidx]
        assert abs(Kl[rows] - ref).max() < 1e-12 * abs(Kg).max()
        # ... and they only reference owned nodes and the halo rows an exchange fills
        lev = part.levels[0]
        cols = np.unique(Kl[rows].tocoo().col // nd) // npx + part.row0
        assert cols.min() >= part.own_begin - (lev['h_up'] if rank > 0 else 0)
        assert cols.max() <= part.own_end - 1 + (lev['h_down'] if rank < world - 1 else 0)
        out_q.put((rank, 'ok', 0))
    except Exception as e:  # pragma: no cover
        import traceback
        out_q.put((rank, 'fail: %s\n%s' % (e, traceback.format_exc()), -1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,kind,nx,ny,n_dist', [(2, 'q4', 12, 16, 2), (2, 't6', 6, 8, 2), (3, 't3', 5, 24, 2)])
def test_strip_local_mesh_gloo(world, kind, nx, ny, n_dist):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_strip_worker, args=(r, world, port, kind, nx, ny, n_dist, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, it in sorted(results):
        assert status == 'ok', 'rank %d: %s' % (rank, status)


# ---- DistributedSIMP_COMP on the device: one rank (always) and two ranks (boxes with 2 GPUs) -------------------------
def _simp_problem(F, kind, nx, ny):
    Lx, Ly = float(nx), float(ny)
    mesh = {'q4': F.mesh.rectangle_Q4, 't3': F.mesh.rectangle_T3, 't6': F.mesh.rectangle_T6}[kind](Lx, Ly, nx, ny)
    gp = F.FEM_Problem('displacement', 'plane', mesh)
    gp.define_materials(F.materials.general_isotropic()); gp.define_tensors(1)
    gp.add_fixed_domain(F.domains.inside_box([0], [[0, Ly]]), ddls=[0, 1])
    gp.add_forces([0, -1.0], F.domains.inside_box([Lx], [[Ly / 2 - 1, Ly / 2 + 1]]))
    gp.solver_rtol = 1e-10
    return gp


def _check_strip_simp(F, fd, kind, nx, ny, n_dist, comm, iters=3):
    opt = fd.DistributedSIMP_COMP(F, _simp_problem(F, kind, nx, ny), comm, volfrac=0.4, penal=3, rmin=1.5, n_dist=n_dist)
    log = []
    for _ in range(iters):
        opt.iterate()
        log.append((opt.c, opt.solver_info['iterations'], opt.solver_info['converged']))
    ref_p = _simp_problem(F, kind, nx, ny)
    ref = F.SIMP_COMP(ref_p, volfrac=0.4, penal=3, rmin=1.5)
    ref.archive = False
    ref_log = []
    for _ in range(iters):
        ref.iterate()
        ref_log.append((ref.c, ref_p.solver_info['iterations']))
    assert all(l[2] for l in log), log
    assert np.abs(opt.density() - ref.density()).max() < 1e-4                        # north star: densities within 1e-4
    assert max(abs(a[0] - b[0]) / abs(b[0]) for a, b in zip(log, ref_log)) < 1e-6     # compliance within 1e-6
    # the distributed V-cycle is the same operator: iteration counts differ by rounding only
    assert max(abs(a[1] - b[1]) for a, b in zip(log, ref_log)) <= 1, (log, ref_log)
    nodes, U = opt.owned_displacement()
    U_ref = ref._U.view(-1, 2).cpu().numpy()[nodes]
    assert np.abs(U - U_ref).max() < 1e-6 * np.abs(U_ref).max()
    # static solve of the solid structure through the same machinery (FEM_Problem.assemble + solve)
    nodes, U = opt.solve_static()
    sp = _simp_problem(F, kind, nx, ny)
    sp.assemble('K')
    U_ref = sp.solve_device().view(-1, 2).cpu().numpy()[nodes]
    assert opt.solver_info['converged'] and np.abs(U - U_ref).max() < 1e-6 * np.abs(U_ref).max()
    return opt


@pytest.mark.gpu
@pytest.mark.parametrize('kind,nx,ny,n_dist', [('q4', 48, 40, 2), ('t6', 20, 24, 2), ('t3', 33, 21, 1)])
def test_strip_simp_single_rank(kind, nx, ny, n_dist):
    """One rank: the truncated strip hierarchy + the gathered, replicated coarse hierarchy reproduce SIMP_COMP."""
    import femol_b200 as F
    from femol_b200 import dist as fd
    comm = fd.create_comm(F)
    try:
        opt = _check_strip_simp(F, fd, kind, nx, ny, n_dist, comm)
        assert opt.part.cuts == [0, opt.Ny + 1]
        del opt
    finally:
        F._lib.load().femol_dist_destroy(comm)


def _gpu_strip_worker(rank, world, port, kind, nx, ny, n_dist, transport, out_q):
    for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), FEMOL_DIST_TRANSPORT=transport)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import femol_b200 as F
        from femol_b200 import dist as fd
        comm = fd.create_comm(F)
        opt = _check_strip_simp(F, fd, kind, nx, ny, n_dist, comm)
        assert opt.transport == transport
        stats = (ctypes_i64 * 3)()
        assert F._lib.load().femol_dist_p2p_status(comm, stats) == 0
        if transport == 'p2p':
            assert stats[0] > 0 and stats[1] > 0 and stats[2] > 0          # halo / all-reduce / gather kernels ran
        X = torch.from_numpy(opt.density()).cuda()
        Xs = [torch.empty_like(X) for _ in range(world)]
        dist.all_gather(Xs, X)
        assert all(torch.equal(Xs[0], x) for x in Xs), 'replicated densities must be bit-identical on all ranks'
        del opt
        torch.cuda.synchronize()
        dist.barrier()
        F._lib.load().femol_dist_destroy(comm)
        out_q.put((rank, 'ok', 0))
    except Exception as e:  # pragma: no cover
        import traceback
        out_q.put((rank, 'fail: %s\n%s' % (e, traceback.format_exc()), -1))
    finally:
        dist.destroy_process_group()


import ctypes as _ctypes  # noqa: E402
ctypes_i64 = _ctypes.c_int64


@pytest.mark.gpu
@pytest.mark.parametrize('kind,nx,ny,n_dist,transport', [('q4', 48, 40, 2, 'p2p'), ('t6', 20, 24, 2, 'p2p'), ('q4', 40, 33, 1, 'nccl')])
def test_strip_simp_two_gpus(kind, nx, ny, n_dist, transport):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_strip_worker, args=(r, 2, port, kind, nx, ny, n_dist, transport, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = []
    try:
        for _ in procs:
            results.append(q.get(timeout=200))
            if results[-1][1] != 'ok':
                break
    finally:
        for p in procs:
            p.join(timeout=5 if (results and results[-1][1] != 'ok') else 60)
            if p.is_alive():
                p.kill()
    assert len(results) == len(procs) and all(r[1] == 'ok' for r in results), results


# ---- the strip-distributed V-cycle, emulated in numpy over gloo ------------------------------------------------------
def _strip_vcycle_worker(rank, world, port, kind, nx, ny, n_dist, out_q):
    """numpy stand-in for csrc/gmg.cu dvcycle + csrc/dist.cu transfers: every rank holds the rows [row0, top] of the
    operators of the distributed levels, exchanges halo rows over gloo exactly where the device code does (e, t, e_c;
    NOT before the post-smoothing product), gathers the level-n_dist residual and runs the remaining levels on the
    whole grid.  After every operation all rows that the device code leaves undefined (everything that is neither
    owned nor just filled by an exchange) are overwritten with random numbers: the owned rows of the result must still
    equal the global V-cycle to rounding, i.e. garbage never reaches an owned row."""
    for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import femol_oracle as o
    import cases
    import numpy_ops
    from femol_b200 import dist as fd
    from femol_b200 import mesh as fmesh
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        nd, omega = 2, 0.7
        gm = {'q4': fmesh.rectangle_Q4, 't6': fmesh.rectangle_T6}[kind](float(nx), float(ny), nx, ny)
        cells = np.asarray(gm.cells[gm.contains[0]])
        pts = gm.points[:, :2]
        hx = np.unique(np.round(pts[:, 0], 9)); hy = np.unique(np.round(pts[:, 1], 9))
        Nx, Ny = hx.shape[0] - 1, hy.shape[0] - 1
        li = np.rint(pts[:, 0] / (hx[1] - hx[0])).astype(np.int64); lj = np.rint(pts[:, 1] / (hy[1] - hy[0])).astype(np.int64)
        node_lat = lj * (Nx + 1) + li
        lat_node = np.empty(gm.N_nodes, dtype=np.int64); lat_node[node_lat] = np.arange(gm.N_nodes)
        # global operator in LATTICE order (what every strip uses locally), clamped on the left edge
        fixed_nodes = np.nonzero(np.isclose(pts[:, 0], 0.0))[0]
        fixed_idx = (fixed_nodes[:, None] * nd + np.arange(nd)[None]).reshape(-1)
        arr = {'etype': {'q4': 'Q4', 't6': 'T6'}[kind], 'points': gm.points, 'cells': cells, 'ndof': nd,
               'tensors': o.iso_tensors(1.0, 0.3, 1.0), 'coat_tensors': None, 'structured': False, 'fixed_idx': fixed_idx}
        X = cases.density_field(cells.shape[0])
        K = o.global_matrices(arr, 'K', X, 3).tocsr()
        dperm = (lat_node[:, None] * nd + np.arange(nd)[None]).reshape(-1)           # lattice dof -> mesh dof
        K = K[dperm][:, dperm].tocsr()
        fixed_lat = np.nonzero(np.isin(dperm, fixed_idx))[0]
        lat = {'Nx': Nx, 'Ny': Ny, 'node_lat': None}
        rng = np.random.default_rng(3)
        r = rng.standard_normal(K.shape[0]); r[fixed_lat] = 0.0
        z_ref, levels, cyc = numpy_ops.numpy_lattice_vcycle(K, lat, fixed_lat, r, omega, nd=nd, return_levels=True)
        assert len(levels) > n_dist + 1, 'mesh too small for %d distributed levels' % n_dist

        h_up0 = fd.strip_halo_rows(cells, Nx, lat_node, gm.N_nodes)
        part = fd.StripPartition(Ny, rank, world, n_dist, h_up0=h_up0)
        garbage = np.random.default_rng(100 + rank)
        nxs = [Nx]
        for _ in range(n_dist):
            nxs.append((nxs[-1] + 1) // 2)

        def dofs(l, row_a, row_b):                       # global (lattice) dofs of the level-l rows [row_a, row_b)
            npx = nxs[l] + 1
            return np.arange(row_a * npx * nd, row_b * npx * nd)

        class Strip(object):
            def __init__(self, l):
                lev = part.levels[l]
                self.l, self.lev = l, lev
                self.rowd = (nxs[l] + 1) * nd
                self.loc = dofs(l, lev['row0'], lev['top'] + 1)
                self.lo = (lev['own_begin'] - lev['row0']) * self.rowd          # owned local range [lo, hi)
                self.hi = (lev['own_end'] - lev['row0']) * self.rowd
                self.up, self.down = lev['h_up'] * self.rowd, lev['h_down'] * self.rowd

            def poison(self, v, keep_halo):
                """rows the device leaves undefined: all non-owned rows, except (keep_halo) the exchanged halo rows"""
                lo_keep = self.lo - (self.up if (keep_halo and rank > 0) else 0)
                hi_keep = self.hi + (self.down if (keep_halo and rank < world - 1) else 0)
                v[:lo_keep] = garbage.standard_normal(lo_keep)
                v[hi_keep:] = garbage.standard_normal(v.shape[0] - hi_keep)

            def exchange(self, v):
                reqs, bufs = [], {}
                if rank > 0:
                    reqs.append(dist.isend(torch.from_numpy(v[self.lo:self.lo + self.down].copy()), rank - 1))
                    bufs['lo'] = torch.empty(self.up, dtype=torch.float64); reqs.append(dist.irecv(bufs['lo'], rank - 1))
                if rank < world - 1:
                    reqs.append(dist.isend(torch.from_numpy(v[self.hi - self.up:self.hi].copy()), rank + 1))
                    bufs['hi'] = torch.empty(self.down, dtype=torch.float64); reqs.append(dist.irecv(bufs['hi'], rank + 1))
                for q in reqs:
                    q.wait()
                if 'lo' in bufs:
                    v[self.lo - self.up:self.lo] = bufs['lo'].numpy()
                if 'hi' in bufs:
                    v[self.hi:self.hi + self.down] = bufs['hi'].numpy()

        strips = [Strip(l) for l in range(n_dist + 1)]

        def jac(l, rr, e=None):
            A, D, _ = levels[l]
            s = strips[l]
            Aloc = A[s.loc][:, s.loc]
            res = rr if e is None else rr - Aloc @ e
            dz = omega * np.einsum('nab,nb->na', D[s.loc[::nd] // nd], res.reshape(-1, nd)).reshape(-1)
            return dz if e is None else e + dz

        def dvcycle(l, rr):
            A, D, P = levels[l]
            s, sc = strips[l], strips[l + 1]
            Aloc = A[s.loc][:, s.loc]
            Ploc = P[s.loc][:, sc.loc]
            e = jac(l, rr)                                  # first sweep
            s.poison(e, False); s.exchange(e)
            t = rr - Aloc @ e
            s.poison(t, False); s.exchange(t)
            rc = Ploc.T @ t
            sc.poison(rc, False)
            if l + 1 < n_dist:
                ec = dvcycle(l + 1, rc)
                sc.poison(ec, False); sc.exchange(ec)
            else:                                           # gather the owned rows, remaining levels on the whole grid
                pieces = [None] * world
                dist.all_gather_object(pieces, rc[sc.lo:sc.hi].copy())
                ec = cyc(l + 1, np.concatenate(pieces))[sc.loc]
            t = e + Ploc @ ec                                # NO exchange: both inputs are valid on the halo rows
            s.poison(t, True)
            out = jac(l, rr, t)
            s.poison(out, False)
            return out

        s0 = strips[0]
        r_loc = r[s0.loc].copy()
        s0.poison(r_loc, False)
        z_loc = dvcycle(0, r_loc)
        err = np.abs(z_loc[s0.lo:s0.hi] - z_ref[s0.loc][s0.lo:s0.hi]).max() / np.abs(z_ref).max()
        assert err < 1e-12, err
        out_q.put((rank, 'ok', err))
    except Exception as e:  # pragma: no cover
        import traceback
        out_q.put((rank, 'fail: %s\n%s' % (e, traceback.format_exc()), -1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,kind,nx,ny,n_dist', [(2, 'q4', 20, 32, 2), (2, 't6', 12, 16, 2), (3, 'q4', 17, 47, 1)])
def test_strip_vcycle_gloo(world, kind, nx, ny, n_dist):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_strip_vcycle_worker, args=(r, world, port, kind, nx, ny, n_dist, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, it in sorted(results):
        assert status == 'ok', 'rank %d: %s' % (rank, status)


def test_strip_partition_invariants_property():
    """Random (rows, ranks, levels): whenever a partition exists its levels tile the rows, every strip coarsens like the
    global grid, and pick_levels never proposes strips thinner than four alignment units."""
    from hypothesis import given, settings, strategies as st
    from femol_b200.dist import StripPartition

    @settings(max_examples=200, deadline=None)
    @given(st.integers(8, 6000), st.integers(1, 8), st.integers(1, 6), st.integers(1, 2))
    def check(Ny, world, n_dist, h_up0):
        try:
            parts = [StripPartition(Ny, r, world, n_dist, h_up0=h_up0) for r in range(world)]
        except ValueError:
            return                                        # strips too thin for this many levels: rejected, fine
        A = 1 << n_dist
        assert all(c % A == 0 for c in parts[0].cuts[1:-1])
        ny_g = Ny
        for l in range(n_dist + 1):
            assert [p.levels[l]['own_begin'] for p in parts][1:] == [p.levels[l]['own_end'] for p in parts][:-1]
            assert parts[0].levels[l]['own_begin'] == 0 and parts[-1].levels[l]['own_end'] == ny_g + 1
            for p in parts:
                lev = p.levels[l]
                assert lev['row0'] <= lev['own_begin'] < lev['own_end'] <= lev['top'] + 1 <= ny_g + 1
                if l < n_dist:
                    nxt = p.levels[l + 1]
                    assert lev['row0'] % 2 == 0 and lev['row0'] // 2 == nxt['row0']
                    assert (lev['top'] - lev['row0'] + 1) // 2 == nxt['top'] - nxt['row0']
            ny_g = (ny_g + 1) // 2
        n = StripPartition.pick_levels(2 * Ny, Ny, world)
        assert n == 1 or (Ny + 1) // world >= 4 * (1 << n)
    check()
