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
