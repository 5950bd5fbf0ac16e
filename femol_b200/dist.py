"""Multi-GPU decomposition of the hot path, one process per GPU under torch.distributed.

Two modes, as the north star lists them:

* **design / load-case batches** — independent problems, one per GPU, no data-path collective
  (`shard_items`; `bench.py --gpus N` reports it as weak scaling).
* **row-block partition of one large problem** — every rank owns a contiguous range of nodes (structured
  meshes: strips of node rows, because nodes are numbered x-fastest) and builds a LOCAL mesh: its owned nodes
  first, then the ghost nodes of the one halo layer of elements, grouped by owning rank.  Assembly is local
  and communication free (halo elements are assembled redundantly); the distributed block-Jacobi PCG
  (`femol_pcg_dist`, csrc/dist.cu) exchanges one halo of the search direction per iteration with
  ncclSend/ncclRecv and all-reduces the CG scalars in stream.

The index bookkeeping below (`Partition`) is pure host logic and is tested on CPU with gloo at world size 2
(tests/test_dist.py), where numpy/scipy stand in for the device kernels.
"""
import ctypes
import os

import numpy as np

__all__ = ['shard_items', 'Partition', 'DistributedStatic', 'init_process_group', 'StripPartition', 'DistributedSIMP_COMP',
           'create_comm', 'coarsen_mask_host', 'owned_gather_index', 'strip_local_mesh', 'strip_halo_rows']


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (RANK / WORLD_SIZE / MASTER_*)."""
    import os
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world == 1:
        return 0, 1
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    else:
        dist.init_process_group(backend)
    return rank, world


def shard_items(items, rank, world):
    """Independent designs / load cases: item k goes to rank k % world (no data-path collective)."""
    return [it for k, it in enumerate(items) if k % world == rank]


class Partition(object):
    """Row-block partition of a mesh for `world` ranks, seen from `rank`.

    owned nodes : the contiguous global range [node_begin, node_end)
    local elements : every element touching an owned node (owned + one halo layer)
    local nodes : owned nodes in global order, then ghost nodes sorted by global id (hence grouped by owner)
    """

    def __init__(self, n_nodes, cells, rank, world, node_ranges=None):
        cells = np.asarray(cells)
        self.rank, self.world = rank, world
        self.n_nodes_global = int(n_nodes)
        if node_ranges is None:
            bounds = np.linspace(0, n_nodes, world + 1).round().astype(np.int64)
        else:
            bounds = np.asarray(node_ranges, dtype=np.int64)
        self.bounds = bounds
        self.node_begin, self.node_end = int(bounds[rank]), int(bounds[rank + 1])
        owned = (cells >= self.node_begin) & (cells < self.node_end)
        self.local_elements = np.nonzero(owned.any(axis=1))[0]            # global element ids, ascending
        lc = cells[self.local_elements]
        nodes = np.unique(lc)
        ghost = nodes[(nodes < self.node_begin) | (nodes >= self.node_end)]
        self.n_owned = self.node_end - self.node_begin
        self.ghost_global = ghost                                         # sorted
        self.local_to_global = np.concatenate([np.arange(self.node_begin, self.node_end, dtype=np.int64), ghost])
        g2l = -np.ones(self.n_nodes_global, dtype=np.int64)
        g2l[self.local_to_global] = np.arange(self.local_to_global.shape[0])
        self.global_to_local = g2l
        self.local_cells = g2l[lc]
        # element ownership (for element-wise quantities): the rank owning the element's lowest node
        self.element_owner_is_me = (lc.min(axis=1) >= self.node_begin) & (lc.min(axis=1) < self.node_end)
        # ghosts grouped by owner: contiguous segments of the ghost block
        owner = np.searchsorted(bounds, ghost, side='right') - 1
        self.recv = []                                                    # (peer, first local node, count)
        for peer in np.unique(owner):
            idx = np.nonzero(owner == peer)[0]
            self.recv.append((int(peer), int(self.n_owned + idx[0]), int(idx.shape[0])))
        self.send = None                                                  # filled by exchange_halo_lists

    def exchange_halo_lists(self, all_ghosts):
        """all_ghosts[q] = sorted global ghost ids of rank q (an all_gather of `ghost_global`).  Computes, for
        every peer, the owned LOCAL nodes this rank has to send, in the order the peer stores them."""
        self.send = []
        for q, ghosts in enumerate(all_ghosts):
            if q == self.rank:
                continue
            ghosts = np.asarray(ghosts, dtype=np.int64)
            mine = ghosts[(ghosts >= self.node_begin) & (ghosts < self.node_end)]
            if mine.shape[0]:
                self.send.append((q, (mine - self.node_begin).astype(np.int32)))
        peers_s = sorted(p for p, _ in self.send)
        peers_r = sorted(p for p, _, _ in self.recv)
        assert peers_s == peers_r, 'halo graph must be symmetric: %s vs %s' % (peers_s, peers_r)
        return self

    def neighbours(self):
        return sorted(p for p, _, _ in self.recv)

    def gather_all_ghosts(self):
        """all_gather of the ghost lists through torch.distributed (any backend)."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return [self.ghost_global]
        dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
        cnt = torch.tensor([self.ghost_global.shape[0]], dtype=torch.int64, device=dev)
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(self.world)]
        dist.all_gather(counts, cnt)
        m = int(max(int(c.item()) for c in counts))
        pad = torch.full((max(m, 1),), -1, dtype=torch.int64, device=dev)
        pad[:self.ghost_global.shape[0]] = torch.from_numpy(self.ghost_global).to(dev)
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad)
        return [b[:int(c.item())].cpu().numpy() for b, c in zip(bufs, counts)]


class DistributedStatic(object):
    """K U = F on `world` GPUs: local assembly + femol_pcg_dist.  Built from a *global* FEM_Problem description
    (every rank generates the same global mesh on the host; only its strip goes to the device)."""

    def __init__(self, F, global_problem, rank, world, comm_handle=None):
        import torch
        self.F_mod = F
        self.rank, self.world = rank, world
        gp = global_problem
        gm = gp.mesh
        ctype = gm.contains[0]
        self.part = Partition(gm.N_nodes, gm.cells[ctype], rank, world)
        self.part.exchange_halo_lists(self.part.gather_all_ghosts())
        p = self.part
        # local mesh through the public Mesh ctor (unstructured path on the device: per-element matrices are
        # not needed for a structured source mesh, so keep its flag to reuse the single element matrix)
        pts = gm.points[p.local_to_global]
        kw = {'quad_element': gm.ElementClasses['quad']} if ctype == 'quad' else {'tri_element': gm.ElementClasses['triangle']}
        lm = F.Mesh(pts, {ctype: p.local_cells}, structured=gm.structured, **kw)
        assert lm.N_nodes == pts.shape[0], 'every local node is used by a local element'
        self.local_mesh = lm
        lp = F.FEM_Problem(gp.physics, gp.model, lm)
        lp.define_materials(*gp.materials)
        lp.tensors, lp.ho = gp.tensors, gp.ho
        if gp.coating:
            lp.coat_tensors, lp.coat_ho, lp.coating = gp.coat_tensors, gp.coat_ho, True
        lp.fixed_domains, lp.fixed_ddls = list(gp.fixed_domains), list(gp.fixed_ddls)
        # global force vector (node counts of the force domains are global), restricted to the local nodes
        Fg = gp.F.copy()
        Fg[gp.fixed_dofs()] = 0
        nd = gp.N_dof
        idx = (p.local_to_global[:, None] * nd + np.arange(nd)[None]).reshape(-1)
        lp.F = Fg[idx]
        lp.forces = list(gp.forces)
        self.local_problem = lp
        self.N_dof = nd
        self.comm = comm_handle
        self._ws = None
        self._U = None
        self.solver_info = {}
        self.torch = torch

    def local_density(self, X_global):
        return np.asarray(X_global)[self.part.local_elements]

    def assemble(self, X_global=None, p=None):
        lp = self.local_problem
        X = None if X_global is None else {self.local_mesh.contains[0]: self.local_density(X_global)}
        lp.assemble('K', X=X, p=p)

    def solve(self, rtol=1e-8, max_iter=200000, precond='block_jacobi'):
        """Returns the owned part of U (device tensor) after the distributed PCG."""
        torch = self.torch
        L = self.F_mod._lib
        lp, part = self.local_problem, self.part
        d = lp.device
        nd = self.N_dof
        if self._U is None:
            self._U = torch.zeros(d.n, dtype=torch.float64, device=d.dev)
        nbrs = part.neighbours()
        send = dict(part.send)
        recv = {peer: (start, cnt) for peer, start, cnt in part.recv}
        n_nbr = len(nbrs)
        nbr_arr = (ctypes.c_int * max(n_nbr, 1))(*nbrs)
        send_ptr = np.zeros(n_nbr + 1, dtype=np.int64)
        send_nodes = []
        for k, q in enumerate(nbrs):
            send_nodes.append(send[q])
            send_ptr[k + 1] = send_ptr[k] + send[q].shape[0]
        send_nodes = np.concatenate(send_nodes) if send_nodes else np.zeros(0, np.int32)
        d_send = torch.from_numpy(np.ascontiguousarray(send_nodes, dtype=np.int32)).to(d.dev) if send_nodes.size else None
        recv_start = np.array([recv[q][0] for q in nbrs], dtype=np.int64)
        recv_count = np.array([recv[q][1] for q in nbrs], dtype=np.int64)
        if self._ws is None:
            nws = int(L.load().femol_pcg_dist_workspace_doubles(d.plan, int(send_ptr[-1])))
            self._ws = torch.empty(nws, dtype=torch.float64, device=d.dev)
        iters, relres = ctypes.c_int(0), ctypes.c_double(0.0)
        kind = {'none': 0, 'jacobi': 1, 'block_jacobi': 2}[precond]
        rc = L.call('femol_pcg_dist', d.plan, self.comm, part.n_owned, n_nbr, nbr_arr,
                    send_ptr.ctypes.data_as(ctypes.c_void_p), L.ptr(d_send), recv_start.ctypes.data_as(ctypes.c_void_p),
                    recv_count.ctypes.data_as(ctypes.c_void_p), L.ptr(lp._Kv), L.ptr(lp._F_dev), L.ptr(self._U),
                    float(rtol), int(max_iter), kind, L.ptr(self._ws), ctypes.byref(iters), ctypes.byref(relres),
                    L.stream())
        self.solver_info = {'iterations': iters.value, 'relres': relres.value, 'status': rc,
                            'converged': rc in (0, L.S_STAGNATED)}
        return self._U[:part.n_owned * nd]


def create_comm(F):
    """NCCL communicator for femol_pcg_dist: rank 0 creates the unique id, torch.distributed broadcasts it.
    Without an initialised process group: a communicator of one rank."""
    import torch
    import torch.distributed as dist
    L = F._lib
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    buf = (ctypes.c_char * 128)()
    if rank == 0:
        L.call('femol_dist_unique_id', buf)
    raw = bytes(buf)
    if world > 1:
        t = torch.tensor(list(raw), dtype=torch.uint8, device='cuda')
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().tolist())
    cbuf = ctypes.create_string_buffer(raw, 128)
    handle = ctypes.c_void_p()
    L.call('femol_dist_init', cbuf, rank, world, ctypes.byref(handle))
    return handle


# =====================================================================================================================
# Strip partition with multigrid: ONE large SIMP_COMP problem on the GPUs of a box (BASELINE config 5)
# =====================================================================================================================
def _lattice_of(problem):
    """(Nx, Ny, lat_node or None) of the problem's mesh: lattice cells per direction and lattice index -> node id
    (None = identity, the rectangle_Q4 ordering)."""
    shape = problem.structured_grid_shape()
    if shape:
        return shape[0], shape[1], None
    lat = problem.node_lattice()
    if lat is None:
        raise ValueError('the strip partition needs a mesh whose nodes form a lattice (rectangle_Q4 / T3 / T6)')
    return lat['Nx'], lat['Ny'], lat['lat_node'].astype(np.int64)


def coarsen_mask_host(mask, nx, ny):
    """Node mask of the next-coarser grid: coarse node (I, J) sits on fine node (min(2I, nx), min(2J, ny))
    (csrc/gmg.cu coarsen_mask).  mask: [(ny+1)(nx+1)] lattice order."""
    nxc, nyc = (nx + 1) // 2, (ny + 1) // 2
    m = np.asarray(mask).reshape(ny + 1, nx + 1)
    ii = np.minimum(2 * np.arange(nxc + 1), nx)
    jj = np.minimum(2 * np.arange(nyc + 1), ny)
    return np.ascontiguousarray(m[jj][:, ii]).reshape(-1), nxc, nyc


def _node_rows(Nx, lat_node, n_nodes):
    """(lattice index of every node or None = identity, lattice row of every node)"""
    if lat_node is None:
        return None, np.arange(n_nodes, dtype=np.int64) // (Nx + 1)
    node_lat = np.empty(n_nodes, dtype=np.int64)
    node_lat[lat_node] = np.arange(n_nodes, dtype=np.int64)
    return node_lat, node_lat // (Nx + 1)


def strip_halo_rows(cells, Nx, lat_node, n_nodes):
    """Lattice rows an element spans (1: Q4 / T3 on their corner lattice, 2: T6) = rows a strip needs from below."""
    _, node_row = _node_rows(Nx, lat_node, n_nodes)
    r = node_row[np.asarray(cells)]
    return int((r.max(axis=1) - r.min(axis=1)).max())


def strip_local_mesh(cells, Nx, lat_node, n_nodes, part):
    """Cut the strip of `part` out of a lattice mesh: every element whose nodes lie on the local rows [row0, top], nodes
    renumbered in lattice order (local id = lattice index - row0 * (Nx + 1)).  Returns dict(elements: global element
    ids (ascending), nodes: global node id of every local node, cells: local connectivity, owned: the elements whose
    lowest row this rank owns - each element has exactly one owner).  Checks that every element with a node on an owned
    row is local and stays inside the rows a halo exchange fills."""
    cells = np.asarray(cells)
    npx = Nx + 1
    node_lat, node_row = _node_rows(Nx, lat_node, n_nodes)
    cell_rows = node_row[cells]
    rmin_e, rmax_e = cell_rows.min(axis=1), cell_rows.max(axis=1)
    sel = (rmin_e >= part.row0) & (rmax_e <= part.top)
    elements = np.nonzero(sel)[0]
    lat0 = part.row0 * npx
    g_nodes = np.arange(lat0, lat0 + (part.top - part.row0 + 1) * npx, dtype=np.int64)
    if lat_node is not None:
        g_nodes = np.asarray(lat_node, dtype=np.int64)[g_nodes]
    lc = cells[elements]
    lc = (lc if node_lat is None else node_lat[lc]) - lat0
    own = (rmin_e[sel] >= part.own_begin) & (rmin_e[sel] < part.own_end)
    lev0 = part.levels[0]
    touch = (rmax_e >= part.own_begin) & (rmin_e <= part.own_end - 1)
    lo_ok = part.own_begin - (lev0['h_up'] if part.rank > 0 else 0)
    hi_ok = part.own_end - 1 + (lev0['h_down'] if part.rank < part.world - 1 else 0)
    if not (sel[touch].all() and (rmin_e[touch] >= lo_ok).all() and (rmax_e[touch] <= hi_ok).all()):
        raise ValueError('an element of an owned row reaches beyond the exchanged halo rows')
    return {'elements': elements, 'nodes': g_nodes, 'cells': lc, 'owned': own}


class StripPartition(object):
    """Strips of lattice rows for `world` ranks with `n_dist` distributed multigrid levels, seen from `rank`.

    Node rows 0 .. Ny; rank r owns the rows [cuts[r], cuts[r+1]); interior cuts are multiples of A = 2^n_dist, so
    that the owned rows of level l are [cuts[r] >> l, cuts[r+1] >> l) and a local grid that starts A rows below the
    first owned row (and ends on the first row of the upper neighbour) coarsens exactly like the global grid:
    level l of the strip = global rows [row0 >> l, top >> l] of global level l."""

    def __init__(self, Ny, rank, world, n_dist, h_up0=1):
        A = 1 << n_dist
        self.Ny, self.rank, self.world, self.n_dist, self.A = int(Ny), rank, world, n_dist, A
        cuts = [0] + [int(round(r * (Ny + 1) / float(world) / A)) * A for r in range(1, world)] + [Ny + 1]
        self.cuts = cuts
        if any(cuts[r + 1] - cuts[r] < max(A, 2) for r in range(world)):
            raise ValueError('strips too thin: %d rows, %d ranks, %d distributed levels' % (Ny + 1, world, n_dist))
        self.row0 = max(0, cuts[rank] - A) if rank > 0 else 0
        self.top = cuts[rank + 1] if rank < world - 1 else Ny        # last local row (inclusive)
        self.own_begin, self.own_end = cuts[rank], cuts[rank + 1]
        ny_l = Ny
        self.levels = []                                              # per level: dict(Ny, row0, top, own_begin, own_end)
        for l in range(n_dist + 1):
            last = rank == world - 1
            lev = {'Ny': ny_l, 'row0': self.row0 >> l, 'own_begin': cuts[rank] >> l,
                   'own_end': (ny_l + 1) if last else (cuts[rank + 1] >> l)}
            lev['top'] = ny_l if last else lev['own_end']
            lev['h_up'] = h_up0 if l == 0 else 1
            lev['h_down'] = 1
            self.levels.append(lev)
            ny_l = (ny_l + 1) // 2

    def local_ny(self, l=0):
        return self.levels[l]['top'] - self.levels[l]['row0']

    @staticmethod
    def pick_levels(Nx, Ny, world, gather_nodes=40000, max_levels=8):
        """Distributed levels: coarsen until the gathered level has at most `gather_nodes` nodes, keeping the strips
        at least four alignment units high."""
        n = 1
        while n < max_levels:
            nx_l, ny_l = Nx, Ny
            for _ in range(n):
                nx_l, ny_l = (nx_l + 1) // 2, (ny_l + 1) // 2
            if (nx_l + 1) * (ny_l + 1) <= gather_nodes:
                break
            if (Ny + 1) // world < 4 * (1 << (n + 1)):
                break
            n += 1
        return n


def _gather_objects(obj, world):
    if world == 1:
        return [obj]
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def owned_gather_index(own_global, n_total, world):
    """Every rank contributes its owned items (global ids `own_global`) in a buffer padded to the largest count; the
    all-gathered buffers are scattered to `index` (padding goes to the dummy slot n_total).  Returns (padded count,
    index [world * padded count])."""
    own_global = np.asarray(own_global, dtype=np.int64)
    counts = _gather_objects(int(own_global.shape[0]), world)
    assert sum(counts) == n_total, 'every item has exactly one owner'
    m = max(counts)
    pad = np.full(m, n_total, dtype=np.int64)
    pad[:own_global.shape[0]] = own_global
    return m, np.concatenate(_gather_objects(pad, world))


def attach_p2p(F, comm, gmg, world):
    """Open this rank's peer-memory window, swap the CUDA IPC handles through torch.distributed and map the peers'
    windows (femol_dist_p2p_open / _attach).  Returns 'p2p', or 'nccl' when some rank could not map a peer (the
    library then keeps using NCCL send/recv; all ranks take the same decision)."""
    import torch
    import torch.distributed as dist
    L = F._lib
    sizes = (ctypes.c_int64 * 2)()
    L.call('femol_gmg_dist_p2p_sizes', gmg, sizes)
    handle = (ctypes.c_char * 64)()
    ok = 1
    try:
        L.call('femol_dist_p2p_open', comm, int(sizes[0]), int(sizes[1]), handle)
    except L.FemolError:
        ok = 0
    handles = _gather_objects((ok, bytes(handle)), world)
    if not all(h[0] for h in handles):
        return 'nccl'
    blob = ctypes.create_string_buffer(b''.join(h[1] for h in handles), 64 * world)
    try:
        L.call('femol_dist_p2p_attach', comm, blob)
    except L.FemolError:
        ok = 0
    oks = _gather_objects(ok, world)
    torch.cuda.synchronize()
    dist.barrier()
    if not any(oks):
        return 'nccl'           # no rank could map its peers (CUDA IPC unavailable): the library keeps using NCCL
    if not all(oks):
        raise L.FemolError('peer-memory transport: some ranks attached and some did not; set FEMOL_DIST_TRANSPORT=nccl')
    return 'p2p'


def _dev_i64(values):
    return (ctypes.c_int64 * len(values))(*[int(v) for v in values])


class DistributedSIMP_COMP(object):
    """SIMP_COMP (TOM_SIMP.py:13-81) on `world` GPUs for one large lattice mesh.

    Every rank builds the same GLOBAL problem description on the host (mesh, domains, loads - never on the device),
    cuts its strip out of it as an ordinary local FEM_Problem, and keeps the global density / sensitivity vectors
    replicated on its GPU: assembly, multigrid set-up and the PCG run on the strip (halo exchanges + scalar
    all-reduces inside the solve); the sensitivities of the owned elements are all-gathered, and the filter and the
    OC update then run redundantly on every rank over the whole element set (they cost < 1 ms at 4 M elements and need
    neighbours / a global volume constraint), so all ranks hold bit-identical densities without a broadcast."""

    def __init__(self, F, global_problem, comm, volfrac=0.5, penal=3, rmin=1.5, n_dist=None, rank=None, world=None):
        import torch
        import torch.distributed as dist
        from . import TOM_SIMP
        self.F_mod, self.torch, self.comm = F, torch, comm
        if rank is None or world is None:
            rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
        self.rank, self.world = rank, world
        gp = self.gp = global_problem
        gm = gp.mesh
        if gp.coating:
            raise NotImplementedError('DistributedSIMP_COMP: coating formulation not distributed yet')
        L = F._lib
        nd = self.N_dof = gp.N_dof
        ctype = self.ctype = gm.contains[0]
        Nx, Ny, lat_node = _lattice_of(gp)
        self.Nx, self.Ny = Nx, Ny
        cells = np.asarray(gm.cells[ctype])
        h_up0 = strip_halo_rows(cells, Nx, lat_node, gm.N_nodes)
        if n_dist is None:
            n_dist = StripPartition.pick_levels(Nx, Ny, self.world)
        part = self.part = StripPartition(Ny, self.rank, self.world, n_dist, h_up0=h_up0)
        cut = strip_local_mesh(cells, Nx, lat_node, gm.N_nodes, part)
        self.local_elements, self.local_to_global = cut['elements'], cut['nodes']
        g_nodes, lc, own = cut['nodes'], cut['cells'], cut['owned']
        n_loc_nodes = g_nodes.shape[0]
        kw = {'quad_element': gm.ElementClasses['quad']} if ctype == 'quad' else {'tri_element': gm.ElementClasses['triangle']}
        lm = self.local_mesh = F.Mesh(gm.points[g_nodes], {ctype: lc}, structured=gm.structured, **kw)
        assert lm.N_nodes == n_loc_nodes, 'every local lattice node is used by a local element'
        lp = self.local_problem = F.FEM_Problem(gp.physics, gp.model, lm)
        lp.define_materials(*gp.materials)
        lp.tensors, lp.ho = gp.tensors, gp.ho
        lp.fixed_domains, lp.fixed_ddls = list(gp.fixed_domains), list(gp.fixed_ddls)
        Fg = np.array(gp.F, dtype=np.float64)
        lp.F = Fg[(g_nodes[:, None] * nd + np.arange(nd)[None]).reshape(-1)]
        lp.forces, lp.force_domains = list(gp.forces), list(gp.force_domains)
        lp.preconditioner, lp.gmg_hourglass = 'gmg', False
        lp.solver_rtol, lp.solver_max_iter = gp.solver_rtol, gp.solver_max_iter
        lp.gmg_smoothing_steps, lp.gmg_omega = 1, gp.gmg_omega
        lp.on_nonconvergence = gp.on_nonconvergence
        d = lp.device
        dev = d.dev
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
        # ---- replicated SIMP state over ALL elements ---------------------------------------------------------
        self.f, self.p, self.rmin = volfrac, penal, rmin
        self.n_elem = cells.shape[0]
        self._x = torch.full((self.n_elem,), float(volfrac), dtype=torch.float64, device=dev)
        self._xn = torch.empty_like(self._x)
        self._dc = torch.zeros_like(self._x)
        self._dcf = torch.empty_like(self._x)
        self._scal = torch.zeros(8, dtype=torch.float64, device=dev)
        self._red_ws = torch.zeros(int(L.load().femol_reduce_workspace_doubles()), dtype=torch.float64, device=dev)
        gm.compute_element_centers()
        self._filter = TOM_SIMP.build_filter(gm, rmin, dev)
        self._mask, self._mask_key = None, (None, None)
        self._loc_elems = up(self.local_elements, np.int64)
        own_local = np.nonzero(own)[0]
        self._own_local = up(own_local, np.int64)
        own_global = self.local_elements[own_local]
        self._own_max, index = owned_gather_index(own_global, self.n_elem, self.world)
        self._gather_index = up(index, np.int64)
        self._send = torch.zeros(self._own_max, dtype=torch.float64, device=dev)
        self._recv = torch.empty(self._own_max * self.world, dtype=torch.float64, device=dev)
        self._dc_pad = torch.zeros(self.n_elem + 1, dtype=torch.float64, device=dev)
        self._x_loc = torch.empty(d.n_elem, dtype=torch.float64, device=dev)
        self._dc_loc = torch.empty(d.n_elem, dtype=torch.float64, device=dev)
        self._scal_loc = torch.zeros(8, dtype=torch.float64, device=dev)
        # ---- hierarchy: local strip levels + the replicated coarse hierarchy ----------------------------------
        self._U = torch.zeros(d.n, dtype=torch.float64, device=dev)
        self._ws = None
        self._gmg = None
        self._coarse = None
        self.solver_info = {}
        self.iteration_log = []
        self.loop, self.change, self.c = 0, 1.0, None
        self.debug_timing, self.timing, self._t_last = False, {}, 0.0
        self._global_mask_lat = self._global_fixed_mask(lat_node)

    # ---- set-up helpers ---------------------------------------------------------------------------------------
    def _global_fixed_mask(self, lat_node):
        gp, gm = self.gp, self.gp.mesh
        mask = np.zeros(gm.N_nodes, dtype=np.uint8)
        for domain, ddls in zip(gp.fixed_domains, gp.fixed_ddls):
            mask[gm.domain_nodes(domain)] |= np.uint8(sum(1 << int(k) for k in ddls))
        return mask if lat_node is None else mask[lat_node]

    def _build_hierarchy(self):
        """Local hierarchy (created and first refreshed by the local problem), truncated to n_dist + 1 levels, plus the
        replicated lattice hierarchy over the global level-n_dist grid."""
        torch, F, L = self.torch, self.F_mod, self.F_mod._lib
        lp, part, nd = self.local_problem, self.part, self.N_dof
        d = lp.device
        g = lp._gmg_ready()
        info = (ctypes.c_int64 * 64)()
        L.call('femol_gmg_info', g, info)
        if int(info[0]) < part.n_dist + 1:
            raise ValueError('the local hierarchy has %d levels, %d needed' % (int(info[0]), part.n_dist + 1))
        for l, lev in enumerate(part.levels):
            assert int(info[2 + 2 * l]) == lev['top'] - lev['row0'], 'local level %d does not match the partition' % l
        L.call('femol_gmg_truncate', g, part.n_dist + 1)
        # global grid of level n_dist and its node mask
        mask, nx, ny = self._global_mask_lat, self.Nx, self.Ny
        for _ in range(part.n_dist):
            mask, nx, ny = coarsen_mask_host(mask, nx, ny)
        assert ny == part.levels[-1]['Ny'] and nx == int(info[1 + 2 * part.n_dist])
        a = np.add.outer(np.arange(nx), (nx + 1) * np.arange(ny)).reshape(-1)           # e = I * ny + J
        conn = np.stack([a, a + 1, a + nx + 2, a + nx + 1], axis=1).astype(np.int32)
        up = lambda arr, dt: torch.from_numpy(np.ascontiguousarray(arr, dtype=dt)).to(d.dev)
        self._c_conn = up(conn, np.int32)
        n_nodes = (nx + 1) * (ny + 1)
        cplan = ctypes.c_void_p()
        L.call('femol_plan_create', nx * ny, 4, nd, n_nodes, L.ptr(self._c_conn), L.stream(), ctypes.byref(cplan))
        self._c_plan = cplan
        self._c_mask = up(mask, np.uint8)
        nxc, nyc = (nx + 1) // 2, (ny + 1) // 2
        e = np.arange(nx * ny)
        parent = (e // ny // 2) * nyc + (e % ny) // 2
        order = np.argsort(parent, kind='stable')
        ptr = np.concatenate([[0], np.cumsum(np.bincount(parent, minlength=nxc * nyc))])
        ident = np.arange(n_nodes, dtype=np.int32)
        self._c_lat = [up(ident, np.int32), up(ident, np.int32), up(ptr, np.int32), up(order, np.int32)]
        ch = ctypes.c_void_p()
        L.call('femol_gmg_create_lattice', nx, ny, nd, cplan, L.ptr(self._c_mask), L.ptr(self._c_lat[0]), L.ptr(self._c_lat[1]),
               L.ptr(self._c_lat[2]), L.ptr(self._c_lat[3]), 1, float(lp.gmg_omega), L.stream(), ctypes.byref(ch))
        torch.cuda.current_stream().synchronize()
        self._coarse = ch
        lv = part.levels
        L.call('femol_gmg_set_dist', g, self.comm, part.n_dist, _dev_i64([v['row0'] for v in lv]),
               _dev_i64([v['own_begin'] for v in lv]), _dev_i64([v['own_end'] for v in lv]),
               (ctypes.c_int * len(lv))(*[v['h_up'] for v in lv]), (ctypes.c_int * len(lv))(*[v['h_down'] for v in lv]),
               ch, cplan, L.ptr(self._c_mask))
        self._gmg = g
        self.transport = 'nccl'
        if self.world > 1 and os.environ.get('FEMOL_DIST_TRANSPORT', 'p2p') != 'nccl':
            self.transport = attach_p2p(F, self.comm, g, self.world)

    def __del__(self):
        try:
            L = self.F_mod._lib.load()
            self.local_problem._destroy_gmg()
            if self._coarse is not None:
                L.femol_gmg_destroy(self._coarse)
                L.femol_plan_destroy(self._c_plan)
        except Exception:
            pass

    # ---- one SIMP iteration ----------------------------------------------------------------------------------------
    def _tick(self, name):
        """debug_timing: host wall time per section, with a device synchronisation at every boundary"""
        if self.debug_timing:
            import time
            if self.debug_timing != 'host':
                self.torch.cuda.synchronize()
            now = time.perf_counter()
            self.timing[name] = self.timing.get(name, 0.0) + now - self._t_last
            self._t_last = now

    def solve_displacement(self):
        """Assemble K(x^p) on the strip, refresh both hierarchies, distributed multigrid PCG -> local U (device)."""
        torch, L = self.torch, self.F_mod._lib
        lp = self.local_problem
        d = lp.device
        if self.debug_timing:
            import time
            torch.cuda.synchronize()
            self._t_last = time.perf_counter()
        torch.index_select(self._x, 0, self._loc_elems, out=self._x_loc)
        self._tick('select')
        lp.assemble('K', X=self._x_loc, p=self.p)
        self._tick('assemble')
        if self._gmg is None:
            self._build_hierarchy()
        else:
            lp._gmg_ready()
        self._tick('gmg_local')
        L.call('femol_gmg_dist_refresh', self._gmg, L.stream())
        self._tick('gmg_refresh')
        if self._ws is None:
            self._ws = torch.empty(int(L.load().femol_pcg_gmg_dist_workspace_doubles(d.plan)), dtype=torch.float64, device=d.dev)
        iters, relres = ctypes.c_int(0), ctypes.c_double(0.0)
        rc = L.call('femol_pcg_gmg_dist', d.plan, self._gmg, L.ptr(lp._Kv), L.ptr(lp._F_dev), L.ptr(self._U),
                    float(lp.solver_rtol), int(lp.solver_max_iter), L.ptr(self._ws), ctypes.byref(iters), ctypes.byref(relres),
                    L.stream())
        self._tick('pcg')
        converged = rc == 0 or (rc == L.S_STAGNATED and relres.value <= lp.stagnation_bound)
        self.solver_info = {'iterations': iters.value, 'relres': relres.value, 'status': rc, 'converged': converged,
                            'preconditioner': 'gmg-dist'}
        if rc == L.S_BREAKDOWN:
            raise L.FemolError('distributed PCG breakdown after %d iterations' % iters.value)
        if not converged:
            lp._report_nonconvergence('distributed PCG did not converge: %d iterations, relative residual %.3e (rtol %.1e)'
                                      % (iters.value, relres.value, lp.solver_rtol))
        return self._U

    def solve_static(self, X=None, p=1.0):
        """K(X^p) U = F on the strip partition - the distributed counterpart of `FEM_Problem.assemble('K', X, p)` +
        `solve()` (FEM.py:160-214).  X: global densities (array over ALL elements, identical on every rank) or None (solid).
        Returns (global node ids of the owned rows, their displacements [n_owned_nodes, N_dof]); `solver_info` as usual."""
        torch = self.torch
        keep_x, keep_p = self._x, self.p
        try:
            if X is None:
                self._x = torch.ones_like(keep_x)
            else:
                self._x = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64)).to(keep_x.device)
            self.p = p
            self.solve_displacement()
        finally:
            self._x, self.p = keep_x, keep_p
        return self.owned_displacement()

    def iterate(self):
        torch, L = self.torch, self.F_mod._lib
        import torch.distributed as dist
        from . import TOM_SIMP
        lp = self.local_problem
        d = lp.device
        self.solve_displacement()
        Eb, pe_b = lp._element_matrices('K', False)
        pc = self.p if lp.mesh.structured else 2 * self.p               # FEM.py:630-631 double penalisation quirk
        L.call('femol_sens_compliance', d.plan, L.ptr(Eb), pe_b, None, 0, L.ptr(self._U), L.ptr(self._x_loc), float(self.p),
               float(pc), float(pc - 1), L.ptr(self._dc_loc), L.ptr(self._scal_loc), L.ptr(d.red_ws), L.stream())
        # owned elements: sensitivities -> all ranks; compliance c = sum_e x^pc q_e = -sum_e x_e dc_e / p
        dc_own = self._dc_loc[self._own_local]
        n_own = dc_own.shape[0]
        self._send[:n_own] = dc_own
        c_loc = -(dc_own * self._x_loc[self._own_local]).sum() / float(self.p)
        self._tick('sens')
        if self.world > 1:
            dist.all_gather_into_tensor(self._recv, self._send)
            dist.all_reduce(c_loc)
        else:
            self._recv.copy_(self._send)
        self._dc_pad[self._gather_index] = self._recv
        self._scal[0] = c_loc
        self._tick('gather')
        TOM_SIMP.apply_filter(self._filter, self._x, self._dc_pad, self._dcf)
        key = (getattr(self, 'solid_domain', None), getattr(self, 'void_domain', None))     # TOM_SIMP.py:299-303, 315-331
        if key != self._mask_key:
            self._mask = TOM_SIMP.domain_mask(self.gp.mesh, key[0], key[1], d.dev)
            self._mask_key = key
        L.call('femol_oc_update', self.n_elem, L.ptr(self._x), L.ptr(self._dcf), 0, float(self.f), None, L.ptr(self._mask),
               L.ptr(self._xn), self._scal.data_ptr() + 4 * 8, L.ptr(self._red_ws), L.stream())
        self._x, self._xn = self._xn, self._x
        s = self._scal.cpu().numpy()
        self._tick('filter_oc')
        self.c, self.change = float(s[0]), float(s[4])
        return self.change, self.c

    def solve(self, converge=0.01, max_iter=100):
        self.change, self.loop = 1.0, 0
        while self.change > converge and self.loop < max_iter:
            self.loop += 1
            self.iterate()
            info = dict(self.solver_info)
            info.update(loop=self.loop, change=self.change, c=self.c)
            self.iteration_log.append(info)
        return self.density()

    def density(self):
        """Global densities (host array; identical on every rank)."""
        return self._x.cpu().numpy()

    def owned_displacement(self):
        """(global node ids of the owned rows, their displacements [n_owned_nodes, N_dof]) as host arrays."""
        part, npx, nd = self.part, self.Nx + 1, self.N_dof
        lo, hi = (part.own_begin - part.row0) * npx, (part.own_end - part.row0) * npx
        U = self._U.view(-1, nd)[lo:hi].cpu().numpy()
        return self.local_to_global[lo:hi], U
