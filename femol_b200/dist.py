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

import numpy as np

__all__ = ['shard_items', 'Partition', 'DistributedStatic', 'init_process_group']


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
    """NCCL communicator for femol_pcg_dist: rank 0 creates the unique id, torch.distributed broadcasts it."""
    import torch
    import torch.distributed as dist
    L = F._lib
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = (ctypes.c_char * 128)()
    if rank == 0:
        L.call('femol_dist_unique_id', buf)
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device='cuda')
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    cbuf = ctypes.create_string_buffer(raw, 128)
    handle = ctypes.c_void_p()
    L.call('femol_dist_init', cbuf, rank, world, ctypes.byref(handle))
    return handle
