"""Row-block partitioned static solve over N GPUs (torchrun), cfg-5-like T6 plane-stress cantilever.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_dist.py [n]          # n x n x 2 T6 elements

Every rank builds the global mesh description on the host, keeps its strip on the device (femol_b200/dist.py),
assembles locally and runs femol_pcg_dist (block-Jacobi PCG, halo exchange + in-stream all-reduces over NCCL).
Rank 0 prints one JSON line: iterations, relative residual, MAX-over-ranks device time of the solve.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import femol_b200 as F  # noqa: E402
from femol_b200.dist import DistributedStatic, create_comm  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    t0 = time.perf_counter()
    mesh = F.mesh.rectangle_T6(float(n), float(n), n, n)
    gp = F.FEM_Problem('displacement', 'plane', mesh)
    gp.define_materials(F.materials.general_isotropic()); gp.define_tensors(1)
    gp.add_fixed_domain(F.domains.inside_box([0], [[0, float(n)]]), ddls=[0, 1])
    gp.add_forces([0, -1.0], F.domains.inside_box([float(n)], [[0, float(n)]]))
    comm = create_comm(F)
    ds = DistributedStatic(F, gp, rank, world, comm)
    x = np.random.default_rng(0).uniform(0.3, 1.0, mesh.N_ele)
    ds.assemble(X_global=x, p=3)
    torch.cuda.synchronize(); dist.barrier()
    t_setup = time.perf_counter() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    U = ds.solve(rtol=1e-8)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    owned = torch.tensor([float(ds.part.n_owned)], dtype=torch.float64, device='cuda')
    dist.all_reduce(owned)
    if rank == 0:
        print(json.dumps({'config': 'T6 plane static %dx%d, row-block partition over %d GPU(s)' % (n, n, world),
                          'ndof': 2 * mesh.N_nodes, 'n_elem': mesh.N_ele, 'owned_nodes_total': int(owned.item()),
                          'setup_s (host mesh + partition + local assembly)': t_setup, 'solve_s': float(t.item()),
                          'pcg': ds.solver_info}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
