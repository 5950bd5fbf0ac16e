"""Short driver for ncu: a few launches of every hot-path kernel at the cfg-2 size (2000 x 1000 Q4 plane).

    python tools/profile_kernels.py [nx ny] [pcg_iters]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import femol_b200 as F  # noqa: E402
import bench  # noqa: E402


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    problem, opt = bench.build_gpu_problem(F, nx, ny, 1e-8, sys.argv[4] if len(sys.argv) > 4 else 'auto')
    problem.solver_max_iter = iters
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    opt.x_device.copy_(torch.rand(problem.device.n_elem, dtype=torch.float64, device='cuda', generator=g) * 0.999 + 0.001)
    for _ in range(3):
        opt.iterate()          # assemble, PCG (bounded), sensitivities, filter, OC
    torch.cuda.synchronize()
    print('ok', problem.solver_info, F._lib.load().femol_launch_count(0))


if __name__ == '__main__':
    main()
