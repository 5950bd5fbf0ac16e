import sys, time, os
sys.path.insert(0, '/root/repo')
import torch
import femol_b200 as F
import bench
problem, opt = bench.build_gpu_problem(F, 2000, 1000, 1e-8, 'auto')
for _ in range(3): opt.iterate()
torch.cuda.synchronize()
n_elem = problem.device.n_elem
x_host = torch.empty(n_elem, dtype=torch.float64).pin_memory(); x_host.copy_(opt.x_device)
out_host = torch.empty(n_elem, dtype=torch.float64).pin_memory()
torch.cuda.synchronize()
for k in range(6):
    t0 = time.perf_counter()
    opt.x_device.copy_(x_host, non_blocking=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    opt.iterate()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    out_host.copy_(opt.x_device, non_blocking=True)
    torch.cuda.current_stream().synchronize(); t3 = time.perf_counter()
    x_host, out_host = out_host, x_host
    print('step %d: h2d %.2f ms, iterate %.2f ms (pcg %d), d2h %.2f ms' % (k, (t1-t0)*1e3, (t2-t1)*1e3, problem.solver_info['iterations'], (t3-t2)*1e3))
for k in range(4):
    torch.cuda.synchronize(); t1 = time.perf_counter()
    opt.iterate()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print('resident step: %.2f ms (pcg %d)' % ((t2-t1)*1e3, problem.solver_info['iterations']))
