"""Multigrid-CG iteration counts on uniform plates: size, boundary conditions and load (diagnostic)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import femol_b200 as F  # noqa: E402


def run(n, bc, load, model='plate', t=0.1):
    mesh = F.mesh.rectangle_Q4(1.0, 1.0, n, n)
    p = F.FEM_Problem('displacement', model, mesh)
    p.define_materials(F.materials.isotropic_bending_benchmark()); p.define_tensors(t)
    box = F.domains.inside_box
    nd = 6 if model == 'plate' else 2
    if bc == 'cantilever':
        p.add_fixed_domain(box([0], [[0, 1.]]))
    elif bc == 'clamped_all':
        p.add_fixed_domain(box([0, 1.], [[0, 1.]])); p.add_fixed_domain(box([[0, 1.]], [0, 1.]))
    elif bc == 'ss_free_inplane':      # simply supported bending BCs, in-plane + drilling dofs fixed on the boundary only
        p.add_fixed_domain(box([0, 1.], [[0, 1.]]), ddls=[0, 1, 5, 2, 4]); p.add_fixed_domain(box([[0, 1.]], [0, 1.]), ddls=[0, 1, 5, 2, 3])
    elif bc == 'cantilever_no_inplane':  # cantilever, in-plane + drilling dofs fixed everywhere
        p.add_fixed_domain(box([[0, 1.]], [[0, 1.]]), ddls=[0, 1, 5]); p.add_fixed_domain(box([0], [[0, 1.]]))
    elif bc == 'ss':
        p.add_fixed_domain(box([[0, 1.]], [[0, 1.]]), ddls=[0, 1, 5])
        p.add_fixed_domain(box([0, 1.], [[0, 1.]]), ddls=[2, 4])
        p.add_fixed_domain(box([[0, 1.]], [0, 1.]), ddls=[2, 3])
    f = [0, 0, -1.0, 0, 0, 0] if model == 'plate' else [0, -1.0]
    dom = box([1.], [[0.45, 0.55]]) if load == 'tip' else box([[0.001, 0.999]], [[0.001, 0.999]])
    p.add_forces(f, dom)
    p.solver_rtol = 1e-8
    p.solver_max_iter = int(os.environ.get('MAXIT', '200000'))
    if os.environ.get('NU'):
        p.gmg_smoothing_steps = int(os.environ['NU'])
    if os.environ.get('HOURGLASS'):
        p.gmg_hourglass = {'0': False, '1': True}.get(os.environ['HOURGLASS'], 'auto')
    p.solve_device()
    print('%-6s n=%4d %-12s %-8s t=%g: %4d CG iterations (relres %.1e)' % (model, n, bc, load, t, p.solver_info['iterations'], p.solver_info['relres']), flush=True)


if len(sys.argv) > 1 and sys.argv[1] == 'thickness':
    for bc in ('ss', 'cantilever'):
        for t in (0.02, 0.1, 0.5, 2.0):
            run(200, bc, 'uniform', t=t)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == 'sizes':
    for n in (128, 160, 192, 200, 256):
        run(n, 'cantilever', 'uniform', t=0.5)
    for n in (128, 200, 256):
        run(n, 'cantilever', 'tip', model='plane')
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == 'bcs':
    for bc in ('ss', 'ss_free_inplane', 'clamped_all', 'cantilever', 'cantilever_no_inplane'):
        run(200, bc, 'uniform', t=0.5)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == 'supported':
    for bc in ('ss', 'clamped_all'):
        for t in (0.02, 0.1, 0.5):
            run(200, bc, 'uniform', t=t)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == 'thin':
    for t in (0.001, 0.0025, 0.005, 0.01, 0.02, 0.05):
        run(200, 'cantilever', 'uniform', t=t)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == 'verythin':
    for t in (0.001, 0.0025, 0.01):
        run(200, 'cantilever', 'uniform', t=t)
        run(200, 'ss', 'uniform', t=t)
    sys.exit(0)
for n in (50, 100, 200, 400):
    run(n, 'cantilever', 'tip')
run(200, 'cantilever', 'uniform')
run(200, 'ss', 'uniform')
run(200, 'ss', 'tip')
run(200, 'clamped_all', 'uniform')
run(200, 'cantilever', 'tip', model='plane')
