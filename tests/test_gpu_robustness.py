"""Behaviour around the solvers that the reference never needed (it uses direct solvers): non-convergence is
reported, a zero right-hand side overrides the warm start, CUDA-graph capture really engages from the default
stream, and the boundary-condition / load caches notice in-place edits.  Needs a B200: `-m gpu`."""
import warnings

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def F():
    import femol_b200
    femol_b200._lib.require_gpu()
    return femol_b200


def _cantilever(F, nx=48, ny=24, model='plane'):
    mesh = F.mesh.rectangle_Q4(float(nx), float(ny), nx, ny)
    p = F.FEM_Problem('displacement', model, mesh)
    p.define_materials(F.materials.general_isotropic()); p.define_tensors(1)
    p.add_fixed_domain(F.domains.inside_box([0], [[0, float(ny)]]))
    p.add_forces([0, -1.0] if model == 'plane' else [0, 0, -1.0, 0, 0, 0], F.domains.inside_box([float(nx)], [[0, float(ny)]]))
    return p


@pytest.mark.parametrize('precond', ['gmg', 'block_jacobi'])
def test_iteration_limit_is_reported(F, precond):
    """max_iter exhausted: status 1 -> FemolConvergenceWarning (default), FemolError ('raise'), silence ('ignore');
    solver_info says converged = False either way and the last iterate is still returned."""
    p = _cantilever(F)
    p.preconditioner = precond
    p.solver_max_iter = 2
    p.assemble('K')
    with pytest.warns(F._lib.FemolConvergenceWarning, match='iteration limit'):
        U = p.solve_device()
    assert p.solver_info['status'] == F._lib.S_NOT_CONVERGED and not p.solver_info['converged']
    assert p.solver_info['relres'] > p.solver_rtol and bool(np.isfinite(U.cpu().numpy()).all())
    p.on_nonconvergence = 'raise'
    with pytest.raises(F._lib.FemolError, match='did not converge'):
        p.solve_device()
    p.on_nonconvergence = 'ignore'
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        p.solve_device()
    # with room to iterate the same problem converges and nothing is reported
    p.on_nonconvergence = 'raise'
    p.solver_max_iter = 20000
    p.solve_device()
    assert p.solver_info['converged'] and p.solver_info['status'] == 0


def test_simp_loop_counts_unconverged_solves(F):
    p = _cantilever(F, 30, 15)
    p.solver_max_iter = 1
    opt = F.SIMP_COMP(p, volfrac=0.4, penal=3, rmin=1.5)
    with pytest.warns(F._lib.FemolConvergenceWarning):
        opt.solve(converge=0.0, max_iter=3, plot=True, save=False)
    assert opt.unconverged_solves == 3 and all(not it['converged'] for it in opt.iteration_log)
    p.solver_max_iter = 20000
    opt2 = F.SIMP_COMP(p, volfrac=0.4, penal=3, rmin=1.5)
    opt2.solve(converge=0.0, max_iter=2, plot=True, save=False)
    assert opt2.unconverged_solves == 0


def test_modal_iteration_limit_is_reported(F):
    case = cases.CASES['modal_q4_plane_10x5']
    p = cases.build_problem(F, case)
    p.n_modes = 4
    p.modal_max_iter = 1
    with pytest.warns(F._lib.FemolConvergenceWarning, match='LOBPCG'):
        p.solve(verbose=False)
    assert not p.solver_info['converged']
    p.modal_max_iter = 500
    p._X_modal = None
    with warnings.catch_warnings():
        warnings.simplefilter('error', F._lib.FemolConvergenceWarning)
        p.solve(verbose=False)
    assert p.solver_info['converged']


@pytest.mark.parametrize('precond', ['gmg', 'block_jacobi'])
def test_zero_rhs_overrides_warm_start(F, precond):
    """K U = 0 has the solution U = 0 even when the previous displacement is kept as the start vector."""
    p = _cantilever(F)
    p.preconditioner = precond
    p.assemble('K')
    U = p.solve_device()
    assert float(U.abs().max()) > 0 and p.warm_start
    p.F = np.zeros_like(np.asarray(p.F))
    p.assemble('K')
    U0 = p.solve_device()
    assert p.solver_info['converged'] and float(U0.abs().max()) == 0.0


def test_graph_capture_engages_from_the_default_stream(F):
    """The Python host passes torch's current stream, normally the legacy default stream, on which capture is not
    permitted: the solve moves to the library's side stream and replays one captured CG iteration."""
    import torch
    p = _cantilever(F, 96, 48)
    p.preconditioner = 'gmg'
    p.assemble('K')
    assert torch.cuda.current_stream().cuda_stream == 0
    F._lib.graph_stats(reset=True)
    U = p.solve_device()
    torch.cuda.synchronize()
    st = F._lib.graph_stats()
    assert p.solver_info['converged']
    assert st['captures'] >= 1 and st['instantiated'] == st['captures'], st
    assert st['graph_launches'] >= p.solver_info['iterations'] and st['side_stream_solves'] == 1, st
    # ordering with the caller's stream: a torch op enqueued right after the call sees the finished solve
    y = torch.empty_like(U)
    F._lib.call('femol_spmv', p.device.plan, p.K_values.data_ptr(), U.data_ptr(), y.data_ptr(), F._lib.stream())
    assert float((y - p._F_dev).norm() / p._F_dev.norm()) < 10 * p.solver_rtol
    # a caller-made side stream is used as is
    s = torch.cuda.Stream()
    F._lib.graph_stats(reset=True)
    with torch.cuda.stream(s):
        p._U_dev = None
        p.solve_device()
    s.synchronize()
    st = F._lib.graph_stats()
    assert st['instantiated'] >= 1 and st['side_stream_solves'] == 0, st


def test_bc_and_load_caches_see_in_place_edits(F):
    """the caches were keyed on list lengths only."""
    nx, ny = 20, 10
    p = _cantilever(F, nx, ny)
    p.assemble('K')
    fixed_left = p.fixed_dofs().copy()
    U_left = p.solve_device().cpu().numpy().copy()
    # replace the fixed domain in place (same list length): clamp the right edge instead, load the left edge
    p.fixed_domains[0] = F.domains.inside_box([float(nx)], [[0, float(ny)]])
    p.forces, p.force_domains, p.F = [], [], np.zeros_like(np.asarray(p.F))
    p.add_forces([0, -1.0], F.domains.inside_box([0], [[0, float(ny)]]))
    p.assemble('K')
    fixed_right = p.fixed_dofs()
    assert not np.array_equal(fixed_left, fixed_right)
    U_right = p.solve_device().cpu().numpy()
    assert np.abs(U_right[fixed_right]).max() == 0.0 and np.abs(U_right[fixed_left]).max() > 0
    # mirror symmetry of the two problems
    uy_l = U_left[1::2].reshape(ny + 1, nx + 1)
    uy_r = U_right[1::2].reshape(ny + 1, nx + 1)
    assert np.allclose(uy_l, uy_r[:, ::-1], rtol=1e-6, atol=1e-9 * np.abs(uy_l).max())
    # in-place edit of the load vector: picked up by the next assembly, boundary conditions re-applied
    assert isinstance(p.F, np.ndarray)
    p.F[:] = 0.0
    p.F[fixed_right[0]] = 5.0                 # on a fixed dof: must be zeroed again
    free = np.setdiff1d(np.arange(p.F.shape[0]), fixed_right)[3]
    p.F[free] = -2.0
    p.assemble('K')
    Fd = p._F_dev.cpu().numpy()
    assert Fd[free] == -2.0 and Fd[fixed_right[0]] == 0.0 and np.count_nonzero(Fd) == 1
    p.F *= 2.0
    p.assemble('F')
    assert p._F_dev.cpu().numpy()[free] == -4.0
    # geometry edited in place: invalidate() refreshes the element matrices
    k0 = p.K_values.clone()
    p.mesh.points[:, 0] *= 2.0
    p.invalidate()
    p.mesh.element = p.mesh.ElementClasses['quad'](p.mesh.points[p.mesh.cells['quad'][0]], N_dof=2)
    p.assemble('K')
    assert float((p.K_values - k0).abs().max()) > 1e-3
