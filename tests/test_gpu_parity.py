"""Parity of the CUDA path (through the C-ABI / the reference-compatible Python API) against the golden
outputs of the unmodified reference and against the oracle on seeded inputs.  Needs a B200: `-m gpu`."""
import ctypes

import numpy as np
import pytest

import cases
import femol_oracle as o

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.fixture(scope='module')
def F():
    import femol_b200
    femol_b200._lib.require_gpu()
    return femol_b200


# ---- K1 / K2 element matrices -------------------------------------------------------------------------
ELEMENT_COMBOS = [('Q4', 2), ('Q4', 6), ('T3', 2), ('T6', 2), ('T6', 6)]


@pytest.mark.parametrize('et,nd', ELEMENT_COMBOS)
def test_element_matrices(F, golden, et, nd):
    """Elem(points, N_dof).Ke / .Me vs the reference classes on seeded distorted elements (1e-12)."""
    g = golden('elements')
    E, mu, rho, t = g['iso_E_mu_rho_t']
    mat = F.materials.IsotropicMaterial(E, mu, rho)
    cls = getattr(F.elements, et)
    for tn in ('iso', 'lam'):
        T = [g['%s_T%d' % (tn, k)] for k in range(4)]
        for i, pts in enumerate(g['xy_' + et]):
            Ke = cls(pts, nd).Ke(*T) if nd == 6 else cls(pts, nd).Ke(T[0])
            assert rel(Ke, g['Ke_%s_%d_%s' % (et, nd, tn)][i]) < 1e-12
    for i, pts in enumerate(g['xy_' + et]):
        assert rel(cls(pts, nd).Me(mat, t), g['Me_%s_%d' % (et, nd)][i]) < 1e-11


def test_element_batch_and_q4_uncoupled(F, golden):
    import torch
    from femol_b200 import elements, _lib
    g = golden('elements')
    xy = g['xy_Q4']
    pts = torch.from_numpy(xy.reshape(-1, 2).copy()).cuda()
    conn = torch.arange(xy.shape[0] * 4, dtype=torch.int32).reshape(-1, 4).cuda()
    T3t = [g['iso_T%d' % k] for k in range(3)]
    Ke = elements.ke_batch_device(_lib.Q4, 6, pts, conn, T3t).cpu().numpy()
    assert rel(Ke, g['Ke_Q4_6_iso3']) < 1e-12
    assert F.elements.T3(g['xy_T3'][0], 6).Ke(*T3t) is None      # reference: T3 plate unsupported


def test_reference_known_answers(F, golden):
    """The reference's own element known-answer tests through the CUDA path (test_Q4_element.py:22-101,
    test_T3_element.py:12-34)."""
    ref = golden('ref_matrices')
    mat = F.materials.IsotropicMaterial(210e6, 0.3, 1)
    Ke = F.elements.Q4(np.array([[0, 0], [0.5, 0], [0.5, 0.25], [0, 0.25]]), 2).Ke(mat.plane_tensor(0.025))
    assert abs(np.sum(Ke - ref['kattan_q4'])) < 1000
    Ke = F.elements.Q4(np.array([[0, 0], [0.25, 0], [0.25, 0.25], [0, 0.25]]), 2).Ke(mat.plane_tensor(0.025))
    assert rel(Ke, ref['kattan_q4']) < 5e-5
    mat = F.materials.IsotropicMaterial(210e9, 0.3, 1)
    Ke = F.elements.T3(np.array([[0, 0, 0], [0.5, 0.25, 0], [0, 0.25, 0]]), 2).Ke(mat.plane_tensor(0.025))
    assert np.allclose(np.around(Ke / 1e9, 4), ref['kattan_t3'], 1e-4)
    mat = F.materials.IsotropicMaterial(10920, 0.3, 1)
    el = F.elements.Q4(np.array([[0, 0], [0.1, 0], [0.1, 0.1], [0, 0.1]]), 6)
    idx = np.array([2, 3, 4, 8, 9, 10, 14, 15, 16, 20, 21, 22])
    Kb = el.Ke(np.zeros((3, 3)), mat.bending_tensor(0.1), np.zeros((2, 2)))
    Ks = el.Ke(np.zeros((3, 3)), np.zeros((3, 3)), mat.shear_tensor(0.1))
    assert np.allclose(ref['kb'], Kb[np.ix_(idx, idx)]) and np.allclose(ref['ks'], Ks[np.ix_(idx, idx)])
    # membrane block of the plate element == plane element on a distorted quad (test_Q4_element.py:34-57)
    pts = np.array([[-0.1, -0.1], [0.7, -0.2], [0.9, 0.3], [0, 0.5]])
    m2 = F.materials.IsotropicMaterial(2, 0.3, 1.1)
    Kp = F.elements.Q4(pts, 2).Ke(m2.plane_tensor(0.1))
    K6 = F.elements.Q4(pts, 6).Ke(m2.plane_tensor(0.1), m2.bending_tensor(0.1), m2.shear_tensor(0.1))
    pi = np.array([0, 1, 6, 7, 12, 13, 18, 19])
    assert np.allclose(Kp, K6[np.ix_(pi, pi)])


def test_q8_element_unpinned(F):
    """Q8 has no working reference (elements.py:1155); checked against the oracle's standard serendipity
    element and basic invariants (symmetry, rigid-body null space, area-consistent mass)."""
    rng = np.random.default_rng(5)
    c = np.array([[0, 0], [2, 0], [2, 1], [0, 1]], float) + rng.uniform(-0.1, 0.1, (4, 2))
    pts = np.vstack([c, (c + np.roll(c, -1, axis=0)) / 2])
    mat = F.materials.IsotropicMaterial(3.0, 0.3, 2.0)
    T = [mat.plane_tensor(0.2), mat.bending_tensor(0.2), mat.shear_tensor(0.2), mat.coupled_tensor()]
    for nd in (2, 6):
        Ke = F.elements.Q8(pts, nd).Ke(*T) if nd == 6 else F.elements.Q8(pts, 2).Ke(T[0])
        assert rel(Ke, o.ke_batch('Q8', pts[None], nd, T)[0]) < 1e-12
        assert np.allclose(Ke, Ke.T, atol=1e-12 * np.abs(Ke).max())
        Me = F.elements.Q8(pts, nd).Me(mat, 0.2)
        assert rel(Me, o.me_batch('Q8', pts[None], nd, 2.0, 0.2)[0]) < 1e-12
    K2 = F.elements.Q8(pts, 2).Ke(T[0])
    tx = np.tile([1.0, 0.0], 8)
    assert np.abs(K2 @ tx).max() < 1e-12 * np.abs(K2).max()
    M2 = F.elements.Q8(pts, 2).Me(mat, 0.2)
    assert np.isclose(tx @ M2 @ tx, 2.0 * 0.2 * F.elements.Q8(pts, 2).area(), rtol=1e-12)


# ---- plan + assembly + static solve ---------------------------------------------------------------------
NON_SIMP = [n for n, c in cases.CASES.items() if 'simp' not in c]


@pytest.mark.parametrize('name', NON_SIMP)
def test_assembly_and_solve(F, golden, name):
    """K / M: sparsity pattern bit-exact (sorted int32 indices, explicit zeros), values 1e-12 relative;
    displacements: relative residual <= 1e-8, compliance 1e-6."""
    case, g = cases.CASES[name], golden(name)
    problem = cases.build_problem(F, case)
    ctype = problem.mesh.contains[0]
    problem.assemble('K')
    K = problem.K
    assert K.indices.dtype == np.int32 and K.indptr.dtype == np.int32
    assert np.array_equal(K.indptr, g['K_indptr']) and np.array_equal(K.indices, g['K_indices'])
    assert rel(K.data, g['K_data']) < 1e-12
    if 'M_data' in g:
        problem.assemble('M')
        M = problem.M
        assert np.array_equal(M.indptr, g['M_indptr']) and np.array_equal(M.indices, g['M_indices'])
        assert rel(M.data, g['M_data']) < 1e-12
        problem.assemble('M', X={ctype: g['X']}, q=1)
        assert rel(problem.M.data, g['MX_data']) < 1e-12
    if 'U' in g:
        assert np.array_equal(problem.F, g['F'])
        problem.solver_rtol = 1e-11
        mesh = problem.solve(verbose=False)
        U = mesh.U
        # relative residual <= 1e-8 — or the reference's own residual level when its direct solve cannot
        # reach that either (ill-conditioned laminate plates: attainable residual ~ eps * cond(K))
        rr = np.linalg.norm(g['F'] - K @ U) / np.linalg.norm(g['F'])
        rr_ref = np.linalg.norm(g['F'] - K @ g['U']) / np.linalg.norm(g['F'])
        assert problem.solver_info['converged'], problem.solver_info
        assert rr < max(1e-8, 20 * rr_ref), (rr, rr_ref, problem.solver_info)
        assert abs(g['F'] @ U - g['F'] @ g['U']) / abs(g['F'] @ g['U']) < 1e-6
        assert rel(U, g['U']) < 1e-5, (rel(U, g['U']), problem.solver_info)
        assert np.allclose(mesh.point_data['Ux'], U[0::problem.N_dof])
    problem.assemble('K', X={ctype: g['X']}, p=3)
    assert rel(problem.K.data, g['KX_data']) < 1e-12
    if 'UX' in g:
        U = problem.solve(verbose=False).U
        assert abs(g['F'] @ U - g['F'] @ g['UX']) / abs(g['F'] @ g['UX']) < 1e-6, problem.solver_info


def test_scatter_map_equals_coo_assembly(F, golden):
    """element -> CSR scatter map (the plan's replacement of mesh.rows / mesh.cols, mesh.py:284-308)"""
    g = golden('plate_q4_jitter_6x5')
    problem = cases.build_problem(F, cases.CASES['plate_q4_jitter_6x5'])
    d = problem.device
    smap = d.scatter_map().cpu().numpy()           # [n_elem][nn][nn] node-block positions
    problem.assemble('K')
    K = problem.K
    nd, nn = 6, 4
    cells = g['cells']
    # rebuild every COO slot's CSR position from the map and check it addresses the right (row, col)
    adjp = np.zeros(d.n_nodes + 1, np.int32); adji = np.zeros(d.nnz_node, np.int32)
    import torch
    ap = torch.empty(d.n_nodes + 1, dtype=torch.int32, device='cuda'); ai = torch.empty(d.nnz_node, dtype=torch.int32, device='cuda')
    F._lib.call('femol_plan_export_adjacency', d.plan, ap.data_ptr(), ai.data_ptr(), F._lib.stream())
    adjp, adji = ap.cpu().numpy(), ai.cpu().numpy()
    for e in range(0, cells.shape[0], 7):
        for a in range(nn):
            for b in range(nn):
                pos = smap[e, a, b]
                na, nb = cells[e, a], cells[e, b]
                assert adjp[na] <= pos < adjp[na + 1] and adji[pos] == nb
                rank = pos - adjp[na]
                for da in range(nd):
                    for db in range(nd):
                        csr_pos = K.indptr[na * nd + da] + rank * nd + db
                        assert K.indices[csr_pos] == nb * nd + db


def test_empty_and_degenerate_inputs(F):
    """Edge cases: no fixed domain (no BC), single element, zero force -> zero displacement at once."""
    mesh = F.mesh.rectangle_Q4(1., 1., 1, 1)
    p = F.FEM_Problem('displacement', 'plane', mesh)
    p.define_materials(F.materials.general_isotropic()); p.define_tensors(1)
    p.assemble('K')
    arr_K = o.ke_batch('Q4', mesh.points[mesh.cells['quad']][:, :, :2], 2, o.iso_tensors(1, 0.3, 1))[0]
    Kd = p.K.toarray()
    order = mesh.cells['quad'][0]
    idx = (np.repeat(order, 2) * 2 + np.tile([0, 1], 4))
    assert rel(Kd[np.ix_(idx, idx)], arr_K) < 1e-13
    p.add_fixed_domain(F.domains.inside_box([0], [[0, 1]]))
    p.assemble('K')
    U = p.solve(verbose=False).U
    assert np.all(U == 0) and p.solver_info['iterations'] == 0
    with pytest.raises(F._lib.FemolError):
        F._lib.call('femol_ke_batch', 1, 3, 1, None, None, None, None, None)


# ---- K4 SpMV ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['plate_lam_8x6', 'plane_t3_9x7', 'plate_t6_6x4', 'modal_q4_plane_10x5'])
def test_spmv(F, golden, name):
    import torch
    problem = cases.build_problem(F, cases.CASES[name])
    problem.assemble('K')
    d = problem.device
    K = problem.K
    x = np.random.default_rng(0).standard_normal(d.n)
    xd = torch.from_numpy(x).cuda()
    y = torch.empty_like(xd)
    F._lib.call('femol_spmv', d.plan, problem.K_values.data_ptr(), xd.data_ptr(), y.data_ptr(), F._lib.stream())
    ref = K @ x
    assert rel(y.cpu().numpy(), ref) < 1e-13
    indptr, indices = d.csr_pattern()
    y2 = torch.empty_like(xd)
    F._lib.call('femol_spmv_csr', d.n, indptr.data_ptr(), indices.data_ptr(), problem.K_values.data_ptr(),
                xd.data_ptr(), y2.data_ptr(), F._lib.stream())
    assert rel(y2.cpu().numpy(), ref) < 1e-13
    # linearity (size-independent property)
    x2 = torch.from_numpy(np.random.default_rng(1).standard_normal(d.n)).cuda()
    ya = torch.empty_like(xd); yb = torch.empty_like(xd)
    F._lib.call('femol_spmv', d.plan, problem.K_values.data_ptr(), x2.data_ptr(), ya.data_ptr(), F._lib.stream())
    z = (2.5 * xd - 0.5 * x2).contiguous()
    F._lib.call('femol_spmv', d.plan, problem.K_values.data_ptr(), z.data_ptr(), yb.data_ptr(), F._lib.stream())
    assert rel((2.5 * y - 0.5 * ya).cpu().numpy(), yb.cpu().numpy()) < 1e-12


@pytest.mark.parametrize('precond', ['none', 'jacobi', 'block_jacobi'])
def test_pcg_preconditioners(F, golden, precond):
    g = golden('plate_lam_8x6')
    problem = cases.build_problem(F, cases.CASES['plate_lam_8x6'])
    problem.preconditioner = precond
    problem.solver_rtol = 1e-7          # attainable on this ill-conditioned laminate (cond ~1e9)
    U = problem.solve(verbose=False).U
    assert problem.solver_info['converged'] and problem.solver_info['relres'] <= 1e-6, problem.solver_info
    assert abs(g['F'] @ U - g['F'] @ g['U']) / abs(g['F'] @ g['U']) < 1e-6
    # warm start: a second solve from the converged solution needs (almost) no work
    it1 = problem.solver_info['iterations']
    problem.solve(verbose=False)
    assert problem.solver_info['iterations'] <= max(1, it1 // 5)


# ---- K7 / K9 / K10 kernel-level parity on seeded inputs -----------------------------------------------------
def _oc_call(F, x, dc, mode, volfrac, areas=None, mask=None):
    import torch
    xd = torch.from_numpy(x.copy()).cuda(); dd = torch.from_numpy(dc.copy()).cuda()
    xn = torch.empty_like(xd); sc = torch.zeros(4, dtype=torch.float64, device='cuda')
    ws = torch.zeros(int(F._lib.load().femol_reduce_workspace_doubles()), dtype=torch.float64, device='cuda')
    ad = torch.from_numpy(areas.copy()).cuda() if areas is not None else None
    md = torch.from_numpy(mask.copy()).cuda() if mask is not None else None
    F._lib.call('femol_oc_update', x.size, xd.data_ptr(), dd.data_ptr(), mode, float(volfrac),
                ad.data_ptr() if ad is not None else None, md.data_ptr() if md is not None else None,
                xn.data_ptr(), sc.data_ptr(), ws.data_ptr(), F._lib.stream())
    return xn.cpu().numpy(), sc.cpu().numpy(), dd.cpu().numpy()


@pytest.mark.parametrize('n', [1, 37, 5000, 200003])
def test_oc_update_comp(F, n):
    rng = np.random.default_rng(n)
    x = rng.uniform(0.001, 1, n); dc = -rng.uniform(1e-6, 5, n)
    xn, sc, _ = _oc_call(F, x, dc, 0, 0.4)
    ref = o.oc_update_comp(x, dc, 0.4)
    assert np.abs(xn - ref).max() < 1e-9
    assert np.isclose(sc[0], np.abs(ref - x).max(), atol=1e-9) and sc[3] == 30
    # solid / void masks (TOM_SIMP.py:299-303)
    if n > 30:
        mask = np.zeros(n, np.uint8); mask[:5] = 1; mask[5:9] = 2
        xn, _, _ = _oc_call(F, x, dc, 0, 0.4, mask=mask)
        ref = o.oc_update_comp(x, dc, 0.4, solid_mask=mask == 1, void_mask=mask == 2)
        assert np.abs(xn - ref).max() < 1e-9 and np.all(xn[:5] == 1) and np.all(xn[5:9] == 0.001)


def test_oc_update_sign_switch_and_nan(F):
    """any(dc > 0) switches the sign convention for the whole vector; the other sign then yields NaN
    through the fractional power and numpy's NaN-propagating min / max (TOM_SIMP.py:289-297)."""
    rng = np.random.default_rng(2)
    x = rng.uniform(0.1, 1, 100)
    dc = rng.uniform(0.1, 2, 100)             # all positive: the "dc > 0" branch
    xn, _, _ = _oc_call(F, x, dc, 0, 0.5)
    assert np.abs(xn - o.oc_update_comp(x, dc, 0.5)).max() < 1e-9
    dc[3] = -0.7                               # mixed signs -> NaN at that element, like the reference
    xn, _, _ = _oc_call(F, x, dc, 2, 0.5)     # mode 2 = the reference's literal rule
    ref = o.oc_update_comp(x, dc, 0.5)
    assert np.isnan(ref[3]) and np.isnan(xn[3])
    ok = ~np.isnan(ref)
    assert np.array_equal(np.isnan(xn), np.isnan(ref)) and np.abs(xn[ok] - ref[ok]).max() < 1e-9
    # default rule: dominant sign decides, the stray entry counts as zero sensitivity (no NaN)
    xn, _, _ = _oc_call(F, x, dc, 0, 0.5)
    dc0 = dc.copy(); dc0[3] = 0.0
    assert np.abs(xn - o.oc_update_comp(x, dc0, 0.5)).max() < 1e-9
    # rounding-level positive noise in an otherwise negative field must not flip the convention
    dcn = -rng.uniform(0.1, 2, 100); dcn[7] = 1e-12
    xn, _, _ = _oc_call(F, x, dcn, 0, 0.5)
    dcz = dcn.copy(); dcz[7] = 0.0
    assert np.all(np.isfinite(xn)) and np.abs(xn - o.oc_update_comp(x, dcz, 0.5)).max() < 1e-9


@pytest.mark.parametrize('structured', [True, False])
def test_oc_update_vibe(F, structured):
    rng = np.random.default_rng(7)
    n = 4000
    x = rng.uniform(0.001, 1, n)
    for scale in (1.0, 1e-6):                  # the second triggers the "sum < 1 -> rescale" branch
        dl = rng.uniform(-0.3, 1, n) * scale
        areas = rng.uniform(0.5, 1.5, n)
        xn, sc, dl_out = _oc_call(F, x, dl, 1, 0.5, areas=None if structured else areas)
        ref, dl_ref = o.oc_update_vibe(x, dl, 0.5, structured, areas)
        assert np.abs(xn - ref).max() < 1e-9
        assert rel(dl_out, dl_ref) < 1e-14


@pytest.mark.parametrize('rmin', [1.2, 1.5, 2.0, 3.1])
def test_filter_grid_and_list(F, rmin):
    import torch
    from femol_b200.TOM_SIMP import apply_filter, build_filter
    mesh = F.mesh.rectangle_Q4(9., 4.5, 18, 9)
    rng = np.random.default_rng(1)
    x = rng.uniform(0.001, 1, mesh.N_ele); dc = -rng.uniform(0, 3, mesh.N_ele)
    ref = o.filter_sens_literal(x, dc, mesh.all_cell_centers, rmin, mesh.element_size())
    xd, dd = torch.from_numpy(x).cuda(), torch.from_numpy(dc).cuda()
    out = torch.empty_like(xd)
    flt = build_filter(mesh, rmin, torch.device('cuda'))
    assert flt['kind'] == 'grid'
    apply_filter(flt, xd, dd, out)
    assert rel(out.cpu().numpy(), ref) < 1e-13
    mesh.structured = False                      # force the neighbour-list path on the same mesh
    h_struct = 0.5
    flt = build_filter(mesh, rmin * h_struct / mesh.element_size(), torch.device('cuda'))
    assert flt['kind'] == 'list'
    ref2 = o.filter_sens_literal(x, dc, mesh.all_cell_centers, rmin * h_struct / mesh.element_size(), mesh.element_size())
    apply_filter(flt, xd, dd, out)
    assert rel(out.cpu().numpy(), ref2) < 1e-12


def test_compliance_sensitivities(F, golden):
    """c and dc of iteration 0 vs the reference (structured, coating, and the unstructured
    double-penalisation quirk)."""
    for name in ('comp_q4_30x15', 'comp_q4_coat_30x15', 'comp_t6_8x4', 'comp_q4_jitter_coat_8x4', 'comp_q4_plate_16x8'):
        case, g = cases.CASES[name], golden(name)
        problem = cases.build_problem(F, case)
        s = case['simp']
        opt = F.SIMP_COMP(problem, volfrac=s['volfrac'], penal=s['penal'], rmin=s['rmin'])
        problem.solver_rtol = 1e-12
        opt._U = opt.FEM_solver(None)
        assert rel(problem.K.data, g['K0_data']) < 1e-12
        assert rel(opt._U.cpu().numpy(), g['U0']) < 1e-6
        c, dc = opt.objective_function(None)
        assert abs(float(c) - g['c_hist'][0]) / abs(g['c_hist'][0]) < 1e-6
        assert rel(dc.cpu().numpy(), g['dc0_raw']) < 1e-6
        opt._filter_sensibility(None)
        assert rel(opt._dc.cpu().numpy(), g['dc0_filtered']) < 1e-6


# ---- full optimiser loops ---------------------------------------------------------------------------------------
COMP = [n for n, c in cases.CASES.items() if c.get('simp', {}).get('kind') == 'comp']


@pytest.mark.parametrize('name', COMP)
def test_simp_comp(F, golden, name):
    """SIMP_COMP density fields within 1e-4 after a fixed iteration count, compliance within 1e-6."""
    case, g = cases.CASES[name], golden(name)
    s = case['simp']
    problem = cases.build_problem(F, case)
    problem.solver_rtol = 1e-11
    opt = F.SIMP_COMP(problem, volfrac=s['volfrac'], penal=s['penal'], rmin=s['rmin'])
    cs, Xs = [], []
    for it in range(s['iters']):
        opt.solve(converge=0.0, max_iter=1, plot=True, save=False)
        cs.append(opt.c); Xs.append(opt.X[problem.mesh.contains[0]].copy())
        assert problem.solver_info['converged'], problem.solver_info
    assert rel(np.array(cs), g['c_hist']) < 1e-6
    err = np.abs(np.array(Xs) - g['X_hist']).max(axis=1)
    assert err.max() < 1e-4, err
    mesh = opt.mesh
    assert 'X' in mesh.cell_data and 'd1_Ux' in mesh.point_data
    assert np.isclose(np.sum(Xs[-1]), np.sum(g['X_hist'][-1]), rtol=1e-6)
    if case['model'] == 'plane' and 'jitter' not in name:
        assert np.isclose(np.sum(Xs[-1]), s['volfrac'] * len(Xs[-1]), rtol=1e-2)   # reference test_SIMP_COMP.py:37


def test_simp_comp_solve_loop_and_domains(F):
    """solve() as a user calls it (test_SIMP_COMP.py:7-37) + solid / void domains."""
    problem = cases.build_problem(F, cases.CASES['comp_q4_30x15'])
    opt = F.SIMP_COMP(problem, volfrac=0.4, penal=3)
    opt.solid_domain = F.domains.inside_box([[27, 30]], [[6, 9]])
    opt.void_domain = F.domains.inside_box([[0, 4]], [[5, 10]])
    mesh = opt.solve(converge=0.03, max_iter=4, plot=True, save=False)
    X = mesh.cell_data['X']['quad']
    c = mesh.all_cell_centers
    solid = (c[:, 0] > 27) & (c[:, 1] > 6) & (c[:, 1] < 9)
    void = (c[:, 0] < 4) & (c[:, 1] > 5) & (c[:, 1] < 10)
    assert np.all(X[solid] == 1.0) and np.all(X[void] == 0.001) and opt.loop == 4
    assert opt.iteration_log[1]['c'] < opt.iteration_log[0]['c']                 # test_SIMP_COMP.py:153 (c2 < c1)


# ---- modal ----------------------------------------------------------------------------------------------------------
MODAL = [n for n, c in cases.CASES.items() if c['physics'] == 'modal' and 'simp' not in c]


def _reference_physical_spectrum(case, g, k):
    """The k lowest eigen-frequencies of the reference's pencil that are neither the spurious lambda = 1 pairs
    of the fixed dofs (FEM.py:496,530) nor pure drilling-rotation modes (elements.py:1071): the dense
    eigen-solve of the oracle on the free, non-drilling dofs.  Each value must also be present in the spectrum
    the unmodified reference returned (golden 'w'), which pins this sub-spectrum to the reference."""
    import scipy.linalg
    arr = o.case_arrays(case, cases.jitter_points)
    K = o.global_matrices(arr, 'K'); M = o.global_matrices(arr, 'M')
    keep = np.ones(K.shape[0], bool)
    keep[arr['fixed_idx']] = False
    if arr['ndof'] == 6:
        keep[5::6] = False
    Kf = K[keep][:, keep].toarray(); Mf = M[keep][:, keep].toarray()
    lam = scipy.linalg.eigh(Kf, Mf, eigvals_only=True, subset_by_index=[0, k - 1])
    f = np.sqrt(lam) / (2 * np.pi)
    for fi in f:
        assert np.min(np.abs(g['w'] - fi) / fi) < 1e-4, (fi, 'not in the reference spectrum')
    return f


@pytest.mark.parametrize('name', MODAL)
def test_modal_frequencies(F, golden, name):
    """Lowest physical eigen-frequencies within 1e-6 of the reference pencil's (LOBPCG, residual 1e-8)."""
    case, g = cases.CASES[name], golden(name)
    problem = cases.build_problem(F, case)
    problem.n_modes = 6
    problem.modal_tol = 1e-9
    w, v = problem.solve(verbose=False)
    ref = _reference_physical_spectrum(case, g, 6)
    assert np.all(problem.solver_info['relres'] < 1e-8), problem.solver_info
    assert np.allclose(w, ref, rtol=1e-6), (w, ref, problem.solver_info)
    M = problem.M
    assert np.allclose(v @ (M @ v.T), np.eye(6), atol=1e-7)                      # test_modal_solver.py:7-22
    if problem.N_dof == 6:
        assert np.abs(v[:, 5::6]).max() == 0.0
    assert np.abs(v[:, problem.fixed_dofs()]).max() == 0.0


def test_mac_kernel(F):
    import torch
    rng = np.random.default_rng(0)
    V = rng.standard_normal((5, 3001)); vr = V[2] * -3.0 + 1e-3 * rng.standard_normal(3001)
    Vd, vd = torch.from_numpy(V).cuda(), torch.from_numpy(vr).cuda()
    out = torch.empty(5, dtype=torch.float64, device='cuda')
    ws = torch.zeros(int(F._lib.load().femol_reduce_workspace_doubles()), dtype=torch.float64, device='cuda')
    F._lib.call('femol_mac', 3001, 5, Vd.data_ptr(), vd.data_ptr(), out.data_ptr(), ws.data_ptr(), F._lib.stream())
    ref = np.array([o.mac(v, vr) for v in V])
    assert np.allclose(out.cpu().numpy(), ref, rtol=1e-12) and int(out.argmax()) == 2
    assert float(ws.abs().max()) == 0.0


def test_simp_vibe(F, golden):
    """SIMP_VIBE (unpinned by the reference's tests; golden = the reference run): tracked frequency 1e-6,
    densities 1e-4 after a fixed iteration count."""
    name = 'vibe_q4_plate_10x10'
    case, g = cases.CASES[name], golden(name)
    s = case['simp']
    problem = cases.build_problem(F, case)
    problem.n_modes = 6
    problem.modal_tol = 1e-10
    opt = F.SIMP_VIBE(problem, volfrac=s['volfrac'], p=s['p'], q=s['q'], rmin=s['rmin'])
    mesh = opt.solve(g['v_ref'], converge=0.0, min_iter=1, max_iter=s['iters'], plot=False, save=False, verbose=False)
    assert np.allclose(opt.lmbds, g['lmbd_hist'], rtol=1e-6), (opt.lmbds, g['lmbd_hist'])
    Xs = [mesh.cell_data['X%d' % i]['quad'] for i in range(1, s['iters'])] + [mesh.cell_data['X']['quad']]
    err = np.abs(np.array(Xs) - g['X_hist']).max(axis=1)
    assert err.max() < 1e-4, err
    assert min(opt.macs) > 0.9


# ---- size-independent properties at a larger size -----------------------------------------------------------------
def test_large_mesh_properties(F):
    """400 x 200 Q4 plane (161k dof): nnz formula, symmetry, rigid-body null space before BCs, SpMV vs PCG
    consistency, filter / OC invariants."""
    import torch
    nx, ny = 400, 200
    mesh = F.mesh.rectangle_Q4(float(nx), float(ny), nx, ny)
    p = F.FEM_Problem('displacement', 'plane', mesh)
    p.define_materials(F.materials.general_isotropic()); p.define_tensors(1)
    x = torch.from_numpy(cases.density_field(mesh.N_ele)).cuda()
    p.assemble('K', X=x, p=3)
    d = p.device
    assert d.nnz == 4 * (3 * nx + 1) * (3 * ny + 1)
    K = p.K
    assert abs(K - K.T).max() < 1e-12 * abs(K).max()
    tx = np.tile([1.0, 0.0], d.n_nodes)
    assert np.abs(K @ tx).max() < 1e-10 * abs(K).max()
    p.add_fixed_domain(F.domains.inside_box([0], [[0, float(ny)]]))
    p.add_forces([0, -1.0], F.domains.inside_box([float(nx)], [[0, float(ny)]]))
    p.assemble('K', X=x, p=3)
    p.solver_rtol = 1e-9
    U = p.solve_device()
    assert p.solver_info['converged'] and p.solver_info['relres'] < 1e-8, p.solver_info
    y = torch.empty_like(U)
    F._lib.call('femol_spmv', d.plan, p.K_values.data_ptr(), U.data_ptr(), y.data_ptr(), F._lib.stream())
    assert float((y - p._F_dev).norm() / p._F_dev.norm()) < 1e-8
    fixed = p.fixed_dofs()
    assert float(U[torch.from_numpy(fixed).cuda()].abs().max()) == 0.0


def test_matrix_free_solve_beyond_one_reduction_grid(F):
    """4200 x 2100 Q4 plane (17.7 M dof): the matrix-free product needs more CTAs (17 358) than the one-channel grid
    reduction has slots (8 192), so it runs in chunks and its fused dot products (p . K p in CG, r . z in the last
    kernel of the V-cycle) become a pass of their own.  A 36 M-dof run found the r . z pass called with a null operand;
    the solve must converge with the iteration count of the small meshes and satisfy K u = f."""
    import torch
    nx, ny = 4200, 2100
    mesh = F.mesh.rectangle_Q4(float(nx), float(ny), nx, ny)
    p = F.FEM_Problem('displacement', 'plane', mesh)
    p.define_materials(F.materials.general_isotropic()); p.define_tensors(1)
    p.add_fixed_domain(F.domains.inside_box([0], [[0, float(ny)]]))
    p.add_forces([0, -1.0], F.domains.inside_box([float(nx)], [[ny / 2.0 - 1, ny / 2.0 + 1]]))
    p.solver_rtol = 1e-8
    p.assemble('K')
    U = p.solve_device()
    info = p.solver_info
    assert info['preconditioner'] == 'gmg' and info['converged'] and info['iterations'] < 40, info
    d = p.device
    y = torch.empty_like(U)
    F._lib.call('femol_spmv', d.plan, p.K_values.data_ptr(), U.data_ptr(), y.data_ptr(), F._lib.stream())   # assembled CSR operator
    assert float((y - p._F_dev).norm() / p._F_dev.norm()) < 2e-8
    del p, U, y
    torch.cuda.empty_cache()


# ---- K6 block kernels (LOBPCG building blocks) vs numpy ---------------------------------------------------------
def test_block_kernels(F, golden):
    import torch
    L = F._lib
    problem = cases.build_problem(F, cases.CASES['modal_q4_plate_cant_8x8'])
    problem.assemble('K'); problem.assemble('M')
    d = problem.device
    K, M = problem.K, problem.M
    n = d.n
    rng = np.random.default_rng(0)
    for m in (1, 2, 4, 8, 12, 16):
        ld = 3 * m
        S = rng.standard_normal((n, ld))
        Sd = torch.from_numpy(S).cuda()
        Yd = torch.zeros_like(Sd)
        # spmm on a column slice with leading dimension 3m
        L.call('femol_spmm', d.plan, problem.K_values.data_ptr(), Sd.data_ptr() + 8 * m, ld, Yd.data_ptr() + 8 * m, ld, m, L.stream())
        Y = Yd.cpu().numpy()
        assert rel(Y[:, m:2 * m], K @ S[:, m:2 * m]) < 1e-13 and np.all(Y[:, :m] == 0) and np.all(Y[:, 2 * m:] == 0)
        # gram of slices
        ws = torch.zeros(int(L.load().femol_gram_workspace_doubles(48, 48)), dtype=torch.float64, device='cuda')
        out = torch.zeros(48 * 48, dtype=torch.float64, device='cuda')
        T = rng.standard_normal((n, ld)); Td = torch.from_numpy(T).cuda()
        L.call('femol_gram', n, 2 * m, ld, Sd.data_ptr() + 8 * m, ld, Td.data_ptr(), ld, out.data_ptr(), ws.data_ptr(), L.stream())
        G = out[:2 * m * ld].cpu().numpy().reshape(2 * m, ld)
        assert rel(G, S[:, m:].T @ T) < 1e-12
        # gemm out of place into a slice of another array, alpha / beta, and in place
        C = rng.standard_normal((2 * m, m)); Cd = torch.from_numpy(C).cuda()
        Zd = torch.from_numpy(T.copy()).cuda()
        L.call('femol_block_gemm', n, 2 * m, m, Sd.data_ptr(), ld, Cd.data_ptr(), Zd.data_ptr() + 8 * 2 * m, ld, 1.0, 0.0, L.stream())
        Z = Zd.cpu().numpy()
        assert rel(Z[:, 2 * m:], S[:, :2 * m] @ C) < 1e-13 and np.array_equal(Z[:, :2 * m], T[:, :2 * m])
        L.call('femol_block_gemm', n, 2 * m, m, Sd.data_ptr(), ld, Cd.data_ptr(), Zd.data_ptr() + 8 * 2 * m, ld, -1.0, 1.0, L.stream())
        assert np.abs(Zd.cpu().numpy()[:, 2 * m:]).max() < 1e-12 * np.abs(Z).max()
        C2 = rng.standard_normal((m, m)); C2d = torch.from_numpy(C2).cuda()
        S2d = torch.from_numpy(S.copy()).cuda()
        L.call('femol_block_gemm', n, m, m, S2d.data_ptr() + 8 * m, ld, C2d.data_ptr(), S2d.data_ptr() + 8 * m, ld, 1.0, 0.0, L.stream())
        S2 = S2d.cpu().numpy()
        assert rel(S2[:, m:2 * m], S[:, m:2 * m] @ C2) < 1e-13 and np.array_equal(S2[:, :m], S[:, :m]) and np.array_equal(S2[:, 2 * m:], S[:, 2 * m:])
        # residual
        lam = rng.uniform(1, 2, m); lamd = torch.from_numpy(lam).cuda()
        Rd = torch.zeros_like(Sd)
        L.call('femol_block_residual', n, m, Sd.data_ptr(), ld, Td.data_ptr(), ld, lamd.data_ptr(), Rd.data_ptr() + 8 * m, ld, L.stream())
        assert rel(Rd.cpu().numpy()[:, m:2 * m], S[:, :m] - T[:, :m] * lam[None]) < 1e-14
    # preconditioners: Jacobi / block-Jacobi inverse applied in place, fixed dofs zeroed
    mask = problem._fixed_mask()
    fixed = problem.fixed_dofs()
    Kd = K.toarray()
    for kind, name in ((1, 'jacobi'), (2, 'block')):
        dinv = torch.empty(d.n_nodes * 36, dtype=torch.float64, device='cuda')
        L.call('femol_precond_setup', d.plan, problem.K_values.data_ptr(), kind, dinv.data_ptr(), L.stream())
        R = rng.standard_normal((n, 8)); Rd = torch.from_numpy(R).cuda()
        L.call('femol_precond_apply', d.plan, dinv.data_ptr(), mask.data_ptr(), Rd.data_ptr(), 8, Rd.data_ptr(), 8, 8, L.stream())
        W = Rd.cpu().numpy()
        Rm = R.copy(); Rm[fixed] = 0
        ref = np.zeros_like(R)
        for i in range(d.n_nodes):
            sl = slice(6 * i, 6 * i + 6)
            B = Kd[sl, sl] if kind == 2 else np.diag(np.diag(Kd[sl, sl]))
            ref[sl] = np.linalg.solve(B, Rm[sl])
        ref[fixed] = 0
        assert rel(W, ref) < 1e-10, name


def test_lobpcg_device_ops_match_numpy_ops(F):
    """The same LOBPCG driver on the CUDA ops and on the numpy stand-in: identical iterates."""
    import torch
    from femol_b200 import lobpcg as lp
    from numpy_ops import NumpyOps
    problem = cases.build_problem(F, cases.CASES['modal_q4_plate_cant_8x8'])
    problem.assemble('K'); problem.assemble('M')
    d = problem.device
    m, k = 8, 6
    X0 = np.random.default_rng(0).standard_normal((d.n, m))
    fixed = problem.fixed_dofs()
    X0[fixed] = 0
    nops = NumpyOps(problem.K, problem.M, fixed, m, X0, precond='jacobi')
    dops = lp.DeviceOps(d.plan, problem.K_values, problem.M_values, problem._fixed_mask(), m, F._lib.PRECOND_JACOBI,
                        torch.from_numpy(X0).cuda())
    ln, rn, _ = lp.lobpcg(nops, m, k, tol=1e-30, max_iter=25)
    ld, rd, _ = lp.lobpcg(dops, m, k, tol=1e-30, max_iter=25)
    assert np.allclose(ln, ld, rtol=1e-6), (ln, ld)
    assert np.allclose(rn, rd, rtol=1e-3, atol=1e-12), (rn, rd)


# ---- geometric multigrid preconditioner -----------------------------------------------------------------------------
def _gmg_problem(F, nx, ny, model, coating=False, seed=0):
    Lx, Ly = float(nx), float(ny)
    mesh = F.mesh.rectangle_Q4(Lx, Ly, nx, ny)
    p = F.FEM_Problem('displacement', model, mesh)
    if model == 'plane':
        mats = [F.materials.general_isotropic()] + ([F.materials.IsotropicMaterial(5, 0.25, 2)] if coating else [])
        p.define_materials(*mats); p.define_tensors(*([1] + ([0.5] if coating else [])))
        p.add_fixed_domain(F.domains.inside_box([0], [[0, Ly]]), ddls=[0, 1])
        p.add_forces([0, -1.0], F.domains.inside_box([Lx], [[0, Ly]]))
    else:
        p.define_materials(F.materials.IsotropicMaterial(10920, 0.3, 1)); p.define_tensors(0.5)
        p.add_fixed_domain(F.domains.inside_box([0], [[0, Ly]]))
        p.add_forces([0, 0, -1e-2, 0, 0, 0], F.domains.inside_box([Lx], [[0, Ly]]))
    x = cases.density_field(mesh.N_ele, seed=seed + 1)
    return p, x


@pytest.mark.parametrize('nx,ny,model,coating', [(32, 16, 'plane', False), (33, 17, 'plane', False), (45, 22, 'plane', True),
                                                 (24, 12, 'plate', False), (21, 13, 'plate', False)])
def test_gmg_pcg_matches_block_jacobi_pcg(F, nx, ny, model, coating):
    """Same solution as the block-Jacobi PCG (and scipy), in far fewer iterations; even / odd grid sizes."""
    import scipy.sparse.linalg as sla
    p, x = _gmg_problem(F, nx, ny, model, coating)
    ctype = 'quad'
    p.solver_rtol = 1e-9 if model == 'plane' else 1e-7
    p.assemble('K', X={ctype: x}, p=3)
    p.preconditioner = 'block_jacobi'
    U_bj = p.solve_device().cpu().numpy().copy()
    it_bj = p.solver_info['iterations']
    p.preconditioner = 'gmg'
    p._U_dev = None
    U_mg = p.solve_device().cpu().numpy().copy()
    info = dict(p.solver_info)
    assert info['converged'] and info['preconditioner'] == 'gmg', info
    # plane stress: an order of magnitude fewer iterations; Reissner-Mindlin plates: the damped block-Jacobi
    # smoother is weaker on the shear-coupled (w, theta) dofs, the gain is about 3x (limits)
    assert info['iterations'] < (max(60, it_bj // 4) if model == 'plane' else it_bj // 2), (info, it_bj)
    U_ref = sla.spsolve(p.K.tocsc(), p.F)
    c_ref = p.F @ U_ref
    assert abs(p.F @ U_mg - c_ref) / abs(c_ref) < 1e-6 and abs(p.F @ U_bj - c_ref) / abs(c_ref) < 1e-6
    assert rel(U_mg, U_ref) < 1e-4, (rel(U_mg, U_ref), info)
    assert p.structured_grid_shape() == (nx, ny)


def test_gmg_vcycle_is_symmetric_positive(F):
    """A V-cycle with equal pre/post damped-Jacobi sweeps is a symmetric positive operator (CG needs that)."""
    import torch
    p, x = _gmg_problem(F, 37, 19, 'plane')
    p.assemble('K', X={'quad': x}, p=3)
    p.preconditioner = 'gmg'
    g = p._gmg_ready()
    d = p.device
    rng = np.random.default_rng(0)
    fixed = p.fixed_dofs()
    R = rng.standard_normal((3, d.n)); R[:, fixed] = 0
    Z = []
    for r in R:
        rd = torch.from_numpy(r).cuda(); zd = torch.empty_like(rd)
        F._lib.call('femol_gmg_apply', g, rd.data_ptr(), zd.data_ptr(), F._lib.stream())
        Z.append(zd.cpu().numpy())
    assert abs(R[0] @ Z[1] - R[1] @ Z[0]) < 1e-10 * abs(R[0] @ Z[1]) + 1e-14 * np.linalg.norm(Z[0]) * np.linalg.norm(R[1])
    assert all(R[i] @ Z[i] > 0 for i in range(3))
    import ctypes
    info = (ctypes.c_int64 * 64)()
    F._lib.call('femol_gmg_info', g, info)
    assert info[0] >= 3 and (info[1], info[2]) == (37, 19) and (info[3], info[4]) == (19, 10)


def _lattice_problem(F, kind, nx, ny, model='plane', seed=5):
    Lx, Ly = float(nx), float(ny)
    if kind == 'Q4perm':
        # a rectangle_Q4 mesh renumbered at random: nodes still fill a lattice, but nothing is x-fastest any more
        m0 = F.mesh.rectangle_Q4(Lx, Ly, nx, ny)
        rng = np.random.default_rng(seed)
        perm = rng.permutation(m0.N_nodes)                 # new id of old node k
        pts = np.empty_like(m0.points); pts[perm] = m0.points
        cells = perm[m0.cells['quad']][rng.permutation(m0.N_ele)]
        mesh = F.Mesh(pts, {'quad': cells}, quad_element=F.elements.Q4)
    else:
        mesh = {'T3': F.mesh.rectangle_T3, 'T6': F.mesh.rectangle_T6}[kind](Lx, Ly, nx, ny)
    p = F.FEM_Problem('displacement', model, mesh)
    if model == 'plane':
        p.define_materials(F.materials.general_isotropic()); p.define_tensors(1)
        p.add_fixed_domain(F.domains.inside_box([0], [[0, Ly]]), ddls=[0, 1])
        p.add_forces([0, -1.0], F.domains.inside_box([Lx], [[0, Ly]]))
    else:
        p.define_materials(F.materials.IsotropicMaterial(10920, 0.3, 1)); p.define_tensors(0.5)
        p.add_fixed_domain(F.domains.inside_box([0], [[0, Ly]]))
        p.add_forces([0, 0, -1e-2, 0, 0, 0], F.domains.inside_box([Lx], [[0, Ly]]))
    x = cases.density_field(mesh.N_ele, seed=seed + 1)
    return p, x


@pytest.mark.parametrize('kind,nx,ny,model', [('T6', 24, 12, 'plane'), ('T6', 17, 9, 'plane'), ('T3', 40, 20, 'plane'),
                                              ('T3', 33, 15, 'plane'), ('Q4perm', 31, 16, 'plane'), ('T6', 10, 6, 'plate'),
                                              ('Q4perm', 12, 9, 'plate')])
def test_lattice_gmg_on_triangle_and_renumbered_meshes(F, kind, nx, ny, model):
    """Lattice hierarchy (femol_gmg_create_lattice): level 0 is the assembled operator of a T3 / T6 mesh or of a Q4
    mesh in arbitrary numbering, coarse levels are bilinear Q4 grids.  Same solution as block-Jacobi PCG and scipy,
    in far fewer iterations; 'auto' selects it."""
    import scipy.sparse.linalg as sla
    p, x = _lattice_problem(F, kind, nx, ny, model)
    ctype = p.mesh.contains[0]
    assert p.structured_grid_shape() is None and p.node_lattice() is not None
    assert p._effective_preconditioner() == 'gmg'
    p.solver_rtol = 1e-9 if model == 'plane' else 1e-7
    p.assemble('K', X={ctype: x}, p=3)
    p.preconditioner = 'block_jacobi'
    U_bj = p.solve_device().cpu().numpy().copy()
    it_bj = p.solver_info['iterations']
    p.preconditioner = 'auto'
    p._U_dev = None
    U_mg = p.solve_device().cpu().numpy().copy()
    info = dict(p.solver_info)
    assert info['converged'] and info['preconditioner'] == 'gmg', info
    assert info['iterations'] < (max(70, it_bj // 4) if model == 'plane' else it_bj // 2), (info, it_bj)
    U_ref = sla.spsolve(p.K.tocsc(), p.F)
    c_ref = p.F @ U_ref
    assert abs(p.F @ U_mg - c_ref) / abs(c_ref) < 1e-6 and abs(p.F @ U_bj - c_ref) / abs(c_ref) < 1e-6
    assert rel(U_mg, U_ref) < 1e-4, (rel(U_mg, U_ref), info)
    # the hierarchy: lattice on level 0, every second lattice line below
    info64 = (ctypes.c_int64 * 64)()
    F._lib.call('femol_gmg_info', p._gmg, info64)
    lat = p.node_lattice()
    assert (info64[1], info64[2]) == (lat['Nx'], lat['Ny']) and (info64[3], info64[4]) == ((lat['Nx'] + 1) // 2, (lat['Ny'] + 1) // 2)


def test_lattice_vcycle_is_symmetric_positive_and_level1_is_galerkin(F):
    """The lattice V-cycle is a symmetric positive operator, and its level-1 operator is exactly P^T K P with the
    bilinear lattice interpolation (checked through the two-level identity on a T6 mesh)."""
    import torch
    p, x = _lattice_problem(F, 'T6', 12, 7)
    p.assemble('K', X={'triangle': x}, p=3)
    g = p._gmg_ready()
    d = p.device
    rng = np.random.default_rng(0)
    fixed = p.fixed_dofs()
    R = rng.standard_normal((3, d.n)); R[:, fixed] = 0
    Z = []
    for r in R:
        rd = torch.from_numpy(r).cuda(); zd = torch.empty_like(rd)
        F._lib.call('femol_gmg_apply', g, rd.data_ptr(), zd.data_ptr(), F._lib.stream())
        Z.append(zd.cpu().numpy())
    assert abs(R[0] @ Z[1] - R[1] @ Z[0]) < 1e-10 * abs(R[0] @ Z[1]) + 1e-14 * np.linalg.norm(Z[0]) * np.linalg.norm(R[1])
    assert all(R[i] @ Z[i] > 0 for i in range(3))
    # the same cycle in numpy with Galerkin operators P^T K P on every level must give the same z
    import scipy.sparse as sp
    K = p.K.tocsr()
    lat = p.node_lattice()
    z_ref = _numpy_lattice_vcycle(K, lat, fixed, R[0], p.gmg_omega)
    # 1e-6: the coarse levels' smoother products read single-precision copies of the operators (FEMOL_GMG_F32)
    assert rel(Z[0], z_ref) < 1e-6, rel(Z[0], z_ref)


from numpy_ops import numpy_lattice_vcycle as _numpy_lattice_vcycle  # noqa: E402


def test_gmg_rejects_unstructured(F):
    p = cases.build_problem(F, cases.CASES['comp_q4_jitter_10x5'])
    p.preconditioner = 'gmg'
    with pytest.raises(F._lib.FemolError):
        p.solve(verbose=False)
    p.preconditioner = 'auto'
    p.solve(verbose=False)
    assert p.solver_info['preconditioner'] == 'block_jacobi'


@pytest.mark.parametrize('nx,ny,model,coating,scaled', [(32, 16, 'plane', False, True), (33, 17, 'plane', False, False),
                                                        (45, 22, 'plane', True, True), (1, 1, 'plane', False, True),
                                                        (24, 12, 'plate', False, True), (21, 13, 'plate', False, False),
                                                        (7, 1, 'plate', False, True), (300, 211, 'plane', False, True),
                                                        # sizes around the 32 x 16 (plane) / 32 x 8 (plate) node tiles and the 4-tile CTA groups
                                                        (31, 15, 'plane', False, True), (64, 33, 'plane', True, True), (65, 31, 'plane', False, False),
                                                        (127, 5, 'plane', False, True), (31, 7, 'plate', False, True), (33, 9, 'plate', False, False),
                                                        (95, 16, 'plate', False, True)])
def test_matrix_free_operator_equals_assembled(F, nx, ny, model, coating, scaled):
    """femol_apply (the stencil the solvers use on structured grids, csrc/stencil.cu) against the assembled CSR
    product and scipy: same K including the Dirichlet rows / columns (FEM.py:902-930)."""
    import torch
    p, x = _gmg_problem(F, nx, ny, model, coating)
    p.assemble('K', X={'quad': x}, p=3) if scaled else p.assemble('K')
    d = p.device
    rng = np.random.default_rng(3)
    xs = rng.standard_normal(d.n)
    xd = torch.from_numpy(xs).cuda()
    y_mf = torch.empty_like(xd); y_csr = torch.empty_like(xd)
    F._lib.call('femol_apply', d.plan, p.K_values.data_ptr(), xd.data_ptr(), y_mf.data_ptr(), F._lib.stream())
    F._lib.call('femol_spmv', d.plan, p.K_values.data_ptr(), xd.data_ptr(), y_csr.data_ptr(), F._lib.stream())
    ref = p.K @ xs
    assert rel(y_csr.cpu().numpy(), ref) < 1e-13
    assert rel(y_mf.cpu().numpy(), ref) < 1e-13
    assert np.array_equal(y_mf.cpu().numpy()[p.fixed_dofs()], xs[p.fixed_dofs()])       # identity rows, exactly
    # a value array the descriptor was not declared for goes through the CSR kernel (bit-identical to femol_spmv)
    other = p.K_values.clone()
    y2 = torch.empty_like(xd)
    F._lib.call('femol_apply', d.plan, other.data_ptr(), xd.data_ptr(), y2.data_ptr(), F._lib.stream())
    assert torch.equal(y2, y_csr)
    # opting out clears the descriptor
    p.matrix_free = False
    p.assemble('K', X={'quad': x}, p=3) if scaled else p.assemble('K')
    F._lib.call('femol_apply', d.plan, p.K_values.data_ptr(), xd.data_ptr(), y2.data_ptr(), F._lib.stream())
    assert torch.equal(y2, y_csr)


def test_matrix_free_and_assembled_solves_agree(F):
    """PCG (multigrid and block-Jacobi) gives the same displacement with either operator."""
    out = {}
    for mf in (True, False):
        p, x = _gmg_problem(F, 60, 31, 'plane')
        p.matrix_free = mf
        p.solver_rtol = 1e-10
        p.assemble('K', X={'quad': x}, p=3)
        for pc in ('gmg', 'block_jacobi'):
            p.preconditioner = pc
            p._U_dev = None
            out[mf, pc] = p.solve_device().cpu().numpy().copy()
            assert p.solver_info['converged']
    for pc in ('gmg', 'block_jacobi'):
        assert rel(out[True, pc], out[False, pc]) < 1e-8


@pytest.mark.parametrize('nx,ny,model,matrix_free', [(40, 21, 'plane', True), (40, 21, 'plane', False), (19, 9, 'plate', True),
                                                     (1, 1, 'plane', True), (130, 70, 'plane', True)])
def test_product_epilogues(F, nx, ny, model, matrix_free):
    """The fused multigrid epilogues of the product (residual and damped block-Jacobi sweep) against numpy."""
    import torch
    p, x = _gmg_problem(F, nx, ny, model)
    p.matrix_free = matrix_free
    p.assemble('K', X={'quad': x}, p=3)
    d = p.device
    nd = p.N_dof
    rng = np.random.default_rng(5)
    xs, rhs = rng.standard_normal(d.n), rng.standard_normal(d.n)
    xd, rd = torch.from_numpy(xs).cuda(), torch.from_numpy(rhs).cuda()
    dinv = torch.empty(d.n_nodes * nd * nd, dtype=torch.float64, device='cuda')
    F._lib.call('femol_precond_setup', d.plan, p.K_values.data_ptr(), 2, dinv.data_ptr(), F._lib.stream())
    K = p.K
    Kx = K @ xs
    y = torch.empty_like(xd)
    F._lib.call('femol_apply_epilogue', d.plan, p.K_values.data_ptr(), xd.data_ptr(), y.data_ptr(), 1, rd.data_ptr(), None, 0.0, F._lib.stream())
    assert rel(y.cpu().numpy(), rhs - Kx) < 1e-13
    D = dinv.cpu().numpy().reshape(d.n_nodes, nd, nd)
    Kd = K.toarray()
    for i in (0, d.n_nodes // 2, d.n_nodes - 1):    # Dinv really is the inverse of the node's diagonal block
        blk = Kd[i * nd:(i + 1) * nd, i * nd:(i + 1) * nd]
        assert np.allclose(D[i] @ blk, np.eye(nd), atol=1e-10)
    ref = xs + 0.7 * np.einsum('nab,nb->na', D, (rhs - Kx).reshape(-1, nd)).reshape(-1)
    F._lib.call('femol_apply_epilogue', d.plan, p.K_values.data_ptr(), xd.data_ptr(), y.data_ptr(), 2, rd.data_ptr(), dinv.data_ptr(), 0.7, F._lib.stream())
    assert rel(y.cpu().numpy(), ref) < 1e-13
    with pytest.raises(F._lib.FemolError):
        F._lib.call('femol_apply_epilogue', d.plan, p.K_values.data_ptr(), xd.data_ptr(), xd.data_ptr(), 0, None, None, 0.0, F._lib.stream())


def test_stress_recovery(F, golden):
    """Mesh.stress_from_displacement (mesh.py:387-425) against the reference's own output on a jittered Q4 mesh:
    plain tensor, tensor * X^3, base + coating * X^3; plus the call pattern after a solve and the oracle on a
    larger random field."""
    import torch
    g = golden('stress_q4')
    mesh = F.Mesh(g['points'], {'quad': g['cells']})
    mesh.add_displacement(g['U'], 2)
    for key, tensors, x in (('plain', (g['T_base'],), None), ('one', (g['T_base'],), g['x']),
                            ('two', (g['T_base'], g['T_coat']), g['x'])):
        mesh.cell_data.pop('X', None)
        if x is not None:
            mesh.cell_data['X'] = {'quad': x}
        mesh.stress_from_displacement(*tensors)
        for comp in ('Sx', 'Sy', 'Sxy', 'Sv'):
            assert rel(mesh.cell_data[comp]['quad'], g['%s_%s' % (key, comp)]) < 1e-12, (key, comp)
    # after a solve of the mirror API (plate problem: 6 dofs per node, Ux / Uy are the first two)
    problem = cases.build_problem(F, cases.CASES['plate_q4_jitter_6x5'])
    m2 = problem.solve(verbose=False)
    m2.stress_from_displacement(problem.tensors[0])
    ref = o.q4_stress(m2.points, np.asarray(m2.cells['quad']), np.stack([m2.point_data['Ux'], m2.point_data['Uy']], 1).reshape(-1),
                      problem.tensors[0])
    for comp, r in zip(('Sx', 'Sy', 'Sxy', 'Sv'), ref):
        assert rel(m2.cell_data[comp]['quad'], r) < 1e-11 or np.abs(r).max() < 1e-300
    # T3 / T6 meshes have no stress method in the reference
    with pytest.raises(AttributeError):
        m3 = F.mesh.rectangle_T3(2., 1., 2, 1)
        m3.add_displacement(np.zeros(2 * m3.N_nodes), 2)
        m3.stress_from_displacement(g['T_base'])


def test_gmg_mid_size_plate_regular_levels(F):
    """96 x 64 Mindlin plate (38 k dof): the level-1 operator (4 MB) is above the cluster-kernel threshold, so this
    runs the regular-level path for 6 dofs per node (single-precision CSR copies, unfused post-smoothing) that the
    small cases never reach.  The solution must satisfy the ASSEMBLED double-precision system."""
    import torch
    p, _ = _gmg_problem(F, 96, 64, 'plate')
    # a smooth density field (like a filtered SIMP design): element-wise random densities with a 1e9 stiffness
    # contrast defeat any geometric coarsening (9 866 multigrid-CG iterations against 14 241 block-Jacobi ones here)
    i, j = np.divmod(np.arange(96 * 64), 64)
    x = 0.55 + 0.45 * np.sin(2 * np.pi * i / 96) * np.cos(2 * np.pi * j / 64)
    p.solver_rtol = 1e-8
    p.assemble('K', X={'quad': x}, p=3)
    p.preconditioner = 'gmg'
    U = p.solve_device()
    assert p.solver_info['converged'], p.solver_info
    d = p.device
    y = torch.empty_like(U)
    F._lib.call('femol_spmv', d.plan, p.K_values.data_ptr(), U.data_ptr(), y.data_ptr(), F._lib.stream())
    # attainable accuracy on this plate with densities down to 1e-3 (cond ~ 1e10) is a few 1e-7, with or without
    # the single-precision level copies (FEMOL_GMG_F32=0 gives 1.7e-7)
    assert float((y - p._F_dev).norm() / p._F_dev.norm()) < 1e-6
    it_mg = p.solver_info['iterations']
    p.preconditioner = 'block_jacobi'
    p._U_dev = None
    U_bj = p.solve_device()
    assert it_mg < p.solver_info['iterations'] // 2, (it_mg, p.solver_info)
    assert rel(U.cpu().numpy(), U_bj.cpu().numpy()) < 1e-4


def test_cluster_vcycle_kernel_opt_in(F, monkeypatch):
    """The thread-block-cluster kernel for the smallest multigrid levels is off by default (plain launches with
    programmatic dependent launch measured faster); FEMOL_GMG_CLUSTER_MB switches it on.  Same solution, same
    iteration count as the default path."""
    out = {}
    for mb in ('0', '0.7'):
        monkeypatch.setenv('FEMOL_GMG_CLUSTER_MB', mb)
        p, x = _gmg_problem(F, 90, 47, 'plane')
        p.solver_rtol = 1e-10
        p.assemble('K', X={'quad': x}, p=3)
        p.preconditioner = 'gmg'
        out[mb] = (p.solve_device().cpu().numpy().copy(), dict(p.solver_info))
        assert out[mb][1]['converged']
    assert abs(out['0'][1]['iterations'] - out['0.7'][1]['iterations']) <= 1
    assert rel(out['0'][0], out['0.7'][0]) < 1e-8


@pytest.mark.parametrize('nx,ny,model,coating', [(17, 9, 'plane', False), (12, 7, 'plane', True), (9, 5, 'plate', False), (1, 1, 'plane', False),
                                                 (64, 3, 'plane', False)])
def test_grid_assembly_is_bit_identical_to_list_assembly(F, nx, ny, model, coating):
    """femol_plan_set_grid: the assembly derives the scatter pattern from grid coordinates instead of reading the
    source lists; same contributions in the same order, so the values are bit-identical (K, M, with BCs, coating)."""
    import torch
    p, x = _gmg_problem(F, nx, ny, model, coating)
    d = p.device
    out = {}
    for grid in (True, False):
        F._lib.call('femol_plan_set_grid', d.plan, nx if grid else 0, ny if grid else 0)
        p.assemble('K', X={'quad': x}, p=3)
        p.assemble('M', X={'quad': x}, q=1)
        out[grid] = (p.K_values.clone(), p.M_values.clone())
    assert torch.equal(out[True][0], out[False][0]) and torch.equal(out[True][1], out[False][1])
    # a mesh whose cells are in another order is refused (the check runs on the device)
    mesh = F.mesh.rectangle_Q4(float(nx), float(ny), nx, ny)
    if nx * ny > 1:
        cells = np.asarray(mesh.cells['quad'])[::-1].copy()
        m2 = F.Mesh(mesh.points, {'quad': cells})
        p2 = F.FEM_Problem('displacement', 'plane', m2)
        assert p2.structured_grid_shape() is None
        with pytest.raises(F._lib.FemolError):
            F._lib.call('femol_plan_set_grid', p2.device.plan, nx, ny)


def test_modal_mid_size_against_shift_invert(F):
    """SURVEY 8(d) cfg 4 at 48 x 48 (14 k dof): NAFEMS FV13-like simply supported plate, lowest 6 modes by LOBPCG with
    the multigrid-CG preconditioner against scipy's shift-invert Lanczos (`eigsh(sigma=0)`, the reference's 'sparse'
    path, FEM.py:234) on the same assembled pencil restricted to the free, non-drilling dofs."""
    import scipy.sparse.linalg as sla
    n = 48
    mesh = F.mesh.rectangle_Q4(10., 10., n, n)
    p = F.FEM_Problem('modal', 'plate', mesh)
    p.define_materials(F.materials.IsotropicMaterial(200e9, 0.3, 8000)); p.define_tensors(0.05)
    p.add_fixed_domain(F.domains.inside_box([[0, 10.]], [[0, 10.]]), ddls=[0, 1, 5])
    p.add_fixed_domain(F.domains.inside_box([0, 10.], [[0, 10.]]), ddls=[2, 4])
    p.add_fixed_domain(F.domains.inside_box([[0, 10.]], [0, 10.]), ddls=[2, 3])
    p.n_modes = 6
    p.modal_tol = 1e-8
    w, v = p.solve(verbose=False)
    assert p.solver_info['iterations'] < p.modal_max_iter
    K, M = p.K, p.M
    keep = np.ones(K.shape[0], bool)
    keep[p.fixed_dofs()] = False
    keep[5::6] = False
    Kf = K[keep][:, keep].tocsc(); Mf = M[keep][:, keep].tocsc()
    lam = sla.eigsh(Kf, k=6, M=Mf, sigma=0, which='LM', return_eigenvectors=False)
    ref = np.sort(np.sqrt(lam) / (2 * np.pi))
    assert np.allclose(np.sort(w)[:6], ref, rtol=1e-6), (w, ref)
    # NAFEMS FV13 (thin-plate target values 2.377, 5.942, 5.942, 9.507, 11.884, 11.884 Hz): within 1 % on this mesh
    assert np.allclose(np.sort(w)[:6], [2.377, 5.942, 5.942, 9.507, 11.884, 11.884], rtol=1e-2)
    # M-orthonormal modes (test_modal_solver.py:7-22)
    G = v[:6] @ (M @ v[:6].T)
    assert np.allclose(G, np.eye(6), atol=1e-6)


@pytest.mark.parametrize('nx,ny,model', [(90, 47, 'plane'), (64, 64, 'plane'), (37, 19, 'plate')])
def test_fast_galerkin_coarsening_matches_generic(F, monkeypatch, nx, ny, model):
    """The register / shuffle Galerkin coarsening of the levels >= 1 against the generic shared-memory kernel
    (FEMOL_GMG_OLD_COARSEN=1): same hierarchy up to rounding -> same iteration count, same solution.  Odd sizes
    exercise the half-width parents, the clamped edge the Dirichlet masking."""
    out = {}
    for old in ('1', '0'):
        monkeypatch.setenv('FEMOL_GMG_OLD_COARSEN', old)
        p, _ = _gmg_problem(F, nx, ny, model)
        i, j = np.divmod(np.arange(nx * ny), ny)
        x = 0.55 + 0.45 * np.sin(2 * np.pi * i / nx) * np.cos(2 * np.pi * j / ny)
        p.solver_rtol = 1e-9 if model == 'plane' else 1e-7
        p.assemble('K', X={'quad': x}, p=3)
        p.preconditioner = 'gmg'
        out[old] = (p.solve_device().cpu().numpy().copy(), dict(p.solver_info))
        assert out[old][1]['converged']
    assert abs(out['0'][1]['iterations'] - out['1'][1]['iterations']) <= 1, (out['0'][1], out['1'][1])
    assert rel(out['0'][0], out['1'][0]) < 1e-6


def test_variable_core_laminate_solves(F, golden):
    """`solve(solve_from_zc=True, h_min)` (FEM.py:201-256, 697-731, 873-896): K and M of per-element shifted laminates
    against the reference's matrices, the static solution against its `spsolve`, the lowest modes against the dense
    eigen-solve of the same pencil."""
    import scipy.linalg
    g = golden('variable_core')
    case = dict(cases.CASES['modal_q4_plate_coat_6x6'], mesh=('Q4J', 1., 1., 6, 6, 0.15))
    h_min = float(g['h_min'])
    p = cases.build_problem(F, case)
    p.mesh.cell_data['zc'] = {'quad': g['zc']}
    p.n_modes = 4
    p.modal_tol = 1e-8
    w, v = p.solve(verbose=False, solve_from_zc=True, h_min=h_min)
    for which, A in (('K', p.K), ('M', p.M)):
        assert np.array_equal(A.indptr, g[which + '_indptr']) and np.array_equal(A.indices, g[which + '_indices'])
        assert rel(A.data, g[which + '_data']) < 1e-12, which
    keep = np.ones(p.K.shape[0], bool)
    keep[p.fixed_dofs()] = False
    keep[5::6] = False
    lam = scipy.linalg.eigh(p.K[keep][:, keep].toarray(), p.M[keep][:, keep].toarray(), eigvals_only=True, subset_by_index=[0, 3])
    assert np.allclose(np.sort(w)[:4], np.sqrt(lam) / (2 * np.pi), rtol=1e-6), (w, np.sqrt(lam) / (2 * np.pi))
    for fi in np.sort(w)[:4]:                       # and they are in the spectrum the reference returned
        assert np.min(np.abs(g['w'] - fi) / fi) < 1e-4
    # static: same K (assembled from zc because none exists yet), the reference's load
    case_s = dict(case, physics='displacement', forces=[([0, 0, -5.0, 0, 0, 0], ([1.], [[0, 1.]]))])
    ps = cases.build_problem(F, case_s)
    ps.mesh.cell_data['zc'] = {'quad': g['zc']}
    ps.solver_rtol = 1e-11
    U = ps.solve(verbose=False, solve_from_zc=True, h_min=h_min).U
    assert ps.solver_info['preconditioner'] == 'block_jacobi'
    assert rel(ps.K.data, g['K_data']) < 1e-12
    r = np.linalg.norm(ps.K @ g['U'] - ps.F) / np.linalg.norm(ps.F)          # the reference's own residual level
    assert np.linalg.norm(ps.K @ U - ps.F) / np.linalg.norm(ps.F) < max(1e-8, 10 * r)
    assert rel(U, g['U']) < 1e-5
    # density -> core height through the mirror class
    vibe = F.SIMP_VIBE(cases.build_problem(F, case), volfrac=0.5, p=3, q=1, rmin=1.5)
    vibe.mesh.cell_data['X'] = {'quad': g['X_dens']}
    assert np.allclose(vibe.density_to_core_height().cell_data['zc']['quad'], g['zc_from_density'], rtol=1e-13, atol=0)


def test_result_files(F, golden, tmp_path, monkeypatch):
    """`save=True` of the optimisers (TOM_SIMP.py:108-116, TOM.py:134-152): the VTK result written through Mesh.save is
    read back by load_vtk with the final densities; SIMP_VIBE also stores the tracked eigenvalue history."""
    import glob
    from femol_b200.mesh import load_vtk
    monkeypatch.chdir(tmp_path)
    problem = cases.build_problem(F, cases.CASES['comp_q4_30x15'])
    opt = F.SIMP_COMP(problem, volfrac=0.4, penal=3, rmin=1.5)
    mesh = opt.solve(converge=0.0, max_iter=2, plot=False, save=True)
    files = glob.glob(str(tmp_path / 'TOM_*.vtk'))
    assert len(files) == 1
    back = load_vtk(files[0])
    assert np.array_equal(back.cell_data['X']['quad'], mesh.cell_data['X']['quad'])
    assert np.array_equal(back.point_data['d2_Ux'], mesh.point_data['d2_Ux'])
    case, g = cases.CASES['vibe_q4_plate_10x10'], golden('vibe_q4_plate_10x10')
    vibe = F.SIMP_VIBE(cases.build_problem(F, case), volfrac=0.5, p=3, q=1, rmin=1.5, saving_root=str(tmp_path) + '/vibe_')
    m2 = vibe.solve(g['v_ref'], converge=0.0, min_iter=1, max_iter=2, plot=False, save=True, verbose=False,
                    mesh_filename='mesh', eigvals_filename='eigs')
    back = load_vtk(str(tmp_path / 'vibe_mesh.vtk'))
    assert np.allclose(back.cell_data['X_real']['quad'], m2.cell_data['X']['quad'] ** 3, rtol=1e-15)
    assert np.array_equal(np.load(str(tmp_path / 'vibe_eigs.npy')), np.array(vibe.lmbds))


def _cantilever_plate(F, n, t=0.25):
    mesh = F.mesh.rectangle_Q4(1.0, 1.0, n, n)
    p = F.FEM_Problem('displacement', 'plate', mesh)
    p.define_materials(F.materials.isotropic_bending_benchmark()); p.define_tensors(t)
    p.add_fixed_domain(F.domains.inside_box([0], [[0, 1.]]))
    p.add_forces([0, 0, -1.0, 0, 0, 0], F.domains.inside_box([1.], [[0.45, 0.55]]))
    p.solver_rtol = 1e-8
    return p


def test_hourglass_hierarchy_on_plates_with_free_edges(F):
    """The reference's Q4 Mindlin plate (shear at one point, elements.py:1029-1054) has a low-energy family
    w = (-1)^(i+j) x smooth envelope that only the boundary conditions restrain: on a cantilever plate multigrid-CG
    needs O(n) iterations.  With the second hierarchy on S K S (femol_gmg_attach_modulated, 'auto' on plates with free
    edges and a thickness of at least 8 element sizes) the count is mesh-independent, the preconditioner stays
    symmetric positive and the solution is the same."""
    import torch
    its = {}
    for hg in (False, 'auto'):
        for n in (40, 80):
            p = _cantilever_plate(F, n)
            p.gmg_hourglass = hg
            U = p.solve_device()
            assert p.solver_info['converged'] and p.solver_info['preconditioner'] == 'gmg'
            its[hg, n] = p.solver_info['iterations']
            if n == 80:
                sol = {**locals().get('sol', {}), hg: U.cpu().numpy().copy()}
    assert its['auto', 80] <= its['auto', 40] + 8, its                      # mesh-independent
    assert its[False, 80] > 1.5 * its[False, 40] and its['auto', 80] < its[False, 80] // 2, its
    assert rel(sol['auto'], sol[False]) < 1e-5
    # symmetric positive preconditioner with the second cycle attached
    p = _cantilever_plate(F, 24)
    p.gmg_hourglass = True            # t / h = 6 here: below the 'auto' threshold
    p.assemble('K')
    g = p._gmg_ready()
    assert p._gmg2 is not None
    thin = _cantilever_plate(F, 24, t=0.01)
    thin.assemble('K'); thin._gmg_ready()
    assert thin._gmg2 is None         # 'auto' stays off far below 8 element sizes
    rng = np.random.default_rng(2)
    R = rng.standard_normal((2, p.device.n)); R[:, p.fixed_dofs()] = 0
    Z = []
    for r in R:
        rd = torch.from_numpy(r).cuda(); zd = torch.empty_like(rd)
        F._lib.call('femol_gmg_apply', g, rd.data_ptr(), zd.data_ptr(), F._lib.stream())
        Z.append(zd.cpu().numpy())
    assert abs(R[0] @ Z[1] - R[1] @ Z[0]) < 1e-9 * abs(R[0] @ Z[1]) + 1e-13 * np.linalg.norm(Z[0]) * np.linalg.norm(R[1])
    assert R[0] @ Z[0] > 0 and R[1] @ Z[1] > 0
    # fully supported plates keep the single hierarchy under 'auto'
    ps, _ = _gmg_problem(F, 24, 12, 'plate')
    ps.add_fixed_domain(F.domains.inside_box([[0, 24.]], [0, 12.]), ddls=[2]); ps.add_fixed_domain(F.domains.inside_box([0, 24.], [[0, 12.]]), ddls=[2])
    ps.assemble('K'); ps._gmg_ready()
    assert ps._gmg2 is None


@pytest.mark.parametrize('kind,nx,ny', [('Q4', 40, 21), ('Q4', 63, 1), ('Q4', 130, 70), ('T6', 24, 12), ('T6', 31, 17), ('T3', 33, 15)])
def test_single_precision_products(F, kind, nx, ny):
    """femol_to_f32 / femol_apply_f32 (the operator copies the multigrid levels multiply with; rows of 4 - 19 node blocks)
    against the FP64 matrix rounded to single precision: y = K x, y = rhs - K x, y = x + omega Dinv (rhs - K x)."""
    import torch
    if kind == 'Q4':
        p, x = _gmg_problem(F, nx, ny, 'plane')
        p.assemble('K', X={'quad': x}, p=3)
    else:
        p, x = _lattice_problem(F, kind, nx, ny)
        p.assemble('K', X={'triangle': x}, p=3)
    d = p.device
    L = F._lib
    K = p.K
    K32 = K.copy(); K32.data = K32.data.astype(np.float32).astype(np.float64)
    v32 = torch.empty(d.nnz, dtype=torch.float32, device='cuda')
    L.call('femol_to_f32', d.plan, p.K_values.data_ptr(), v32.data_ptr(), L.stream())
    rng = np.random.default_rng(11)
    xs, rhs = rng.standard_normal(d.n), rng.standard_normal(d.n)
    xd, rd = torch.from_numpy(xs).cuda(), torch.from_numpy(rhs).cuda()
    dinv = torch.empty(d.n_nodes * 4, dtype=torch.float64, device='cuda')
    L.call('femol_precond_setup', d.plan, p.K_values.data_ptr(), 2, dinv.data_ptr(), L.stream())
    D = dinv.cpu().numpy().reshape(d.n_nodes, 2, 2)
    Kx = K32 @ xs
    refs = [Kx, rhs - Kx, xs + 0.7 * np.einsum('nab,nb->na', D, (rhs - Kx).reshape(-1, 2)).reshape(-1)]
    for mode in range(3):
        y = torch.full_like(xd, float('nan'))
        L.call('femol_apply_f32', d.plan, v32.data_ptr(), xd.data_ptr(), y.data_ptr(), mode, rd.data_ptr(), dinv.data_ptr(), 0.7, L.stream())
        assert rel(y.cpu().numpy(), refs[mode]) < 1e-13, mode
    with pytest.raises(L.FemolError):
        L.call('femol_apply_f32', d.plan, v32.data_ptr(), xd.data_ptr(), xd.data_ptr(), 0, None, None, 0.0, L.stream())


@pytest.mark.parametrize('model', ['plane', 'plate'])
def test_gmg_workspace_views(F, model):
    """femol_gmg_clone_workspace: a view applies the same V-cycle as its hierarchy (bit for bit: same kernels, same
    operators), follows a later set-up for new densities, and two views on two streams run concurrently without
    disturbing each other (LOBPCG solves the columns of its preconditioner this way)."""
    import torch
    p, x = _gmg_problem(F, 41, 23, model)
    p.assemble('K', X={'quad': x}, p=3)
    g = p._gmg_ready()
    L = F._lib
    d = p.device
    views = []
    for _ in range(2):
        h = ctypes.c_void_p()
        L.call('femol_gmg_clone_workspace', g, ctypes.byref(h))
        views.append(h)
    try:
        rng = np.random.default_rng(2)
        R = [torch.from_numpy(rng.standard_normal(d.n)).cuda() for _ in range(2)]

        def apply(h, r, stream=None):
            z = torch.empty_like(r)
            L.call('femol_gmg_apply', h, r.data_ptr(), z.data_ptr(), stream if stream is not None else L.stream())
            return z

        for round_ in range(2):
            ref = [apply(g, r) for r in R]
            torch.cuda.synchronize()
            streams = [torch.cuda.Stream() for _ in views]
            out = []
            for h, r, s in zip(views, R, streams):
                with torch.cuda.stream(s):
                    out.append(apply(h, r, s.cuda_stream))
            torch.cuda.synchronize()
            for a, b in zip(ref, out):
                assert torch.equal(a, b)
            # new densities: the hierarchy is refreshed, the views follow
            x2 = cases.density_field(p.mesh.N_ele, seed=9 + round_)
            p.assemble('K', X={'quad': x2}, p=3)
            p._gmg_ready()
        with pytest.raises(L.FemolError):
            L.call('femol_gmg_clone_workspace', views[0], ctypes.byref(ctypes.c_void_p()))    # no views of views
    finally:
        for h in views:
            L.load().femol_gmg_destroy(h)
