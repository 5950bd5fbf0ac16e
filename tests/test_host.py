"""Host-side logic and the C-ABI surface, without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import cases
import femol_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports every function include/femol_b200.h declares."""
    from femol_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'femol_b200.h')).read()
    declared = set(re.findall(r'\b(femol_[a-z0-9_]+)\s*\(', header))
    declared -= {'femol_plan'}
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b'sm_100a' in _lib.load().femol_version()
    assert _lib.load().femol_reduce_workspace_doubles() > 0


def test_no_cpu_fallback():
    """Compute entry points fail loudly without a CUDA device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import femol_b200 as F
    mesh = F.mesh.rectangle_Q4(4, 2, 4, 2)
    p = F.FEM_Problem('displacement', 'plane', mesh)
    p.define_materials(F.materials.general_isotropic())
    p.define_tensors(1)
    with pytest.raises(F._lib.FemolError):
        p.assemble('K')
    with pytest.raises(F._lib.FemolError):
        F.elements.Q4(np.array([[0, 0], [1, 0], [1, 1], [0, 1.]]), 2).Ke(np.eye(3))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'femol_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'femol_oracle' not in text and 'import oracle' not in text and '/root/reference' not in text, f


def test_laminate_tensors_match_reference(golden):
    import femol_b200 as F
    g = golden('laminate_tensors')
    specs = [('general_carbon', [0, 90, 45, -45], True, 0, 0), ('general_carbon', [45, 45], True, 0, 0),
             ('T300_N5208', [0, 0, 0, 20, -20, 90], True, 0, 0.01), ('T300_N5208', [0, 0, 0, 20, -20, 90], False, 0, 0),
             ('general_flax', [0, 90, 0, 90], True, 0, 0), ('general_carbon', [0, 90], False, 0.002, 0),
             ('abaqus_benchmark', [0, 90, 0, 90, 0, 90, 0], False, 0, 0), ('AS4_PEEK', [10, -35, 60], True, -0.001, 0)]
    for i, (m, plies, sym, zc, hc) in enumerate(specs):
        L = F.laminate.Layup(getattr(F.materials, m)(), plies=plies, symetric=sym, z_core=zc, h_core=hc)
        for k in 'ABDG':
            assert np.array_equal(getattr(L, k + '_mat'), g[k + str(i)]), (i, k)
        assert L.hA == g['hA%d' % i] and np.array_equal(L.z_values, g['z%d' % i])


@pytest.mark.parametrize('name', list(cases.CASES))
def test_mesh_and_problem_setup_match_reference(golden, name):
    """points, cells (incl. the T6 mid-node numbering), tensors, fixed dofs and F of every case."""
    import femol_b200 as F
    case, g = cases.CASES[name], golden(name)
    problem = cases.build_problem(F, case)
    mesh = problem.mesh
    assert np.array_equal(mesh.points, g['points'])
    assert np.array_equal(mesh.cells[mesh.contains[0]], g['cells'])
    for k, T in enumerate(problem.tensors):
        assert np.array_equal(np.asarray(T, float), g['T%d' % k])
    if problem.coating:
        for k, T in enumerate(problem.coat_tensors):
            assert np.array_equal(np.asarray(T, float), g['Tc%d' % k])
    arr = o.case_arrays(case, cases.jitter_points)
    assert np.array_equal(problem.fixed_dofs(), arr['fixed_idx'])
    if 'F' in g:
        Fv = problem.F.copy(); Fv[problem.fixed_dofs()] = 0
        assert np.array_equal(Fv, g['F'])


def test_mesh_helpers():
    import femol_b200 as F
    m = F.mesh.rectangle_Q4(6., 3., 12, 6)
    assert m.structured and m.N_ele == 72 and m.N_nodes == 13 * 7
    assert np.isclose(m.element_size(), 0.5)
    assert np.allclose(m.areas['quad'], 0.25)
    assert np.allclose(m.all_cell_centers[0], [0.25, 0.25]) and np.allclose(m.all_cell_centers[1], [0.25, 0.75])
    m.global_matrix_indexes(2)
    r, c = o.coo_pattern(m.cells['quad'], 2)
    assert np.array_equal(m.rows, r) and np.array_equal(m.cols, c)
    # unused nodes are dropped and cells renumbered (mesh.py:209-240)
    pts = np.array([[9., 9.], [0, 0], [1, 0], [1, 1], [0, 1]])
    mm = F.Mesh(pts, {'quad': np.array([[1, 2, 3, 4]])}, quad_element=F.elements.Q4)
    assert mm.N_nodes == 4 and np.array_equal(mm.cells['quad'], [[0, 1, 2, 3]])
    # inverted quads are flipped (mesh.py:105-109)
    mm = F.Mesh(pts[1:], {'quad': np.array([[0, 3, 2, 1]])}, quad_element=F.elements.Q4)
    assert (mm.areas['quad'] > 0).all()
    # scalar (reference-style) and vectorised domain predicates agree
    dom = F.domains.inside_box([0], [[0, 1.5]])
    scalar = lambda x, y: dom(x, y)
    assert np.array_equal(m.domain_nodes(dom), m.domain_nodes(scalar))
    assert dom(0.0, 1.0) is True and dom(1.0, 1.0) is False


def test_filter_stencil_construction():
    import torch
    import femol_b200 as F
    from femol_b200.TOM_SIMP import build_filter
    m = F.mesh.rectangle_Q4(8., 3., 16, 6)   # h = 0.5
    f = build_filter(m, 1.5, torch.device('cpu'))
    assert f['kind'] == 'grid' and (f['n_slow'], f['n_fast'], f['rx'], f['ry']) == (16, 6, 1, 1)
    w = f['w'].numpy()
    assert w[1, 1] == 0 and np.isclose(w[0, 1], 0.5) and np.isclose(w[0, 0], 1.5 - np.sqrt(2))
    # stencil == literal reference filter
    rng = np.random.default_rng(0)
    x = rng.uniform(0.01, 1, m.N_ele); dc = -rng.uniform(0, 1, m.N_ele)
    ref = o.filter_sens_literal(x, dc, m.all_cell_centers, 1.5, m.element_size())
    X = x.reshape(16, 6); D = (x * dc).reshape(16, 6)
    num = np.zeros((16, 6)); den = np.zeros((16, 6))
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            ww = w[di + 1, dj + 1]
            for i in range(16):
                for j in range(6):
                    if 0 <= i + di < 16 and 0 <= j + dj < 6:
                        num[i, j] += ww * D[i + di, j + dj]; den[i, j] += ww
    assert np.allclose((num / (X * den)).reshape(-1), ref, rtol=1e-13)
    ft = build_filter(F.mesh.rectangle_T3(4., 2., 4, 2), 1.5, torch.device('cpu'))
    assert ft['kind'] == 'list'


def test_lobpcg_driver_on_numpy_ops(golden):
    """The LOBPCG outer iteration (femol_b200/lobpcg.py) against the dense generalised eigen-solve."""
    from femol_b200.lobpcg import lobpcg, pick_block_size
    from numpy_ops import NumpyOps
    case = cases.CASES['modal_q4_plate_ss_8x8']
    arr = o.case_arrays(case)
    K = o.global_matrices(arr, 'K'); M = o.global_matrices(arr, 'M')
    w_ref, _ = o.modal_solve_dense(K, M)
    w_ref = w_ref[np.abs(w_ref - 1 / (2 * np.pi)) > 1e-3]
    k = 6
    m = pick_block_size(k)
    assert m == 8
    X0 = np.random.default_rng(0).standard_normal((K.shape[0], m))
    ops = NumpyOps(K, M, arr['fixed_idx'], m, X0, precond='exact')
    lam, rel, its = lobpcg(ops, m, k, tol=1e-9, max_iter=60)
    assert np.all(rel[:k] <= 1e-9) and its < 40, (its, rel)
    f = np.sqrt(lam[:k]) / (2 * np.pi)
    assert np.allclose(f, w_ref[:k], rtol=1e-7), (f, w_ref[:k])
    V = ops.vectors()[:, :k]
    assert np.allclose(V.T @ (M @ V), np.eye(k), atol=1e-8)      # M-orthonormal like the reference (test_modal_solver.py:7-22)
    assert np.abs(V[arr['fixed_idx']]).max() == 0
    # the Jacobi-preconditioned iteration converges too (slowly: the Mindlin plate pencil is stiff)
    ops = NumpyOps(K, M, arr['fixed_idx'], m, X0, precond='jacobi')
    lam, rel, its = lobpcg(ops, m, k, tol=1e-6, max_iter=1500, stall_window=10 ** 6)   # slow is not stalled
    assert np.all(rel[:k] <= 1e-6), (its, rel)
    assert np.allclose(np.sqrt(lam[:k]) / (2 * np.pi), w_ref[:k], rtol=1e-6)


def test_modal_filters_match_reference(golden):
    """FEM_Problem._filter_eigenvalues_1/2 (FEM.py:335-418) on a synthetic spectrum against the reference's output."""
    import femol_b200 as F
    g = golden('modal_filters')
    mesh = F.mesh.rectangle_Q4(2., 2., 2, 2)
    prob = F.FEM_Problem('modal', 'plate', mesh)
    for k, fn in ((1, prob._filter_eigenvalues_1), (2, prob._filter_eigenvalues_2)):
        w, v = fn(g['w'].copy(), g['v'].copy())
        assert np.array_equal(w.shape, g['w%d' % k].shape)
        assert np.allclose(w, g['w%d' % k], rtol=1e-14) and np.array_equal(v, g['v%d' % k])
    plane = F.FEM_Problem('modal', 'plane', mesh)
    with pytest.raises(TypeError):                      # the reference's mesh-mode filter returns None for plane models
        plane._filter_eigenvalues_1(g['w'].copy(), g['v'][:, :18].copy())


def test_variable_core_host_pieces(golden):
    """Host side of the variable-core laminate path: the vectorised per-element tensors (shift formulas) against
    the oracle's per-element laminates, and SIMP_VIBE.density_to_core_height (TOM.py:497-540) against the reference."""
    import types
    import femol_b200 as F
    from femol_b200.elements import pack_tensors
    g = golden('variable_core')
    case = dict(cases.CASES['modal_q4_plate_coat_6x6'], mesh=('Q4J', 1., 1., 6, 6, 0.15))
    p = cases.build_problem(F, case)
    zc, h_min = g['zc'], float(g['h_min'])
    Tb, Tc, above = p._variable_core_tensors(zc, h_min)
    par = [o.MATERIAL_LIB['general_flax'], o.MATERIAL_LIB['general_carbon']]
    lb, lc = p.layups

    def lam(k, lay, z):
        q = par[k]
        return pack_tensors(o.layup_tensors(q['Ex'], q['Ey'], q['Es'], q['vx'], q['ho'], lay.angles, False, z, 0,
                                            q.get('Gyz'), q.get('Gxz'))[0])
    for e in range(zc.shape[0]):
        h = zc[e] + lb.hA + lc.hA
        if above[e]:
            assert np.allclose(Tb[e], lam(0, lb, -h / 2 + lb.hA / 2), rtol=1e-12, atol=1e-12 * np.abs(Tb[e]).max())
            assert np.allclose(Tc[e], lam(1, lc, h / 2 - lc.hA / 2), rtol=1e-12, atol=1e-12 * np.abs(Tc[e]).max())
        else:
            assert np.allclose(Tb[e], lam(0, lb, 0.0), rtol=1e-12, atol=1e-12 * np.abs(Tb[e]).max()) and not Tc[e].any()
    p.mesh.cell_data['X'] = {'quad': g['X_dens']}
    fake = types.SimpleNamespace(FEM=p, mesh=p.mesh, p=3)
    mesh = F.SIMP_VIBE.density_to_core_height(fake)
    assert np.allclose(mesh.cell_data['zc']['quad'], g['zc_from_density'], rtol=1e-13, atol=0)


@pytest.mark.parametrize('kind', ['Q4', 'T3', 'T6'])
def test_vtk_round_trip(tmp_path, kind):
    """Mesh.save / load_vtk (mesh.py:242-267, 827-844, without meshio): points, cells (6-node triangles through the
    VTK node order and back), point data and cell data survive a round trip; the file is a legacy-VTK grid."""
    import femol_b200 as F
    from femol_b200.mesh import load_vtk
    mesh = {'Q4': F.mesh.rectangle_Q4, 'T3': F.mesh.rectangle_T3, 'T6': F.mesh.rectangle_T6}[kind](3., 2., 3, 2)
    rng = np.random.default_rng(0)
    ctype = mesh.contains[0]
    mesh.add_displacement(rng.standard_normal(2 * mesh.N_nodes), 2)
    mesh.cell_data['X'] = {ctype: rng.uniform(0, 1, mesh.N_ele)}
    mesh.cell_data['zc'] = {ctype: rng.uniform(0, 1, mesh.N_ele)}
    path = str(tmp_path / 'result')
    mesh.save(path)
    head = open(path + '.vtk', 'rb').read(80)
    assert head.startswith(b'# vtk DataFile Version') and b'BINARY' in head
    back = load_vtk(path + '.vtk')
    assert np.array_equal(back.points, mesh.points)
    assert np.array_equal(np.asarray(back.cells[ctype]), np.asarray(mesh.cells[ctype]))
    assert back.ElementClasses[ctype] is mesh.ElementClasses[ctype]
    for k in ('Ux', 'Uy'):
        assert np.array_equal(back.point_data[k], mesh.point_data[k])
    for k in ('X', 'zc'):
        assert np.array_equal(back.cell_data[k][ctype], mesh.cell_data[k][ctype])


def test_load_vtk_ascii(tmp_path):
    """load_vtk also reads ASCII legacy files (e.g. written by other tools): one quad and one triangle with data."""
    from femol_b200.mesh import load_vtk
    text = '''# vtk DataFile Version 3.0
hand written
ASCII
DATASET UNSTRUCTURED_GRID
POINTS 5 double
0 0 0  1 0 0  1 1 0
0 1 0  2 0.5 0
CELLS 2 9
4 0 1 2 3
3 1 4 2
CELL_TYPES 2
9
5
POINT_DATA 5
SCALARS Ux double 1
LOOKUP_TABLE default
0.0 0.1 0.2 0.3 0.4
CELL_DATA 2
SCALARS X double 1
LOOKUP_TABLE default
0.25 0.75
'''
    path = tmp_path / 'ascii.vtk'
    path.write_text(text)
    mesh = load_vtk(str(path))
    assert mesh.points.shape == (5, 3) and np.array_equal(mesh.cells['quad'], [[0, 1, 2, 3]])
    assert np.array_equal(mesh.cells['triangle'], [[1, 4, 2]])
    assert np.allclose(mesh.point_data['Ux'], [0.0, 0.1, 0.2, 0.3, 0.4])
    vals = {k: mesh.cell_data['X'][k][0] for k in mesh.contains}
    assert vals == {'quad': 0.25, 'triangle': 0.75}


def test_tracked_force_vector_flags_in_place_edits():
    """problem.F is handed back as an ndarray subclass whose writes raise a dirty flag (in-place edits of
    the load vector were silently ignored, then made to raise; now they are picked up by the next assembly)."""
    from femol_b200.FEM import _TrackedVector as T
    F = T(np.arange(6.0))
    assert isinstance(F, np.ndarray) and F.dtype == np.float64 and not F.dirty
    assert type(F + 1.0) is np.ndarray and type(np.sum(F)) is np.float64 and not F.dirty   # reads do not flag
    assert float(F @ F) == 55.0 and not F.dirty
    F[2] = 7.0
    assert F.dirty and F[2] == 7.0
    F = T(np.zeros(6)); v = F[1:4]; v[0] = 3.0                     # write through a view
    assert F.dirty and F[1] == 3.0
    F = T(np.ones(6)); F *= 2.0                                    # in-place operator keeps the tracked type
    assert isinstance(F, T) and F.dirty and F[0] == 2.0
    F = T(np.ones(6)); np.add(F, 1.0, out=F)
    assert F.dirty and F[5] == 2.0
    G = np.array(F, dtype=np.float64)                              # private copies are plain arrays
    assert type(G) is np.ndarray


def test_unique_rows_equals_numpy_unique():
    """mesh._unique_rows (T3 -> T6 mid-side nodes, mesh.py:485-520 of the reference uses np.unique(axis=0)): same unique
    rows, same inverse, on lattice mid-points and on random rows with many duplicates; empty-free edge cases."""
    from femol_b200 import mesh as fmesh
    rng = np.random.default_rng(0)
    for a in (np.round(rng.random((5000, 3)) * 12) / 4, rng.standard_normal((300, 3)), np.zeros((7, 3)), np.array([[1.0, 2.0, 0.0]])):
        u0, i0 = np.unique(a, axis=0, return_inverse=True)
        u1, i1 = fmesh._unique_rows(a)
        assert np.array_equal(u0, u1) and np.array_equal(np.asarray(i0).reshape(-1), i1)
    m = fmesh.rectangle_T6(3.0, 2.0, 6, 4)
    assert m.N_nodes == (2 * 6 + 1) * (2 * 4 + 1) and m.cells['triangle'].shape == (48, 6)
    # node and cell numbering of the T6 mesh = the oracle's restatement of Mesh.make_triangles_quad (pinned to the reference)
    import femol_oracle as o
    pts, tri = o.rectangle_points_tris(3.0, 2.0, 6, 4)
    pts6, tri6 = o.t3_to_t6(pts, tri)
    assert np.array_equal(np.asarray(m.cells['triangle']), tri6) and np.allclose(m.points[:, :2], np.asarray(pts6)[:, :2], rtol=0, atol=0)


def test_filter_neighbour_list_equals_sparse_distance_matrix():
    """TOM_SIMP.neighbour_distances (neighbour list of the mesh-independency filter on unstructured meshes,
    TOM_SIMP.py:237-267) against the direct statement `sparse_distance_matrix` minus its zeros: same pattern, same
    distances - including an empty list and two coincident centres (a zero distance off the diagonal is dropped)."""
    from scipy.spatial import cKDTree
    import femol_b200 as F
    from femol_b200 import TOM_SIMP
    rng = np.random.default_rng(1)
    for mesh in (F.mesh.rectangle_T6(7., 5., 7, 5), F.mesh.rectangle_T3(9., 4., 9, 4)):
        mesh.compute_element_centers()
        c = mesh.all_cell_centers + 0.05 * rng.standard_normal(mesh.all_cell_centers.shape)
        c[3] = c[2]
        for r in (0.05, 0.8, 1.5, 2.7):
            tree = cKDTree(c)
            D = tree.sparse_distance_matrix(tree, r, output_type='coo_matrix').tocsr()
            D.setdiag(0); D.eliminate_zeros(); D.sort_indices()
            E = TOM_SIMP.neighbour_distances(c, r)
            assert D.shape == E.shape and np.array_equal(D.indptr, E.indptr) and np.array_equal(D.indices, E.indices)
            assert D.nnz == 0 or np.abs(D.data - E.data).max() < 1e-15
