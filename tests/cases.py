"""Problem specifications shared by the golden-vector generator (oracle/make_golden.py, which
feeds them to the UNMODIFIED reference `FEMOL`) and by the parity tests (which feed the same
specs to `femol_b200`).  Because femol_b200 mirrors the reference's Python API, one builder
(`build_problem(F, case)`) serves both: `F` is either module.

Case templates follow the reference's own tests / examples:
  comp_*  : test/test_SIMP_COMP.py:7-37, 68-104, 138-153 ; Examples/Topopt/quantilever_plane.py
  plate_* : test/test_bending_solutions.py:51-78, 115-151
  modal_* : test/test_Q4_element.py:145-204 ; test/test_modal_solver.py:7-42
"""
import numpy as np


def _box(xs, ys):
    return (xs, ys)


CASES = {}


def _cantilever_comp(nx, ny, iters, **kw):
    Lx, Ly = float(nx), float(ny)
    c = dict(mesh=('Q4', Lx, Ly, nx, ny), physics='displacement', model='plane',
             materials=[('iso', 1, 0.3, 1)], tensors=[1],
             fixed=[(_box([0], [[0, Ly / 4]]), [0, 1]), (_box([0], [[3 * Ly / 4, Ly]]), [0, 1])],
             forces=[([0, -0.1], _box([Lx], [[Ly / 2 - 1, Ly / 2 + 1]]))],
             simp=dict(kind='comp', volfrac=0.4, penal=3, rmin=1.5, iters=iters))
    c.update(kw)
    return c


# cfg 1 and the reference's own 30x15 test
CASES['comp_q4_30x15'] = _cantilever_comp(30, 15, 5)
CASES['comp_q4_60x30'] = _cantilever_comp(60, 30, 20)

# coating formulation, laminate tensors (test_SIMP_COMP.py:68-104)
CASES['comp_q4_coat_30x15'] = dict(
    mesh=('Q4', 30., 15., 30, 15), physics='displacement', model='plane',
    materials=[('lib', 'general_carbon'), ('lib', 'general_carbon')],
    tensors=[('layup', [0, 90, 45, -45], True, 0, 0), ('layup', [45, 45], True, 0, 0)],
    fixed=[(_box([0], [[0, 15 / 4]]), [0, 1]), (_box([0], [[3 * 15 / 4, 15]]), [0, 1])],
    forces=[([0, -1e6], _box([30.], [[15 / 2 - 1, 15 / 2 + 1]]))],
    simp=dict(kind='comp', volfrac=0.4, penal=3, rmin=1.5, iters=3))

# 6-dof plate SIMP_COMP (test_SIMP_COMP.py:138-153 pattern), isotropic plate
CASES['comp_q4_plate_16x8'] = dict(
    mesh=('Q4', 16., 8., 16, 8), physics='displacement', model='plate',
    materials=[('iso', 10920, 0.3, 1)], tensors=[0.5],
    fixed=[(_box([0], [[0, 8.]]), None)],
    forces=[([0, 0, -1e-2, 0, 0, 0], _box([16.], [[0, 8.]]))],
    simp=dict(kind='comp', volfrac=0.5, penal=3, rmin=1.5, iters=3))

# unstructured element paths (per-element Ke; double-penalised objective quirk, FEM.py:630-631)
CASES['comp_t3_12x6'] = dict(
    mesh=('T3', 12., 6., 12, 6), physics='displacement', model='plane',
    materials=[('iso', 1, 0.3, 1)], tensors=[1],
    fixed=[(_box([0], [[0, 6.]]), [0, 1])],
    forces=[([0, -0.1], _box([12.], [[0, 6.]]))],
    simp=dict(kind='comp', volfrac=0.5, penal=3, rmin=1.5, iters=3))
CASES['comp_t6_8x4'] = dict(
    mesh=('T6', 8., 4., 8, 4), physics='displacement', model='plane',
    materials=[('iso', 1, 0.3, 1)], tensors=[1],
    fixed=[(_box([0], [[0, 4.]]), [0, 1])],
    forces=[([0, -0.1], _box([8.], [[0, 4.]]))],
    simp=dict(kind='comp', volfrac=0.5, penal=3, rmin=1.5, iters=3))
CASES['comp_q4_jitter_10x5'] = dict(
    mesh=('Q4J', 10., 5., 10, 5, 0.2), physics='displacement', model='plane',
    materials=[('iso', 1, 0.3, 1)], tensors=[1],
    fixed=[(_box([0], [[0, 5.]]), [0, 1])],
    forces=[([0, -0.1], _box([10.], [[0, 5.]]))],
    simp=dict(kind='comp', volfrac=0.5, penal=3, rmin=1.5, iters=3))
CASES['comp_q4_jitter_coat_8x4'] = dict(
    mesh=('Q4J', 8., 4., 8, 4, 0.15), physics='displacement', model='plane',
    materials=[('iso', 1, 0.3, 1), ('iso', 5, 0.25, 2)], tensors=[1, 0.5],
    fixed=[(_box([0], [[0, 4.]]), [0, 1])],
    forces=[([0, -0.1], _box([8.], [[0, 4.]]))],
    simp=dict(kind='comp', volfrac=0.5, penal=3, rmin=1.5, iters=2))

# static plates: simply supported isotropic plate under uniform pressure (test_bending_solutions.py:51-78)
CASES['plate_iso_ss_10x10'] = dict(
    mesh=('Q4', 1., 1., 10, 10), physics='displacement', model='plate',
    materials=[('lib', 'isotropic_bending_benchmark')], tensors=[0.1],
    fixed=[(_box([[0, 1.]], [[0, 1.]]), [0, 1, 5]),
           (_box([0, 1.], [[0, 1.]]), [2, 4]),
           (_box([[0, 1.]], [0, 1.]), [2, 3])],
    forces=[([0, 0, -1.0, 0, 0, 0], _box([[0.01, 0.99]], [[0.01, 0.99]]))])
# laminate with core (test_bending_solutions.py:115-151 layup) -> nonzero B coupling when unsymmetric
CASES['plate_lam_8x6'] = dict(
    mesh=('Q4', 0.8, 0.6, 8, 6), physics='displacement', model='plate',
    materials=[('lib', 'T300_N5208')], tensors=[('layup', [0, 0, 0, 20, -20, 90], False, 0, 0)],
    fixed=[(_box([0], [[0, 0.6]]), None)],
    forces=[([0, 0, -10.0, 0, 0, 0], _box([0.8], [[0, 0.6]]))])
CASES['plate_lam_core_sym_6x6'] = dict(
    mesh=('Q4', 0.6, 0.6, 6, 6), physics='displacement', model='plate',
    materials=[('lib', 'T300_N5208')], tensors=[('layup', [0, 0, 0, 20, -20, 90], True, 0, 0.01)],
    fixed=[(_box([0], [[0, 0.6]]), None)],
    forces=[([0, 0, -10.0, 0, 0, 0], _box([0.6], [[0, 0.6]]))])
CASES['plate_t6_6x4'] = dict(
    mesh=('T6', 0.6, 0.4, 6, 4), physics='displacement', model='plate',
    materials=[('iso', 10920, 0.3, 1)], tensors=[0.05],
    fixed=[(_box([0], [[0, 0.4]]), None)],
    forces=[([0, 0, -1.0, 0, 0, 0], _box([0.6], [[0, 0.4]]))])
CASES['plate_q4_jitter_6x5'] = dict(
    mesh=('Q4J', 0.6, 0.5, 6, 5, 0.2), physics='displacement', model='plate',
    materials=[('lib', 'T300_N5208')], tensors=[('layup', [0, 45, -45, 90], False, 0, 0)],
    fixed=[(_box([0], [[0, 0.5]]), None)],
    forces=[([0, 0, -1.0, 0, 0, 0], _box([0.6], [[0, 0.5]]))])
CASES['plane_t3_9x7'] = dict(
    mesh=('T3', 9., 7., 9, 7), physics='displacement', model='plane',
    materials=[('iso', 3.0, 0.25, 1)], tensors=[0.7],
    fixed=[(_box([0], [[0, 7.]]), [0, 1])],
    forces=[([1.0, -0.5], _box([9.], [[0, 7.]]))])

# modal
CASES['modal_q4_plate_ss_8x8'] = dict(  # NAFEMS FV13 set-up (test_Q4_element.py:145-173)
    mesh=('Q4', 10., 10., 8, 8), physics='modal', model='plate',
    materials=[('iso', 200e9, 0.3, 8000)], tensors=[0.05],
    fixed=[(_box([[0, 10.]], [[0, 10.]]), [0, 1, 5]),
           (_box([0, 10.], [[0, 10.]]), [2, 4]),
           (_box([[0, 10.]], [0, 10.]), [2, 3])])
CASES['modal_q4_plate_cant_8x8'] = dict(  # NAFEMS 11a set-up (test_Q4_element.py:175-204)
    mesh=('Q4', 10., 10., 8, 8), physics='modal', model='plate',
    materials=[('iso', 200e9, 0.3, 8000)], tensors=[0.05],
    fixed=[(_box([[0, 10.]], [[0, 10.]]), [0, 1]), (_box([0], [[0, 10.]]), None)])
CASES['modal_q4_plane_10x5'] = dict(
    mesh=('Q4', 1.0, 0.5, 10, 5), physics='modal', model='plane',
    materials=[('iso', 71e9, 0.33, 2700)], tensors=[0.005],
    fixed=[(_box([0], [[0, 0.5]]), None)])
CASES['modal_t6_plane_6x3'] = dict(
    mesh=('T6', 1.0, 0.5, 6, 3), physics='modal', model='plane',
    materials=[('iso', 71e9, 0.33, 2700)], tensors=[0.005],
    fixed=[(_box([0], [[0, 0.5]]), None)])
CASES['modal_t3_plane_8x4'] = dict(
    mesh=('T3', 1.0, 0.5, 8, 4), physics='modal', model='plane',
    materials=[('iso', 71e9, 0.33, 2700)], tensors=[0.005],
    fixed=[(_box([0], [[0, 0.5]]), None)])
CASES['modal_t6_plate_4x4'] = dict(
    mesh=('T6', 1.0, 1.0, 4, 4), physics='modal', model='plate',
    materials=[('iso', 200e9, 0.3, 8000)], tensors=[0.05],
    fixed=[(_box([[0, 1.]], [[0, 1.]]), [0, 1]), (_box([0], [[0, 1.]]), None)])
CASES['modal_q4_plate_coat_6x6'] = dict(  # coating split (test_modal_solver.py:24-42 pattern)
    mesh=('Q4', 1., 1., 6, 6), physics='modal', model='plate',
    materials=[('lib', 'general_flax'), ('lib', 'general_carbon')],
    tensors=[('layup', [0, 90, 0, 90], True, 0, 0), ('layup', [0, 90], False, 0.002, 0)],
    fixed=[(_box([[0, 1.]], [[0, 1.]]), [0, 1]), (_box([0], [[0, 1.]]), None)])

# SIMP_VIBE (unpinned by the reference's tests; golden = the reference run itself)
CASES['vibe_q4_plate_10x10'] = dict(
    mesh=('Q4', 10., 10., 10, 10), physics='modal', model='plate',
    materials=[('iso', 200e9, 0.3, 8000)], tensors=[0.05],
    fixed=[(_box([[0, 10.]], [[0, 10.]]), [0, 1, 5]),
           (_box([0, 10.], [[0, 10.]]), [2, 4]),
           (_box([[0, 10.]], [0, 10.]), [2, 3])],
    simp=dict(kind='vibe', volfrac=0.5, p=3, q=1, rmin=1.5, iters=4))
CASES['vibe_q4_plate_coat_8x8'] = dict(
    mesh=('Q4', 1., 1., 8, 8), physics='modal', model='plate',
    materials=[('lib', 'general_flax'), ('lib', 'general_carbon')],
    tensors=[('layup', [0, 90, 0, 90], True, 0, 0), ('layup', [0, 90], False, 0.002, 0)],
    fixed=[(_box([[0, 1.]], [[0, 1.]]), [0, 1]), (_box([0], [[0, 1.]]), None)],
    simp=dict(kind='vibe', volfrac=0.5, p=3, q=1, rmin=1.5, iters=3))


def jitter_points(points, nx, ny, Lx, Ly, amp, seed=0):
    """Deterministic interior-node jitter (<= amp * h) for distorted-mesh parity."""
    rng = np.random.default_rng(seed)
    pts = np.array(points, dtype=float, copy=True)
    hx, hy = Lx / nx, Ly / ny
    x, y = pts[:, 0], pts[:, 1]
    interior = (x > 1e-9) & (x < Lx - 1e-9) & (y > 1e-9) & (y < Ly - 1e-9)
    d = rng.uniform(-1, 1, size=(pts.shape[0], 2)) * amp
    pts[interior, 0] += d[interior, 0] * hx
    pts[interior, 1] += d[interior, 1] * hy
    return pts


def make_mesh(F, spec):
    kind = spec[0]
    if kind in ('Q4', 'T3', 'T6'):
        return getattr(F.mesh, 'rectangle_' + kind)(*spec[1:5])
    if kind == 'Q4J':  # jittered quads through the public Mesh ctor, unstructured path
        Lx, Ly, nx, ny, amp = spec[1:6]
        base = F.mesh.rectangle_Q4(Lx, Ly, nx, ny)
        pts = jitter_points(base.points, nx, ny, Lx, Ly, amp)
        return F.Mesh(pts, {'quad': np.array(base.cells['quad'])}, structured=False, quad_element=F.elements.Q4)
    raise ValueError(kind)


def make_material(F, m):
    if m[0] == 'iso':
        return F.materials.IsotropicMaterial(*m[1:])
    return getattr(F.materials, m[1])()


def make_tensor(F, t, material):
    if isinstance(t, tuple) and t[0] == 'layup':
        _, plies, sym, zc, hc = t
        return F.laminate.Layup(material, plies=plies, symetric=sym, z_core=zc, h_core=hc)
    return t


def build_problem(F, case):
    """Build an FEM_Problem through the (reference-compatible) public API of module F."""
    mesh = make_mesh(F, case['mesh'])
    problem = F.FEM_Problem(case['physics'], case['model'], mesh)
    mats = [make_material(F, m) for m in case['materials']]
    problem.define_materials(*mats)
    problem.define_tensors(*[make_tensor(F, t, m) for t, m in zip(case['tensors'], mats)])
    for (xs, ys), ddls in case.get('fixed', []):
        problem.add_fixed_domain(F.domains.inside_box(xs, ys), ddls=ddls)
    for force, (xs, ys) in case.get('forces', []):
        problem.add_forces(force, F.domains.inside_box(xs, ys))
    return problem


def density_field(n, seed=1):
    """Deterministic non-uniform density x ~ U[0.001, 1]"""
    return np.random.default_rng(seed).uniform(0.001, 1.0, size=n)
