"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/FEMOL).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference tree does not exist on
the GPU box):

    python oracle/make_golden.py            # all cases
    python oracle/make_golden.py comp_q4_30x15 elements

The reference imports five modules that are absent from this image; `oracle/stubs` provides
inert stand-ins (matplotlib, IPython), raising stand-ins (meshio, pygmsh, gmsh) and a
restatement of the two meshzoo==0.8.0 generators it calls (see oracle/stubs/meshzoo.py).
All finite-element arithmetic is the reference's own numpy/scipy code.
"""
import contextlib
import io
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, 'stubs'))
sys.path.insert(1, '/root/reference')
sys.path.insert(2, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402
import scipy.io  # noqa: E402
import FEMOL  # noqa: E402  (the reference)
import FEMOL.test_utils  # noqa: E402
import cases  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def csr_parts(prefix, K):
    return {prefix + '_indptr': K.indptr.astype(np.int32), prefix + '_indices': K.indices.astype(np.int32),
            prefix + '_data': K.data.copy()}


def flatX(problem, X):
    return {key: X[sl] for key, sl in zip(problem.mesh.contains, [slice(None)])}


def gen_elements():
    """Ke / Me of the reference element classes on seeded distorted elements."""
    rng = np.random.default_rng(0)
    out = {}
    iso = FEMOL.materials.IsotropicMaterial(2.5, 0.3, 1.7)
    t = 0.3
    iso_t = [iso.plane_tensor(t), iso.bending_tensor(t), iso.shear_tensor(t), iso.coupled_tensor()]
    lam = FEMOL.laminate.Layup(FEMOL.materials.general_carbon(), plies=[0, 30, -45, 90], symetric=False)
    lam_t = [lam.get_A_FEM(), lam.D_mat, lam.G_mat, lam.B_mat]
    out['iso_E_mu_rho_t'] = np.array([2.5, 0.3, 1.7, t])
    for k, T in enumerate(iso_t):
        out['iso_T%d' % k] = T
    for k, T in enumerate(lam_t):
        out['lam_T%d' % k] = T
    n = 6
    q4 = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)
    t3 = np.array([[0, 0], [1, 0], [0, 1]], float)
    xyq = q4[None] * rng.uniform(0.5, 2, (n, 1, 2)) + rng.uniform(-0.15, 0.15, (n, 4, 2)) + rng.uniform(-1, 1, (n, 1, 2))
    xyt = t3[None] * rng.uniform(0.5, 2, (n, 1, 2)) + rng.uniform(-0.15, 0.15, (n, 3, 2)) + rng.uniform(-1, 1, (n, 1, 2))
    # straight-sided T6: corner, mid, corner, mid, corner, mid
    xy6 = np.zeros((n, 6, 2))
    xy6[:, 0::2] = xyt
    xy6[:, 1] = (xyt[:, 0] + xyt[:, 1]) / 2
    xy6[:, 3] = (xyt[:, 1] + xyt[:, 2]) / 2
    xy6[:, 5] = (xyt[:, 2] + xyt[:, 0]) / 2
    out['xy_Q4'], out['xy_T3'], out['xy_T6'] = xyq, xyt, xy6
    E = FEMOL.elements
    for name, cls, xy, dofs in (('Q4', E.Q4, xyq, (2, 6)), ('T3', E.T3, xyt, (2,)), ('T6', E.T6, xy6, (2, 6))):
        for nd in dofs:
            for tn, T in (('iso', iso_t), ('lam', lam_t)):
                Ks = []
                for pts in xy:
                    el = cls(pts, nd)
                    Ks.append(el.Ke(*T) if nd == 6 else el.Ke(T[0]))
                out['Ke_%s_%d_%s' % (name, nd, tn)] = np.array(Ks)
            Ms = [cls(pts, nd).Me(iso, t) for pts in xy]
            out['Me_%s_%d' % (name, nd)] = np.array(Ms)
    # Q4 plate, three-tensor (uncoupled) branch, elements.py:1056-1069
    out['Ke_Q4_6_iso3'] = np.array([E.Q4(pts, 6).Ke(*iso_t[:3]) for pts in xyq])
    np.savez_compressed(os.path.join(GOLD, 'elements.npz'), **out)

    # the reference's own known-answer vectors
    ref = {'kattan_q4': FEMOL.test_utils.reference_2dof_stiffness_matrix(),
           'kattan_t3': FEMOL.test_utils.reference_T3_2dof_stiffnes_matrix(),
           'kb': scipy.io.loadmat('/root/reference/test/reference_matrices/element_bending.mat')['kb'],
           'ks': scipy.io.loadmat('/root/reference/test/reference_matrices/element_shear.mat')['ks']}
    np.savez_compressed(os.path.join(GOLD, 'ref_matrices.npz'), **ref)

    # laminate tensors (test_stiffness_tensors.py patterns) for several library materials / layups
    lt = {}
    specs = [('general_carbon', [0, 90, 45, -45], True, 0, 0), ('general_carbon', [45, 45], True, 0, 0),
             ('T300_N5208', [0, 0, 0, 20, -20, 90], True, 0, 0.01), ('T300_N5208', [0, 0, 0, 20, -20, 90], False, 0, 0),
             ('general_flax', [0, 90, 0, 90], True, 0, 0), ('general_carbon', [0, 90], False, 0.002, 0),
             ('abaqus_benchmark', [0, 90, 0, 90, 0, 90, 0], False, 0, 0), ('AS4_PEEK', [10, -35, 60], True, -0.001, 0)]
    for i, (mat, plies, sym, zc, hc) in enumerate(specs):
        L = FEMOL.laminate.Layup(getattr(FEMOL.materials, mat)(), plies=plies, symetric=sym, z_core=zc, h_core=hc)
        lt['A%d' % i], lt['B%d' % i], lt['D%d' % i], lt['G%d' % i] = L.A_mat, L.B_mat, L.D_mat, L.G_mat
        lt['hA%d' % i] = np.array(L.hA)
        lt['z%d' % i] = L.z_values
    np.savez_compressed(os.path.join(GOLD, 'laminate_tensors.npz'), **lt)


def gen_case(name):
    case = cases.CASES[name]
    problem = quiet(cases.build_problem, FEMOL, case)
    mesh = problem.mesh
    out = {}
    ctype = mesh.contains[0]
    out['points'] = mesh.points
    out['cells'] = np.array(mesh.cells[ctype]).astype(np.int64)
    for k, T in enumerate(problem.tensors):
        out['T%d' % k] = np.asarray(T, dtype=float)
    if problem.coating:
        for k, T in enumerate(problem.coat_tensors):
            out['Tc%d' % k] = np.asarray(T, dtype=float)
    X = cases.density_field(mesh.N_ele)
    Xd = {ctype: X}
    out['X'] = X

    if case['physics'] == 'displacement' and 'simp' not in case:
        quiet(problem.assemble, 'K')
        out.update(csr_parts('K', problem.K))
        out['F'] = problem.F.copy()
        m = quiet(problem.solve, verbose=False)
        out['U'] = m.U.copy()
        quiet(problem.assemble, 'K', X=Xd, p=3)
        out['KX_data'] = problem.K.data.copy()
        out['UX'] = quiet(problem.solve, verbose=False).U.copy()

    if case['physics'] == 'modal' and 'simp' not in case:
        quiet(problem.assemble, 'K')
        quiet(problem.assemble, 'M')
        out.update(csr_parts('K', problem.K))
        out.update(csr_parts('M', problem.M))
        w, v = quiet(problem.solve, verbose=False, filtre=0)
        out['w'] = w
        phys = np.nonzero(w > 1.0)[0]
        out['v_first'] = v[phys[0]]
        quiet(problem.assemble, 'K', X=Xd, p=3)
        quiet(problem.assemble, 'M', X=Xd, q=1)
        out['KX_data'] = problem.K.data.copy()
        out['MX_data'] = problem.M.data.copy()
        wX, _ = quiet(problem.solve, verbose=False, filtre=0)
        out['wX'] = wX

    if 'simp' in case:
        s = case['simp']
        if s['kind'] == 'comp':
            opt = FEMOL.SIMP_COMP(problem, volfrac=s['volfrac'], penal=s['penal'], rmin=s['rmin'])
            cs, Xs, chg = [], [], []
            for it in range(s['iters']):
                # one reference iteration == solve(max_iter=1) continuing from opt.X (TOM_SIMP.py:48-60)
                Xold = {k: v.copy() for k, v in opt.X.items()}
                quiet(opt.solve, converge=0.0, max_iter=1, plot=False, save=False)
                if it == 0:
                    out['U0'] = opt.U.copy()
                    out['dc0_filtered'] = opt.dc.copy()
                    out.update(csr_parts('K0', problem.K))
                    out['F'] = problem.F.copy()
                    # unfiltered sensitivities of iteration 0
                    opt.U = out['U0']
                    c0, dc0 = opt.objective_function(Xold)
                    out['dc0_raw'] = np.asarray(dc0)
                cs.append(opt.c); Xs.append(np.hstack([opt.X[k] for k in mesh.contains])); chg.append(opt.change)
            out['c_hist'] = np.array(cs); out['X_hist'] = np.array(Xs); out['change_hist'] = np.array(chg)
        else:
            # v_ref = first physical mode (f > 1 Hz) of the solid structure
            w, v = quiet(problem.solve, verbose=False, filtre=0)
            phys = np.nonzero(w > 1.0)[0]
            v_ref = v[phys[0]].copy()
            out['v_ref'] = v_ref
            out['w_solid'] = w
            opt = FEMOL.SIMP_VIBE(problem, volfrac=s['volfrac'], p=s['p'], q=s['q'], rmin=s['rmin'])
            quiet(opt.solve, v_ref, converge=0.0, min_iter=1, max_iter=s['iters'], plot=False, save=False, verbose=False)
            out['lmbd_hist'] = np.array(opt.lmbds)
            Xs = [np.hstack([mesh.cell_data['X%d' % i][k] for k in mesh.contains]) for i in range(1, s['iters'])]
            Xs.append(np.hstack([mesh.cell_data['X'][k] for k in mesh.contains]))
            out['X_hist'] = np.array(Xs)
            out['v_hist0'] = opt.eigen_vectors[0]
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **out)
    return out


def gen_stress():
    """Mesh.stress_from_displacement (mesh.py:387-425) of the reference on a jittered Q4 plane-stress mesh: no
    densities, one tensor with densities, base + coating tensors with densities."""
    case = cases.CASES['comp_q4_jitter_coat_8x4']
    problem = cases.build_problem(FEMOL, case)
    mesh = quiet(problem.solve, verbose=False)
    U = mesh.U.copy()
    Tb, Tc = problem.tensors[0], problem.coat_tensors[0]
    out = {'points': mesh.points.copy(), 'cells': np.asarray(mesh.cells['quad']), 'U': U, 'T_base': Tb, 'T_coat': Tc}
    x = cases.density_field(mesh.N_ele, seed=7)
    variants = {'plain': ((Tb,), None), 'one': ((Tb,), x), 'two': ((Tb, Tc), x)}
    for key, (tensors, dens) in variants.items():
        mesh.cell_data.pop('X', None)
        if dens is not None:
            mesh.cell_data['X'] = {'quad': dens}
        mesh.stress_from_displacement(*tensors)
        for comp in ('Sx', 'Sy', 'Sxy', 'Sv'):
            out['%s_%s' % (key, comp)] = np.asarray(mesh.cell_data[comp]['quad']).copy()
    out['x'] = x
    np.savez_compressed(os.path.join(GOLD, 'stress_q4.npz'), **out)
    return out


VARCORE_CASE = dict(cases.CASES['modal_q4_plate_coat_6x6'], mesh=('Q4J', 1., 1., 6, 6, 0.15))
VARCORE_STATIC = dict(VARCORE_CASE, physics='displacement',
                      forces=[([0, 0, -5.0, 0, 0, 0], ([1.], [[0, 1.]]))])


def gen_variable_core():
    """`solve(solve_from_zc=True, h_min)` of the reference (FEM.py:201-256, 697-731, 873-896): per-element core
    heights on a jittered two-laminate plate; K, M, the static solution and the modal spectrum."""
    out = {}
    h_min = 0.003
    problem = cases.build_problem(FEMOL, VARCORE_CASE)
    zc = np.random.default_rng(5).uniform(0.0, 0.012, problem.mesh.N_ele)
    problem.mesh.cell_data['zc'] = {'quad': zc}
    w, v = quiet(problem.solve, verbose=False, filtre=0, solve_from_zc=True, h_min=h_min, modal_solver='fast')
    out.update(csr_parts('K', problem.K)); out.update(csr_parts('M', problem.M))
    out['w'] = np.asarray(w); out['zc'] = zc; out['h_min'] = np.array(h_min)
    static = cases.build_problem(FEMOL, VARCORE_STATIC)
    static.mesh.cell_data['zc'] = {'quad': zc}
    mesh = quiet(static.solve, verbose=False, solve_from_zc=True, h_min=h_min)
    out['U'] = mesh.U.copy(); out['F'] = static.F.copy()
    # SIMP_VIBE.density_to_core_height (TOM.py:497-540) on a seeded density field
    vibe = FEMOL.SIMP_VIBE(cases.build_problem(FEMOL, VARCORE_CASE), volfrac=0.5, p=3, q=1, rmin=1.5)
    X = cases.density_field(vibe.mesh.N_ele, seed=9)
    vibe.mesh.cell_data['X'] = {'quad': X}
    m2 = vibe.density_to_core_height()
    out['X_dens'] = X; out['zc_from_density'] = np.asarray(m2.cell_data['zc']['quad'])
    np.savez_compressed(os.path.join(GOLD, 'variable_core.npz'), **out)


def gen_modal_filters():
    """The eigen-filters of the reference (FEM.py:335-418) applied to synthetic spectra."""
    rng = np.random.default_rng(11)
    mesh = FEMOL.mesh.rectangle_Q4(2., 2., 2, 2)
    prob = FEMOL.FEM_Problem('modal', 'plate', mesh)
    n = 6 * mesh.N_nodes
    out = {}
    # spectrum with low duplicates, in-plane (zero z-std) modes, closely spaced groups, a peaked vector
    f = np.array([0.5, 0.5, 0.5, 0.9, 3.0, 3.00001, 40.0, 41.0, 90.0, 200.0, 260.0, 400.0, 900.0])
    w = (2 * np.pi * f) ** 2
    v = rng.standard_normal((f.size, n))
    v[4, 2::6] = 0.0                 # a "mesh mode": no z component
    v[9] = 0.0; v[9, 7] = 1.0        # a peaked vector
    out['w'], out['v'] = w, v
    for k in (1, 2):
        wf, vf = {1: prob._filter_eigenvalues_1, 2: prob._filter_eigenvalues_2}[k](w.copy(), v.copy())
        out['w%d' % k], out['v%d' % k] = np.asarray(wf), np.asarray(vf)
    np.savez_compressed(os.path.join(GOLD, 'modal_filters.npz'), **out)


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    todo = sys.argv[1:] or (['elements', 'stress', 'modal_filters', 'variable_core'] + list(cases.CASES))
    for name in todo:
        if name == 'stress':
            gen_stress()
        elif name == 'variable_core':
            gen_variable_core()
        elif name == 'modal_filters':
            gen_modal_filters()
        elif name == 'elements':
            gen_elements()
        else:
            gen_case(name)
        print('golden:', name)
