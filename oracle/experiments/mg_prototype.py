"""CPU multigrid prototype behind two round-2 design decisions: TEST INFRASTRUCTURE / analysis only.

  (A) lattice multigrid for T3 / T6 meshes: the fine level is the assembled operator of the triangle mesh, whose nodes
      form a complete lattice; coarse levels are bilinear Q4 grids (Galerkin).  PCG iteration counts.
  (B) smoothers on the Q4 SIMP cantilever at high density contrast: damped block-Jacobi against a 4-colour block
      Gauss-Seidel (forward before, backward after the coarse correction).

    python oracle/experiments/mg_prototype.py [A|B]
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), 'tests'))
import femol_oracle as o


def prolong_1d(n):
    nc = (n + 1) // 2
    P = np.zeros((n + 1, nc + 1))
    for i in range(n + 1):
        if i % 2 == 0:
            P[i, i // 2] = 1
        elif i == n:
            P[i, (i + 1) // 2] = 1
        else:
            P[i, i // 2] = .5
            P[i, i // 2 + 1] = .5
    return sp.csr_matrix(P)


class Level(object):
    pass


def block_diag_inverse(K, nd):
    n = K.shape[0] // nd
    Kb = K.tobsr(blocksize=(nd, nd))
    Kb.sort_indices()
    D = np.zeros((n, nd, nd))
    for i in range(n):
        for k in range(Kb.indptr[i], Kb.indptr[i + 1]):
            if Kb.indices[k] == i:
                D[i] = Kb.data[k]
    return np.linalg.inv(D)


def build(K, Nx, Ny, nd, fixed, lat_node=None, coarsest=100):
    """Hierarchy of Galerkin operators; level 0 nodes may be permuted (lat_node: lattice index -> node id)."""
    levels = []
    A = K.tocsr()
    fx = fixed.copy()
    perm = lat_node
    while True:
        L = Level()
        L.A, L.Nx, L.Ny, L.fixed = A, Nx, Ny, fx
        L.Dinv = block_diag_inverse(A, nd)
        L.colour = None
        levels.append(L)
        if A.shape[0] <= coarsest or (Nx == 1 and Ny == 1):
            L.Ainv = np.linalg.inv(A.toarray())
            break
        P2 = sp.kron(prolong_1d(Ny), prolong_1d(Nx)).tocsr()          # lattice index j * (Nx + 1) + i
        if perm is not None:
            inv = np.empty_like(perm)
            inv[perm] = np.arange(perm.shape[0])
            P2 = P2[inv]                                            # row of node id = row of its lattice index
        P = sp.kron(P2, sp.identity(nd)).tocsr()
        P = sp.diags((~fx).astype(float)) @ P
        nxc, nyc = (Nx + 1) // 2, (Ny + 1) // 2
        I, J = np.meshgrid(np.arange(nxc + 1), np.arange(nyc + 1))
        fi, fj = np.minimum(2 * I, Nx).reshape(-1), np.minimum(2 * J, Ny).reshape(-1)
        fine_lat = fj * (Nx + 1) + fi
        fine_node = perm[fine_lat] if perm is not None else fine_lat
        fc = fx.reshape(-1, nd)[fine_node].reshape(-1)
        Ac = (P.T @ A @ P).tocsr()
        # fixed coarse dofs: identity rows
        keep = sp.diags((~fc).astype(float))
        Ac = keep @ Ac @ keep + sp.diags(fc.astype(float))
        L.P = P @ keep
        # colours (lattice parity) for the Gauss-Seidel smoother of THIS level
        ii, jj = np.meshgrid(np.arange(Nx + 1), np.arange(Ny + 1))
        col = ((ii % 2) + 2 * (jj % 2)).reshape(-1)
        if perm is not None:
            c2 = np.empty_like(col)
            c2[perm] = col
            col = c2
        L.colour = col
        A, Nx, Ny, fx, perm = Ac, nxc, nyc, fc, None
    return levels


def jacobi(L, r, e, nd, omega):
    res = r if e is None else r - L.A @ e
    d = np.einsum('nab,nb->na', L.Dinv, res.reshape(-1, nd)).reshape(-1)
    return omega * d if e is None else e + omega * d


def gs_colour(L, r, e, nd, order):
    """Multi-colour block Gauss-Seidel: nodes of one colour are updated together (they are coupled only through
    other colours on a Q4 lattice; on other meshes it is a colour-wise block-Jacobi, still a convergent splitting
    when damped by the caller - here plain)."""
    e = np.zeros_like(r) if e is None else e.copy()
    for c in order:
        nodes = np.flatnonzero(L.colour == c)
        dofs = (nodes[:, None] * nd + np.arange(nd)[None]).reshape(-1)
        res = r[dofs] - L.A[dofs] @ e
        e[dofs] += np.einsum('nab,nb->na', L.Dinv[nodes], res.reshape(-1, nd)).reshape(-1)
    return e


def vcycle(levels, l, r, nd, smoother, omega=0.7, gs_levels=99):
    L = levels[l]
    if l == len(levels) - 1:
        return L.Ainv @ r
    use_gs = smoother == 'gs' and l < gs_levels
    e = gs_colour(L, r, None, nd, (0, 1, 2, 3)) if use_gs else jacobi(L, r, None, nd, omega)
    rc = L.P.T @ (r - L.A @ e)
    e = e + L.P @ vcycle(levels, l + 1, rc, nd, smoother, omega, gs_levels)
    return gs_colour(L, r, e, nd, (3, 2, 1, 0)) if use_gs else jacobi(L, r, e, nd, omega)


def pcg(A, b, M, rtol=1e-8, maxit=2000, x0=None):
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - A @ x
    z = M(r)
    p = z.copy()
    rz = r @ z
    bn = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        q = A @ p
        a = rz / (p @ q)
        x += a * p
        r -= a * q
        if np.linalg.norm(r) <= rtol * bn:
            return x, it
        z = M(r)
        rz2 = r @ z
        p = z + (rz2 / rz) * p
        rz = rz2
    return x, maxit


def simp_density(nx, ny, iters):
    P = o.CompProblem('Q4', *o.rectangle_points_quads(float(nx), float(ny), nx, ny), 2, o.iso_tensors(1, 0.3, 1),
                      [(o.inside_box([0], [[0, float(ny)]]), [0, 1])],
                      [([0, -1.0], o.inside_box([float(nx)], [[ny / 2 - 1, ny / 2 + 1]]))], structured=True,
                      volfrac=0.4, penal=3, rmin=1.5)
    for _ in range(iters):
        P.iterate()
    return P.X.copy()


def experiment_A():
    print('(A) lattice multigrid on T6 / T3 plane-stress cantilevers: PCG iterations to 1e-8')
    for et in ('T6', 'T3'):
        for (nx, ny) in ((16, 8), (32, 16), (64, 32), (96, 48)):
            case = dict(mesh=(et, float(nx), float(ny), nx, ny), physics='displacement', model='plane',
                        materials=[('iso', 1.0, 0.3, 1)], tensors=[1.0],
                        fixed=[(([0], [[0, float(ny)]]), None)], forces=[([0, -1.0], ([float(nx)], [[0, float(ny)]]))])
            arr = o.case_arrays(case)
            pts = arr['points']
            hx, hy = (0.5, 0.5) if et == 'T6' else (1.0, 1.0)
            Nx, Ny = int(round(nx / hx)), int(round(ny / hy))
            li = np.rint(pts[:, 0] / hx).astype(int)
            lj = np.rint(pts[:, 1] / hy).astype(int)
            lat_node = np.empty((Nx + 1) * (Ny + 1), dtype=int)
            lat_node[lj * (Nx + 1) + li] = np.arange(pts.shape[0])
            n_el = arr['cells'].shape[0]
            fixed = np.zeros(2 * pts.shape[0], bool)
            fixed[arr['fixed_idx']] = True
            out = []
            for label, X in (('solid', None), ('random x^3', np.random.default_rng(0).uniform(0.05, 1.0, n_el))):
                K = o.global_matrices(arr, 'K', X=X, p=3 if X is not None else None).tocsr()
                lv = build(K, Nx, Ny, 2, fixed, lat_node)
                for om in (0.5, 0.6, 0.7):
                    _, it = pcg(K, arr['F'], lambda r: vcycle(lv, 0, r, 2, 'jacobi', om))
                    out.append('%s w=%.1f: %d' % (label, om, it))
                _, it = pcg(K, arr['F'], lambda r: jacobi(lv[0], r, None, 2, 1.0), maxit=20000)
                out.append('%s block-Jacobi PCG: %d' % (label, it))
            print('  %s %3d x %3d (%6d dof, %d levels): %s' % (et, nx, ny, K.shape[0], len(lv), ' | '.join(out)))


def experiment_B():
    print('(B) Q4 SIMP cantilever, densities after k SIMP iterations: PCG iterations (Jacobi w=0.7 | GS all levels | GS level 0 only)')
    for (nx, ny) in ((96, 48), (160, 80)):
        for k in (1, 10, 25, 40):
            X = simp_density(nx, ny, k)
            case = dict(mesh=('Q4', float(nx), float(ny), nx, ny), physics='displacement', model='plane',
                        materials=[('iso', 1.0, 0.3, 1)], tensors=[1.0],
                        fixed=[(([0], [[0, float(ny)]]), [0, 1])],
                        forces=[([0, -1.0], ([float(nx)], [[ny / 2 - 1, ny / 2 + 1]]))])
            arr = o.case_arrays(case)
            K = o.global_matrices(arr, 'K', X=X, p=3).tocsr()
            fixed = np.zeros(K.shape[0], bool)
            fixed[arr['fixed_idx']] = True
            lv = build(K, nx, ny, 2, fixed)
            res = []
            for sm, gl in (('jacobi', 99), ('gs', 99), ('gs', 1), ('gs', 2)):
                _, it = pcg(K, arr['F'], lambda r: vcycle(lv, 0, r, 2, sm, 0.7, gl))
                res.append(it)
            print('  %3d x %3d after %2d SIMP iterations (min x %.3f): jacobi %d | gs %d | gs on level 0 %d | gs on levels 0-1 %d'
                  % (nx, ny, k, X.min(), *res))


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'AB'
    if 'A' in which:
        experiment_A()
    if 'B' in which:
        experiment_B()
