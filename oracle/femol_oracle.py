"""CPU oracle: numpy/scipy restatement of FEMOL's per-iteration finite-element hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`femol_b200/`) may import
this module; only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline / reference
legs of `bench.py` use it, and only as the checker / the thing timed on host cores.

Every function restates one piece of the reference (olivecha/FEMOL, read-only tree at
/root/reference) and cites the file:line it follows.  The restatement is vectorised over
elements (the reference loops in Python) but keeps the reference's arithmetic, index
conventions and quirks so results agree to rounding.

Pinning (see tests/test_oracle_golden.py):
  * the reference's own known-answer vectors: Kattan Q4/T3 element matrices
    (FEMOL/test_utils.py:7-43), test/reference_matrices/element_{bending,shear}.mat values
    (committed as tests/golden/ref_matrices.npz), NAFEMS / analytical static and modal answers
    (test/test_bending_solutions.py, test/test_Q4_element.py:145-204);
  * outputs of the reference itself, executed in the build container through oracle/stubs
    (oracle/make_golden.py -> tests/golden/*.npz): Ke/Me for T3/T6/Q4 plane+plate, assembled
    K/M/F, displacements, eigen-frequencies, SIMP_COMP and SIMP_VIBE density histories.
Unpinned: Q8 (the reference's Q8 is a non-functional stub, elements.py:1155) -> "parity unpinned".
"""
import numpy as np
import scipy.sparse
import scipy.sparse.linalg
import scipy.linalg

# ----------------------------------------------------------------------------------------
# Gauss rules (elements.py:50-59 T3, 279-288 T6, 694-712 Q4, 1164-1180 Q8)
# ----------------------------------------------------------------------------------------
TRI_PTS_1 = [(1 / 3, 1 / 3)]
TRI_WTS_1 = [1]
TRI_PTS_2 = [(2 / 3, 1 / 6), (1 / 6, 1 / 6), (1 / 6, 2 / 3)]
TRI_WTS_2 = [1 / 3, 1 / 3, 1 / 3]
TRI_PTS_3 = [(1 / 3, 1 / 3), (1 / 5, 1 / 5), (3 / 5, 1 / 5), (1 / 5, 3 / 5)]
TRI_WTS_3 = [-27 / 96, 25 / 96, 25 / 96, 25 / 96]
_v = np.sqrt(3) / 3
QUAD_PTS_1 = [(0, 0)]
QUAD_WTS_1 = [4]
QUAD_PTS_2 = [(-_v, -_v), (_v, -_v), (_v, _v), (-_v, _v)]
QUAD_WTS_2 = [1, 1, 1, 1]
_v3 = np.sqrt(3 / 5)
_w1, _w2 = 8 / 9, 5 / 9
Q8_PTS_3 = [(-_v3, -_v3), (0, -_v3), (_v3, -_v3), (-_v3, 0), (0, 0), (_v3, 0), (-_v3, _v3), (0, _v3), (_v3, _v3)]
Q8_WTS_3 = [_w2 * _w2, _w1 * _w2, _w2 * _w2, _w1 * _w2, _w1 * _w1, _w1 * _w2, _w2 * _w2, _w1 * _w2, _w2 * _w2]

N_NODES = {'T3': 3, 'Q4': 4, 'T6': 6, 'Q8': 8}


# ----------------------------------------------------------------------------------------
# Shape functions
# ----------------------------------------------------------------------------------------
def q4_shape(xi, eta):
    """elements.py:715-725"""
    return 0.25 * np.array([(1 - xi) * (1 - eta), (1 + xi) * (1 - eta), (1 + xi) * (1 + eta), (1 - xi) * (1 + eta)])


def q4_dshape(xi, eta):
    """elements.py:728-747 -> (dN/dxi, dN/deta)"""
    dxi = 0.25 * np.array([(eta - 1), (1 - eta), (eta + 1), (-eta - 1)])
    deta = 0.25 * np.array([(xi - 1), (-xi - 1), (xi + 1), (1 - xi)])
    return dxi, deta


def t6_shape(xi, eta):
    """elements.py:291-302"""
    lm = 1 - xi - eta
    return np.array([-lm * (1 - 2 * lm), 4 * xi * lm, -xi * (1 - 2 * xi), 4 * xi * eta, -eta * (1 - 2 * eta), 4 * eta * lm])


def t6_dshape(xi, eta):
    """elements.py:305-330"""
    lm = 1 - xi - eta
    dxi = np.array([1 - 4 * lm, 4 * (lm - xi), -1 + 4 * xi, 4 * eta, 0, -4 * eta])
    deta = np.array([1 - 4 * lm, -4 * xi, 0, 4 * xi, -1 + 4 * eta, 4 * (lm - eta)])
    return dxi, deta


def q8_shape(xi, eta):
    """elements.py:1183-1197"""
    return np.array([0.25 * (1 - xi) * (1 - eta) * (-xi - eta - 1),
                     0.25 * (1 + xi) * (1 - eta) * (xi - eta - 1),
                     0.25 * (1 + xi) * (1 + eta) * (xi + eta - 1),
                     0.25 * (1 - xi) * (1 + eta) * (-xi + eta - 1),
                     0.50 * (1 - eta) * (1 + xi) * (1 - xi),
                     0.50 * (1 + xi) * (1 + eta) * (1 - eta),
                     0.50 * (1 + eta) * (1 + xi) * (1 - xi),
                     0.50 * (1 - xi) * (1 + eta) * (1 - eta)])


def q8_dshape(xi, eta):
    """elements.py:1200-1233 (the derivative tables themselves are correct serendipity forms)"""
    dxi = np.array([-(0.25 - 0.25 * xi) * (1 - eta) + (0.25 * eta - 0.25) * (-eta - xi - 1),
                    (0.25 - 0.25 * eta) * (-eta + xi - 1) + (1 - eta) * (0.25 * xi + 0.25),
                    (0.25 * eta + 0.25) * (eta + xi - 1) + (eta + 1) * (0.25 * xi + 0.25),
                    -(0.25 - 0.25 * xi) * (eta + 1) + (-0.25 * eta - 0.25) * (eta - xi - 1),
                    (0.5 - 0.5 * eta) * (1 - xi) + (0.5 - 0.5 * eta) * (-xi - 1),
                    0.5 * (1 - eta) * (eta + 1),
                    (1 - xi) * (0.5 * eta + 0.5) + (0.5 * eta + 0.5) * (-xi - 1),
                    -0.5 * (1 - eta) * (eta + 1)])
    deta = np.array([-(0.25 - 0.25 * xi) * (1 - eta) + (0.25 * xi - 0.25) * (-eta - xi - 1),
                     -(1 - eta) * (0.25 * xi + 0.25) + (-0.25 * xi - 0.25) * (-eta + xi - 1),
                     (eta + 1) * (0.25 * xi + 0.25) + (0.25 * xi + 0.25) * (eta + xi - 1),
                     (0.25 - 0.25 * xi) * (eta + 1) + (0.25 - 0.25 * xi) * (eta - xi - 1),
                     -0.5 * (1 - xi) * (xi + 1),
                     (1 - eta) * (0.5 * xi + 0.5) + (-eta - 1) * (0.5 * xi + 0.5),
                     0.5 * (1 - xi) * (xi + 1),
                     (0.5 - 0.5 * xi) * (1 - eta) + (0.5 - 0.5 * xi) * (-eta - 1)])
    return dxi, deta


# ----------------------------------------------------------------------------------------
# Geometry helpers, batched over elements.  xy: (N, nn, 2)
# ----------------------------------------------------------------------------------------
def _q4_detJ(x, y, xi, eta):
    """elements.py:872-882 closed form; x, y: (N, 4)"""
    A0 = (1 / 8) * (x[:, 0] * y[:, 1] - x[:, 0] * y[:, 3] - x[:, 1] * y[:, 0] + x[:, 1] * y[:, 2]
                    - x[:, 2] * y[:, 1] + x[:, 2] * y[:, 3] + x[:, 3] * y[:, 0] - x[:, 3] * y[:, 2])
    A1 = (1 / 8) * (-x[:, 0] * y[:, 2] + x[:, 0] * y[:, 3] + x[:, 1] * y[:, 2] - x[:, 1] * y[:, 3]
                    + x[:, 2] * y[:, 0] - x[:, 2] * y[:, 1] - x[:, 3] * y[:, 0] + x[:, 3] * y[:, 1])
    A2 = (1 / 8) * (-x[:, 0] * y[:, 1] + x[:, 0] * y[:, 2] + x[:, 1] * y[:, 0] - x[:, 1] * y[:, 3]
                    - x[:, 2] * y[:, 0] + x[:, 2] * y[:, 3] + x[:, 3] * y[:, 1] - x[:, 3] * y[:, 2])
    return A0 + A1 * xi + A2 * eta


def _q4_grads(x, y, xi, eta):
    """elements.py:884-915 -> dN/dx, dN/dy (N, 4) and detJ (N,)"""
    d = 0.25 * (y[:, 0] * (xi - 1) - y[:, 1] * (xi + 1) + y[:, 2] * (xi + 1) - y[:, 3] * (xi - 1))
    b = 0.25 * (y[:, 0] * (eta - 1) - y[:, 1] * (eta - 1) + y[:, 2] * (eta + 1) - y[:, 3] * (eta + 1))
    a = 0.25 * (x[:, 0] * (eta - 1) - x[:, 1] * (eta - 1) + x[:, 2] * (eta + 1) - x[:, 3] * (eta + 1))
    c = 0.25 * (x[:, 0] * (xi - 1) - x[:, 1] * (xi + 1) + x[:, 2] * (xi + 1) - x[:, 3] * (xi - 1))
    dxi, deta = q4_dshape(xi, eta)
    det = _q4_detJ(x, y, xi, eta)
    dNdx = (1 / det)[:, None] * (d[:, None] * dxi[None] - b[:, None] * deta[None])
    dNdy = (1 / det)[:, None] * (-c[:, None] * dxi[None] + a[:, None] * deta[None])
    return dNdx, dNdy, det


def _iso_grads(x, y, dxi, deta):
    """Isoparametric Jacobian from all nodes (T6: elements.py:446-491; Q8: standard, unpinned).
    J = [[dx/dxi, dy/dxi], [dx/deta, dy/deta]]"""
    J00 = x @ dxi
    J01 = y @ dxi
    J10 = x @ deta
    J11 = y @ deta
    det = J00 * J11 - J10 * J01
    dNdx = (1 / det)[:, None] * (J11[:, None] * dxi[None] - J01[:, None] * deta[None])
    dNdy = (1 / det)[:, None] * (-J10[:, None] * dxi[None] + J00[:, None] * deta[None])
    return dNdx, dNdy, det


def _B_plane(dNdx, dNdy, ndof, rows):
    """plane_B_matrix (elements.py:193-216, 493-518, 917-945): rows 0..2 of a (rows x nn*ndof) B"""
    N, nn = dNdx.shape
    B = np.zeros((N, rows, nn * ndof))
    B[:, 0, 0::ndof] = dNdx
    B[:, 1, 1::ndof] = dNdy
    B[:, 2, 0::ndof] = dNdy
    B[:, 2, 1::ndof] = dNdx
    return B


def _B_bending(dNdx, dNdy):
    """bending_B_matrix (elements.py:520-537, 947-961)"""
    N, nn = dNdx.shape
    B = np.zeros((N, 8, nn * 6))
    B[:, 3, 3::6] = -dNdx
    B[:, 4, 4::6] = -dNdy
    B[:, 5, 4::6] = -dNdx
    B[:, 5, 3::6] = -dNdy
    return B


def _B_shear(dNdx, dNdy, Nsh):
    """shear_B_matrix (elements.py:539-558, 963-977)"""
    N, nn = dNdx.shape
    B = np.zeros((N, 8, nn * 6))
    B[:, 6, 2::6] = dNdx
    B[:, 7, 2::6] = dNdy
    B[:, 6, 3::6] = -Nsh[None]
    B[:, 7, 4::6] = -Nsh[None]
    return B


def _BtCB(B, C):
    return np.einsum('nki,kl,nlj->nij', B, C, B)


def _tensor8(tensors):
    """8x8 block tensors (elements.py:1016-1025, 1037-1040, 588-597)"""
    A = np.zeros((8, 8)); A[:3, :3] = tensors[0]
    D = np.zeros((8, 8)); D[3:6, 3:6] = tensors[1]
    G = np.zeros((8, 8)); G[6:, 6:] = tensors[2]
    Bc = np.zeros((8, 8))
    if len(tensors) > 3:
        Bc[:3, 3:6] = tensors[3]
        Bc[3:6, :3] = tensors[3]
    return A, D, G, Bc


# ----------------------------------------------------------------------------------------
# Element stiffness matrices (a3-a7)
# ----------------------------------------------------------------------------------------
def ke_batch(etype, xy, ndof, tensors):
    """Batched Element.Ke(*tensors).  xy: (N, nn, 2); tensors: [A, D, G, B] (plate) or [C, ...] (plane).
    Returns (N, nn*ndof, nn*ndof)."""
    xy = np.asarray(xy, dtype=np.float64)
    x, y = xy[:, :, 0], xy[:, :, 1]
    N = xy.shape[0]
    if etype == 'Q4':
        if ndof == 2:  # elements.py:1002-1012
            Ke = np.zeros((N, 8, 8))
            for (xi, eta) in QUAD_PTS_2:
                dNdx, dNdy, det = _q4_grads(x, y, xi, eta)
                B = _B_plane(dNdx, dNdy, 2, 3)
                Ke += _BtCB(B, np.asarray(tensors[0])) * det[:, None, None]
            return Ke
        A, D, G, Bc = _tensor8(tensors)  # elements.py:1015-1072
        Ks = 0
        for (xi, eta), w in zip(QUAD_PTS_1, QUAD_WTS_1):
            dNdx, dNdy, det = _q4_grads(x, y, xi, eta)
            Bs = _B_shear(dNdx, dNdy, q4_shape(xi, eta))
            Ks = Ks + w * (_BtCB(Bs, G) * det[:, None, None])
        K = 0
        if len(tensors) > 3:  # coupled branch (always taken by FEM_Problem, FEM.py:148)
            C = A + Bc + D
            for (xi, eta), w in zip(QUAD_PTS_2, QUAD_WTS_2):
                dNdx, dNdy, det = _q4_grads(x, y, xi, eta)
                B = _B_plane(dNdx, dNdy, 6, 8) + _B_bending(dNdx, dNdy)
                K = K + w * (_BtCB(B, C) * det[:, None, None])
        else:
            for (xi, eta), w in zip(QUAD_PTS_2, QUAD_WTS_2):
                dNdx, dNdy, det = _q4_grads(x, y, xi, eta)
                Bp = _B_plane(dNdx, dNdy, 6, 8)
                Bb = _B_bending(dNdx, dNdy)
                K = K + w * (_BtCB(Bp, A) * det[:, None, None]) + w * (_BtCB(Bb, D) * det[:, None, None])
        K = K + Ks
        idx = np.arange(5, 24, 6)
        K[:, idx, idx] = 1  # drilling dof (elements.py:1071)
        return K
    if etype == 'T3':
        if ndof != 2:
            raise ValueError('T3 supports plane stress only (elements.py:231-239 returns None)')
        # elements.py:91-98, 129-168, 218-239
        area = 0.5 * (x[:, 0] * (y[:, 1] - y[:, 2]) + x[:, 1] * (y[:, 2] - y[:, 0]) + x[:, 2] * (y[:, 0] - y[:, 1]))
        J00 = x[:, 1] - x[:, 0]; J01 = y[:, 1] - y[:, 0]
        J10 = x[:, 2] - x[:, 0]; J11 = y[:, 2] - y[:, 0]
        idet = 1 / (2 * area)
        j00, j01, j10, j11 = idet * J11, -idet * J01, -idet * J10, idet * J00
        dxi = np.array([-1, 1, 0]); deta = np.array([-1, 0, 1])
        dNdx = j00[:, None] * dxi[None] + j01[:, None] * deta[None]
        dNdy = j10[:, None] * dxi[None] + j11[:, None] * deta[None]
        B = _B_plane(dNdx, dNdy, 2, 3)
        return area[:, None, None] * _BtCB(B, np.asarray(tensors[0]))
    if etype == 'T6':
        if ndof == 2:  # elements.py:575-584
            Ke = np.zeros((N, 12, 12))
            for (xi, eta), w in zip(TRI_PTS_3, TRI_WTS_3):
                dNdx, dNdy, det = _iso_grads(x, y, *t6_dshape(xi, eta))
                B = _B_plane(dNdx, dNdy, 2, 3)
                Ke += w * (_BtCB(B, np.asarray(tensors[0])) * det[:, None, None])
            return Ke
        A, D, G, _ = _tensor8(tensors[:3])  # elements.py:587-621 (B coupling ignored)
        Kp = Kb = Ks = 0
        for (xi, eta), w in zip(TRI_PTS_3, TRI_WTS_3):
            dNdx, dNdy, det = _iso_grads(x, y, *t6_dshape(xi, eta))
            Kp = Kp + w * (_BtCB(_B_plane(dNdx, dNdy, 6, 8), A) * det[:, None, None])
            Kb = Kb + w * (_BtCB(_B_bending(dNdx, dNdy), D) * det[:, None, None])
        for (xi, eta), w in zip(TRI_PTS_2, TRI_WTS_2):
            dNdx, dNdy, det = _iso_grads(x, y, *t6_dshape(xi, eta))
            Ks = Ks + w * (_BtCB(_B_shear(dNdx, dNdy, t6_shape(xi, eta)), G) * det[:, None, None])
        K = Kp + Kb + Ks
        idx = np.arange(5, 36, 6)
        K[:, idx, idx] = 1  # elements.py:620
        return K
    if etype == 'Q8':
        # PARITY UNPINNED: the reference Q8.Ke raises TypeError (elements.py:1269).  Standard
        # serendipity element built from the reference's shape tables (elements.py:1183-1233),
        # isoparametric Jacobian over the 8 nodes, 3x3 Gauss rule (elements.py:1174-1180).
        if ndof == 2:
            Ke = np.zeros((N, 16, 16))
            for (xi, eta), w in zip(Q8_PTS_3, Q8_WTS_3):
                dNdx, dNdy, det = _iso_grads(x, y, *q8_dshape(xi, eta))
                Ke += w * (_BtCB(_B_plane(dNdx, dNdy, 2, 3), np.asarray(tensors[0])) * det[:, None, None])
            return Ke
        A, D, G, Bc = _tensor8(tensors)
        C = A + Bc + D
        K = Ks = 0
        for (xi, eta), w in zip(Q8_PTS_3, Q8_WTS_3):
            dNdx, dNdy, det = _iso_grads(x, y, *q8_dshape(xi, eta))
            B = _B_plane(dNdx, dNdy, 6, 8) + _B_bending(dNdx, dNdy)
            K = K + w * (_BtCB(B, C) * det[:, None, None])
        for (xi, eta), w in zip(QUAD_PTS_2, QUAD_WTS_2):  # reduced 2x2 shear
            dNdx, dNdy, det = _iso_grads(x, y, *q8_dshape(xi, eta))
            Ks = Ks + w * (_BtCB(_B_shear(dNdx, dNdy, q8_shape(xi, eta)), G) * det[:, None, None])
        K = K + Ks
        idx = np.arange(5, 48, 6)
        K[:, idx, idx] = 1
        return K
    raise ValueError(etype)


# ----------------------------------------------------------------------------------------
# Element mass matrices (a8)
# ----------------------------------------------------------------------------------------
def _mass_tensor(ndof, rho, t, fudge_plane=1.0, fudge_plate=1.0):
    if ndof == 2:
        return fudge_plane * np.identity(2) * rho * t
    V = np.zeros((6, 6))
    V[:3, :3] = fudge_plate * np.identity(3) * rho * t
    V[3:, 3:] = np.identity(3) * rho * (t ** 3) / 12
    return V


def _N_matrix(Nk, ndof):
    """hstack([I*N1, I*N2, ...]) batched: Nk (N, nn) -> (N, ndof, nn*ndof)"""
    N, nn = Nk.shape
    out = np.zeros((N, ndof, nn * ndof))
    for d in range(ndof):
        out[:, d, d::ndof] = Nk
    return out


def _t6_shape_xy(x, y, px, py):
    """Physical-coordinate T6 shape functions (elements.py:342-402). x,y (N,6); px,py (N,) -> (N,6)
    in element node order [N1, N4, N2, N5, N3, N6]."""
    x1, x2, x3 = x[:, 0], x[:, 2], x[:, 4]
    x4, x5, x6 = x[:, 1], x[:, 3], x[:, 5]
    y1, y2, y3 = y[:, 0], y[:, 2], y[:, 4]
    y4, y5, y6 = y[:, 1], y[:, 3], y[:, 5]

    def line(xa, ya, xb, yb, X, Y):  # (xa-xb)*(Y-yb) - (ya-yb)*(X-xb)
        return (xa - xb) * (Y - yb) - (ya - yb) * (X - xb)

    def prod(l1, l2, xn, yn):
        a = line(*l1, px, py); b = line(*l2, px, py)
        c = line(*l1, xn, yn); d = line(*l2, xn, yn)
        return (a * b) / (c * d)
    N1 = prod((x2, y2, x3, y3), (x4, y4, x6, y6), x1, y1)
    N2 = prod((x3, y3, x1, y1), (x5, y5, x4, y4), x2, y2)
    N3 = prod((x2, y2, x1, y1), (x5, y5, x6, y6), x3, y3)
    N4 = prod((x3, y3, x1, y1), (x2, y2, x3, y3), x4, y4)
    N5 = prod((x3, y3, x1, y1), (x2, y2, x1, y1), x5, y5)
    N6 = prod((x2, y2, x1, y1), (x2, y2, x3, y3), x6, y6)
    return np.stack([N1, N4, N2, N5, N3, N6], axis=1)


def me_batch(etype, xy, ndof, rho, thickness, V=None):
    """Batched Element.Me(material, thickness[, V]).  Returns (N, size, size)."""
    xy = np.asarray(xy, dtype=np.float64)
    x, y = xy[:, :, 0], xy[:, :, 1]
    N = xy.shape[0]
    if etype == 'Q4':  # elements.py:1074-1142, make_shape_xy 792-825
        P = np.stack([np.ones_like(x), x, y, x * y], axis=2)  # (N, 4, 4) rows [1, xi, yi, xi*yi]
        Pinv = np.linalg.inv(P)
        Vt = _mass_tensor(ndof, rho, thickness) if V is None else np.asarray(V)
        Me = np.zeros((N, 4 * ndof, 4 * ndof))
        for (xi, eta) in QUAD_PTS_2:
            sh = q4_shape(xi, eta)
            px = x @ sh; py = y @ sh
            mono = np.stack([np.ones_like(px), px, py, px * py], axis=1)  # (N,4)
            Nk = np.einsum('nm,nmk->nk', mono, Pinv)  # coefs = Pinv[:, k]
            Nm = _N_matrix(Nk, ndof)
            Me += _BtCB(Nm, Vt) * _q4_detJ(x, y, xi, eta)[:, None, None]
        return Me
    if etype == 'T3':  # elements.py:241-274, 170-191
        if ndof != 2:
            raise ValueError('T3 Me: plane only')
        area = 0.5 * (x[:, 0] * (y[:, 1] - y[:, 2]) + x[:, 1] * (y[:, 2] - y[:, 0]) + x[:, 2] * (y[:, 0] - y[:, 1]))
        Vt = _mass_tensor(2, rho, thickness)
        Me = np.zeros((N, 6, 6))
        xi_, yi_ = x[:, 0], y[:, 0]; xj, yj = x[:, 1], y[:, 1]; xk, yk = x[:, 2], y[:, 2]
        for (xi, eta), w in zip(TRI_PTS_3, TRI_WTS_3):
            sh = np.array([1 - xi - eta, xi, eta])
            px = x @ sh; py = y @ sh
            N1 = 1 / (2 * area) * ((yk - yj) * (xj - px) - (xk - xj) * (yj - py))
            N2 = 1 / (2 * area) * ((yi_ - yk) * (xk - px) - (xi_ - xk) * (yk - py))
            N3 = 1 / (2 * area) * ((yj - yi_) * (xi_ - px) - (xj - xi_) * (yi_ - py))
            Nm = _N_matrix(np.stack([N1, N2, N3], 1), 2)
            Me += w * _BtCB(Nm, Vt) * (2 * area)[:, None, None]
        return Me
    if etype == 'T6':  # elements.py:623-683 (0.45 / 0.5 fudge factors, sum-1 weights)
        Vt = _mass_tensor(ndof, rho, thickness, fudge_plane=0.45, fudge_plate=0.5)
        Me = np.zeros((N, 6 * ndof, 6 * ndof))
        for (xi, eta), w in zip(TRI_PTS_2, TRI_WTS_2):
            sh = t6_shape(xi, eta)
            px = x @ sh; py = y @ sh
            Nk = _t6_shape_xy(x, y, px, py)
            _, _, det = _iso_grads(x, y, *t6_dshape(xi, eta))
            Me += w * (_BtCB(_N_matrix(Nk, ndof), Vt) * det[:, None, None])
        return Me
    if etype == 'Q8':  # PARITY UNPINNED (no Q8.Me in the reference): consistent isoparametric mass, 3x3 Gauss
        Vt = _mass_tensor(ndof, rho, thickness) if V is None else np.asarray(V)
        Me = np.zeros((N, 8 * ndof, 8 * ndof))
        for (xi, eta), w in zip(Q8_PTS_3, Q8_WTS_3):
            _, _, det = _iso_grads(x, y, *q8_dshape(xi, eta))
            Nm = _N_matrix(np.broadcast_to(q8_shape(xi, eta), (N, 8)).copy(), ndof)
            Me += w * (_BtCB(Nm, Vt) * det[:, None, None])
        return Me
    raise ValueError(etype)


def element_area(etype, xy):
    """T3.area elements.py:91-98; T6.area :404-412; Q4.area :781-790"""
    x, y = xy[:, :, 0], xy[:, :, 1]
    if etype == 'T3':
        return 0.5 * (x[:, 0] * (y[:, 1] - y[:, 2]) + x[:, 1] * (y[:, 2] - y[:, 0]) + x[:, 2] * (y[:, 0] - y[:, 1]))
    if etype == 'T6':
        x1, x2, x3 = x[:, 0], x[:, 2], x[:, 4]
        y1, y2, y3 = y[:, 0], y[:, 2], y[:, 4]
        return (x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2)) / 2
    if etype in ('Q4', 'Q8'):
        A1 = x[:, 0] * (y[:, 1] - y[:, 2]) + x[:, 1] * (y[:, 2] - y[:, 0]) + x[:, 2] * (y[:, 0] - y[:, 1])
        A2 = x[:, 0] * (y[:, 2] - y[:, 3]) + x[:, 2] * (y[:, 3] - y[:, 0]) + x[:, 3] * (y[:, 0] - y[:, 2])
        return A1 / 2 + A2 / 2
    raise ValueError(etype)


# ----------------------------------------------------------------------------------------
# Constitutive tensors (a2): materials.py:137-187, 66-97; laminate.py:8-165
# ----------------------------------------------------------------------------------------
def iso_tensors(E, mu, t):
    """[A, D, G, B] for an isotropic material of thickness t (materials.py:145-187, FEM.py:139-148)"""
    base = np.array([[1, mu, 0], [mu, 1, 0], [0, 0, (1 - mu) / 2]])
    A = E / (1 - mu ** 2) * base * t
    D = (t ** 3 / 12) * E / (1 - mu ** 2) * base
    Gs = (5 / 6) * t * 0.5 * E / (1 + mu)
    return [A, D, np.array([[Gs, 0], [0, Gs]]), np.zeros((3, 3))]


def ortho_Q(Ex, Ey, Es, vx):
    """materials.py:31-56"""
    vy = min(1, vx * (Ey / Ex))
    m = 1 / (1 - vx * vy)
    return dict(Qxx=m * Ex, Qyy=m * Ey, Qyx=m * vx * Ey, Qxy=m * vy * Ex, Qss=Es)


def ortho_Q_mat(q, angle):
    """materials.py:66-97"""
    U_1 = (1 / 8) * (3 * q['Qxx'] + 3 * q['Qyy'] + 2 * q['Qxy'] + 4 * q['Qss'])
    U_2 = (1 / 2) * (q['Qxx'] - q['Qyy'])
    U_3 = (1 / 8) * (q['Qxx'] + q['Qyy'] - 2 * q['Qxy'] - 4 * q['Qss'])
    U_4 = (1 / 8) * (q['Qxx'] + q['Qyy'] + 6 * q['Qxy'] - 4 * q['Qss'])
    U_5 = (1 / 8) * (q['Qxx'] + q['Qyy'] - 2 * q['Qxy'] + 4 * q['Qss'])
    U_vec = np.array([1, U_2, U_3])
    th = np.radians(angle)
    U_mat = np.array([[U_1, np.cos(2 * th), np.cos(4 * th)],
                      [U_1, -np.cos(2 * th), np.cos(4 * th)],
                      [U_4, 0, -np.cos(4 * th)],
                      [U_5, 0, -np.cos(4 * th)],
                      [0, 0.5 * np.sin(2 * th), np.sin(4 * th)],
                      [0, 0.5 * np.sin(2 * th), -np.sin(4 * th)]])
    Q = U_mat @ U_vec
    return np.array([[Q[0], Q[2], Q[4]], [Q[2], Q[1], Q[5]], [Q[4], Q[5], Q[3]]])


def layup_tensors(Ex, Ey, Es, vx, hi, plies, symetric=False, z_core=0, h_core=0, Gyz=None, Gxz=None):
    """[A, D, G, B], hA for a laminate (laminate.py:8-165, FEM.py:130-137).  Returns (tensors, hA)."""
    if Gyz and Gxz:
        pass
    else:
        Gyz, Gxz = Ey, Es  # materials.py:23-30
    q = ortho_Q(Ex, Ey, Es, vx)
    angles = np.hstack([np.array(plies), np.flip(np.array(plies))]) if symetric else np.array(plies)
    n = angles.shape[0]
    hA = n * hi
    h = hA + h_core
    if h_core == 0:  # laminate.py:144-155
        z = np.arange(-h / 2, h / 2 + hi / 2, hi) + z_core
    else:
        zb = np.arange(-h / 2, -h_core / 2 + hi / 2, hi)
        zt = np.arange(h_core / 2, h / 2 + hi / 2, hi)
        z = np.hstack([zb, zt]) + z_core
    A = B = D = 0
    for j, th in enumerate(angles[:n // 2]):
        Q = ortho_Q_mat(q, th)
        z1, z2 = z[j], z[j + 1]
        A = A + Q * (z2 - z1)
        B = B + Q * (z2 ** 2 - z1 ** 2) * (1 / 2)
        D = D + Q * (z2 ** 3 - z1 ** 3) * (1 / 3)
    for i, th in enumerate(angles[n // 2:]):
        i += n // 2 + 1 * (h_core > 0)
        Q = ortho_Q_mat(q, th)
        z1, z2 = z[i], z[i + 1]
        A = A + Q * (z2 - z1)
        B = B + Q * (z2 ** 2 - z1 ** 2) * (1 / 2)
        D = D + Q * (z2 ** 3 - z1 ** 3) * (1 / 3)
    G = (5 / 6) * h * np.array([[Gxz, 0], [0, Gyz]])
    return [A, D, G, B], hA


# ----------------------------------------------------------------------------------------
# Meshes (meshzoo==0.8.0 ordering restated, see oracle/stubs/meshzoo.py) and mesh helpers
# ----------------------------------------------------------------------------------------
def rectangle_points_quads(Lx, Ly, nx, ny):
    """mesh.py:852-860 via meshzoo.rectangle_quad: nodes x-fastest, cells e = i*ny + j"""
    xs = np.linspace(0, Lx, nx + 1); ys = np.linspace(0, Ly, ny + 1)
    pts = np.array(np.meshgrid(xs, ys)).reshape(2, -1).T
    a = np.add.outer(np.arange(nx), (nx + 1) * np.arange(ny))
    cells = np.array([a, a + 1, a + nx + 2, a + nx + 1]).reshape(4, -1).T
    return pts, cells


def rectangle_points_tris(Lx, Ly, nx, ny):
    """mesh.py:863-882 via meshzoo.rectangle_tri (zigzag split; triangle order parity-unpinned)"""
    pts, q = rectangle_points_quads(Lx, Ly, nx, ny)
    i, j = np.divmod(np.arange(nx * ny), ny)
    up = ((i + j) % 2 == 0)
    n00, n10, n11, n01 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    t0 = np.where(up[:, None], np.stack([n00, n10, n11], 1), np.stack([n00, n10, n01], 1))
    t1 = np.where(up[:, None], np.stack([n00, n11, n01], 1), np.stack([n10, n11, n01], 1))
    cells = np.empty((2 * nx * ny, 3), dtype=np.int64)
    cells[0::2] = t0; cells[1::2] = t1
    return pts, cells


def t3_to_t6(points, tri):
    """Mesh.make_triangles_quad (mesh.py:485-520): mid-side nodes appended in np.unique(axis=0) order of
    their coordinates; cell = [c0, m01, c1, m12, c2, m20]."""
    pts3 = points if points.shape[1] == 3 else np.hstack([points, np.zeros((points.shape[0], 1))])
    c = np.hstack([tri, tri[:, :1]])
    a = c[:, :-1].reshape(-1); b = c[:, 1:].reshape(-1)
    mids = np.stack([(pts3[a, 0] + pts3[b, 0]) / 2, (pts3[a, 1] + pts3[b, 1]) / 2, np.zeros(a.shape[0])], 1)
    uniq, inv = np.unique(mids, axis=0, return_inverse=True)
    inv = inv.reshape(-1) + pts3.shape[0]
    inv = inv.reshape(-1, 3)
    cells = np.stack([tri[:, 0], inv[:, 0], tri[:, 1], inv[:, 1], tri[:, 2], inv[:, 2]], 1)
    return np.vstack([pts3, uniq]), cells


def inside_box(xs, ys):
    """domains.py:12-33 vectorised: returns f(x_array, y_array) -> bool array"""
    def dom(x, y):
        cx = np.zeros(x.shape, bool); cy = np.zeros(y.shape, bool)
        for v in xs:
            cx |= ((v[0] <= x) & (x <= v[1])) if type(v) == list else np.isclose(x, v, 1e-5)
        for v in ys:
            cy |= ((v[0] <= y) & (y <= v[1])) if type(v) == list else np.isclose(y, v, 1e-5)
        return cx & cy
    return dom


def element_size_structured(points, cells):
    """Mesh.element_size structured branch (mesh.py:324-328): x-extent of cell 0 (y_size is negative)"""
    x = points[cells[0], 0]; y = points[cells[0], 1]
    return max(x.max() - x.min(), y.min() - y.max())


def element_size_unstructured(points):
    """mesh.py:331-334"""
    n = points.shape[0]
    Lx = points[:, 0].max() - points[:, 0].min(); Ly = points[:, 1].max() - points[:, 1].min()
    return np.mean([Lx / np.sqrt(n), Ly / np.sqrt(n)])


def cell_centers(points, cells):
    """mesh.py:437-465"""
    return points[cells][:, :, :2].mean(axis=1)


# ----------------------------------------------------------------------------------------
# Global pattern, assembly, boundary conditions (a1, a9-a14)
# ----------------------------------------------------------------------------------------
def coo_pattern(cells, ndof):
    """Mesh.global_matrix_indexes (mesh.py:284-308): element-major, row-major within Ke."""
    cells = np.asarray(cells)
    nn = cells.shape[1]
    size = nn * ndof
    base = (np.repeat(cells, ndof, axis=1) * ndof + np.tile(np.arange(ndof), nn)[None]).astype(np.int64)
    rows = np.repeat(base, size, axis=1).reshape(-1)
    cols = np.tile(base, (1, size)).reshape(-1)
    return rows, cols


def fixed_dofs(points, ndof, fixed):
    """fixed = list of (domain_fn(x_arr, y_arr)->bool, ddls).  FEM.py:424-438, mesh.py:269-282"""
    out = []
    for dom, ddls in fixed:
        nodes = np.nonzero(dom(points[:, 0], points[:, 1]))[0]
        ddls = np.arange(ndof) if ddls is None else np.asarray(ddls)
        out.append((nodes[:, None] * ndof + ddls[None]).reshape(-1))
    return np.unique(np.concatenate(out)) if out else np.zeros(0, np.int64)


def force_vector(points, ndof, forces, fixed_dof_idx):
    """add_forces (FEM.py:440-464): force split equally over the domain nodes, then zeroed at fixed dofs
    (_assemble_F, FEM.py:536-544)."""
    n = points.shape[0]
    F = np.zeros(ndof * n)
    for force, dom in forces:
        nodes = np.nonzero(dom(points[:, 0], points[:, 1]))[0]
        if len(nodes) > 0:
            f = np.array(force, dtype=float) / len(nodes)
            Fi = np.zeros(ndof * n)
            for d in range(ndof):
                Fi[ndof * nodes + d] = f[d]
            F = F + Fi
    F[fixed_dof_idx] = 0
    return F


def apply_bc_data(data, rows, cols, fixed_dof_idx, n_dof_total):
    """_apply_boundary_conditions_to_matrix_data (FEM.py:902-930): zero every COO entry in a fixed row or
    column; the first COO diagonal entry of each fixed dof is set to 1."""
    if len(fixed_dof_idx) == 0:
        return data
    mask = np.zeros(n_dof_total, bool); mask[fixed_dof_idx] = True
    fr = mask[rows]; fc = mask[cols]
    data = data.copy()
    data[fr | fc] = 0
    diag = np.nonzero(fr & (rows == cols))[0]
    _, first = np.unique(rows[diag], return_index=True)
    data[diag[first]] = 1
    return data


def assemble_csr(rows, cols, data, n_dof_total):
    """FEM.py:498-501: csr_matrix((data,(rows,cols))) + sum_duplicates -> sorted int32 indices, explicit
    zeros kept."""
    K = scipy.sparse.csr_matrix((data, (rows, cols)), shape=(n_dof_total, n_dof_total))
    K.sum_duplicates()
    return K


def matrix_data(Ke, N_e, X=None, p=None, Ke_coat=None):
    """a9-a11 (FEM.py:550-606, 608-695, 737-871).  Ke: (size,size) [structured: one matrix tiled] or
    (N_e,size,size) [per element].  Solid: data = Ke * x^p.  Coating: data = Ke + Ke_coat * x^p.
    X=None means no density scaling (FEM.py:563-564)."""
    def expand(M):
        M = np.asarray(M)
        return np.broadcast_to(M, (N_e,) + M.shape[-2:])
    s = 1 if X is None else (np.asarray(X) ** p)[:, None, None]
    if Ke_coat is None:
        return (expand(Ke) * s).reshape(-1)
    return (expand(Ke) + expand(Ke_coat) * s).reshape(-1)


# ----------------------------------------------------------------------------------------
# Solvers (a15, a16)
# ----------------------------------------------------------------------------------------
def solve_static(K, F):
    """FEM.py:258-265"""
    return scipy.sparse.linalg.spsolve(K, F)


def modal_solve_dense(K, M):
    """_scipy_modal_solve + _modal_solve clean-up, filtre=0 (FEM.py:216-256, 267-279):
    returns f[Hz] ascending and v (rows = modes, M-orthonormal)."""
    w, v = scipy.linalg.eigh(K.toarray(), M.toarray())
    v = v.T
    keep = ~np.isnan(w); v = v[keep]; w = w[keep]
    keep = ~np.isclose(w, 0); v = v[keep]; w = w[keep]
    keep = w > 0; v = v[keep]; w = w[keep]
    return np.sqrt(w) / (2 * np.pi), v


def modal_solve_sparse(K, M, k=10, sigma=0.0):
    """modal_solver='sparse' path (FEM.py:234): ARPACK shift-invert"""
    w, v = scipy.sparse.linalg.eigsh(K.tocsc(), M=M.tocsc(), k=k, sigma=sigma)
    order = np.argsort(w)
    w = w[order]; v = v[:, order].T
    keep = (w > 0) & ~np.isclose(w, 0)
    return np.sqrt(w[keep]) / (2 * np.pi), v[keep]


def q4_stress(points, cells, U, T_base, x=None, T_coat=None):
    """Centroid stresses of Q4 plane-stress elements (mesh.py:387-425 + elements.py:1144-1152).

    S = T @ B(0, 0) @ u_e with T = T_base (no densities), T_base * x^3 (one tensor) or T_base + T_coat * x^3
    (two tensors): the exponent 3 is hard-coded in the reference (mesh.py:412-414).  Returns Sx, Sy, Sxy and the
    reference's "von Mises" value sqrt(((Sx-Sy)^2 + (Sy-Sxy)^2 + (Sxy-Sx)^2) / 2) (mesh.py:420), which treats
    Sxy like a normal stress - reproduced as is."""
    P = np.asarray(points)[:, :2]
    S = np.zeros((len(cells), 3))
    dxi = 0.25 * np.array([-1.0, 1.0, 1.0, -1.0])      # dN/dxi, dN/deta at the centroid
    det = 0.25 * np.array([-1.0, -1.0, 1.0, 1.0])
    for e, cell in enumerate(cells):
        X, Y = P[cell, 0], P[cell, 1]
        J = np.array([[dxi @ X, dxi @ Y], [det @ X, det @ Y]])
        d = np.linalg.solve(J, np.stack([dxi, det]))   # rows: dN/dx, dN/dy
        B = np.zeros((3, 8))
        B[0, 0::2] = d[0]; B[1, 1::2] = d[1]; B[2, 0::2] = d[1]; B[2, 1::2] = d[0]
        u = np.empty(8); u[0::2] = U[2 * cell]; u[1::2] = U[2 * cell + 1]
        if x is None:
            T = T_base
        elif T_coat is None:
            T = T_base * x[e] ** 3
        else:
            T = T_base + T_coat * x[e] ** 3
        S[e] = T @ B @ u
    Sv = np.sqrt(((S[:, 0] - S[:, 1]) ** 2 + (S[:, 1] - S[:, 2]) ** 2 + (S[:, 2] - S[:, 0]) ** 2) / 2)
    return S[:, 0], S[:, 1], S[:, 2], Sv


def mac(v1, v2):
    """utils.py:436-447"""
    return abs(v1 @ v2) ** 2 / ((v1 @ v1) * (v2 @ v2))


def pick_mode(w, v, v_ref):
    """TOM.py:383-386"""
    m = np.array([mac(vi, v_ref) for vi in v])
    i = int(np.argmax(m))
    return w[i], v[i], m[i]


# ----------------------------------------------------------------------------------------
# Sensitivities (a17, a18)
# ----------------------------------------------------------------------------------------
def _gather(U, cells, ndof):
    base = (np.repeat(cells, ndof, axis=1) * ndof + np.tile(np.arange(ndof), cells.shape[1])[None])
    return U[base]


def compliance_sens(U, cells, ndof, X, p, Ke, Ke_coat=None, structured=True):
    """SIMP_COMP objective (TOM_SIMP.py:126-226).
    structured: Ke (size,size) unscaled.  Solid: c = sum x^p q, dc = -p x^(p-1) q with q = Ue'Ke Ue.
    Coating: c = sum Ue'(Kb + x^p Kc)Ue, dc = -p x^(p-1) Ue'Kc Ue.
    unstructured: the reference reuses FEM.element_Ke, which is ALREADY scaled by x^p (FEM.py:630-631,
    673-675) -> c = sum x^(2p) q, dc = -p x^(2p-1) q (and likewise for the coat part).  Reproduced."""
    Ue = _gather(U, cells, ndof)
    Kb = np.asarray(Ke)
    quad = (lambda Kx: np.einsum('ni,ij,nj->n', Ue, Kx, Ue)) if Kb.ndim == 2 else \
           (lambda Kx: np.einsum('ni,nij,nj->n', Ue, Kx, Ue))
    X = np.asarray(X)
    extra = 0 if structured else p
    if Ke_coat is None:
        q = quad(Kb)
        c = np.sum(X ** (p + extra) * q)
        dc = -p * X ** (p - 1 + extra) * q
    else:
        qb = quad(Kb); qc = quad(np.asarray(Ke_coat))
        c = np.sum(qb + X ** (p + extra) * qc)
        dc = -p * X ** (p - 1 + extra) * qc
    return c, dc


def eig_sens(v, lmbd, cells, ndof, X, p, q, Ke, Me, structured=True):
    """SIMP_VIBE objective (TOM.py:408-495): dl_e = Ve'(p x^(p-1) Ke - lmbd q x^(q-1) Me)Ve, lmbd in Hz.
    (coating: pass the coat matrices.)  Unstructured solid reuses x^p / x^q scaled element matrices
    (FEM.py:630, 779) -> exponents 2p-1 and 2q-1."""
    Ve = _gather(v, cells, ndof)
    Kb = np.asarray(Ke); Mb = np.asarray(Me)
    qk = np.einsum('ni,ij,nj->n', Ve, Kb, Ve) if Kb.ndim == 2 else np.einsum('ni,nij,nj->n', Ve, Kb, Ve)
    qm = np.einsum('ni,ij,nj->n', Ve, Mb, Ve) if Mb.ndim == 2 else np.einsum('ni,nij,nj->n', Ve, Mb, Ve)
    X = np.asarray(X)
    ek = 0 if structured else p
    em = 0 if structured else q
    return p * X ** (p - 1 + ek) * qk - lmbd * q * X ** (q - 1 + em) * qm


# ----------------------------------------------------------------------------------------
# Filter (a19) and OC updates (a20, a21)
# ----------------------------------------------------------------------------------------
def filter_sens_literal(X, dc, centers, rmin, h):
    """_filter_sensibility exactly as written (TOM_SIMP.py:228-277; TOM.py:305-354) - O(N^2)"""
    N = len(X)
    out = np.zeros(N)
    sd = rmin * h
    for i in range(N):
        xe, ye = centers[i]
        D = np.sqrt((centers[:, 0] - xe) ** 2 + (centers[:, 1] - ye) ** 2)
        D[i] = D.max()
        nb = np.nonzero(D <= sd)[0]
        sH = 0
        for j in nb:
            dist = ((centers[j, 0] - xe) ** 2 + (centers[j, 1] - ye) ** 2) ** 0.5
            Hf = max(0, rmin - dist / h)
            sH += Hf
            out[i] += Hf * X[j] * dc[j]
        out[i] /= (X[i] * sH)
    return out


def filter_sens(X, dc, centers, rmin, h):
    """Same filter through a KD-tree neighbour search (weights at the search radius are 0, so the
    inclusive/exclusive boundary does not matter)."""
    from scipy.spatial import cKDTree
    tree = cKDTree(centers)
    Dm = tree.sparse_distance_matrix(tree, rmin * h * (1 + 1e-9), output_type='coo_matrix')
    keep = Dm.row != Dm.col
    r, c, d = Dm.row[keep], Dm.col[keep], Dm.data[keep]
    Hw = np.maximum(0, rmin - d / h)
    H = scipy.sparse.csr_matrix((Hw, (r, c)), shape=(len(X), len(X)))
    return (H @ (X * dc)) / (X * (H @ np.ones(len(X))))


def oc_update_comp(X, dc, volfrac, solid_mask=None, void_mask=None):
    """SIMP_COMP._get_new_x (TOM_SIMP.py:279-313)"""
    l1, l2, move = 0, 100000, 0.5
    N = len(X)
    with np.errstate(invalid='ignore', divide='ignore'):
        while (l2 - l1) > 1e-4:
            lmid = 0.5 * (l1 + l2)
            X1 = X + move
            if np.any(dc > 0):
                X2 = X * (dc / lmid) ** 0.3
            else:
                X2 = X * (-dc / lmid) ** 0.3
            Xn = np.min([X1, X2], axis=0)
            Xn = np.min([np.ones(N), Xn], axis=0)
            Xn = np.max([X - move, Xn], axis=0)
            Xn = np.max([0.001 * np.ones(N), Xn], axis=0)
            if solid_mask is not None:
                Xn[solid_mask] = 1
            if void_mask is not None:
                Xn[void_mask] = 0.001
            if (np.sum(Xn) - volfrac * N) > 0:
                l1 = lmid
            else:
                l2 = lmid
    return Xn


def oc_update_vibe(X, dl, volfrac, structured=True, areas=None, solid_mask=None, void_mask=None):
    """SIMP_VIBE._get_new_x (TOM.py:251-303).  Returns (X_new, dl_after_clamp_and_rescale)."""
    l1, l2, move = 0, 100000, 0.3
    N = len(X)
    dl = dl.copy()
    dl[dl < 0] = 0
    if np.sum(dl) < 1:
        dl *= 1 / dl.max()
    while (l2 - l1) > 1e-4:
        lmid = 0.5 * (l1 + l2)
        X1 = X + move
        X2 = X * (dl / lmid) ** 0.3
        Xn = np.min([X1, X2], axis=0)
        Xn = np.min([np.ones(N), Xn], axis=0)
        Xn = np.max([X - move, Xn], axis=0)
        Xn = np.max([0.001 * np.ones(N), Xn], axis=0)
        if solid_mask is not None:
            Xn[solid_mask] = 1
        if void_mask is not None:
            Xn[void_mask] = 0.001
        if structured:
            up = (np.sum(Xn) - volfrac * N) > 0
        else:
            up = np.sum(Xn * areas) / np.sum(areas) > volfrac
        if up:
            l1 = lmid
        else:
            l2 = lmid
    return Xn, dl


# ----------------------------------------------------------------------------------------
# Whole-iteration drivers (a22): SIMP_COMP.solve / SIMP_VIBE.solve on plain arrays
# ----------------------------------------------------------------------------------------
class CompProblem(object):
    """A SIMP_COMP problem on plain arrays (single cell type).  Mirrors TOM_SIMP.py:13-81 + FEM.py."""

    def __init__(self, etype, points, cells, ndof, tensors, fixed, forces, structured=True,
                 coat_tensors=None, volfrac=0.5, penal=3, rmin=1.5):
        self.etype, self.points, self.cells, self.ndof = etype, np.asarray(points, float), np.asarray(cells), ndof
        self.tensors, self.coat_tensors = tensors, coat_tensors
        self.structured = structured
        self.f, self.p, self.rmin = volfrac, penal, rmin
        self.N_e = self.cells.shape[0]
        self.ntot = ndof * self.points.shape[0]
        self.rows, self.cols = coo_pattern(self.cells, ndof)
        self.fixed_idx = fixed_dofs(self.points, ndof, fixed)
        self.F = force_vector(self.points, ndof, forces, self.fixed_idx)
        xy = self.points[self.cells][:, :, :2]
        if structured:  # one Ke from cell 0 (mesh.py:295-305, FEM.py:571)
            self.Ke = ke_batch(etype, xy[:1], ndof, tensors)[0]
            self.Kc = ke_batch(etype, xy[:1], ndof, coat_tensors)[0] if coat_tensors is not None else None
            self.h = element_size_structured(self.points, self.cells)
        else:
            self.Ke = ke_batch(etype, xy, ndof, tensors)
            self.Kc = ke_batch(etype, xy, ndof, coat_tensors) if coat_tensors is not None else None
            self.h = element_size_unstructured(self.points)
        self.centers = cell_centers(self.points, self.cells)
        self.X = np.ones(self.N_e) * volfrac

    def assemble_K(self, X=None):
        data = matrix_data(self.Ke, self.N_e, X, self.p, self.Kc)
        data = apply_bc_data(data, self.rows, self.cols, self.fixed_idx, self.ntot)
        return assemble_csr(self.rows, self.cols, data, self.ntot)

    def iterate(self, fast_filter=True):
        X = self.X
        self.K = self.assemble_K(X)
        self.U = solve_static(self.K, self.F)
        self.c, dc = compliance_sens(self.U, self.cells, self.ndof, X, self.p, self.Ke, self.Kc, self.structured)
        self.dc_raw = dc
        flt = filter_sens if fast_filter else filter_sens_literal
        self.dc = flt(X, dc, self.centers, self.rmin, self.h)
        self.X = oc_update_comp(X, self.dc, self.f)
        self.change = np.max(np.abs(self.X - X))
        return self.X

    def run(self, iters, fast_filter=True):
        cs, Xs = [], []
        for _ in range(iters):
            self.iterate(fast_filter)
            cs.append(self.c); Xs.append(self.X.copy())
        return np.array(cs), np.array(Xs)


# ----------------------------------------------------------------------------------------
# Material library (materials.py:195-527): name -> parameters
# ----------------------------------------------------------------------------------------
def _resin_like(E, rho, mu=0.3, ho=0.01):
    G = 0.5 * E / (1 + mu)
    return dict(kind='ortho', Ex=E, Ey=E, Es=E * (1 - mu) / 2 / (1 - mu ** 2), vx=mu, ho=ho, rho=rho, Gxz=G, Gyz=G)


MATERIAL_LIB = {
    'general_carbon': dict(kind='ortho', Ex=130e9, Ey=3.5e9, Es=5e9, vx=0.28, ho=0.00015, rho=1550),
    'general_flax': dict(kind='ortho', Ex=19e9, Ey=2.9e9, Es=1.6e9, vx=0.28, ho=0.00033, rho=1100, Gyz=2.9e9, Gxz=1.6e9),
    'laminate_resin': _resin_like(2.9e9, 1100),
    'laminate_core': _resin_like(4.8e9, 1240),
    'T300_N5208': dict(kind='ortho', Ex=181e9, Ey=10.3e9, Es=7.17e9, vx=0.28, ho=0.000125, rho=1600),
    'Eglass_epoxy': dict(kind='ortho', Ex=38.6e9, Ey=8.27e9, Es=4.14e9, vx=0.3, ho=0.000125, rho=1800),
    'Kev49_epoxy': dict(kind='ortho', Ex=76e9, Ey=5.5e9, Es=2.3e9, vx=0.34, ho=0.000125, rho=1460),
    'AS_H3501': dict(kind='ortho', Ex=138e9, Ey=8.96e9, Es=7.10e9, vx=0.3, ho=0.000125, rho=1600),
    'flax_epoxy': dict(kind='ortho', Ex=28.2e9, Ey=3.31e9, Es=5.2e9, vx=0.34, ho=0.000303, rho=1330),
    'AS4_PEEK': dict(kind='ortho', Ex=134e9, Ey=8.9e9, Es=5.10e9, vx=0.28, ho=0.000125, rho=1600),
    'abaqus_benchmark': dict(kind='ortho', Ex=100e9, Ey=5e9, Es=3e9, vx=0.4, ho=0.0001, rho=1000, Gxz=2e9, Gyz=2e9),
    'isotropic_laminate': _resin_like(10920, 1000),
    'orthotropic_steel': _resin_like(190e9, 7840, mu=0.28, ho=0.001),
    'general_isotropic': dict(kind='iso', E=1, mu=0.3, rho=1),
    'isotropic_bending_benchmark': dict(kind='iso', E=10920, mu=0.3, rho=1),
}


def case_arrays(case, jitter_fn=None):
    """Plain-array version of a tests/cases.py specification (same construction order as
    cases.build_problem through the reference API)."""
    kind = case['mesh'][0]
    Lx, Ly, nx, ny = case['mesh'][1:5]
    ndof = {'plane': 2, 'plate': 6}[case['model']]
    if kind in ('Q4', 'Q4J'):
        pts, cells = rectangle_points_quads(Lx, Ly, nx, ny)
        etype = 'Q4'
        if kind == 'Q4J':
            pts = jitter_fn(np.hstack([pts, np.zeros((pts.shape[0], 1))]), nx, ny, Lx, Ly, case['mesh'][5])
    elif kind == 'T3':
        pts, cells = rectangle_points_tris(Lx, Ly, nx, ny); etype = 'T3'
    elif kind == 'T6':
        pts, cells = rectangle_points_tris(Lx, Ly, nx, ny)
        pts, cells = t3_to_t6(pts, cells); etype = 'T6'
    pts = np.asarray(pts, float)
    if pts.shape[1] == 2:
        pts = np.hstack([pts, np.zeros((pts.shape[0], 1))])
    tens, hos, rhos = [], [], []
    for m, t in zip(case['materials'], case['tensors']):
        par = dict(kind='iso', E=m[1], mu=m[2], rho=m[3]) if m[0] == 'iso' else MATERIAL_LIB[m[1]]
        if par['kind'] == 'iso':
            tens.append(iso_tensors(par['E'], par['mu'], t)); hos.append(t)
        else:
            _, plies, sym, zc, hc = t
            T, hA = layup_tensors(par['Ex'], par['Ey'], par['Es'], par['vx'], par['ho'], plies, sym, zc, hc,
                                  par.get('Gyz'), par.get('Gxz'))
            tens.append(T); hos.append(hA)
        rhos.append(par['rho'])
    fixed = [(inside_box(xs, ys), ddls) for (xs, ys), ddls in case.get('fixed', [])]
    forces = [(f, inside_box(xs, ys)) for f, (xs, ys) in case.get('forces', [])]
    fidx = fixed_dofs(pts, ndof, fixed)
    return dict(etype=etype, points=pts, cells=cells, ndof=ndof, tensors=tens[0],
                coat_tensors=tens[1] if len(tens) > 1 else None, rho=rhos, ho=hos,
                fixed=fixed, forces=forces, fixed_idx=fidx, F=force_vector(pts, ndof, forces, fidx),
                structured=(kind == 'Q4'))


def global_matrices(arr, which='K', X=None, p=None):
    """Assemble K or M for a case_arrays() dict (structured: one element matrix from cell 0)."""
    pts, cells, ndof, et = arr['points'], arr['cells'], arr['ndof'], arr['etype']
    xy = pts[cells][:, :, :2]
    sel = xy[:1] if arr['structured'] else xy
    if which == 'K':
        E0 = ke_batch(et, sel, ndof, arr['tensors'])
        E1 = ke_batch(et, sel, ndof, arr['coat_tensors']) if arr['coat_tensors'] is not None else None
    else:
        E0 = me_batch(et, sel, ndof, arr['rho'][0], arr['ho'][0])
        E1 = me_batch(et, sel, ndof, arr['rho'][1], arr['ho'][1]) if arr['coat_tensors'] is not None else None
    if arr['structured']:
        E0 = E0[0]; E1 = None if E1 is None else E1[0]
    rows, cols = coo_pattern(cells, ndof)
    ntot = ndof * pts.shape[0]
    data = matrix_data(E0, cells.shape[0], X, p, E1)
    data = apply_bc_data(data, rows, cols, arr['fixed_idx'], ntot)
    return assemble_csr(rows, cols, data, ntot)


def variable_core_matrices(case, zc, h_min=0.0, which='K', jitter_fn=None):
    """K / M of a two-laminate plate whose core height varies per element (FEM.py:697-731, 873-896, used by
    `solve(solve_from_zc=True)`): for zc > h_min the element is Ke(base laminate shifted to -h/2 + hA0/2) +
    Ke(coating laminate shifted to +h/2 - hA1/2) with h = zc + hA0 + hA1, each with the FULL ply list and
    symetric=False; otherwise the unshifted base laminate alone.  Mass: Me(base) (+ Me(coat) if zc > h_min)."""
    arr = case_arrays(case, jitter_fn)
    pts, cells, ndof, et = arr['points'], arr['cells'], arr['ndof'], arr['etype']
    xy = pts[cells][:, :, :2]
    par = [MATERIAL_LIB[m[1]] for m in case['materials']]
    angles, hA = [], []
    for t, pr in zip(case['tensors'], par):
        _, plies, sym, _, _ = t
        full = np.hstack([np.array(plies), np.flip(np.array(plies))]) if sym else np.array(plies)
        angles.append(full); hA.append(full.shape[0] * pr['ho'])

    def lam(k, z_core):
        pr = par[k]
        return layup_tensors(pr['Ex'], pr['Ey'], pr['Es'], pr['vx'], pr['ho'], angles[k], False, z_core, 0,
                             pr.get('Gyz'), pr.get('Gxz'))[0]

    size = xy.shape[1] * ndof
    E = np.zeros((cells.shape[0], size, size))
    if which == 'M':
        Mb = me_batch(et, xy, ndof, arr['rho'][0], arr['ho'][0])
        Mc = me_batch(et, xy, ndof, arr['rho'][1], arr['ho'][1])
    for e in range(cells.shape[0]):
        above = zc[e] > h_min
        if which == 'K':
            if above:
                h = zc[e] + hA[0] + hA[1]
                E[e] = ke_batch(et, xy[e:e + 1], ndof, lam(0, -h / 2 + hA[0] / 2))[0] + \
                    ke_batch(et, xy[e:e + 1], ndof, lam(1, h / 2 - hA[1] / 2))[0]
            else:
                E[e] = ke_batch(et, xy[e:e + 1], ndof, lam(0, 0))[0]
        else:
            E[e] = Mb[e] + (Mc[e] if above else 0.0)
    rows, cols = coo_pattern(cells, ndof)
    ntot = ndof * pts.shape[0]
    data = apply_bc_data(E.reshape(-1), rows, cols, arr['fixed_idx'], ntot)
    return assemble_csr(rows, cols, data, ntot), arr


class VibeProblem(object):
    """SIMP_VIBE on plain arrays (TOM.py:37-132, 374-386) with the dense modal solve (FEM.py:277)."""

    def __init__(self, arr, volfrac=0.5, p=3, q=1, rmin=1.5):
        self.arr = arr
        self.f, self.p, self.q, self.rmin = volfrac, p, q, rmin
        pts, cells, ndof, et = arr['points'], arr['cells'], arr['ndof'], arr['etype']
        self.cells, self.ndof = cells, ndof
        xy = pts[cells][:, :, :2]
        sel = xy[:1] if arr['structured'] else xy
        coat = arr['coat_tensors'] is not None
        # the objective uses the coat matrices for coating problems (TOM.py:461-462) else the base ones
        T = arr['coat_tensors'] if coat else arr['tensors']
        k = 1 if coat else 0
        self.Ke = ke_batch(et, sel, ndof, T)
        self.Me = me_batch(et, sel, ndof, arr['rho'][k], arr['ho'][k])
        if arr['structured']:
            self.Ke, self.Me = self.Ke[0], self.Me[0]
            self.h = element_size_structured(pts, cells)
        else:
            self.h = element_size_unstructured(pts)
        self.centers = cell_centers(pts, cells)
        self.areas = element_area(et, xy)
        self.X = np.ones(cells.shape[0]) * volfrac
        self.lmbds = []

    def iterate(self, v_ref):
        X = self.X
        K = global_matrices(self.arr, 'K', X, self.p)
        M = global_matrices(self.arr, 'M', X, self.q)
        w, v = modal_solve_dense(K, M)
        self.lmbd, self.v, self.mac = pick_mode(w, v, v_ref)
        self.lmbds.append(self.lmbd)
        structured = self.arr['structured'] or self.arr['coat_tensors'] is not None and False
        dl = eig_sens(self.v, self.lmbd, self.cells, self.ndof, X, self.p, self.q, self.Ke, self.Me,
                      self.arr['structured'])
        self.dl_raw = dl
        dl = filter_sens(X, dl, self.centers, self.rmin, self.h)
        self.X, self.dl = oc_update_vibe(X, dl, self.f, self.arr['structured'], self.areas)
        self.change = np.max(np.abs(self.X - X))
        return self.X
