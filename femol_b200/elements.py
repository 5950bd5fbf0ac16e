"""Element classes T3 / Q4 / T6 / Q8 — host wrappers over the batched device kernels.

Mirror of the reference's FEMOL/elements.py public interface: `Elem(points, N_dof)` with `.Ke(*tensors)`,
`.Me(material, thickness[, V])`, `.area()`, `.center()`, `.N_nodes`, `.size`.  The arithmetic runs in
femol_ke_batch / femol_me_batch (csrc/elements.cu) — one element here, the whole mesh in FEM_Problem.
Q8 is a non-functional stub in the reference (elements.py:1155); ours is the standard serendipity element
built from the reference's shape tables (parity unpinned).
"""
import numpy as np
import torch

from . import _lib

__all__ = ['T3', 'Q4', 'T6', 'Q8']


def pack_tensors(tensors):
    """[A, D, G, B] (any prefix) -> the 31-double block femol_ke_batch expects."""
    out = np.zeros(31)
    shapes = [(0, 9), (9, 18), (18, 22), (22, 31)]
    for T, (a, b) in zip(tensors, shapes):
        out[a:b] = np.asarray(T, dtype=np.float64).reshape(-1)
    return out


def mass_tensor(etype, ndof, material=None, thickness=None, V=None):
    """The ndof x ndof inertia tensor, including the T6 factors 0.45 / 0.5 (elements.py:649, 673)."""
    if V is not None:
        return np.asarray(V, dtype=np.float64)
    rho = material.rho
    if ndof == 2:
        f = 0.45 if etype == _lib.T6 else 1.0
        return f * np.identity(2) * rho * thickness
    f = 0.5 if etype == _lib.T6 else 1.0
    out = np.zeros((6, 6))
    out[:3, :3] = f * np.identity(3) * rho * thickness
    out[3:, 3:] = np.identity(3) * rho * (thickness ** 3) / 12
    return out


def ke_batch_device(etype, ndof, xy_dev, conn_dev, tensors):
    """Element stiffness matrices of all cells: torch [n_elem, size, size] on the device."""
    n_elem, nn = conn_dev.shape
    size = nn * ndof
    t = torch.from_numpy(pack_tensors(tensors)).to(xy_dev.device)
    out = torch.empty((n_elem, size, size), dtype=torch.float64, device=xy_dev.device)
    _lib.call('femol_ke_batch', etype, ndof, n_elem, _lib.ptr(xy_dev), _lib.ptr(conn_dev), _lib.ptr(t),
              _lib.ptr(out), _lib.stream())
    return out


def me_batch_device(etype, ndof, xy_dev, conn_dev, V):
    n_elem, nn = conn_dev.shape
    size = nn * ndof
    v = torch.from_numpy(np.ascontiguousarray(V, dtype=np.float64)).to(xy_dev.device)
    out = torch.empty((n_elem, size, size), dtype=torch.float64, device=xy_dev.device)
    _lib.call('femol_me_batch', etype, ndof, n_elem, _lib.ptr(xy_dev), _lib.ptr(conn_dev), _lib.ptr(v),
              _lib.ptr(out), _lib.stream())
    return out


class _Element(object):
    ETYPE = None
    N_nodes = 0

    def __init__(self, points, N_dof=2):
        points = np.asarray(points, dtype=np.float64)
        self.x, self.y = points.transpose()[:2]
        self.N_dof = N_dof
        self.size = self.N_nodes * N_dof

    def _device_inputs(self):
        _lib.require_gpu()
        xy = torch.tensor(np.stack([self.x, self.y], axis=1), dtype=torch.float64, device='cuda')
        conn = torch.arange(self.N_nodes, dtype=torch.int32, device='cuda').reshape(1, -1)
        return xy, conn

    def Ke(self, *tensors):
        if self.ETYPE == _lib.T3 and self.N_dof != 2:
            return None  # the reference's T3.Ke only handles plane stress (elements.py:231-239)
        xy, conn = self._device_inputs()
        return ke_batch_device(self.ETYPE, self.N_dof, xy, conn, tensors)[0].cpu().numpy()

    def Me(self, material=None, thickness=None, V=None):
        if self.ETYPE == _lib.T3 and self.N_dof != 2:
            return None
        xy, conn = self._device_inputs()
        Vt = mass_tensor(self.ETYPE, self.N_dof, material, thickness, V)
        return me_batch_device(self.ETYPE, self.N_dof, xy, conn, Vt)[0].cpu().numpy()

    def center(self):
        return np.mean(self.x), np.mean(self.y)


class T3(_Element):
    ETYPE, N_nodes = _lib.T3, 3

    def area(self):
        x, y = self.x, self.y
        return 0.5 * (x[0] * (y[1] - y[2]) + x[1] * (y[2] - y[0]) + x[2] * (y[0] - y[1]))


class T6(_Element):
    ETYPE, N_nodes = _lib.T6, 6

    def area(self):
        x1, x2, x3 = self.x[[0, 2, 4]]
        y1, y2, y3 = self.y[[0, 2, 4]]
        return (x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2)) / 2


class Q4(_Element):
    ETYPE, N_nodes = _lib.Q4, 4

    def area(self):
        x, y = self.x, self.y
        A1 = x[0] * (y[1] - y[2]) + x[1] * (y[2] - y[0]) + x[2] * (y[0] - y[1])
        A2 = x[0] * (y[2] - y[3]) + x[2] * (y[3] - y[0]) + x[3] * (y[0] - y[2])
        return A1 / 2 + A2 / 2


class Q8(Q4):
    ETYPE, N_nodes = _lib.Q8, 8
