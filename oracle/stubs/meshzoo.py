"""Restatement of the two meshzoo==0.8.0 generators the reference calls
(reference mesh.py:857 rectangle_quad, mesh.py:868/879 rectangle_tri).

TEST INFRASTRUCTURE ONLY (oracle harness).  meshzoo is a pinned third-party
dependency (requirements.txt:5) that is absent from /root/reference and not
installable here; its published node/cell ordering is restated:

* nodes x-fastest:  n = j*(nx+1) + i,  points = meshgrid(linspace x, linspace y)
* quad cells [a, a+1, a+nx+2, a+nx+1], a = i + (nx+1)*j, enumerated x-slow /
  y-fast (e = i*ny + j).  Pinned indirectly by the reference's own tests
  (test_SIMP_COMP.py:102-104, test_plane_stress_solutions.py:151-152).
* triangles: each quad split in two, diagonal direction alternating
  ("zigzag").  The exact meshzoo triangle order is parity-unpinned, and is
  irrelevant for result parity: oracle and build consume the same
  points/cells arrays through Mesh(points, cell_dict, ...).
"""
import numpy as np


def _grid(a0, a1, n):
    nx, ny = n
    x = np.linspace(a0[0], a0[1], nx + 1)
    y = np.linspace(a1[0], a1[1], ny + 1)
    pts = np.array(np.meshgrid(x, y)).reshape(2, -1).T
    a = np.add.outer(np.arange(nx), (nx + 1) * np.arange(ny))
    return pts, a, nx, ny


def rectangle_quad(a0, a1, n, cell_type="quad4"):
    pts, a, nx, ny = _grid(a0, a1, n)
    cells = np.array([a, a + 1, a + nx + 2, a + nx + 1]).reshape(4, -1).T
    return pts, cells


def rectangle_tri(a0, a1, n, variant="zigzag"):
    pts, a, nx, ny = _grid(a0, a1, n)
    a = a.reshape(-1)
    i, j = np.divmod(np.arange(nx * ny), ny)
    up = ((i + j) % 2 == 0)
    n00, n10, n11, n01 = a, a + 1, a + nx + 2, a + nx + 1
    t0 = np.where(up[:, None], np.stack([n00, n10, n11], 1), np.stack([n00, n10, n01], 1))
    t1 = np.where(up[:, None], np.stack([n00, n11, n01], 1), np.stack([n10, n11, n01], 1))
    cells = np.empty((2 * nx * ny, 3), dtype=np.int64)
    cells[0::2] = t0
    cells[1::2] = t1
    return pts, cells


def disk_quad(*a, **k):
    raise ImportError("meshzoo.disk_quad is not restated (oracle stub)")


def disk(*a, **k):
    raise ImportError("meshzoo.disk is not restated (oracle stub)")
