"""Mesh container: points / cells, element centres and areas, node look-ups, result attachment.

Mirror of the reference's FEMOL/mesh.py `Mesh` for the methods on the hot path (__init__
:19-88, element_areas :90-103, clean :209-240, domain_nodes :269-282, global_matrix_indexes :284-308,
element_size :319-336, add_displacement :378-385, add_mode :427-435, compute_element_centers :437-465,
make_triangles_quad :485-520) with every per-cell / per-node Python loop replaced by array operations,
plus meshzoo-free `rectangle_Q4/T3/T6` generators that keep meshzoo 0.8.0's node and cell ordering
(mesh.py:852-882).  Plotting, VTK IO and the gmsh / guitar generators are out of scope.
"""
import numpy as np

from . import elements

__all__ = ['Mesh', 'load_vtk', 'rectangle_Q4', 'rectangle_T3', 'rectangle_T6', 'rectangle_Q8']


class Mesh(object):
    point_variables = ['Ux', 'Uy', 'Uz', 'Tx', 'Ty', 'Tz']

    def __init__(self, points, cell_dict, structured=False, quad_element=None, tri_element=None):
        self.element = None
        self.rows = None
        self.cols = None
        self.wrap_factor = None
        self.point_data = {}
        self.cell_data = {}
        self.cell_centers = {}
        self.ElementClasses = {}
        points = np.asarray(points, dtype=np.float64)
        if points.shape[1] == 2:
            points = np.hstack([points, np.zeros((points.shape[0], 1))])
        self.points = points
        self.cells = {k: np.asarray(v) for k, v in cell_dict.items()}
        self.contains = [key for key in cell_dict if key in ['triangle', 'quad']]
        self.structured = structured
        if self.structured:
            self.cell_type = self.contains[0]
            self.N_ele = self.cells[self.cell_type].shape[0]
        else:
            self.N_ele = int(np.sum([self.cells[c].shape[0] for c in self.contains]))
        self.ElementClasses['triangle'] = tri_element if tri_element else elements.T3
        self.ElementClasses['quad'] = quad_element if quad_element else elements.Q4
        if self.ElementClasses['triangle'] == elements.T6 and 'triangle' in self.contains:
            if self.cells['triangle'].shape[1] == 3:
                self.make_triangles_quad()
        self.clean()
        self.N_nodes = self.points.shape[0]
        self.compute_element_centers()
        self.areas = self.element_areas()
        if 'quad' in self.contains and (self.areas['quad'] < 0).any():
            self.fix_quadrilaterals()
            self.areas = self.element_areas()

    # ---- result IO ---------------------------------------------------------------------------------
    # VTK cell types and the permutation local node -> VTK node (6-node triangles are corner, mid, corner, mid, ...
    # here, mesh.py:105-138, and corners-then-mids in VTK; 8-node quads are already corners-then-mids)
    _VTK = {3: (5, [0, 1, 2]), 4: (9, [0, 1, 2, 3]), 6: (22, [0, 2, 4, 1, 3, 5]), 8: (23, list(range(8)))}

    def save(self, file):
        """mesh.py:242-267 without meshio: `file`.vtk, legacy VTK (binary, big-endian) UNSTRUCTURED_GRID with the
        points, the cells of every type in `contains`, `point_data` (per-node arrays) and `cell_data`
        ({name: {cell_type: array}}, concatenated in `contains` order like the reference does)."""
        cells = [np.asarray(self.cells[c], dtype=np.int64) for c in self.contains]
        n_cells = sum(c.shape[0] for c in cells)
        with open(file + '.vtk', 'wb') as f:
            f.write(b'# vtk DataFile Version 4.2\nfemol_b200 mesh\nBINARY\nDATASET UNSTRUCTURED_GRID\n')
            f.write(('POINTS %d double\n' % self.points.shape[0]).encode())
            f.write(np.ascontiguousarray(self.points, dtype='>f8').tobytes()); f.write(b'\n')
            f.write(('CELLS %d %d\n' % (n_cells, sum(c.shape[0] * (c.shape[1] + 1) for c in cells))).encode())
            for c in cells:
                perm = self._VTK[c.shape[1]][1]
                rec = np.hstack([np.full((c.shape[0], 1), c.shape[1], dtype=np.int64), c[:, perm]])
                f.write(rec.astype('>i4').tobytes())
            f.write(('\nCELL_TYPES %d\n' % n_cells).encode())
            for c in cells:
                f.write(np.full(c.shape[0], self._VTK[c.shape[1]][0], dtype='>i4').tobytes())
            f.write(b'\n')

            def scalars(name, arr):
                f.write(('SCALARS %s double 1\nLOOKUP_TABLE default\n' % str(name).replace(' ', '_')).encode())
                f.write(np.ascontiguousarray(arr, dtype='>f8').tobytes()); f.write(b'\n')
            if self.point_data:
                f.write(('POINT_DATA %d\n' % self.points.shape[0]).encode())
                for name, arr in self.point_data.items():
                    scalars(name, np.asarray(arr, dtype=np.float64).reshape(-1))
            if self.cell_data:
                f.write(('CELL_DATA %d\n' % n_cells).encode())
                for name, per_type in self.cell_data.items():
                    scalars(name, np.concatenate([np.asarray(per_type[c], dtype=np.float64).reshape(-1) for c in self.contains]))

    # ---- geometry -------------------------------------------------------------------------------
    @property
    def all_cells(self):
        return np.array([nodes for c in self.contains for nodes in self.cells[c]], dtype=object)

    def element_areas(self):
        """mesh.py:90-103, vectorised with the element classes' own area formulas"""
        areas = {}
        for ctype in self.contains:
            cls = self.ElementClasses[ctype]
            xy = self.points[self.cells[ctype]]
            x, y = xy[:, :, 0], xy[:, :, 1]
            if cls.N_nodes == 3:
                a = 0.5 * (x[:, 0] * (y[:, 1] - y[:, 2]) + x[:, 1] * (y[:, 2] - y[:, 0]) + x[:, 2] * (y[:, 0] - y[:, 1]))
            elif cls.N_nodes == 6:
                a = (x[:, 0] * (y[:, 2] - y[:, 4]) + x[:, 2] * (y[:, 4] - y[:, 0]) + x[:, 4] * (y[:, 0] - y[:, 2])) / 2
            else:
                A1 = x[:, 0] * (y[:, 1] - y[:, 2]) + x[:, 1] * (y[:, 2] - y[:, 0]) + x[:, 2] * (y[:, 0] - y[:, 1])
                A2 = x[:, 0] * (y[:, 2] - y[:, 3]) + x[:, 2] * (y[:, 3] - y[:, 0]) + x[:, 3] * (y[:, 0] - y[:, 2])
                a = A1 / 2 + A2 / 2
            areas[ctype] = a
        return areas

    def fix_quadrilaterals(self):
        """mesh.py:105-109"""
        bad = self.areas['quad'] < 0
        self.cells['quad'][bad] = np.flip(self.cells['quad'][bad], axis=1)

    def compute_element_centers(self):
        """mesh.py:437-465"""
        for ctype in self.contains:
            self.cell_centers[ctype] = self.points[self.cells[ctype]][:, :, :2].mean(axis=1)
        self.all_cell_centers = np.vstack([self.cell_centers[c] for c in self.contains])

    def dimensions(self):
        x, y = self.points.T[:2]
        return x.max() - x.min(), y.max() - y.min()

    def element_size(self):
        """mesh.py:319-336.  Structured: the x-extent of cell 0 (the reference's y_size is negative)."""
        if self.structured:
            x, y = self.points[self.cells[self.contains[0]][0]].T[:2]
            return np.max([x.max() - x.min(), y.min() - y.max()])
        dx = self.dimensions()[0] / np.sqrt(self.N_nodes)
        dy = self.dimensions()[1] / np.sqrt(self.N_nodes)
        return np.mean([dx, dy])

    # ---- node bookkeeping -----------------------------------------------------------------------
    def clean(self):
        """Drop nodes no cell uses and renumber (mesh.py:209-240)."""
        used = np.unique(np.hstack([self.cells[c].reshape(-1) for c in self.contains]))
        self.full_nodes = used
        if used.shape[0] == self.points.shape[0]:
            return
        remap = np.zeros(int(used.max()) + 1, dtype=np.int64)
        remap[used] = np.arange(used.shape[0])
        for c in self.contains:
            self.cells[c] = remap[self.cells[c]]
        self.points = self.points[used]

    def make_triangles_quad(self):
        """T3 -> T6: mid-side nodes appended in np.unique(axis=0) order of their coordinates, cell =
        [c0, m01, c1, m12, c2, m20] (mesh.py:485-520)."""
        tri = self.cells['triangle']
        pts = self.points
        a = tri.reshape(-1)
        b = np.roll(tri, -1, axis=1).reshape(-1)
        mids = np.stack([(pts[a, 0] + pts[b, 0]) / 2, (pts[a, 1] + pts[b, 1]) / 2, np.zeros(a.shape[0])], axis=1)
        uniq, inv = _unique_rows(mids)           # == np.unique(mids, axis=0, return_inverse=True), 9 x faster
        inv = inv.reshape(-1, 3) + pts.shape[0]
        self.cells['triangle'] = np.stack([tri[:, 0], inv[:, 0], tri[:, 1], inv[:, 1], tri[:, 2], inv[:, 2]], axis=1)
        self.points = np.vstack([pts, uniq])

    def domain_nodes(self, domain):
        """Nodes inside `domain(x, y)` (mesh.py:269-282); one vectorised call for array-aware
        predicates (femol_b200.domains), the reference's per-node loop otherwise."""
        x, y = self.points[:, 0], self.points[:, 1]
        if getattr(domain, 'vectorized', False):
            return np.asarray(domain(x, y)).nonzero()[0]
        return np.array([domain(*coord[:2]) for coord in self.points]).nonzero()[0]

    def global_matrix_indexes(self, N_dof):
        """COO pattern (mesh.py:284-308).  The device plan never needs these N_e*size^2 arrays; they are
        produced for API compatibility on request (`mesh.rows`, `mesh.cols`) and for small meshes."""
        rows, cols = [], []
        element = None
        for ctype in self.contains:
            cells = self.cells[ctype]
            element = self.ElementClasses[ctype](self.points[cells[0]], N_dof=N_dof)
            base = np.repeat(cells, N_dof, axis=1) * N_dof + np.tile(np.arange(N_dof), cells.shape[1])[None]
            rows.append(np.repeat(base, element.size, axis=1).reshape(-1))
            cols.append(np.tile(base, (1, element.size)).reshape(-1))
        if self.structured:
            self.element = element
        self.rows = np.hstack(rows)
        self.cols = np.hstack(cols)

    # ---- results --------------------------------------------------------------------------------
    def wrap(self, name='', factor=1):
        kx, ky = ('Ux', 'Uy') if name == '' else (name + '_Ux', name + '_Uy')
        self.wrap_data = {'x': kx, 'y': ky}
        self.wrap_factor = factor
        self.wrapped = True
        self.wrapped_points_2D = np.vstack([self.points.T[0] + factor * self.point_data[kx],
                                            self.points.T[1] + factor * self.point_data[ky]]).T

    def add_displacement(self, U, N_dof):
        """mesh.py:378-385"""
        self.U = U
        for dof in range(N_dof):
            self.point_data[self.point_variables[dof]] = U[dof::N_dof]
        self.wrap()

    def stress_from_displacement(self, *tensors, which=None, N_dof=2):
        """mesh.py:387-425: centroid stresses Sx, Sy, Sxy and Sv as cell data (plane stress, Q4 cells - the only
        element class of the reference with a `stress` method, elements.py:1144-1152).  With densities in
        `cell_data['X']` the tensor is tensors[0] + tensors[1] * X^3 (two tensors) or tensors[0] * X^3 (one); the
        exponent and the Sv formula are the reference's.  One kernel (femol_stress_q4) instead of a Python loop."""
        if N_dof != 2:
            return
        import torch
        from . import _lib
        from .FEM import _Device
        ctype = self.contains[0]
        if len(self.contains) != 1 or self.ElementClasses[ctype].N_nodes != 4:
            raise AttributeError("'%s' object has no attribute 'stress'" % self.ElementClasses[ctype].__name__)
        dev = getattr(self, '_femol_device', None)
        if dev is None:
            dev = self._femol_device = _Device(self, 2)
        ux = np.asarray(self.point_data[(which + '_Ux') if which else 'Ux'], dtype=np.float64)
        uy = np.asarray(self.point_data[(which + '_Uy') if which else 'Uy'], dtype=np.float64)
        u = np.zeros((self.N_nodes, dev.ndof))
        u[:, 0], u[:, 1] = ux, uy
        u_dev = torch.from_numpy(u).to(dev.dev)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev.dev)
        t_base = up(tensors[0])
        x = t_coat = None
        if 'X' in self.cell_data.keys():
            x = up(self.cell_data['X'][ctype])
            if len(tensors) > 1:
                t_coat = up(tensors[1])
        out = torch.empty((4, dev.n_elem), dtype=torch.float64, device=dev.dev)
        _lib.call('femol_stress_q4', dev.plan, _lib.ptr(dev.xy), _lib.ptr(u_dev), _lib.ptr(t_base), _lib.ptr(t_coat),
                  _lib.ptr(x), _lib.ptr(out), _lib.stream())
        S = out.cpu().numpy()
        for k, name in enumerate(('Sx', 'Sy', 'Sxy', 'Sv')):
            self.cell_data[name] = {ctype: S[k].copy()}

    def add_mode(self, name, V, N_dof):
        """mesh.py:427-435"""
        for dof in range(N_dof):
            self.point_data[name + '_' + self.point_variables[dof]] = V[dof::N_dof]
        self.wrap(name)


def _unique_rows(a):
    """np.unique(a, axis=0, return_inverse=True) for a float [n, 3] array.  numpy sorts the rows as structured scalars
    with a generic comparison (12 s for the 12 M edge mid-points of a 4 M-element T6 mesh); a lexicographic sort of the
    columns gives the same order, hence the same unique rows and the same inverse."""
    order = np.lexsort((a[:, 2], a[:, 1], a[:, 0]))
    s = a[order]
    first = np.empty(s.shape[0], dtype=bool)
    first[:1] = True
    np.any(s[1:] != s[:-1], axis=1, out=first[1:])
    inv = np.empty(a.shape[0], dtype=np.intp)
    inv[order] = np.cumsum(first) - 1
    return s[first], inv


def _grid(Lx, Ly, nelx, nely):
    """meshzoo 0.8.0 rectangle ordering: nodes x-fastest (n = j*(nx+1) + i); cell index e = i*ny + j."""
    x = np.linspace(0, Lx, nelx + 1)
    y = np.linspace(0, Ly, nely + 1)
    points = np.array(np.meshgrid(x, y)).reshape(2, -1).T
    a = np.add.outer(np.arange(nelx), (nelx + 1) * np.arange(nely))
    return points, a


def _quads(Lx, Ly, nelx, nely):
    points, a = _grid(Lx, Ly, nelx, nely)
    cells = np.array([a, a + 1, a + nelx + 2, a + nelx + 1]).reshape(4, -1).T
    return points, cells


def _tris(Lx, Ly, nelx, nely):
    points, q = _quads(Lx, Ly, nelx, nely)
    i, j = np.divmod(np.arange(nelx * nely), nely)
    up = ((i + j) % 2 == 0)[:, None]
    n00, n10, n11, n01 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    t0 = np.where(up, np.stack([n00, n10, n11], 1), np.stack([n00, n10, n01], 1))
    t1 = np.where(up, np.stack([n00, n11, n01], 1), np.stack([n10, n11, n01], 1))
    cells = np.empty((2 * nelx * nely, 3), dtype=np.int64)
    cells[0::2] = t0
    cells[1::2] = t1
    return points, cells


def rectangle_Q4(Lx, Ly, nelx, nely):
    """Structured rectangular quadrilateral mesh (mesh.py:852-860)."""
    points, cells = _quads(Lx, Ly, nelx, nely)
    return Mesh(points, {'quad': cells}, structured=True, quad_element=elements.Q4)


def rectangle_T3(Lx, Ly, nelx, nely):
    """mesh.py:863-871"""
    points, cells = _tris(Lx, Ly, nelx, nely)
    return Mesh(points, {'triangle': cells}, tri_element=elements.T3)


def rectangle_T6(Lx, Ly, nelx, nely):
    """mesh.py:874-882"""
    points, cells = _tris(Lx, Ly, nelx, nely)
    return Mesh(points, {'triangle': cells}, tri_element=elements.T6)


def rectangle_Q8(Lx, Ly, nelx, nely):
    """Serendipity quadrilaterals on a rectangle (no reference counterpart: its Q8 is a stub).
    Corner nodes keep the Q4 numbering; mid-side nodes follow (horizontal edges, then vertical)."""
    points, q = _quads(Lx, Ly, nelx, nely)
    nc = points.shape[0]
    hx, hy = Lx / nelx, Ly / nely
    ih, jh = np.meshgrid(np.arange(nelx), np.arange(nely + 1))          # horizontal edge (i, j): nodes (i,j)-(i+1,j)
    hmid = np.stack([(ih.reshape(-1) + 0.5) * hx, jh.reshape(-1) * hy], 1)
    iv, jv = np.meshgrid(np.arange(nelx + 1), np.arange(nely))          # vertical edge (i, j): nodes (i,j)-(i,j+1)
    vmid = np.stack([iv.reshape(-1) * hx, (jv.reshape(-1) + 0.5) * hy], 1)
    e = np.arange(nelx * nely)
    i, j = np.divmod(e, nely)
    h_id = lambda ii, jj: nc + jj * nelx + ii
    v_id = lambda ii, jj: nc + nelx * (nely + 1) + jj * (nelx + 1) + ii
    cells = np.stack([q[:, 0], q[:, 1], q[:, 2], q[:, 3], h_id(i, j), v_id(i + 1, j), h_id(i, j + 1), v_id(i, j)], 1)
    return Mesh(np.vstack([points, hmid, vmid]), {'quad': cells}, structured=True, quad_element=elements.Q8)


def load_vtk(file):
    """mesh.py:827-844 without meshio: read a legacy-VTK UNSTRUCTURED_GRID (binary as written by `Mesh.save`, or
    ASCII) with triangle / quad / 6-node triangle / 8-node quad cells back into a Mesh with its point and cell data."""
    raw = open(file, 'rb').read()
    pos = 0

    def line():
        nonlocal pos
        while True:
            end = raw.index(b'\n', pos)
            out = raw[pos:end].decode().strip()
            pos = end + 1
            if out:
                return out
    line(); line()
    binary = line().upper() == 'BINARY'
    assert line().upper().endswith('UNSTRUCTURED_GRID'), 'only UNSTRUCTURED_GRID files are supported'

    def numbers(count, dtype):
        nonlocal pos
        if binary:
            out = np.frombuffer(raw, dtype=dtype, count=count, offset=pos)
            pos += count * np.dtype(dtype).itemsize
            return out.astype(np.float64 if 'f' in dtype else np.int64)
        vals = []
        while len(vals) < count:
            vals += line().split()
        return np.array(vals[:count], dtype=np.float64 if 'f' in dtype else np.int64)
    points = cells_flat = types = None
    point_data, cell_data_flat = {}, {}
    target = None
    n_points = n_cells = 0
    while pos < len(raw):
        try:
            head = line().split()
        except ValueError:
            break
        key = head[0].upper()
        if key == 'POINTS':
            n_points = int(head[1])
            points = numbers(3 * n_points, '>f8' if head[2] == 'double' else '>f4').reshape(-1, 3)
        elif key == 'CELLS':
            n_cells = int(head[1])
            cells_flat = numbers(int(head[2]), '>i4')
        elif key == 'CELL_TYPES':
            types = numbers(int(head[1]), '>i4')
        elif key == 'POINT_DATA':
            target = (point_data, n_points)
        elif key == 'CELL_DATA':
            target = (cell_data_flat, n_cells)
        elif key == 'SCALARS':
            line()                                   # LOOKUP_TABLE default
            target[0][head[1]] = numbers(target[1], '>f8' if head[2] == 'double' else '>f4')
    inverse = {code: (n, np.argsort(perm)) for n, (code, perm) in Mesh._VTK.items()}
    per_type, order = {}, []
    o = 0
    for t in types:
        n, inv = inverse[int(t)]
        nodes = cells_flat[o + 1:o + 1 + n][inv]
        o += n + 1
        key = 'quad' if n in (4, 8) else 'triangle'
        if key not in per_type:
            per_type[key] = []
            order.append(key)
        per_type[key].append(nodes)
    cell_dict = {k: np.array(per_type[k], dtype=np.int64) for k in order}
    kw = {}
    if 'triangle' in cell_dict and cell_dict['triangle'].shape[1] == 6:
        kw['tri_element'] = elements.T6
    if 'quad' in cell_dict and cell_dict['quad'].shape[1] == 8:
        kw['quad_element'] = elements.Q8
    mesh = Mesh(points, cell_dict, **kw)
    mesh.point_data = point_data
    bounds = np.cumsum([0] + [cell_dict[k].shape[0] for k in mesh.contains])
    mesh.cell_data = {name: {k: arr[bounds[i]:bounds[i + 1]] for i, k in enumerate(mesh.contains)}
                      for name, arr in cell_data_flat.items()}
    return mesh
