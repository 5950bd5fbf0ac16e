"""Domain predicates `domain(x, y) -> bool` for boundary conditions, forces and solid/void regions.

Mirror of the reference's FEMOL/domains.py (:12-86).  Every predicate accepts scalars (exactly as
the reference) and also numpy arrays, so Mesh.domain_nodes evaluates a whole mesh in one vectorised
call instead of the reference's Python loop over nodes (mesh.py:281).
"""
import numpy as np

__all__ = ['inside_box', 'outside_box', 'inside_circle', 'outside_circle']


def inside_box(xs, ys):
    """xs / ys: lists whose items are either [lo, hi] intervals or scalars matched with
    np.isclose(value, rtol=1e-5) (domains.py:12-33)."""

    def _match(coord, values):
        coord = np.asarray(coord)
        hit = np.zeros(coord.shape, dtype=bool)
        for value in values:
            if type(value) == list:
                hit = hit | ((value[0] <= coord) & (coord <= value[1]))
            else:
                hit = hit | np.isclose(coord, value, 1e-5)
        return hit

    def domain(x, y):
        out = _match(x, xs) & _match(y, ys)
        return bool(out) if out.ndim == 0 else out

    domain.vectorized = True
    return domain


def outside_box(x1, x2, y1, y2):
    """domains.py:36-45"""

    def box(x, y):
        return ~((np.asarray(x) > x1) & (np.asarray(x) < x2) & (np.asarray(y) > y1) & (np.asarray(y) < y2))

    box.vectorized = True
    return box


def inside_circle(x_pos, y_pos, R):
    """domains.py:48-65"""

    def circle(x, y):
        out = (np.asarray(x) - x_pos) ** 2 + (np.asarray(y) - y_pos) ** 2 < R ** 2
        return bool(out) if out.ndim == 0 else out

    circle.vectorized = True
    return circle


def outside_circle(x_pos, y_pos, R):
    """domains.py:68-86"""

    def circle(x, y):
        out = (np.asarray(x) - x_pos) ** 2 + (np.asarray(y) - y_pos) ** 2 >= R ** 2
        return bool(out) if out.ndim == 0 else out

    circle.vectorized = True
    return circle
