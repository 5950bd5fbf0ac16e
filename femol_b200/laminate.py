"""Classical-laminate-theory tensors of a ply stack (host side).

Mirror of the reference's FEMOL/laminate.py Layup (:8-165): same constructor arguments
(`material, plies, symetric, z_core, h_core`) and attributes (`A_mat, B_mat, D_mat, G_mat, hA,
h, z_values, angles, a_mat, d_mat`).  O(plies) host arithmetic.
"""
import numpy as np

__all__ = ['Layup']


class Layup(object):

    def __init__(self, material, plies=None, symetric=False, z_core=0, h_core=0):
        self.curvature = 0
        self.eps_off = 0
        self.mtr = material
        self.plies = plies
        self.sym = symetric
        self.hc = h_core
        self.zc = z_core
        if self.sym:
            self.angles = np.hstack([np.array(plies), np.flip(np.array(plies))])
            self.angles_s = np.flip(np.array(plies))
        else:
            self.angles = np.array(plies)
        if (not self.sym) & (len(self.angles) % 2 != 0) & (self.hc > 0):
            raise ValueError('unable to build layup: a core needs an even ply count')
        self.Nb_plies = self.angles.shape[0]
        self.hA = self.Nb_plies * self.mtr.hi
        self.h = self.hA + self.hc
        self.z_values = self.get_z_values()
        self.Q_off = [self.mtr.Q_mat(a) for a in self.angles]
        self.S_off = [self.mtr.S_mat(a) for a in self.angles]
        self.A_mat, self.a_mat = self.get_A(return_a=True)
        self.B_mat = self.get_B()
        self.D_mat, self.d_mat = self.get_D(return_d=True)
        self.get_G()

    def _ply_integral(self, order):
        """sum_plies Q(theta) (z2^order - z1^order)/order ; plies above the mid-plane skip the core
        interval (laminate.py:67-134)."""
        n = len(self.angles)
        out = 0
        for idx, theta in enumerate(self.angles):
            i = idx if idx < n // 2 else idx + 1 * (self.hc > 0)
            z1, z2 = self.z_values[i], self.z_values[i + 1]
            if order == 1:
                out = out + self.mtr.Q_mat(theta) * (z2 - z1)
            else:
                out = out + self.mtr.Q_mat(theta) * (z2 ** order - z1 ** order) * (1 / order)
        return out

    def get_A(self, return_a=False):
        A = self._ply_integral(1)
        return (A, np.linalg.inv(A)) if return_a else A

    def get_B(self, return_b=False):
        B = self._ply_integral(2)
        return (B, np.linalg.inv(B)) if return_b else B

    def get_D(self, return_d=False):
        D = self._ply_integral(3)
        return (D, np.linalg.inv(D)) if return_d else D

    def get_G(self):
        """laminate.py:136-142"""
        self.G_mat = (5 / 6) * self.h * np.array([[self.mtr.Gxz, 0], [0, self.mtr.Gyz]])

    def get_z_values(self):
        """laminate.py:144-155"""
        hi = self.mtr.hi
        if self.hc == 0:
            return np.arange(-self.h / 2, self.h / 2 + hi / 2, hi) + self.zc
        zb = np.arange(-self.h / 2, -self.hc / 2 + hi / 2, hi)
        zt = np.arange(self.hc / 2, self.h / 2 + hi / 2, hi)
        return np.hstack([zb, zt]) + self.zc

    def get_A_FEM(self):
        """laminate.py:157-165 (multiplies by a matrix of ones: a no-op kept for API parity)"""
        return self.A_mat

    def apply_load(self, load):
        self.eps_off += self.a_mat @ load

    def apply_moment(self, moment):
        self.curvature += self.d_mat @ moment
