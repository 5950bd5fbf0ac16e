"""Constitutive data: isotropic / orthotropic materials and the material library.

Host-side mirror of the reference's FEMOL/materials.py (IsotropicMaterial :137-187,
OrthotropicMaterial :4-134, library functions :195-527): same class names, attributes and
argument meaning.  These are O(1) host computations that feed four small FP64 tensors to
the device kernels.
"""
import numpy as np

__all__ = ['OrthotropicMaterial', 'IsotropicMaterial']


class OrthotropicMaterial(object):
    """Orthotropic ply: on-axis Q / S matrices and their off-axis rotations (materials.py:4-134)."""

    def __init__(self, name=None, Ex=None, Ey=None, Es=None, vx=None, Xt=None, Xc=None, Yt=None, Yc=None,
                 Sc=None, ho=None, rho=None, Gyz=None, Gxz=None):
        self.name = name
        self.kind = 'orthotropic'
        self.Ex, self.Ey, self.Es = Ex, Ey, Es
        if Gyz and Gxz:  # out-of-plane shear given explicitly
            self.Gyz, self.Gxz = Gyz, Gxz
            self.out_of_plane_shear = True
        else:            # estimated from the in-plane moduli
            self.Gyz, self.Gxz = self.Ey, self.Es
        self.vx = vx
        self.vy = np.min([1, vx * (Ey / Ex)])
        self.hi = ho
        self.rho = rho
        # compliance
        self.Sxx, self.Sxy, self.Syx = 1 / Ex, -self.vy / Ey, -self.vx / Ex
        self.Syy, self.Sss = 1 / Ey, 1 / Es
        self.S = np.array([[self.Sxx, self.Sxy, 0], [self.Syx, self.Syy, 0], [0, 0, self.Sss]])
        # stiffness
        self.m = 1 / (1 - vx * self.vy)
        self.Qxx, self.Qyy = self.m * Ex, self.m * Ey
        self.Qyx, self.Qxy = self.m * vx * Ey, self.m * self.vy * Ex
        self.Qss = Es
        self.Q = np.array([[self.Qxx, self.Qxy, 0], [self.Qyx, self.Qyy, 0], [0, 0, self.Qss]])
        self.Xt, self.Xc, self.Yt, self.Yc, self.Sc = Xt, Xc, Yt, Yc, Sc

    @staticmethod
    def _rotate(U1, U2, U3, U4, U5, angle, k5, ks1, ks2):
        th = np.radians(angle)
        c2, c4, s2, s4 = np.cos(2 * th), np.cos(4 * th), np.sin(2 * th), np.sin(4 * th)
        U_mat = np.array([[U1, c2, c4], [U1, -c2, c4], [U4, 0, -c4], [U5, 0, -k5 * c4],
                          [0, ks1 * s2, ks2 * s4], [0, ks1 * s2, -ks2 * s4]])
        V = U_mat @ np.array([1, U2, U3])
        return np.array([[V[0], V[2], V[4]], [V[2], V[1], V[5]], [V[4], V[5], V[3]]])

    def Q_mat(self, angle):
        """Off-axis stiffness of a ply at `angle` degrees (materials.py:66-97)."""
        Qxx, Qyy, Qxy, Qss = self.Qxx, self.Qyy, self.Qxy, self.Qss
        U1 = (1 / 8) * (3 * Qxx + 3 * Qyy + 2 * Qxy + 4 * Qss)
        U2 = (1 / 2) * (Qxx - Qyy)
        U3 = (1 / 8) * (Qxx + Qyy - 2 * Qxy - 4 * Qss)
        U4 = (1 / 8) * (Qxx + Qyy + 6 * Qxy - 4 * Qss)
        U5 = (1 / 8) * (Qxx + Qyy - 2 * Qxy + 4 * Qss)
        return self._rotate(U1, U2, U3, U4, U5, angle, 1, 0.5, 1)

    def S_mat(self, angle):
        """Off-axis compliance (materials.py:99-124)."""
        Sxx, Syy, Sxy, Sss = self.Sxx, self.Syy, self.Sxy, self.Sss
        U1 = (1 / 8) * (3 * Sxx + 3 * Syy + 2 * Sxy + Sss)
        U2 = (1 / 2) * (Sxx - Syy)
        U3 = (1 / 8) * (Sxx + Syy - 2 * Sxy - Sss)
        U4 = (1 / 8) * (Sxx + Syy + 6 * Sxy - Sss)
        U5 = (1 / 2) * (Sxx + Syy - 2 * Sxy + Sss)
        return self._rotate(U1, U2, U3, U4, U5, angle, 4, 1, 2)

    def G_mat(self, angle):
        """Out-of-plane shear projected on the ply angle (materials.py:126-134)."""
        th = np.radians(angle)
        Gxz = self.Gxz * np.cos(th) + self.Gyz * np.sin(th)
        Gyz = self.Gyz * np.cos(th) + self.Gxz * np.sin(th)
        return np.array([[Gxz, 0], [0, Gyz]])


class IsotropicMaterial(object):
    """materials.py:137-187"""

    def __init__(self, E, mu, rho):
        self.kind = 'isotropic'
        self.E, self.mu, self.rho = E, mu, rho

    def _base(self):
        mu = self.mu
        return self.E / (1 - mu ** 2) * np.array([[1, mu, 0], [mu, 1, 0], [0, 0, (1 - mu) / 2]])

    def plane_tensor(self, t):
        return self._base() * t

    def bending_tensor(self, t):
        mu = self.mu
        return (t ** 3 / 12) * self.E / (1 - mu ** 2) * np.array([[1, mu, 0], [mu, 1, 0], [0, 0, (1 - mu) / 2]])

    @staticmethod
    def coupled_tensor():
        return np.zeros((3, 3))

    def shear_tensor(self, t):
        G = (5 / 6) * t * 0.5 * self.E / (1 + self.mu)
        return np.array([[G, 0], [0, G]])


# ---- library (materials.py:195-527): name -> constructor arguments ---------------------------------
def _ortho(name, Ex, Ey, Es, vx, ho, rho, **kw):
    return lambda: OrthotropicMaterial(name=name, Ex=Ex, Ey=Ey, Es=Es, vx=vx, ho=ho, rho=rho, **kw)


def _quasi_iso(name, E, rho, mu=0.3, ho=0.01):
    G = 0.5 * E / (1 + mu)
    return lambda: OrthotropicMaterial(name, E, E, E * (1 - mu) / 2 / (1 - mu ** 2), mu, ho=ho, rho=rho, Gxz=G, Gyz=G)


general_carbon = _ortho('general_carbon', 130e9, 3.5e9, 5e9, 0.28, 0.00015, 1550)
general_flax = _ortho('general_flax', 19e9, 2.9e9, 1.6e9, 0.28, 0.00033, 1100, Gyz=2.9e9, Gxz=1.6e9)
laminate_resin = _quasi_iso('laminate resin', 2.9e9, 1100)
T300_N5208 = _ortho('T300_N5208', 181e9, 10.3e9, 7.17e9, 0.28, 0.000125, 1600, Xt=1500e6, Xc=1500e6, Yt=40e6, Yc=246e6, Sc=68e6)
Eglass_epoxy = _ortho('Eglass_epoxy', 38.6e9, 8.27e9, 4.14e9, 0.3, 0.000125, 1800, Xt=1062e6, Xc=610e6, Yt=31e6, Yc=118e6, Sc=72e6)
Kev49_epoxy = _ortho('Kev49_epoxy', 76e9, 5.5e9, 2.3e9, 0.34, 0.000125, 1460, Xt=1400e6, Xc=235e6, Yt=12e6, Yc=53e6, Sc=34e6)
AS_H3501 = _ortho('AS_H3501', 138e9, 8.96e9, 7.10e9, 0.3, 0.000125, 1600, Xt=1447e6, Xc=1447e6, Yt=51.7e6, Yc=206e6, Sc=93e6)
flax_epoxy = _ortho('flax_epoxy', 28.2e9, 3.31e9, 5.2e9, 0.34, 0.000303, 1330, Xt=286e6, Xc=96e6, Yt=12e6, Yc=41e6, Sc=27e6)
AS4_PEEK = _ortho('AS4_PEEK', 134e9, 8.9e9, 5.10e9, 0.28, 0.000125, 1600, Xt=2130e6, Xc=1100e6, Yt=80e6, Yc=200e6, Sc=160e6)
abaqus_benchmark = _ortho('abacus_benchmark', 100e9, 5e9, 3e9, 0.4, 0.0001, 1000, Gxz=2e9, Gyz=2e9)
isotropic_laminate = _quasi_iso('isotropic laminate', 10920, 1000)
orthotropic_steel = _quasi_iso('orthotropic steel', 190e9, 7840, mu=0.28, ho=0.001)


def laminate_core(infill_density=1):
    return _quasi_iso('laminate core', infill_density * 4.8e9, 1240)()


def ALL_LAMINATE_MTR():
    return [f() for f in (general_carbon, general_flax, T300_N5208, Eglass_epoxy, Kev49_epoxy, AS_H3501,
                          flax_epoxy, AS4_PEEK, abaqus_benchmark)]


def random_laminate_material():
    mats = ALL_LAMINATE_MTR()
    return mats[np.random.randint(len(mats))]


def general_isotropic():
    return IsotropicMaterial(1, 0.3, 1)


def isotropic_bending_benchmark():
    return IsotropicMaterial(10920, 0.3, 1)


def solid_PLA():
    return IsotropicMaterial(3500, 0.36, 1.252)


def printed_PLA(infill=1):
    pla = solid_PLA()
    return IsotropicMaterial(infill * pla.E, infill * pla.mu, infill * pla.rho)
