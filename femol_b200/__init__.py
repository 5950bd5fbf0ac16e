"""femol_b200 — B200-native (sm_100a) implementation of FEMOL's per-iteration finite-element hot path.

Drop-in surface of the reference package (FEMOL/__init__.py:1-10): `Mesh`, `FEM_Problem`, `Layup`,
`SIMP_COMP`, `SIMP_VIBE` and the submodules `materials, utils, laminate, elements, mesh, domains`.
All heavy arithmetic runs in hand-written CUDA kernels behind the C-ABI of include/femol_b200.h
(libfemol_b200.so, loaded with ctypes); torch only owns device memory, streams and torch.distributed.
"""
from . import _lib  # noqa: F401
from . import materials, utils, laminate, elements, mesh, domains  # noqa: F401
from .mesh import Mesh
from .laminate import Layup
from .FEM import FEM_Problem
from .TOM_SIMP import SIMP_COMP
from .TOM import SIMP_VIBE

__all__ = ['Mesh', 'Layup', 'FEM_Problem', 'SIMP_COMP', 'SIMP_VIBE',
           'materials', 'utils', 'laminate', 'elements', 'mesh', 'domains']
__version__ = '0.1.0'
