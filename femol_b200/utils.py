"""Modal helpers on host arrays (reference FEMOL/utils.py:436-461, :272-278)."""
import time

import numpy as np

__all__ = ['MAC', 'MAC_mat', 'unique_time_string']


def MAC(v1, v2):
    """Modal assurance criterion between two vectors (utils.py:436-447)."""
    v1 = np.asarray(v1); v2 = np.asarray(v2)
    return abs(v1.T @ v2) ** 2 / ((v1.T @ v1) * (v2.T @ v2))


def MAC_mat(modes1, modes2):
    """utils.py:450-461"""
    return np.array([[MAC(v1, v2) for v2 in modes2] for v1 in modes1])


def unique_time_string():
    return time.strftime('%Y%m%d_%H%M%S')
