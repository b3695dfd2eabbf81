"""Oracle: Philox4x32-10 noise stream (ctypes over oracle/philox_ref.c + a pure-Python
restatement of the integer rounds for the Random123 known-answer vectors).

Test infrastructure -- see ``oracle/__init__.py`` and the header of ``philox_ref.c``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# column order and sigmas: generation_type1.py:25-32,295-306 / generation_type2.py:31-43,190-200
NOISE_STD = (0.05, 0.05, 0.003, 0.010, 0.003, 0.030)   # X, Y, phi, vx, vy, omega
NOISE_SEED_BASE = 12345                                 # generation_type1.py:255, generation_type2.py:191


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_philox.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.tgo_noise_block.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p]
        L.tgo_philox_stream.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
        L.tgo_box_muller.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        L.tgo_box_muller_vec.argtypes = [ctypes.c_uint32] + [ctypes.c_void_p] * 4
        _LIB = L
    return _LIB


def philox4x32_10_py(ctr, key):
    """Pure-Python rounds (Salmon et al. SC'11) -- only for the KAT cross-check."""
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return (c0, c1, c2, c3)


def philox_stream(seed, first, block, n):
    out = np.zeros((n, 4), dtype=np.uint32)
    lib().tgo_philox_stream(int(seed), int(first), int(block), int(n), out.ctypes.data)
    return out


def uniform01(r):
    """u = (r + 0.5) 2^-32 of raw 32-bit words (exact in fp64) -- the stream's uniform convention."""
    return (np.asarray(r, dtype=np.float64) + 0.5) * 2.0 ** -32


def box_muller(r0, r1):
    """Vectorised tgo_box_muller: (r0[i], r1[i]) -> (n0[i], n1[i])."""
    r0 = np.ascontiguousarray(r0, dtype=np.uint32).ravel()
    r1 = np.ascontiguousarray(r1, dtype=np.uint32).ravel()
    n0 = np.zeros(r0.size)
    n1 = np.zeros(r0.size)
    lib().tgo_box_muller_vec(r0.size, r0.ctypes.data, r1.ctypes.data, n0.ctypes.data, n1.ctypes.data)
    return n0, n1


def standard_normals(seed, n_rows):
    """[n_rows, 6] standard normals of trajectory seed ``seed`` (rows 0..n_rows-1)."""
    out = np.zeros((n_rows, 6), dtype=np.float64)
    lib().tgo_noise_block(int(seed), int(n_rows), out.ctypes.data)
    return out


def sensor_noise(traj_id, n_rows, seed_base=NOISE_SEED_BASE, std=NOISE_STD):
    """noise[n_rows, 6] = sigma_c * n(seed_base + traj_id, row, c)."""
    return standard_normals(seed_base + traj_id, n_rows) * np.asarray(std)[None, :]
