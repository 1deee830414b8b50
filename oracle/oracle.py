"""ctypes wrapper over oracle/liboracle.so (CPU restatement) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


class OGrid(C.Structure):
    _fields_ = [("cs", C.c_int * 3), ("ncells", C.c_int), ("cell_length", C.c_float), ("hash_rcp", C.c_float)]


class OScene(C.Structure):
    _fields_ = [("nF", C.c_int), ("nB", C.c_int),
                ("posF", C.c_void_p), ("massF", C.c_void_p), ("csF", C.c_void_p),
                ("posB", C.c_void_p), ("massB", C.c_void_p), ("csB", C.c_void_p),
                ("g", OGrid), ("R", C.c_float)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sph_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, _LIB_PATH], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_system_create.restype = C.c_void_p
        _lib.oracle_system_field.restype = C.c_void_p
        _lib.oracle_abs_sum.restype = C.c_float
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def grid(cell_size, cell_length, hash_rcp=0.0) -> OGrid:
    g = OGrid()
    g.cs[:] = [int(c) for c in cell_size]
    g.ncells = int(cell_size[0]) * int(cell_size[1]) * int(cell_size[2])
    g.cell_length = float(cell_length)
    g.hash_rcp = float(hash_rcp)
    return g


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class SceneArrays:
    """Sorted fluid + boundary arrays with their cell ranges: the inputs every sweep kernel takes."""

    def __init__(self, posF, massF, csF, posB, massB, csB, g: OGrid, R: float):
        self.posF, self.massF, self.csF = f32(posF), f32(massF), i32(csF)
        self.posB, self.massB, self.csB = f32(posB), f32(massB), i32(csB)
        self.g, self.R = g, float(R)
        sc = OScene()
        sc.nF, sc.nB = self.posF.shape[0], self.posB.shape[0]
        sc.posF, sc.massF, sc.csF = _p(self.posF), _p(self.massF), _p(self.csF)
        sc.posB, sc.massB, sc.csB = _p(self.posB), _p(self.massB), _p(self.csB)
        sc.g, sc.R = g, self.R
        self.c = sc

    @property
    def n(self):
        return self.posF.shape[0]


# ---- neighbour search ------------------------------------------------------------------------
def map_cells(pos, g: OGrid):
    pos = f32(pos)
    out = np.empty(pos.shape[0], np.int32)
    lib().oracle_map_cells(_p(pos), C.c_int(pos.shape[0]), C.byref(g), _p(out))
    return out


def neighbor_search(pos, vel, g: OGrid):
    """Returns (sorted pos, sorted vel, p2c [pre-sort order], cell_start, perm)."""
    pos = f32(pos).copy()
    vel = None if vel is None else f32(vel).copy()
    n = pos.shape[0]
    p2c = np.empty(n, np.int32)
    cs = np.empty(g.ncells + 1, np.int32)
    perm = np.empty(n, np.int32)
    lib().oracle_neighbor_search(_p(pos), _p(vel), C.c_int(n), C.byref(g), _p(p2c), _p(cs), _p(perm))
    return pos, vel, p2c, cs, perm


def boundary_mass(pos, cs, g: OGrid, rhoB, R):
    pos = f32(pos)
    mass = np.zeros(pos.shape[0], np.float32)
    lib().oracle_boundary_mass(_p(mass), _p(pos), C.c_int(pos.shape[0]), _p(i32(cs)), C.byref(g),
                               C.c_float(rhoB), C.c_float(R))
    return mass


# ---- sweeps -----------------------------------------------------------------------------------
def density(sc: SceneArrays):
    out = np.zeros(sc.n, np.float32)
    lib().oracle_density(C.byref(sc.c), _p(out))
    return out


def pressure(dens, rho0, stiff):
    dens = f32(dens)
    out = np.zeros_like(dens)
    lib().oracle_pressure(_p(dens), _p(out), C.c_int(dens.shape[0]), C.c_float(rho0), C.c_float(stiff))
    return out


def pressure_force(sc, dens, pres, vel, dt):
    vel = f32(vel).copy()
    lib().oracle_pressure_force(C.byref(sc.c), _p(f32(dens)), _p(f32(pres)), _p(vel), C.c_float(dt))
    return vel


def viscosity(sc, vel, rho0, visc, dt):
    out = np.zeros((sc.n, 3), np.float32)
    lib().oracle_viscosity(C.byref(sc.c), _p(f32(vel)), _p(out), C.c_float(rho0), C.c_float(visc), C.c_float(dt))
    return out


def color_grad(sc, rho0, rhoB):
    out = np.zeros((sc.n, 3), np.float32)
    lib().oracle_color_grad(C.byref(sc.c), _p(out), C.c_float(rho0), C.c_float(rhoB))
    return out


def surface(sc, cgrad, vel, dt, rho0, kappa, airP):
    vel = f32(vel).copy()
    lib().oracle_surface(C.byref(sc.c), _p(f32(cgrad)), _p(vel), C.c_float(dt), C.c_float(rho0),
                         C.c_float(kappa), C.c_float(airP))
    return vel


def advect(pos, vel, dt, space):
    pos, vel = f32(pos).copy(), f32(vel).copy()
    sp = f32(space)
    lib().oracle_advect(_p(pos), _p(vel), C.c_int(pos.shape[0]), C.c_float(dt), _p(sp))
    return pos, vel


def dfsph_density_alpha(sc):
    d, a = np.zeros(sc.n, np.float32), np.zeros(sc.n, np.float32)
    lib().oracle_dfsph_density_alpha(C.byref(sc.c), _p(d), _p(a))
    return d, a


def dfsph_error(sc, vel, dens, alpha, dt, rho0, kind):
    e, k = np.zeros(sc.n, np.float32), np.zeros(sc.n, np.float32)
    fn = lib().oracle_dfsph_div_error if kind == "div" else lib().oracle_dfsph_den_error
    fn(C.byref(sc.c), _p(f32(vel)), _p(f32(dens)), _p(f32(alpha)), _p(e), _p(k), C.c_float(dt), C.c_float(rho0))
    return e, k


def dfsph_correct(sc, stiff, vel, div_by_dt):
    vel = f32(vel).copy()
    lib().oracle_dfsph_correct(C.byref(sc.c), _p(f32(stiff)), _p(vel), C.c_float(div_by_dt))
    return vel


def pbd_density_lambda(sc, rho0, relaxation):
    d, l = np.zeros(sc.n, np.float32), np.zeros(sc.n, np.float32)
    lib().oracle_pbd_density_lambda(C.byref(sc.c), _p(d), _p(l), C.c_float(rho0), C.c_float(relaxation))
    return d, l


def pbd_delta_pos(sc, lam, rho0):
    out = np.zeros((sc.n, 3), np.float32)
    lib().oracle_pbd_delta_pos(C.byref(sc.c), _p(f32(lam)), _p(out), C.c_float(rho0))
    return out


def pbd_xsph(sc, vel, c, rho0):
    out = np.zeros((sc.n, 3), np.float32)
    lib().oracle_pbd_xsph(C.byref(sc.c), _p(f32(vel)), _p(out), C.c_float(c), C.c_float(rho0))
    return out


def export_dots(pos, dens):
    pos, dens = f32(pos), f32(dens)
    dot, col = np.zeros_like(pos), np.zeros_like(pos)
    lib().oracle_export_dots(_p(pos), _p(dens), C.c_int(pos.shape[0]), _p(dot), _p(col))
    return dot, col


# ---- whole system -----------------------------------------------------------------------------
_FIELDS = {"pos": (0, 3, np.float32), "vel": (1, 3, np.float32), "mass": (2, 1, np.float32),
           "density": (3, 1, np.float32), "pressure": (4, 1, np.float32), "p2c": (5, 1, np.int32),
           "cell_start": (6, 0, np.int32), "posB": (7, 3, np.float32), "massB": (8, 1, np.float32),
           "p2cB": (9, 1, np.int32), "cell_startB": (10, 0, np.int32), "alpha": (11, 1, np.float32),
           "kappa": (12, 1, np.float32), "warm": (13, 1, np.float32), "lambda": (14, 1, np.float32),
           "pos_last": (15, 3, np.float32), "buf3": (16, 3, np.float32)}


class OracleSystem:
    """CPU mirror of SPHSystem (SPHSystem.cu:33-158): the constructor performs step 0 (Q3)."""

    def __init__(self, scene, hash_rcp: float = 0.0):
        p = scene.params
        self.nF, self.nB, self.ncells = scene.fluid.shape[0], scene.boundary.shape[0], p.ncells
        fl, bd = f32(scene.fluid), f32(scene.boundary)
        sp, G, cs = f32(p.space), f32(p.gravity), i32(p.cell_size)
        self.h = C.c_void_p(lib().oracle_system_create(
            _p(fl), C.c_int(self.nF), _p(bd), C.c_int(self.nB), _p(sp), C.c_float(p.cell_length),
            C.c_float(p.radius), C.c_float(p.dt), C.c_float(p.m0), C.c_float(p.rho0), C.c_float(p.rho_boundary),
            C.c_float(p.stiff), C.c_float(p.visc), C.c_float(p.surface_tension), C.c_float(p.air_pressure),
            _p(G), _p(cs), C.c_int(p.solver_id), C.c_int(p.max_iter), C.c_float(p.density_error_threshold),
            C.c_float(p.divergence_error_threshold), C.c_float(hash_rcp)))

    def step(self):
        lib().oracle_system_step(self.h)

    def field(self, name):
        which, width, dt = _FIELDS[name]
        ptr = lib().oracle_system_field(self.h, C.c_int(which))
        boundary = name.endswith("B")
        n = (self.ncells + 1) if width == 0 else (self.nB if boundary else self.nF)
        count = n * max(width, 1)
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float if dt == np.float32 else C.c_int)), (count,)).copy()
        return arr.reshape(n, 3) if width == 3 else arr

    def close(self):
        if self.h:
            lib().oracle_system_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def use_all_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every host core."""
    lib().oracle_set_num_threads(C.c_int(os.cpu_count() or 1))
    return num_threads()


def num_threads() -> int:
    return int(lib().oracle_num_threads())
