"""ctypes bindings of the two in-tree native libraries.

  libsphk.so     the C-ABI of include/sphk.h (CUDA kernels, sm_100a)
  libsphhost.so  the reference-shaped C++ classes + headless facade of include/sph_app.h

There is no CPU fallback anywhere in this package: a missing library raises ImportError-like
RuntimeError immediately, and sphk_create fails with SPHK_ERR_NO_DEVICE on a box without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBSPHK = os.path.join(HERE, "libsphk.so")
LIBHOST = os.path.join(HERE, "libsphhost.so")

SPHK_FUNCTIONS = [
    "sphk_create", "sphk_destroy", "sphk_set_option", "sphk_set_grid", "sphk_synchronize", "sphk_error_string", "sphk_launch_count",
    "sphk_device_rcp", "sphk_neighbor_search", "sphk_permute", "sphk_refresh", "sphk_boundary_mass", "sphk_fill",
    "sphk_gravity", "sphk_viscosity", "sphk_color_grad", "sphk_surface", "sphk_density", "sphk_pressure",
    "sphk_pressure_force", "sphk_advect", "sphk_dfsph_density_alpha", "sphk_dfsph_div_error", "sphk_dfsph_div_correct",
    "sphk_dfsph_den_error", "sphk_dfsph_den_correct", "sphk_reduce_abs_sum", "sphk_copy", "sphk_pbd_density_lambda",
    "sphk_pbd_delta_pos_apply", "sphk_pbd_velocity_from_positions", "sphk_pbd_xsph", "sphk_get_permutation",
    "sphk_list_stats", "sphk_get_skin_displacement", "sphk_set_active_range", "sphk_push_range", "sphk_build_neighbor_list", "sphk_get_neighbor_list", "sphk_fused_density_color_grad",
    "sphk_fused_dfsph_density_alpha_color_grad", "sphk_fused_dfsph_density_alpha_div_error", "sphk_fused_pbd_xsph_color_grad", "sphk_scene_fluid_block", "sphk_scene_boundary_count", "sphk_scene_boundary_shell", "sphk_loop_begin", "sphk_loop_next", "sphk_loop_end", "sphk_loop_iterations", "sphk_fused_viscosity_surface", "sphk_export_dots", "sphk_particles_advect", "sphk_add_launches",
    "sphk_mg_unique_id", "sphk_mg_init", "sphk_mg_destroy", "sphk_mg_ipc_handle", "sphk_mg_ipc_connect", "sphk_mg_set_transport",
    "sphk_mg_exchange_ints", "sphk_mg_exchange_ints_async", "sphk_mg_plane_ranges", "sphk_mg_halo_device", "sphk_mg_check_async", "sphk_set_active_range_device", "sphk_mg_allreduce_sum", "sphk_mg_exchange_slices", "sphk_mg_halo", "sphk_mg_check", "sphk_mg_stats",
    "sphk_strays_block_floats", "sphk_strays_collect", "sphk_strays_append", "sphk_mg_strays_route", "sphk_strays_counts",
]
SPH_APP_FUNCTIONS = [
    "sph_app_create", "sph_app_destroy", "sph_app_step", "sph_app_fluid_size", "sph_app_boundary_size",
    "sph_app_download_fluid", "sph_app_download_boundary", "sph_app_upload_fluid", "sph_app_engine",
    "sph_app_submit", "sph_app_wait", "sph_app_dfsph_iterations", "sph_app_set_option",
]

OPT_NEIGHBOR_LIST, OPT_LIST_CAPACITY, OPT_LIST_SKIN, OPT_SIMPLE_LIST_BUILD, OPT_TILE, OPT_STAGED_LIST_BUILD, OPT_PATCH = 1, 2, 5, 6, 9, 10, 11


class SphkGrid(C.Structure):
    _fields_ = [("cell_size", C.c_int * 3), ("cell_length", C.c_float), ("origin", C.c_int * 3)]


class SphkParticles(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("vel", C.c_void_p), ("mass", C.c_void_p), ("density", C.c_void_p),
                ("pressure", C.c_void_p), ("particle2cell", C.c_void_p), ("n", C.c_int)]


class SphkScene(C.Structure):
    _fields_ = [("fluid", SphkParticles), ("boundary", SphkParticles), ("cell_start_fluid", C.c_void_p),
                ("cell_start_boundary", C.c_void_p), ("radius", C.c_float)]


class SphAppParams(C.Structure):
    _fields_ = [("space", C.c_float * 3), ("cell_length", C.c_float), ("radius", C.c_float), ("dt", C.c_float),
                ("m0", C.c_float), ("rho0", C.c_float), ("rho_boundary", C.c_float), ("stiff", C.c_float),
                ("visc", C.c_float), ("surface_tension", C.c_float), ("air_pressure", C.c_float),
                ("gravity", C.c_float * 3), ("cell_size", C.c_int * 3), ("solver", C.c_int), ("max_iter", C.c_int),
                ("density_error_threshold", C.c_float), ("divergence_error_threshold", C.c_float)]


_sphk = None
_apps = {}


def _load(path: str):
    if not os.path.exists(path):
        raise RuntimeError(f"{os.path.basename(path)} is not built ({path}); run __graft_entry__.build() -- "
                           "this package has no CPU or pure-python fallback")
    # RTLD_LOCAL: libsphhost.so and the reference build oracle/_ref/libsphref.so export the SAME C++ class
    # symbols (SPHSystem::step, ...) by design; they must never interpose each other in one process
    return C.CDLL(path, mode=C.RTLD_LOCAL)


def sphk():
    """libsphk.so with argument/return types set."""
    global _sphk
    if _sphk is None:
        L = _load(LIBSPHK)
        L.sphk_error_string.restype = C.c_char_p
        L.sphk_launch_count.restype = C.c_longlong
        L.sphk_scene_boundary_count.restype = C.c_longlong
        L.sphk_strays_block_floats.restype = C.c_longlong
        L.sphk_destroy.restype = None
        L.sphk_mg_destroy.restype = None
        _sphk = L
    return _sphk


def app_lib(path: str = LIBHOST):
    """A library exporting the sph_app_* facade: libsphhost.so (product) or oracle/_ref/libsphref.so."""
    if path not in _apps:
        if path == LIBHOST:
            sphk()
        L = _load(path)
        L.sph_app_create.restype = C.c_void_p
        L.sph_app_step.restype = C.c_float
        L.sph_app_engine.restype = C.c_char_p
        L.sph_app_destroy.restype = None
        _apps[path] = L
    return _apps[path]


def check(rc: int, what: str = "sphk"):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {sphk().sphk_error_string(C.c_int(rc)).decode()} ({rc})")


def app_params(p) -> SphAppParams:
    """SceneParams (scene.py) -> sph_app_params."""
    q = SphAppParams()
    q.space[:] = [float(x) for x in p.space]
    q.cell_length, q.radius, q.dt, q.m0 = p.cell_length, p.radius, p.dt, p.m0
    q.rho0, q.rho_boundary, q.stiff, q.visc = p.rho0, p.rho_boundary, p.stiff, p.visc
    q.surface_tension, q.air_pressure = p.surface_tension, p.air_pressure
    q.gravity[:] = [float(x) for x in p.gravity]
    q.cell_size[:] = [int(x) for x in p.cell_size]
    q.solver, q.max_iter = p.solver_id, int(p.max_iter)
    q.density_error_threshold = p.density_error_threshold
    q.divergence_error_threshold = p.divergence_error_threshold
    return q


class SphApp:
    """One headless SPHSystem driven through the sph_app facade (numpy in / numpy out).

    `lib_path` selects the engine underneath: this repo's libsphhost.so (default) or the reference's own
    CUDA build oracle/_ref/libsphref.so -- the facade source is the same file for both."""

    def __init__(self, scene, lib_path: str = LIBHOST):
        self.L = app_lib(lib_path)
        self.scene = scene
        fl = np.ascontiguousarray(scene.fluid, np.float32)
        bd = np.ascontiguousarray(scene.boundary, np.float32)
        self.nF, self.nB = fl.shape[0], bd.shape[0]
        prm = app_params(scene.params)
        self.h = self.L.sph_app_create(fl.ctypes.data_as(C.c_void_p), C.c_int(self.nF), bd.ctypes.data_as(C.c_void_p),
                                       C.c_int(self.nB), C.byref(prm))
        if not self.h:
            raise RuntimeError("sph_app_create failed (no CUDA device? there is no CPU fallback)")
        self.h = C.c_void_p(self.h)

    @property
    def engine(self) -> str:
        return self.L.sph_app_engine().decode()

    def step(self) -> float:
        return float(self.L.sph_app_step(self.h))

    def download(self) -> dict:
        out = {"pos": np.empty((self.nF, 3), np.float32), "vel": np.empty((self.nF, 3), np.float32),
               "density": np.empty(self.nF, np.float32), "pressure": np.empty(self.nF, np.float32),
               "mass": np.empty(self.nF, np.float32), "p2c": np.empty(self.nF, np.int32)}
        rc = self.L.sph_app_download_fluid(self.h, *[out[k].ctypes.data_as(C.c_void_p) for k in
                                                     ("pos", "vel", "density", "pressure", "mass", "p2c")])
        if rc:
            raise RuntimeError("sph_app_download_fluid failed")
        return out

    def download_into(self, pos=None, vel=None, density=None):
        """D2H into caller-provided (pinned) buffers; None skips a field."""
        def p(a):
            return C.c_void_p(a.ctypes.data) if a is not None else None
        rc = self.L.sph_app_download_fluid(self.h, p(pos), p(vel), p(density), None, None, None)
        if rc:
            raise RuntimeError("sph_app_download_fluid failed")

    def download_boundary(self) -> dict:
        out = {"pos": np.empty((self.nB, 3), np.float32), "mass": np.empty(self.nB, np.float32),
               "p2c": np.empty(self.nB, np.int32)}
        rc = self.L.sph_app_download_boundary(self.h, *[out[k].ctypes.data_as(C.c_void_p) for k in ("pos", "mass", "p2c")])
        if rc:
            raise RuntimeError("sph_app_download_boundary failed")
        return out

    def upload(self, pos=None, vel=None):
        def p(a):
            return C.c_void_p(a.ctypes.data) if a is not None else None
        if self.L.sph_app_upload_fluid(self.h, p(pos), p(vel)):
            raise RuntimeError("sph_app_upload_fluid failed")

    def submit(self, pos_in, vel_in, pos_out=None, vel_out=None, density_out=None):
        """Pipelined host-buffer step (include/sph_app.h): numpy arrays over pinned host memory."""
        def p(a):
            return C.c_void_p(a.ctypes.data) if a is not None else None
        if self.L.sph_app_submit(self.h, p(pos_in), p(vel_in), p(pos_out), p(vel_out), p(density_out)):
            raise RuntimeError("sph_app_submit failed")

    def wait(self):
        if self.L.sph_app_wait(self.h):
            raise RuntimeError("sph_app_wait failed")

    def dfsph_iterations(self):
        """(divergence, density) iteration counts of the last DFSPH step."""
        a, b = C.c_int(0), C.c_int(0)
        if self.L.sph_app_dfsph_iterations(self.h, C.byref(a), C.byref(b)):
            raise RuntimeError("sph_app_dfsph_iterations: not a DFSPH system of this engine")
        return int(a.value), int(b.value)

    def set_option(self, option: int, value: int):
        """1: device-side loop tests (adaptive DFSPH), 2: fused sweeps, 3: step graph."""
        if self.L.sph_app_set_option(self.h, C.c_int(option), C.c_int(value)):
            raise RuntimeError("sph_app_set_option failed")

    def close(self):
        if getattr(self, "h", None):
            self.L.sph_app_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
