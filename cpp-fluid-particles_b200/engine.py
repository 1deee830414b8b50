"""torch-tensor front end of the libsphk C-ABI (include/sphk.h).

PyTorch is used for device memory and streams only -- every kernel that runs is a libsphk kernel; there
is no torch arithmetic on the hot path and no CPU fallback.  `SphkSystem` mirrors the reference's
SPHSystem + solver step sequences (SPHSystem.cu:33-158, BasicSPHSolver.cu:237-260, DFSPHSolver.cu:33-72,
PBDSolver.cu:34-73) call for call, like cpp-fluid-particles_b200/host/sph_api.cpp does in C++; it exists
so that python drivers (tests, the multi-GPU slab driver) can interleave their own work (halo exchange)
between the C-ABI calls.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi
from .capi import SphkGrid, SphkParticles, SphkScene, check

EPSILON = 1e-6


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def device_fluid_block(L, lattice, device, stream, j_begin: int = 0, j_count: int | None = None, capacity: int | None = None) -> torch.Tensor:
    """The fluid block of a device-side scene (sphk_scene_fluid_block): (capacity, 3) float32 tensor whose first
    ny * j_count * nz rows are the lattice columns [j_begin, j_begin + j_count) in the reference's push order."""
    (nx, ny, nz), origin = lattice
    j_count = nx - j_begin if j_count is None else j_count
    n = ny * j_count * nz
    out = torch.zeros((max(capacity or n, n, 1), 3), dtype=torch.float32, device=device)
    from .scene import SPACING
    check(L.sphk_scene_fluid_block(_ptr(out), C.c_int(nx), C.c_int(ny), C.c_int(nz), (C.c_float * 3)(*origin), C.c_float(float(SPACING)),
                                   C.c_int(j_begin), C.c_int(j_count), C.c_void_p(stream.cuda_stream)), "sphk_scene_fluid_block")
    return out


def device_boundary_shell(L, params, device, stream) -> torch.Tensor:
    """The six-face boundary shell of a device-side scene (sphk_scene_boundary_shell)."""
    cs = (C.c_int * 3)(*[int(c) for c in params.cell_size])
    n = int(L.sphk_scene_boundary_count(cs))
    out = torch.empty((n, 3), dtype=torch.float32, device=device)
    check(L.sphk_scene_boundary_shell(_ptr(out), cs, (C.c_float * 3)(*[float(x) for x in params.space]), C.c_void_p(stream.cuda_stream)),
          "sphk_scene_boundary_shell")
    return out


class ParticleSet:
    """Device arrays of one SPHParticles object (SPHParticles.h:56-59)."""

    def __init__(self, pos_host, device):
        """pos_host: (n, 3) host array, or a (n, 3) float32 tensor already on `device` (device-side scenes)."""
        n = pos_host.shape[0]
        self.n = n
        if isinstance(pos_host, torch.Tensor):
            self.pos = pos_host
        else:
            self.pos = torch.from_numpy(np.ascontiguousarray(pos_host, np.float32)).to(device)
        self.vel = torch.zeros((n, 3), dtype=torch.float32, device=device)
        self.mass = torch.zeros(n, dtype=torch.float32, device=device)
        self.density = torch.zeros(n, dtype=torch.float32, device=device)
        self.pressure = torch.zeros(n, dtype=torch.float32, device=device)
        self.p2c = torch.zeros(n, dtype=torch.int32, device=device)

    def abi(self) -> SphkParticles:
        p = SphkParticles()
        p.pos, p.vel, p.mass = _ptr(self.pos), _ptr(self.vel), _ptr(self.mass)
        p.density, p.pressure, p.particle2cell = _ptr(self.density), _ptr(self.pressure), _ptr(self.p2c)
        p.n = self.n
        return p


class SphkOps:
    """The C-ABI calls and the three solver sequences over torch tensors.  Subclasses provide the tensors
    (fluid, boundary, cs_fluid, cs_boundary, solver buffers), ctx and params.  The sync_* hooks are no-ops on
    one GPU; the slab driver overrides them with halo exchanges."""

    # ---- multi-GPU hooks ---------------------------------------------------------------------------------
    def sync_vel(self):
        pass

    def sync_scalar(self, t):
        pass

    def sync_array(self, t):
        pass

    def sync_positions(self):
        pass

    def reduce_sum(self, x: float) -> float:
        return x

    def owned(self, t):
        """The slice of a per-particle tensor this rank owns (all of it on one GPU)."""
        return t

    def n_total(self) -> int:
        return self.fluid.n

    def _alloc_solver_buffers(self, n):
        dev = self.device
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)  # noqa: E731
        self.buffer3 = f(n, 3)          # BasicSPHSolver::bufferFloat3
        self.color_grad_buf = f(n, 3)   # second buffer: the fused viscosity+surface sweep needs deltaV and colour gradient at once
        self.fused = True               # fused sweeps (default, like the C++ class layer) or one kernel per launch site
        if self.solver == "dfsph":      # DFSPHSolver.h:58-62
            self.alpha, self.kappa, self.error, self.warm = f(n), f(n), f(n), f(n)
            self.max_iter = self.p.max_iter if self.p.max_iter > 0 else 20
            self.den_thr = self.p.density_error_threshold if self.p.max_iter > 0 else 1e-3
            self.div_thr = self.p.divergence_error_threshold if self.p.max_iter > 0 else 1e-3
        if self.solver == "pbd":        # PBDSolver.h:77-84
            self.pos_last, self.dpos, self.lam = f(n, 3), f(n, 3), f(n)
            self.pos_last_init = False
            self.max_iter = self.p.max_iter if self.p.max_iter > 0 else 20
            self.xsph_c, self.relaxation = 0.05, 0.75
        self.it_div = self.it_den = 0

    # ---- C-ABI plumbing ---------------------------------------------------------------------------
    def scene_abi(self) -> SphkScene:
        if self._scene is None:
            s = SphkScene()
            s.fluid, s.boundary = self.fluid.abi(), self.boundary.abi()
            s.cell_start_fluid, s.cell_start_boundary = _ptr(self.cs_fluid), _ptr(self.cs_boundary)
            s.radius = self.p.radius
            self._scene = s
        return self._scene

    def _s(self):
        return C.byref(self.scene_abi())

    def set_use_list(self, on: bool, skin_permille: int = 0):
        check(self.L.sphk_set_option(self.ctx, capi.OPT_NEIGHBOR_LIST, 1 if on else 0))
        check(self.L.sphk_set_option(self.ctx, capi.OPT_LIST_SKIN, int(skin_permille)))

    def build_neighbor_list(self):
        check(self.L.sphk_build_neighbor_list(self.ctx, self._s()), "sphk_build_neighbor_list")

    def search_boundary(self):
        p = self.boundary.abi()
        check(self.L.sphk_neighbor_search(self.ctx, 1, C.byref(p), _ptr(self.cs_boundary)), "sphk_neighbor_search(b)")

    def search_fluid(self):
        p = self.fluid.abi()
        check(self.L.sphk_neighbor_search(self.ctx, 0, C.byref(p), _ptr(self.cs_fluid)), "sphk_neighbor_search(f)")

    def boundary_mass(self):
        p = self.boundary.abi()
        check(self.L.sphk_boundary_mass(self.ctx, C.byref(p), _ptr(self.cs_boundary), C.c_float(self.p.rho_boundary),
                                        C.c_float(self.p.radius)), "sphk_boundary_mass")

    def permutation(self) -> torch.Tensor:
        out = torch.empty(self.fluid.n, dtype=torch.int32, device=self.device)
        check(self.L.sphk_get_permutation(self.ctx, _ptr(out), C.c_int(self.fluid.n)))
        return out

    def device_rcp(self, x: float) -> float:
        out = C.c_float()
        check(self.L.sphk_device_rcp(self.ctx, C.c_float(x), C.byref(out)))
        return float(out.value)

    def list_stats(self):
        out = (C.c_longlong * 3)()
        check(self.L.sphk_list_stats(self.ctx, self._s(), out))
        return {"max": int(out[0]), "overflow": int(out[1]), "total": int(out[2])}

    def neighbor_list(self, capacity: int = 96):
        """(counts[n], entries[capacity/4, cap, 4]) of the current list, as torch tensors."""
        n, cap = self.fluid.n, self.fluid.pos.shape[0]
        cnt = torch.zeros(n, dtype=torch.int32, device=self.device)
        ent = torch.zeros(capacity * cap, dtype=torch.int32, device=self.device)
        check(self.L.sphk_get_neighbor_list(self.ctx, self._s(), _ptr(cnt), _ptr(ent)))
        return cnt, ent.view(capacity // 4, cap, 4)

    def set_option(self, opt: int, value: int):
        check(self.L.sphk_set_option(self.ctx, int(opt), int(value)))

    def export_dots(self):
        """generate_dots (vbo.cu:26-51): (dot, colour) device tensors for the fluid set."""
        n = self.fluid.n
        dot = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        col = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        p = self.fluid.abi()
        check(self.L.sphk_export_dots(self.ctx, C.byref(p), _ptr(dot), _ptr(col)), "sphk_export_dots")
        return dot, col

    def launch_count(self) -> int:
        return int(self.L.sphk_launch_count(self.ctx))

    def synchronize(self):
        check(self.L.sphk_synchronize(self.ctx))

    def refresh(self):
        check(self.L.sphk_refresh(self.ctx, self._s()))

    # ---- one entry per reference launch site --------------------------------------------------------
    def gravity(self):
        check(self.L.sphk_gravity(self.ctx, self._s(), C.c_float(self.p.dt), self._G), "sphk_gravity")

    def viscosity(self):
        check(self.L.sphk_viscosity(self.ctx, self._s(), _ptr(self.buffer3), C.c_float(self.p.rho0),
                                    C.c_float(self.p.visc), C.c_float(self.p.dt)), "sphk_viscosity")

    def color_grad(self):
        check(self.L.sphk_color_grad(self.ctx, self._s(), _ptr(self.buffer3), C.c_float(self.p.rho0),
                                     C.c_float(self.p.rho_boundary)), "sphk_color_grad")

    def surface(self):
        check(self.L.sphk_surface(self.ctx, self._s(), _ptr(self.buffer3), C.c_float(self.p.dt), C.c_float(self.p.rho0),
                                  C.c_float(self.p.surface_tension), C.c_float(self.p.air_pressure)), "sphk_surface")

    def density(self):
        check(self.L.sphk_density(self.ctx, self._s()), "sphk_density")

    def fused_density_color_grad(self, with_alpha: bool):
        if with_alpha:
            check(self.L.sphk_fused_dfsph_density_alpha_color_grad(self.ctx, self._s(), _ptr(self.alpha), _ptr(self.color_grad_buf),
                                                                   C.c_float(self.p.rho0), C.c_float(self.p.rho_boundary)),
                  "sphk_fused_dfsph_density_alpha_color_grad")
        else:
            check(self.L.sphk_fused_density_color_grad(self.ctx, self._s(), _ptr(self.color_grad_buf), C.c_float(self.p.rho0),
                                                       C.c_float(self.p.rho_boundary)), "sphk_fused_density_color_grad")

    def fused_density_alpha_div_error(self, with_color_grad: bool):
        check(self.L.sphk_fused_dfsph_density_alpha_div_error(self.ctx, self._s(), _ptr(self.alpha),
                                                              _ptr(self.color_grad_buf) if with_color_grad else None,
                                                              C.c_float(self.p.rho0), C.c_float(self.p.rho_boundary), _ptr(self.error),
                                                              _ptr(self.kappa), C.c_float(self.p.dt)),
              "sphk_fused_dfsph_density_alpha_div_error")

    def fused_viscosity_surface(self):
        check(self.L.sphk_fused_viscosity_surface(self.ctx, self._s(), _ptr(self.buffer3), _ptr(self.color_grad_buf),
                                                  C.c_float(self.p.rho0), C.c_float(self.p.visc), C.c_float(self.p.dt),
                                                  C.c_float(self.p.surface_tension), C.c_float(self.p.air_pressure)),
              "sphk_fused_viscosity_surface")

    def pressure(self):
        check(self.L.sphk_pressure(self.ctx, self._s(), C.c_float(self.p.rho0), C.c_float(self.p.stiff)), "sphk_pressure")

    def pressure_force(self):
        check(self.L.sphk_pressure_force(self.ctx, self._s(), C.c_float(self.p.dt)), "sphk_pressure_force")

    def advect(self):
        check(self.L.sphk_advect(self.ctx, self._s(), C.c_float(self.p.dt), self._space), "sphk_advect")

    def dfsph_density_alpha(self):
        check(self.L.sphk_dfsph_density_alpha(self.ctx, self._s(), _ptr(self.alpha)), "sphk_dfsph_density_alpha")

    def dfsph_div_error(self):
        check(self.L.sphk_dfsph_div_error(self.ctx, self._s(), _ptr(self.alpha), _ptr(self.error), _ptr(self.kappa),
                                          C.c_float(self.p.dt), C.c_float(self.p.rho0)), "sphk_dfsph_div_error")

    def dfsph_div_correct(self, stiff=None):
        check(self.L.sphk_dfsph_div_correct(self.ctx, self._s(), _ptr(self.kappa if stiff is None else stiff)),
              "sphk_dfsph_div_correct")

    def dfsph_den_error(self, accumulate_warm: bool):
        check(self.L.sphk_dfsph_den_error(self.ctx, self._s(), _ptr(self.alpha), _ptr(self.error), _ptr(self.kappa),
                                          C.c_float(self.p.dt), C.c_float(self.p.rho0),
                                          _ptr(self.warm) if accumulate_warm else None), "sphk_dfsph_den_error")

    def dfsph_den_correct(self, stiff=None):
        check(self.L.sphk_dfsph_den_correct(self.ctx, self._s(), _ptr(self.kappa if stiff is None else stiff),
                                            C.c_float(self.p.dt)), "sphk_dfsph_den_correct")

    def reduce_abs_sum(self, x: torch.Tensor) -> float:
        out = C.c_float()
        check(self.L.sphk_reduce_abs_sum(self.ctx, _ptr(x), C.c_int(x.numel()), C.byref(out)))
        return float(out.value)

    def permute(self, arr: torch.Tensor, width: int):
        check(self.L.sphk_permute(self.ctx, _ptr(arr), C.c_int(width), C.c_int(self.fluid.n)), "sphk_permute")

    def copy(self, dst: torch.Tensor, src: torch.Tensor):
        check(self.L.sphk_copy(self.ctx, _ptr(dst), _ptr(src), C.c_int(src.numel())), "sphk_copy")

    def pbd_density_lambda(self):
        check(self.L.sphk_pbd_density_lambda(self.ctx, self._s(), _ptr(self.lam), C.c_float(self.p.rho0),
                                             C.c_float(self.relaxation)), "sphk_pbd_density_lambda")

    def pbd_delta_pos_apply(self):
        check(self.L.sphk_pbd_delta_pos_apply(self.ctx, self._s(), _ptr(self.lam), _ptr(self.dpos),
                                              C.c_float(self.p.rho0), self._space), "sphk_pbd_delta_pos_apply")

    def pbd_velocity_from_positions(self):
        check(self.L.sphk_pbd_velocity_from_positions(self.ctx, self._s(), _ptr(self.pos_last), C.c_float(self.p.dt)),
              "sphk_pbd_velocity_from_positions")

    def fused_pbd_xsph_color_grad(self):
        check(self.L.sphk_fused_pbd_xsph_color_grad(self.ctx, self._s(), C.c_float(self.xsph_c), C.c_float(self.p.rho0), _ptr(self.buffer3),
                                                    C.c_float(self.p.rho_boundary)), "sphk_fused_pbd_xsph_color_grad")

    def pbd_xsph(self):
        check(self.L.sphk_pbd_xsph(self.ctx, self._s(), C.c_float(self.xsph_c), C.c_float(self.p.rho0)), "sphk_pbd_xsph")

    # ---- solver sequences ----------------------------------------------------------------------------
    def _surface_enabled(self) -> bool:
        return self.p.surface_tension > EPSILON or self.p.air_pressure > EPSILON

    def _handle_surface(self):       # BasicSPHSolver.cu:262-275
        if self._surface_enabled():
            self.color_grad()
            self.surface()

    def _run(self, op, sync=None, tensor=None, split=True):
        """One sweep followed by the halo refresh of what it produced (`sync`: None | "vel" | "scalar" | "array").
        On one GPU the refresh is a no-op; the slab driver overrides this to overlap the exchange with the sweep
        (split=False: the sweep reads the field it writes -- Jacobi velocity updates -- and must run in one piece)."""
        op()
        if sync == "vel":
            self.sync_vel()
        elif sync == "scalar":
            self.sync_scalar(tensor)
        elif sync == "array":
            self.sync_array(tensor)

    def step_wcsph(self):            # BasicSPHSolver.cu:237-260
        self.set_use_list(self.use_list)
        self.gravity()
        if self.fused and self._surface_enabled():
            self._run(lambda: self.fused_density_color_grad(False), "array", self.color_grad_buf)
            self._run(self.fused_viscosity_surface, "vel", split=False)
        else:
            self._run(self.viscosity, "vel", split=False)
            if self._surface_enabled():  # handleSurface, :262-275
                self._run(self.color_grad, "array", self.buffer3)
                self._run(self.surface, "vel")
            self.density()
        self.pressure()
        self.sync_array(self.fluid.density); self.sync_array(self.fluid.pressure)
        self.pressure_force()
        self.advect()

    def step_dfsph(self):            # DFSPHSolver.cu:33-72
        self.set_use_list(self.use_list)
        n, rho0 = self.n_total(), self.p.rho0
        fused = self.fused and self._surface_enabled()
        if self.fused:               # density/alpha (+ colour gradient) + the first divergence error (:341) in one sweep
            self._run(lambda: self.fused_density_alpha_div_error(fused), "scalar", self.kappa)
            if fused:
                self.sync_array(self.color_grad_buf)
        else:
            self.dfsph_density_alpha()
            self._run(self.dfsph_div_error, "scalar", self.kappa)
        total, it = 3.4e38, 0        # correctDivergenceError :331-363
        while (it < 1 or total > self.div_thr * n * rho0) and it < self.max_iter:
            self._run(self.dfsph_div_correct, "vel")
            self._run(self.dfsph_div_error, "scalar", self.kappa)
            it += 1
            if self.div_thr >= 0:    # negative threshold: the test cannot depend on the sum (Q11) -> no host sync
                total = self.reduce_sum(self.reduce_abs_sum(self.owned(self.error)))
        self.it_div = it
        self.gravity()
        if fused:
            self._run(self.fused_viscosity_surface, "vel", split=False)
        else:
            self._run(self.viscosity, "vel", split=False)
            if self._surface_enabled():
                self._run(self.color_grad, "array", self.buffer3)
                self._run(self.surface, "vel")
        total, it = 3.4e38, 0        # project :160-210
        self.permute(self.warm, 1)
        self._run(lambda: self.dfsph_den_correct(self.warm), "vel")
        self._run(lambda: self.dfsph_den_error(False), "scalar", self.kappa)
        self.copy(self.warm, self.kappa)
        while (it < 2 or total > self.den_thr * n * rho0) and it < self.max_iter:
            self._run(self.dfsph_den_correct, "vel")
            self._run(lambda: self.dfsph_den_error(True), "scalar", self.kappa)
            it += 1
            if it >= 2 and self.den_thr >= 0:
                total = self.reduce_sum(self.reduce_abs_sum(self.owned(self.error)))
        self.it_den = it
        self.advect()

    def step_pbd(self) -> bool:      # PBDSolver.cu:34-73; returns False for the "throwing" first call (Q6)
        if not self.pos_last_init:
            self.copy(self.pos_last, self.fluid.pos)
            self.pos_last_init = True
            return False
        self.set_use_list(self.use_list, 150 if self.use_list else 0)   # skin list: positions move inside the step
        self.permute(self.pos_last, 3)
        for _ in range(self.max_iter):
            self._run(self.pbd_density_lambda, "scalar", self.lam)
            self.pbd_delta_pos_apply(); self.sync_positions()
        self.pbd_velocity_from_positions()
        if self.fused and self._surface_enabled():   # XSPH + colour gradient in one sweep, then the surface sweep
            self._run(self.fused_pbd_xsph_color_grad, "vel", split=False)
            self.sync_array(self.buffer3)
            self._run(self.surface, "vel")
        else:
            self._run(self.pbd_xsph, "vel", split=False)
            if self._surface_enabled():
                self._run(self.color_grad, "array", self.buffer3)
                self._run(self.surface, "vel")
        self.gravity()
        self.copy(self.pos_last, self.fluid.pos)
        self.advect()
        return True

    def step(self):                  # SPHSystem.cu:129-158 (no host sync here; callers time with CUDA events)
        self.search_fluid()
        if self.use_list:            # (the C++ layer builds the list lazily in the first sweep; same kernel)
            self.set_use_list(True, 150 if self.solver == "pbd" else 0)
            self.build_neighbor_list()
        if self.solver == "dfsph":
            self.step_dfsph()
        elif self.solver == "pbd":
            self.step_pbd()
        else:
            self.step_wcsph()

    def state(self) -> dict:
        self.synchronize()
        g = lambda t: t.detach().cpu().numpy()  # noqa: E731
        return {"pos": g(self.fluid.pos), "vel": g(self.fluid.vel), "density": g(self.fluid.density),
                "pressure": g(self.fluid.pressure), "mass": g(self.fluid.mass), "p2c": g(self.fluid.p2c),
                "cell_start": g(self.cs_fluid), "posB": g(self.boundary.pos), "massB": g(self.boundary.mass),
                "p2cB": g(self.boundary.p2c), "cell_startB": g(self.cs_boundary)}

    def close(self):
        if self.ctx:
            self.L.sphk_synchronize(self.ctx)
            self.L.sphk_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SphkSystem(SphkOps):
    """SPHSystem + solver over the C-ABI with torch tensors.  construct(step0=True) reproduces the
    reference constructor including its implicit first step (Q3)."""

    def __init__(self, scene, device="cuda:0", use_list: bool | None = None, list_capacity: int | None = None,
                 step0: bool = True):
        self.L = capi.sphk()
        self.p = scene.params
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.current_stream(self.device)
        if scene.fluid is None:      # device-side scene: no host particle arrays
            self.fluid = ParticleSet(device_fluid_block(self.L, scene.lattice, self.device, self.stream), self.device)
            self.boundary = ParticleSet(device_boundary_shell(self.L, self.p, self.device, self.stream), self.device)
        else:
            self.fluid = ParticleSet(scene.fluid, self.device)
            self.boundary = ParticleSet(scene.boundary, self.device)
        nc = self.p.ncells
        self.cs_fluid = torch.zeros(nc + 1, dtype=torch.int32, device=self.device)
        self.cs_boundary = torch.zeros(nc + 1, dtype=torch.int32, device=self.device)
        g = SphkGrid()
        g.cell_size[:] = [int(c) for c in self.p.cell_size]
        g.cell_length = self.p.cell_length
        self.ctx = C.c_void_p()
        check(self.L.sphk_create(C.byref(self.ctx), C.c_int(self.fluid.n), C.c_int(self.boundary.n), C.byref(g),
                                 C.c_void_p(self.stream.cuda_stream)), "sphk_create")
        if list_capacity is not None:
            check(self.L.sphk_set_option(self.ctx, capi.OPT_LIST_CAPACITY, int(list_capacity)))
        self.solver = self.p.solver
        self.use_list = True if use_list is None else bool(use_list)
        n = self.fluid.n
        self._alloc_solver_buffers(n)
        self._scene = None
        self._G = (C.c_float * 3)(*[float(x) for x in self.p.gravity])
        self._space = (C.c_float * 3)(*[float(x) for x in self.p.space])
        # SPHSystem.cu:68-76
        self.search_boundary()
        self.boundary_mass()
        check(self.L.sphk_fill(self.ctx, _ptr(self.fluid.mass), C.c_int(n), C.c_float(self.p.m0)), "sphk_fill")
        self.search_fluid()
        if step0:
            self.step()
