"""Synthetic dam-break scenes: a parameterised restatement of initSPHSystem().

Follows /root/reference/src/main.cpp:54-67 (constants) and :73-116 (fluid lattice, six-face boundary
lattice, in the reference's push order).  The reference has no RNG: scenes are deterministic lattices;
`jitter` adds U(-j, j) per coordinate (numpy Generator, seed given) to exercise off-lattice parity.

All arithmetic is float32, evaluated in the order the host code of main.cpp evaluates it.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

F = np.float32


@dataclasses.dataclass
class SceneParams:
    """The 16 scalar arguments of the SPHSystem constructor (SPHSystem.h:22-38) + solver choice."""
    space: tuple          # spaceSize                     main.cpp:54
    cell_length: float    # sphCellLength                 main.cpp:57
    radius: float         # sphSmoothingRadius            main.cpp:56
    dt: float             #                               main.cpp:58
    m0: float             # sphM0                         main.cpp:61
    rho0: float           # sphRho0                       main.cpp:59
    rho_boundary: float   # sphRhoBoundary                main.cpp:60
    stiff: float          # sphStiff                      main.cpp:62
    visc: float           # sphVisc                       main.cpp:64
    surface_tension: float  # sphSurfaceTensionIntensity  main.cpp:65
    air_pressure: float   # sphAirPressure                main.cpp:66
    gravity: tuple        # sphG                          main.cpp:63
    cell_size: tuple      # cellSize                      main.cpp:67
    solver: str = "wcsph"     # "wcsph" | "dfsph" | "pbd"  (main.cpp:69-71)
    max_iter: int = 0         # 0 -> constructor default (20)
    density_error_threshold: float = 1e-3
    divergence_error_threshold: float = 1e-3

    @property
    def ncells(self) -> int:
        return int(self.cell_size[0]) * int(self.cell_size[1]) * int(self.cell_size[2])

    @property
    def solver_id(self) -> int:
        return {"wcsph": 0, "sph": 0, "dfsph": 1, "pbd": 2}[self.solver]


SPACING = F(0.02)                                   # sphSpacing main.cpp:55
RADIUS = F(2.0) * SPACING                           # main.cpp:56
CELL_LENGTH = F(1.01) * RADIUS                      # main.cpp:57
M0 = F(76.596750762082e-6)                          # main.cpp:61


def default_params(box, solver: str = "wcsph", dt: float | None = None, max_iter: int = 0,
                   den_thr: float = 1e-3, div_thr: float = 1e-3) -> SceneParams:
    """box: edge length (the reference's cubic spaceSize, main.cpp:54) or a per-axis (Lx, Ly, Lz) tuple
    (weak-scaling scenes stretch the box; the SPHSystem API takes float3 spaceSize / int3 cellSize anyway)."""
    Ls = tuple(F(b) for b in (box if isinstance(box, (tuple, list)) else (box,) * 3))
    ncells = tuple(int(math.ceil(float(L / CELL_LENGTH))) for L in Ls)   # main.cpp:67 (float divide, then ceil)
    if dt is None:
        dt = 0.001 if solver in ("wcsph", "sph") else 0.004   # README.md:7-9 / BASELINE.md
    return SceneParams(space=tuple(float(L) for L in Ls), cell_length=float(CELL_LENGTH), radius=float(RADIUS),
                       dt=float(F(dt)), m0=float(M0), rho0=1.0, rho_boundary=float(F(1.4) * F(1.0)),
                       stiff=10.0, visc=float(F(5e-4)), surface_tension=float(F(1e-4)),
                       air_pressure=float(F(1e-4)), gravity=(0.0, float(F(-9.8)), 0.0),
                       cell_size=ncells, solver=solver, max_iter=max_iter,
                       density_error_threshold=den_thr, divergence_error_threshold=div_thr)


def fluid_lattice(nx: int, ny: int, nz: int, origin) -> np.ndarray:
    """main.cpp:75-85: i (y) outermost, then j (x), k (z) innermost."""
    ox, oy, oz = (F(o) for o in origin)
    i, j, k = np.meshgrid(np.arange(ny, dtype=F), np.arange(nx, dtype=F), np.arange(nz, dtype=F), indexing="ij")
    pos = np.empty((ny, nx, nz, 3), dtype=F)
    pos[..., 0] = ox + SPACING * j
    pos[..., 1] = oy + SPACING * i
    pos[..., 2] = oz + SPACING * k
    return pos.reshape(-1, 3)


def boundary_lattice(cell_size, space) -> np.ndarray:
    """main.cpp:89-116: six faces of a (2*cellSize)^3 lattice, mapped by 0.99*x + 0.005*space."""
    cx, cy, cz = (2 * int(c) for c in cell_size)
    sp = np.asarray(space, dtype=F)
    den = np.array([cx - 1, cy - 1, cz - 1], dtype=F)

    def emit(ijk: np.ndarray) -> np.ndarray:
        x = ijk.astype(F) / den * sp
        return F(0.99) * x + F(0.005) * sp

    def pairs(a: np.ndarray, b: np.ndarray) -> np.ndarray:
        out = np.empty((a.shape[0] * 2, 3), dtype=F)
        out[0::2] = emit(a)
        out[1::2] = emit(b)
        return out

    chunks = []
    # front and back (:91-98)
    i, j = np.meshgrid(np.arange(cx), np.arange(cy), indexing="ij")
    i, j = i.ravel(), j.ravel()
    chunks.append(pairs(np.stack([i, j, np.zeros_like(i)], 1), np.stack([i, j, np.full_like(i, cz - 1)], 1)))
    # top and bottom (:100-107)
    i, j = np.meshgrid(np.arange(cx), np.arange(cz - 2), indexing="ij")
    i, j = i.ravel(), j.ravel()
    chunks.append(pairs(np.stack([i, np.zeros_like(i), j + 1], 1), np.stack([i, np.full_like(i, cy - 1), j + 1], 1)))
    # left and right (:109-116)
    i, j = np.meshgrid(np.arange(cy - 2), np.arange(cz - 2), indexing="ij")
    i, j = i.ravel(), j.ravel()
    chunks.append(pairs(np.stack([np.zeros_like(i), i + 1, j + 1], 1), np.stack([np.full_like(i, cx - 1), i + 1, j + 1], 1)))
    return np.ascontiguousarray(np.concatenate(chunks, 0))


@dataclasses.dataclass
class Scene:
    name: str
    fluid: np.ndarray | None      # (n_fluid, 3) float32; None for a device-side scene (generated on the GPU from `lattice`)
    boundary: np.ndarray | None   # (n_boundary, 3) float32; None for a device-side scene
    params: SceneParams
    lattice: tuple | None = None  # ((nx, ny, nz), origin): the fluid block of a device-side scene (sphk_scene_fluid_block)

    @property
    def n_fluid(self) -> int:
        if self.fluid is not None:
            return int(self.fluid.shape[0])
        (nx, ny, nz), _ = self.lattice
        return nx * ny * nz

    @property
    def n_boundary(self) -> int:
        if self.boundary is not None:
            return int(self.boundary.shape[0])
        return boundary_count(self.params.cell_size)


def boundary_count(cell_size) -> int:
    """Number of points of the six-face shell (main.cpp:89-116)."""
    cx, cy, cz = (2 * int(c) for c in cell_size)
    return 2 * cx * cy + 2 * cx * (cz - 2) + 2 * (cy - 2) * (cz - 2)


def lattice_column_planes(nx: int, ox: float, cell_length: float, n_planes: int) -> np.ndarray:
    """Cell plane (x index) of each of the nx lattice columns of the fluid block -- the host-side partition key of
    the slab driver, from nx numbers instead of a per-particle array."""
    x = F(ox) + SPACING * np.arange(nx, dtype=F)
    return np.clip((x / F(cell_length)).astype(np.int64), 0, n_planes - 1)


# name -> (box, (nx, ny, nz), origin).  SURVEY.md section 8(d); origins chosen so that no lattice plane
# sits within a few ulp of a cell face (config0 keeps the reference's own origin, main.cpp:79-81).
_CONFIGS = {
    "mini": (0.5, (10, 14, 10), (0.135, 0.055, 0.135)),
    "config0": (1.0, (24, 36, 24), (0.27, 0.10, 0.27)),
    "200k": (2.0, (60, 60, 60), (0.365, 0.105, 0.365)),
    "2m": (4.0, (128, 128, 128), (0.725, 0.105, 0.725)),
    # weak-scaling family: 2M fluid particles per GPU (x, then z, then y doubled); 16m is BASELINE configs[4]
    "4m": ((8.0, 4.0, 4.0), (256, 128, 128), (0.725, 0.105, 0.725)),
    "8m": ((8.0, 4.0, 8.0), (256, 128, 256), (0.725, 0.105, 0.725)),
    "16m": (8.0, (256, 256, 256), (1.445, 0.105, 1.445)),
    # small multi-rank test scene (x-long block so that 2-4 slabs all own fluid)
    "slabtest": ((2.0, 1.0, 1.0), (64, 24, 24), (0.285, 0.105, 0.285)),
}


def make_scene(name: str = "config0", solver: str = "wcsph", dt: float | None = None, max_iter: int = 0,
               den_thr: float = 1e-3, div_thr: float = 1e-3, jitter: float = 0.0, seed: int = 42, device_init: bool = False) -> Scene:
    """device_init: no host-side particle arrays -- the consumer generates the lattice and the boundary shell on the GPU
    (sphk_scene_fluid_block / sphk_scene_boundary_shell, bit-identical positions)."""
    box, (nx, ny, nz), origin = _CONFIGS[name]
    params = default_params(box, solver, dt, max_iter, den_thr, div_thr)
    if device_init:
        assert jitter == 0.0, "device-side scenes are the reference's deterministic lattices"
        return Scene(name=name, fluid=None, boundary=None, params=params, lattice=((nx, ny, nz), tuple(float(o) for o in origin)))
    fluid = fluid_lattice(nx, ny, nz, origin)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        fluid = (fluid + rng.uniform(-jitter, jitter, fluid.shape).astype(F)).astype(F)
    boundary = boundary_lattice(params.cell_size, params.space)
    return Scene(name=name, fluid=np.ascontiguousarray(fluid), boundary=boundary, params=params)


def benchmark_scene(name: str, solver: str, device_init: bool = False) -> Scene:
    """BASELINE.md fixed-work settings: WCSPH dt=0.001; DFSPH dt=0.004 with exactly 4+4 iterations
    (negative thresholds, Q11); PBD dt=0.004 with exactly 4 iterations (Q12)."""
    if solver == "dfsph":
        return make_scene(name, "dfsph", dt=0.004, max_iter=4, den_thr=-1.0, div_thr=-1.0, device_init=device_init)
    if solver == "pbd":
        return make_scene(name, "pbd", dt=0.004, max_iter=4, device_init=device_init)
    return make_scene(name, "wcsph", dt=0.001, device_init=device_init)


def near_face_count(pos: np.ndarray, cell_length: float, ulps: float = 8.0) -> int:
    """How many particles have a coordinate within `ulps` float32 ulps of a cell face (where the GPU's
    MUFU.RCP multiply and an IEEE divide may truncate to different cells)."""
    q = pos.astype(np.float64) / float(cell_length)
    frac = np.abs(q - np.rint(q))
    tol = ulps * np.spacing(np.maximum(np.abs(q), 1.0).astype(F)).astype(np.float64)
    return int(np.count_nonzero((frac < tol).any(axis=1)))
