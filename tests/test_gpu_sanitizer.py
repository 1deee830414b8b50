"""compute-sanitizer runs of the hot path (SURVEY section 5 asks for memcheck / racecheck; the reference has neither).
Opt-in -- SPHK_SANITIZER=1 -- because a sanitizer run takes a minute per tool: the C++ headless application (the class
layer over the C-ABI, no python in the sanitized process) steps the mini dam-break with each solver under memcheck,
and the shared-memory paths (the staged list builder, and the tile sweeps with SPHK_TILE=1) under racecheck too."""
import os
import shutil
import subprocess

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "cpp-fluid-particles_b200", "sph_headless")
MINI = ["--frames", "2", "--box", "0.5", "--block", "10", "14", "10", "--origin", "0.135", "0.055", "0.135", "--quiet"]


def _sanitize(tool, solver, extra_env=None):
    if os.environ.get("SPHK_SANITIZER") != "1":
        pytest.skip("opt-in: SPHK_SANITIZER=1")
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    cmd = [exe, "--tool", tool, "--error-exitcode", "9", CLI, "--solver", solver] + MINI
    if solver != "sph":
        cmd += ["--iters", "2"]
    env = dict(os.environ, SPHK_STEP_GRAPH="0", **(extra_env or {}))     # plain launches: the sanitizer sees every kernel
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-4000:]
    clean = "ERROR SUMMARY: 0 errors" in r.stdout or "RACECHECK SUMMARY: 0 hazards displayed (0 errors, 0 warnings)" in r.stdout
    assert clean, r.stdout[-4000:]


@pytest.mark.parametrize("solver", ["sph", "dfsph", "pbd"])
def test_memcheck(built, solver):
    _sanitize("memcheck", solver)


@pytest.mark.parametrize("solver", ["dfsph", "pbd"])
def test_memcheck_tile_lists(built, solver):
    _sanitize("memcheck", solver, {"SPHK_TILE": "1"})


def test_racecheck_staged_builder(built):
    _sanitize("racecheck", "dfsph")


def test_racecheck_tile_lists(built):
    _sanitize("racecheck", "dfsph", {"SPHK_TILE": "1"})
