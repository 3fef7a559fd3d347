"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/pomcp_b200.h
declares, the ctypes mirror matches the C struct layout, the product never routes through the
oracle, and — with no device — every compute entry point fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pomcp_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pomcp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(gpu_lib):
    declared = _declared_symbols()
    assert len(declared) >= 30
    assert sorted(abi.EXPORTS) == declared
    for name in declared:
        assert hasattr(gpu_lib, name), name
    assert gpu_lib.pomcp_abi_version() == abi.ABI_VERSION
    assert gpu_lib.pomcp_status_string(abi.ERR_PARTICLE_DEPLETION) == b"PARTICLE_DEPLETION"


def test_struct_layout_matches_header(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pomcp_b200.h"\n'
                    'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(pomcp_problem), '
                    'sizeof(pomcp_solver_params), sizeof(pomcp_counters), sizeof(pomcp_lightdark_state), '
                    'offsetof(pomcp_problem, rs_half_eff_dist), offsetof(pomcp_solver_params, max_nodes), '
                    'offsetof(pomcp_problem, ld_correct_r), offsetof(pomcp_solver_params, n_trees), '
                    'offsetof(pomcp_solver_params, dpw_solver), sizeof(pomcp_episode_params));return 0;}\n')
    exe = tmp_path / "sz"
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(abi.Problem), C.sizeof(abi.SolverParams), C.sizeof(abi.Counters), C.sizeof(abi.LightDarkState),
            abi.Problem.rs_half_eff_dist.offset, abi.SolverParams.max_nodes.offset, abi.Problem.ld_correct_r.offset,
            abi.SolverParams.n_trees.offset, abi.SolverParams.dpw_solver.offset, C.sizeof(abi.EpisodeParams)]
    assert got == want


def test_defaults_match_reference_constructor(gpu_lib):
    """src/constructor.jl:8-18."""
    sp = abi.SolverParams()
    gpu_lib.pomcp_default_solver_params(C.byref(sp))
    assert (sp.eps, sp.c, sp.tree_queries, sp.init_V, sp.init_N, sp.num_sparse_actions) == (0.01, 1.0, 100, 0.0, 0, 0)
    assert sp.max_depth == 2 ** 63 - 1
    s = pj.POMCPSolver()
    assert (s.eps, s.c, s.tree_queries, s.init_V, s.init_N, s.num_sparse_actions) == (0.01, 1.0, 100, 0.0, 0, 0)
    assert isinstance(s.default_action, pj.ExceptionRethrow) and isinstance(s.estimate_value, pj.RolloutEstimator)
    assert isinstance(s.node_belief_updater, pj.DefaultReinvigoratorStub)
    for pomdp, nA, sb in ((pj.TigerPOMDP(), 3, 4), (pj.BabyPOMDP(), 2, 4), (pj.RockSamplePOMDP(7, 8), 13, 4),
                          (pj.RockSamplePOMDP(11, 11), 16, 4), (pj.LightDark1D(), 3, 16)):
        p = pomdp.c_struct()
        assert gpu_lib.pomcp_num_actions(C.byref(p)) == nA == pomdp.n_actions
        assert gpu_lib.pomcp_state_bytes(C.byref(p)) == sb == pomdp.state_dtype.itemsize


def test_rocksample_layout_is_frozen():
    """rock positions from Philox(seed) without replacement, start = (0, n//2)."""
    rs = pj.RockSamplePOMDP(7, 8, seed=7008)
    assert rs.rocks == [(4, 2), (5, 6), (1, 2), (2, 0), (1, 4), (0, 3), (4, 3), (0, 1)] and rs.start == (0, 3)
    rs = pj.RockSamplePOMDP(11, 11, seed=11011)
    assert len(set(rs.rocks)) == 11 and rs.n_actions == 16 and rs.start == (0, 5)


def test_unsupported_estimators_are_rejected():
    with pytest.raises(NotImplementedError):
        pj.POMCPSolver(estimate_value=lambda *a: 0.0).c_params(abi.BELIEF_UNWEIGHTED)
    with pytest.raises(NotImplementedError):
        pj.POMCPSolver(init_V=lambda *a: 0.0).c_params(abi.BELIEF_UNWEIGHTED)
    assert pj.POMCPSolver(estimate_value=3.5).c_params(0).estimator == abi.EST_CONSTANT
    assert pj.POMCPSolver(estimate_value=pj.FORollout(pj.RandomSolver())).c_params(0).estimator == abi.EST_RANDOM_ROLLOUT
    assert pj.POMCPSolver(estimate_value=pj.FORollout(pj.FeedWhenCrying())).c_params(0).estimator == abi.EST_FO_FWC_ROLLOUT
    assert pj.POMCPSolver(estimate_value=pj.FOValue([0.0, 1.0])).c_params(0).estimator == abi.EST_STATE_VALUE


def test_default_action_dispatch():
    """src/exceptions.jl:23-26."""
    ex = pj.AllSamplesTerminal("b")
    with pytest.raises(pj.AllSamplesTerminal):
        pj.default_action(pj.ExceptionRethrow(), None, ex)
    assert pj.default_action(lambda b, e: 7, None, ex) == 7
    assert pj.default_action(3, None, ex) == 3
    assert isinstance(ex, pj.NoDecision) and "all states sampled from the belief were terminal" in str(ex)
    assert "Particle Depletion!" in str(pj.ParticleDepletion())


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "pomcp.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                bad = re.findall(r"import\s+oracle|from\s+oracle|pomcp_oracle|\borc_\w+|oracle/|libpomcp_oracle", txt)
                assert not bad, (os.path.join(dirpath, f), bad)


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="GPU present: the no-device contract is not observable")
def test_no_device_fails_loudly_no_cpu_fallback(gpu_lib):
    with pytest.raises(pj.POMCPError) as ei:
        pj.solve(pj.POMCPSolver(tree_queries=10), pj.BabyPOMDP())
    assert ei.value.status == abi.ERR_CUDA
    assert b"no CPU fallback" in gpu_lib.pomcp_last_error(None)


def test_missing_library_fails_loudly(tmp_path):
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import pomcp_jl_b200._lib as L\n"
            "L.SO_PATH = %r\n"
            "try:\n    L.lib()\nexcept L.LibraryMissing as e:\n    print('MISSING', 'no CPU fallback' in str(e).replace('There is no CPU fallback', 'no CPU fallback'))\n"
            % (ROOT, str(tmp_path / "nope.so")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "MISSING True" in out.stdout, out.stdout + out.stderr


def _build_c_client(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-std=c11", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "abi_smoke.c"), "-o", exe, "-L" + os.path.join(ROOT, "pomcp.jl_b200"),
           "-lpomcp_b200", "-Wl,-rpath," + os.path.join(ROOT, "pomcp.jl_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_a_c_client_links(gpu_lib, tmp_path):
    """include/pomcp_b200.h compiles as C11 and examples/abi_smoke.c (no Python, no torch) links against the
    shared library: the boundary is a C ABI."""
    exe = _build_c_client(tmp_path)
    if not os.path.exists("/dev/nvidiactl"):
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 1 and "no CUDA device (there is no CPU fallback)" in r.stderr


@pytest.mark.gpu
def test_c_client_runs_a_decision_and_an_update(gpu_lib, tmp_path):
    exe = _build_c_client(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi smoke ok" in r.stdout
