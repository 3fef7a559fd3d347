"""Loader for libpomcp_b200.so (the CUDA library behind include/pomcp_b200.h).

There is no CPU fallback: if the shared library is missing or was not built,
importing symbols raises immediately.
"""
import ctypes as C
import os

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# POMCP_B200_LIB: developer override (A/B timing of two builds); the default is the in-tree library
SO_PATH = os.environ.get("POMCP_B200_LIB") or os.path.join(_HERE, "libpomcp_b200.so")
_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise LibraryMissing(
            f"{SO_PATH} not found: build it with `python -m pomcp_jl_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback.")
    L = C.CDLL(SO_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    P = C.POINTER
    L.pomcp_abi_version.restype = i32
    L.pomcp_status_string.restype = C.c_char_p
    L.pomcp_status_string.argtypes = [i32]
    L.pomcp_last_error.restype = C.c_char_p
    L.pomcp_last_error.argtypes = [vp]
    L.pomcp_default_solver_params.argtypes = [P(abi.SolverParams)]
    L.pomcp_default_solver_params.restype = None
    L.pomcp_default_problem.argtypes = [i32, P(abi.Problem)]
    L.pomcp_default_problem.restype = None
    L.pomcp_num_actions.argtypes = [P(abi.Problem)]
    L.pomcp_state_bytes.argtypes = [P(abi.Problem)]
    L.pomcp_create.argtypes = [P(abi.Problem), P(abi.SolverParams), i32, P(vp)]
    L.pomcp_destroy.argtypes = [vp]
    L.pomcp_set_belief_particles.argtypes = [vp, vp, i64]
    L.pomcp_set_belief_initial.argtypes = [vp]
    L.pomcp_reset_tree.argtypes = [vp]
    L.pomcp_set_belief_exact2.argtypes = [vp, dbl]
    L.pomcp_get_belief_exact2.argtypes = [vp, P(dbl)]
    L.pomcp_update_exact2.argtypes = [vp, i32, u64]
    L.pomcp_set_state_values.argtypes = [vp, vp, i64]
    L.pomcp_get_belief_particles.argtypes = [vp, vp, i64, P(i64)]
    L.pomcp_search.argtypes = [vp, i64, P(i32), vp, vp, i32]
    L.pomcp_search_replay.argtypes = [vp, i64, vp, i64, vp, i64, P(i32), vp, vp, i32, P(i64), P(i64)]
    L.pomcp_update_unweighted.argtypes = [vp, i32, u64]
    L.pomcp_update_reinvigorated.argtypes = [vp, i32, u64, u64, i64, P(i32)]
    L.pomcp_pf_update_sir.argtypes = [vp, i32, u64, u64, i64]
    L.pomcp_resample_systematic.argtypes = [vp, vp, i64, dbl, i64, i32, vp]
    L.pomcp_rollout_batch.argtypes = [vp, vp, i64, i32, u64, vp, vp]
    L.pomcp_pf_weights.argtypes = [vp, vp, i64, i32, u64, u64, vp, vp]
    L.pomcp_get_root_stats.argtypes = [vp, P(i64), vp, vp, i32]
    L.pomcp_get_tree_result.argtypes = [vp, i32, P(i32), P(i64), vp, vp, i32]
    L.pomcp_get_node_stats.argtypes = [vp, vp, vp, i32, P(i64), P(i64), vp, vp, i32]
    L.pomcp_dump_tree.argtypes = [vp, i64, P(i64), vp, vp, vp, vp]
    L.pomcp_get_counters.argtypes = [vp, P(abi.Counters)]
    L.pomcp_generate_sor.argtypes = [vp, vp, i32, u64, C.c_uint32, vp, P(u64), P(dbl), P(i32)]
    L.pomcp_initial_state.argtypes = [vp, u64, vp]
    L.pomcp_run_episodes.argtypes = [vp, P(abi.EpisodeParams), i32, vp, vp, vp]
    L.pomcp_root_node.argtypes = [vp, P(i64)]
    L.pomcp_reset_counters.argtypes = [vp]
    L.pomcp_worker_cycles.argtypes = [vp, vp]
    L.pomcp_nccl_unique_id.argtypes = [vp]
    L.pomcp_comm_init.argtypes = [vp, vp, i32, i32]
    L.pomcp_search_rootparallel.argtypes = [vp, i64, P(i32), vp, vp, i32]
    L.pomcp_last_phase_ms.argtypes = [vp, i32, P(dbl), P(i64)]
    L.pomcp_event_record.argtypes = [vp, i32]
    L.pomcp_event_elapsed_ms.argtypes = [vp, i32, i32, P(dbl)]
    L.pomcp_search_async.argtypes = [vp, i64]
    L.pomcp_sync.argtypes = [vp]
    L.pomcp_flush_l2.argtypes = [vp]
    if L.pomcp_abi_version() != abi.ABI_VERSION:
        raise RuntimeError("libpomcp_b200.so ABI version mismatch")
    _lib = L
    return L
