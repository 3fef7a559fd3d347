"""ctypes mirror of include/pomcp_b200.h (POD descriptors, status codes).

Layout must match the header byte for byte; tests/test_abi_layout.py checks the
sizes against the values the C side reports.
"""
import ctypes as C

ABI_VERSION = 2
MAX_ROCKS = 16

# pomcp_status
OK = 0
ERR_ALL_SAMPLES_TERMINAL = 1
ERR_PARTICLE_DEPLETION = 2
ERR_POOL_EXHAUSTED = 3
ERR_UNSUPPORTED = 4
ERR_INVALID_ARG = 5
ERR_CUDA = 6
ERR_NCCL = 7
ERR_NAN_VALUE = 8

# pomcp_problem_id
TIGER, BABY, ROCKSAMPLE, LIGHTDARK1D = 0, 1, 2, 3
RS_TERMINAL = 0x80000000

EST_RANDOM_ROLLOUT, EST_CONSTANT, EST_FWC_ROLLOUT, EST_FO_FWC_ROLLOUT, EST_STATE_VALUE = 0, 1, 2, 3, 4
BELIEF_UNWEIGHTED, BELIEF_EXTERNAL, BELIEF_EXACT2 = 0, 1, 2
MODE_EXACT, MODE_BATCHED = 0, 1
PHASE_SEARCH, PHASE_ROLLOUT, PHASE_PF, PHASE_REROOT, PHASE_MERGE = 0, 1, 2, 3, 4


class LightDarkState(C.Structure):
    _fields_ = [("status", C.c_int64), ("y", C.c_double)]


class Problem(C.Structure):
    _fields_ = [
        ("problem_id", C.c_int32), ("_pad0", C.c_int32),
        ("discount", C.c_double),
        ("tiger_r_listen", C.c_double), ("tiger_r_findtiger", C.c_double),
        ("tiger_r_escapetiger", C.c_double), ("tiger_p_listen_correctly", C.c_double),
        ("baby_r_feed", C.c_double), ("baby_r_hungry", C.c_double), ("baby_p_become_hungry", C.c_double),
        ("baby_p_cry_when_hungry", C.c_double), ("baby_p_cry_when_not_hungry", C.c_double),
        ("rs_n", C.c_int32), ("rs_k", C.c_int32),
        ("rs_rock_x", C.c_int32 * MAX_ROCKS), ("rs_rock_y", C.c_int32 * MAX_ROCKS),
        ("rs_start_x", C.c_int32), ("rs_start_y", C.c_int32),
        ("rs_half_eff_dist", C.c_double), ("rs_r_good", C.c_double), ("rs_r_bad", C.c_double),
        ("rs_r_exit", C.c_double),
        ("ld_correct_r", C.c_double), ("ld_incorrect_r", C.c_double), ("ld_init_mean", C.c_double),
        ("ld_init_std", C.c_double),
    ]


class SolverParams(C.Structure):
    _fields_ = [
        ("eps", C.c_double), ("max_depth", C.c_int64), ("c", C.c_double), ("tree_queries", C.c_int64),
        ("seed", C.c_uint64), ("init_V", C.c_double), ("init_N", C.c_int64),
        ("num_sparse_actions", C.c_int32), ("estimator", C.c_int32), ("estimator_value", C.c_double),
        ("belief_mode", C.c_int32), ("search_mode", C.c_int32),
        ("max_nodes", C.c_int64), ("max_log", C.c_int64),
        ("inflight_min", C.c_int32), ("inflight_max", C.c_int32), ("inflight_ramp_div", C.c_int32),
        ("rollouts_per_leaf", C.c_int32), ("rank", C.c_int32), ("dpw_observation", C.c_int32),
        ("k_observation", C.c_double), ("alpha_observation", C.c_double),
        ("dpw_action", C.c_int32), ("dpw_solver", C.c_int32), ("k_action", C.c_double), ("alpha_action", C.c_double),
        ("n_trees", C.c_int32), ("_pad3", C.c_int32),
    ]


class EpisodeParams(C.Structure):
    _fields_ = [("max_steps", C.c_int32), ("initial_state", C.c_int32), ("sim_eps", C.c_double),
                ("env_seed", C.c_uint64), ("sir_particles", C.c_int64), ("initial_p1", C.c_double)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "simulations", "terminal_skips", "rollouts", "rollout_steps", "tree_steps", "nodes", "log_entries",
        "max_depth_seen", "waves", "kernel_launches")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


STATUS_NAMES = {
    OK: "OK", ERR_ALL_SAMPLES_TERMINAL: "ALL_SAMPLES_TERMINAL", ERR_PARTICLE_DEPLETION: "PARTICLE_DEPLETION",
    ERR_POOL_EXHAUSTED: "POOL_EXHAUSTED", ERR_UNSUPPORTED: "UNSUPPORTED", ERR_INVALID_ARG: "INVALID_ARG",
    ERR_CUDA: "CUDA_ERROR", ERR_NCCL: "NCCL_ERROR", ERR_NAN_VALUE: "NAN_VALUE",
}

# Every symbol include/pomcp_b200.h declares (checked by the not-gpu tests).
EXPORTS = [
    "pomcp_abi_version", "pomcp_status_string", "pomcp_last_error", "pomcp_default_solver_params",
    "pomcp_default_problem", "pomcp_num_actions", "pomcp_state_bytes", "pomcp_create", "pomcp_destroy",
    "pomcp_set_belief_particles", "pomcp_set_belief_initial", "pomcp_set_belief_exact2", "pomcp_get_belief_exact2", "pomcp_update_exact2", "pomcp_reset_tree", "pomcp_set_state_values", "pomcp_get_belief_particles", "pomcp_search",
    "pomcp_search_replay", "pomcp_update_unweighted", "pomcp_update_reinvigorated", "pomcp_pf_update_sir", "pomcp_resample_systematic",
    "pomcp_rollout_batch", "pomcp_pf_weights", "pomcp_get_root_stats", "pomcp_get_tree_result", "pomcp_get_node_stats", "pomcp_dump_tree",
    "pomcp_generate_sor", "pomcp_initial_state", "pomcp_run_episodes", "pomcp_root_node", "pomcp_get_counters", "pomcp_reset_counters", "pomcp_worker_cycles", "pomcp_nccl_unique_id", "pomcp_comm_init",
    "pomcp_search_rootparallel", "pomcp_last_phase_ms", "pomcp_event_record", "pomcp_event_elapsed_ms", "pomcp_search_async", "pomcp_sync", "pomcp_flush_l2",
]
