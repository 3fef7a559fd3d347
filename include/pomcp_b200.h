/*
 * pomcp_b200.h — C ABI of libpomcp_b200.so (B200 / sm_100a).
 *
 * Drop-in boundary for the data-parallel hot path of JuliaPOMDP/POMCP.jl.
 * Every entry point names the reference interface it replaces (file:line is
 * relative to the POMCP.jl source tree).  All functions are extern "C", take
 * plain pointers and sizes, return an int32 status (pomcp_status) and never
 * throw.  Host buffers are caller-owned and are copied in/out before the call
 * returns; no device pointer ever crosses this boundary.
 *
 * One handle = one CUDA device + one stream + one search tree.  Calls on one
 * handle must be serialised by the caller; different handles are independent
 * (root-parallel search uses one handle per GPU / per process).
 *
 * There is no CPU fallback: every compute entry point fails with
 * POMCP_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef POMCP_B200_H
#define POMCP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POMCP_B200_ABI_VERSION 2
#define POMCP_MAX_ROCKS 16
#define POMCP_MAX_ACTIONS 32
#define POMCP_MAX_TREES 128

/* Status codes <-> reference exceptions (src/exceptions.jl:1-26,
 * src/particle_filter.jl:94-101, src/solver.jl:56-58,62). */
typedef enum pomcp_status {
  POMCP_OK = 0,
  POMCP_ERR_ALL_SAMPLES_TERMINAL = 1, /* AllSamplesTerminal, src/solver.jl:56-58 */
  POMCP_ERR_PARTICLE_DEPLETION = 2,   /* particle_depletion_error(), src/particle_filter.jl:94-101 */
  POMCP_ERR_POOL_EXHAUSTED = 3,       /* device node pool / particle log full (no reference counterpart) */
  POMCP_ERR_UNSUPPORTED = 4,          /* arbitrary Julia callbacks cannot run on device */
  POMCP_ERR_INVALID_ARG = 5,          /* incl. discount == 1 (UndefVarError in src/solver.jl:97-102) */
  POMCP_ERR_CUDA = 6,
  POMCP_ERR_NCCL = 7,
  POMCP_ERR_NAN_VALUE = 8             /* @assert !isnan(best_V), src/solver.jl:62 */
} pomcp_status;

/* Problems whose generative model is compiled into the library
 * (replaces the POMDPs.jl callbacks generate_sor / isterminal / discount /
 * actions used at src/solver.jl:82-83,126 and src/actions.jl:30-31). */
typedef enum pomcp_problem_id {
  POMCP_TIGER = 0,       /* S=u32{0 tiger-right,1 tiger-left}, A={0 listen,1 open-left,2 open-right}, O={0,1} */
  POMCP_BABY = 1,        /* S=u32 hungry, A={0 wait,1 feed}, O={0 quiet,1 crying} */
  POMCP_ROCKSAMPLE = 2,  /* S=u32 packed, A={0 N,1 E,2 S,3 W,4 sample,5+i check_i}, O={0 good,1 bad,2 none} */
  POMCP_LIGHTDARK1D = 3  /* S={int64 status,double y}, A={0:-1,1:0(stop),2:+1}, O=double */
} pomcp_problem_id;

/* RockSample packed state: bits 0-3 x, 4-7 y, 8-23 rock goodness, bit 31 terminal. */
#define POMCP_RS_TERMINAL 0x80000000u

typedef struct pomcp_lightdark_state {
  int64_t status; /* < 0 terminal (notebooks/Minimal_Example.ipynb:50-77) */
  double y;
} pomcp_lightdark_state;

typedef struct pomcp_problem {
  int32_t problem_id;
  int32_t _pad0;
  double discount;
  /* Tiger */
  double tiger_r_listen, tiger_r_findtiger, tiger_r_escapetiger, tiger_p_listen_correctly;
  /* Baby */
  double baby_r_feed, baby_r_hungry, baby_p_become_hungry, baby_p_cry_when_hungry,
      baby_p_cry_when_not_hungry;
  /* RockSample(n,k) */
  int32_t rs_n, rs_k;
  int32_t rs_rock_x[POMCP_MAX_ROCKS], rs_rock_y[POMCP_MAX_ROCKS];
  int32_t rs_start_x, rs_start_y;
  double rs_half_eff_dist, rs_r_good, rs_r_bad, rs_r_exit;
  /* LightDark1D (notebooks/Minimal_Example.ipynb:68-146) */
  double ld_correct_r, ld_incorrect_r, ld_init_mean, ld_init_std;
} pomcp_problem;

/* estimate_value kinds (src/rollout.jl:8-69). */
typedef enum pomcp_estimator {
  POMCP_EST_RANDOM_ROLLOUT = 0, /* RolloutEstimator(RandomSolver()), src/constructor.jl:14 */
  POMCP_EST_CONSTANT = 1,       /* estimate_value(n::Number,...), src/rollout.jl:10 */
  POMCP_EST_FWC_ROLLOUT = 2,    /* Baby only: RolloutEstimator(FeedWhenCrying()) with the previous-observation
                                   "belief" of src/rollout.jl:105-108 (test/runtests.jl:13-19) */
  POMCP_EST_FO_FWC_ROLLOUT = 3, /* Baby only: FORollout(FeedWhenCrying()) (src/rollout.jl:24-31,86-92, test/runtests.jl:25-31):
                                   fully observable rollout, the policy sees the state, i.e. feeds iff hungry */
  POMCP_EST_STATE_VALUE = 4     /* FOValue(policy) (src/rollout.jl:33-39,67-69): value(policy, s) read from the table
                                   given to pomcp_set_state_values (discrete-state problems) */
} pomcp_estimator;

/* node_belief_updater kinds (src/solver.jl:135-139). */
typedef enum pomcp_belief_mode {
  POMCP_BELIEF_UNWEIGHTED = 0, /* DeadReinvigorator: states pushed at visited nodes, src/solver.jl:146-148 */
  POMCP_BELIEF_EXTERNAL = 1,   /* belief maintained outside the tree (e.g. SIR filter): nothing stored */
  POMCP_BELIEF_EXACT2 = 2      /* node_belief_updater = the problem's exact discrete updater (`updater(problem)`,
                                  notebooks/Sanity_Checks.ipynb:47,70; generic branch src/solver.jl:138 and
                                  src/updater.jl:30-39): every node carries P(s = 1).  Two-state problems (Tiger, Baby). */
} pomcp_belief_mode;

typedef enum pomcp_search_mode {
  POMCP_MODE_EXACT = 0,  /* strictly sequential simulations, incremental-mean backup: bit-parity mode */
  POMCP_MODE_BATCHED = 1 /* waves of simulations in flight, atomics on the node pool: throughput mode */
} pomcp_search_mode;

/* Mirrors the POMCPSolver fields (src/POMCP.jl:137-152) with the defaults of
 * the keyword constructor (src/constructor.jl:8-18); the trailing block holds
 * the device-side knobs that have no reference counterpart. */
typedef struct pomcp_solver_params {
  double eps;           /* 0.01 */
  int64_t max_depth;    /* typemax(Int) */
  double c;             /* 1.0 */
  int64_t tree_queries; /* 100 */
  uint64_t seed;        /* replaces rng::AbstractRNG */
  double init_V;        /* 0.0 (numeric form only, src/tree.jl:34) */
  int64_t init_N;       /* 0   (numeric form only, src/tree.jl:43) */
  int32_t num_sparse_actions; /* must be <= 0: full action space, src/actions.jl:30-31 */
  int32_t estimator;          /* pomcp_estimator */
  double estimator_value;     /* for POMCP_EST_CONSTANT */
  int32_t belief_mode;        /* pomcp_belief_mode */
  int32_t search_mode;        /* pomcp_search_mode */
  /* ---- device knobs ---- */
  int64_t max_nodes;     /* belief-node pool capacity (0 = derive from tree_queries) */
  int64_t max_log;       /* particle-log capacity in entries (0 = derive) */
  int32_t inflight_min;      /* batched: simulations in flight while the tree is empty
                              * (0 = default: max(8, tree_queries / 512) of each search) */
  int32_t inflight_max;      /* batched: cap on simulations (workers) in flight (0 = default: 16 per SM on RockSample, 8 per SM elsewhere; raised to n_trees when below it: every tree keeps one worker) */
  int32_t inflight_ramp_div; /* batched: in flight = clamp(root visits / div, min, max) (0 = default) */
  int32_t rollouts_per_leaf; /* R rollouts averaged per leaf evaluation (0 = default: 1 exact, 64 batched).  Batched: 1..128, a
                              * worker is a team of ceil(R/32) warps that roll out together; exact: 1..64 on one warp */
  int32_t rank;              /* root-parallel rank, mixed into the Philox counter */
  /* ---- observation progressive widening of POMCPDPWSolver with enable_action_pw = false
   * (src/solver.jl:161-274, fields src/POMCP.jl:233-253, defaults src/constructor.jl:40-73).  At an action
   * node a new observation is generated while #children <= k_observation * N^alpha_observation
   * (src/solver.jl:223); otherwise one of the earlier generated (child, state) pairs is re-used, chosen with
   * probability proportional to the child's N (src/solver.jl:249-255).  Device path: continuous and discrete observations, with an
   * external belief updater. */
  int32_t dpw_observation;   /* 0 = POMCPSolver (default) */
  double k_observation;      /* 10.0 */
  double alpha_observation;  /* 0.5 */
  /* ---- action progressive widening (enable_action_pw = true, src/solver.jl:172-186): while
   * #children <= k_action * (sum of the children's N)^alpha_action a uniformly drawn action
   * (RandomActionGenerator, src/actions.jl:38-40) is added; a node with at most one child after that step
   * returns its leaf estimate; UCB runs over the existing children with total_N = sum of their N. */
  int32_t dpw_action;        /* 0 = every action is expanded at once (default) */
  int32_t dpw_solver;        /* 1 = simulate() of POMCPDPWSolver even where nothing widens: the leaf estimate runs
                              * `depth` rollout steps (estimate_value(..., s, h, depth), src/solver.jl:182,199) and an
                              * untried action scores its V while total_N <= 1 (:210).  Implied by dpw_observation /
                              * dpw_action. */
  double k_action;           /* 10.0 */
  double alpha_action;       /* 0.5 */
  /* ---- root parallelism inside one GPU: n_trees independent trees of tree_queries simulations each are searched
   * side by side in ONE launch from the same belief (stream id of tree t: rank * n_trees + t), the workers and the
   * in-flight cap are divided among them, and pomcp_search returns the merged decision: V_a = sum N_a V_a / sum N_a
   * over the trees, then the arg-max of src/solver.jl:60-70 - the same merge pomcp_search_rootparallel then applies
   * across GPUs.  Batched mode, external belief (0 / 1 = one tree). */
  int32_t n_trees;
  int32_t _pad3;
} pomcp_solver_params;

typedef struct pomcp_counters {
  int64_t simulations;    /* queries that ran a simulation */
  int64_t terminal_skips; /* queries consumed by a terminal root sample, src/solver.jl:49 */
  int64_t rollouts;       /* leaf evaluations */
  int64_t rollout_steps;  /* generative steps inside rollouts */
  int64_t tree_steps;     /* generative steps inside the tree */
  int64_t nodes;          /* belief nodes allocated in the pool */
  int64_t log_entries;    /* (node,state) particle-log entries */
  int64_t max_depth_seen;
  int64_t waves;          /* batched mode: simulations in flight allowed per tree in the last search (the admission cap in effect); 0 in exact mode */
  int64_t kernel_launches;/* kernels launched by this handle since creation */
} pomcp_counters;

typedef struct pomcp_handle pomcp_handle;

int32_t pomcp_abi_version(void);
const char* pomcp_status_string(int32_t status);
/* Last error text for a handle (NULL handle: last creation error). */
const char* pomcp_last_error(const pomcp_handle* h);

void pomcp_default_solver_params(pomcp_solver_params* out); /* src/constructor.jl:8-18 */
void pomcp_default_problem(int32_t problem_id, pomcp_problem* out);
int32_t pomcp_num_actions(const pomcp_problem* p);
int32_t pomcp_state_bytes(const pomcp_problem* p);

/* solve(solver, pomdp) -> planner: src/solver.jl:5-14,30-32.  Allocates the
 * node pool, path buffers and particle arrays on `device`. */
int32_t pomcp_create(const pomcp_problem* problem, const pomcp_solver_params* params, int32_t device,
                     pomcp_handle** out);
int32_t pomcp_destroy(pomcp_handle* h);

/* initialize_belief(up, b) -> RootNode(0, b, {}): src/updater.jl:45-47,
 * src/solver.jl:278-281.  Fresh root over a particle collection ... */
int32_t pomcp_set_belief_particles(pomcp_handle* h, const void* states, int64_t n);
/* ... or over the problem's initial state distribution. */
int32_t pomcp_set_belief_initial(pomcp_handle* h);
/* ... or over BoolDistribution(p1) when the planner's node_belief_updater is the exact two-state updater
 * (POMCP_BELIEF_EXACT2; p1 = P(s = 1): tiger behind the left door / baby hungry). */
int32_t pomcp_set_belief_exact2(pomcp_handle* h, double p1);
int32_t pomcp_get_belief_exact2(pomcp_handle* h, double* p1);
/* The same for the belief already on the device: a fresh RootNode over it (what every action() call on a belief
 * that is not a BeliefNode does, src/solver.jl:278-281), without sending the particles again. */
int32_t pomcp_reset_tree(pomcp_handle* h);
/* Table for POMCP_EST_STATE_VALUE: v[state_index(s)], state_index = s for Tiger/Baby (n = 2) and the packed
 * x | y<<4 | rocks<<8 for RockSample (n = 2^(8+k)); terminal states are worth 0. */
int32_t pomcp_set_state_values(pomcp_handle* h, const double* values, int64_t n);
/* Copy the root belief's particles out (ParticleCollection of the root node). */
int32_t pomcp_get_belief_particles(pomcp_handle* h, void* out, int64_t cap, int64_t* n);

/* action(planner, belief) -> search(): src/solver.jl:16-23,41-71.  Runs
 * tree_queries simulations from the current root and returns the arg-max-V
 * root action together with the root children statistics.  tree_queries < 0
 * uses the value given at creation. */
int32_t pomcp_search(pomcp_handle* h, int64_t tree_queries, int32_t* action_out, int64_t* root_N_out,
                     double* root_V_out, int32_t n_actions);

/* Replay hook for bit-parity tests: the search consumes `uniforms` in call
 * order instead of Philox draws and `rollout_returns` instead of running
 * rollouts (exact mode only). */
int32_t pomcp_search_replay(pomcp_handle* h, int64_t tree_queries, const double* uniforms, int64_t n_uniforms,
                            const double* rollout_returns, int64_t n_returns, int32_t* action_out,
                            int64_t* root_N_out, double* root_V_out, int32_t n_actions,
                            int64_t* uniforms_used, int64_t* returns_used);

/* update(::RootUpdater{<:ParticleReinvigorator}, b_old, a, o): src/updater.jl:13-27.
 * Re-roots the tree at child (a,o); POMCP_ERR_PARTICLE_DEPLETION if absent or empty.
 * obs_bits: the observation index, or the IEEE bits of a continuous observation. */
int32_t pomcp_update_unweighted(pomcp_handle* h, int32_t a, uint64_t obs_bits);
/* The same update with a device-side ParticleReinvigorator (the hooks of src/particle_filter.jl:37-73,
 * called from src/updater.jl:14-23): a child (a,o) that is missing or empty is replaced by n_min
 * particles drawn from the SIR posterior of the old root belief (handle_unseen_observation -> fresh
 * root), a child with fewer than n_min particles is topped up from it (reinvigorate!).
 * outcome: 0 plain re-root, 1 topped up, 2 rebuilt.  PARTICLE_DEPLETION only if the posterior has no mass. */
int32_t pomcp_update_reinvigorated(pomcp_handle* h, int32_t a, uint64_t obs_bits, uint64_t seed, int64_t n_min,
                                   int32_t* outcome);

/* update(::RootUpdater, b_old, a, o) with a user-supplied belief updater (src/updater.jl:30-39), here the exact
 * two-state updater of POMCP_BELIEF_EXACT2: re-roots at child (a,o); a child the tree does not hold is replaced by
 * a new node carrying update(node_belief_updater, b_old.B, a, o) - no depletion with this updater. */
int32_t pomcp_update_exact2(pomcp_handle* h, int32_t a, uint64_t obs_bits);

/* update(::SIRParticleFilter, b::ParticleCollection, a, o) (call site test/runtests.jl:57):
 * propagate, weight by the observation density, systematic resample to n_out
 * particles; the result becomes a fresh root belief (no tree reuse, src/solver.jl:278-281). */
int32_t pomcp_pf_update_sir(pomcp_handle* h, int32_t a, uint64_t obs_bits, uint64_t seed, int64_t n_out);

/* Parity hooks on the individual kernels. */
/* Systematic (low-variance) resampling of given weights.  mode 0: fixed-point
 * parallel kernels (production); mode 1: sequential fp64 chain on one thread
 * (verification of the upstream running-sum semantics). */
int32_t pomcp_resample_systematic(pomcp_handle* h, const double* weights, int64_t n, double r, int64_t n_out,
                                  int32_t mode, int64_t* idx_out);
/* Random-policy rollouts from given start states: one thread per trajectory. */
int32_t pomcp_rollout_batch(pomcp_handle* h, const void* states, int64_t n, int32_t steps, uint64_t seed,
                            double* returns_out, int32_t* steps_out);
/* Per-particle SIR weights w_i = pdf(observation(a, s'_i), o) and propagated states. */
int32_t pomcp_pf_weights(pomcp_handle* h, const void* states, int64_t n, int32_t a, uint64_t obs_bits,
                         uint64_t seed, double* weights_out, void* next_states_out);

/* Tree inspection (replaces poking at BeliefNode fields; src/tree.jl:7-26). */
int32_t pomcp_get_root_stats(pomcp_handle* h, int64_t* root_N, int64_t* act_N, double* act_V, int32_t n_actions);
/* Tree t's own result of the last search, before the merge over trees (n_trees > 1) and ranks
 * (pomcp_search_rootparallel); returns that tree's status. */
int32_t pomcp_get_tree_result(pomcp_handle* h, int32_t tree, int32_t* action_out, int64_t* root_N, int64_t* act_N,
                              double* act_V, int32_t n_actions);
/* Children statistics of the belief node reached from the root by the given
 * (action, observation) history; *node_N = -1 if the node does not exist. */
int32_t pomcp_get_node_stats(pomcp_handle* h, const int32_t* actions, const uint64_t* obs_bits, int32_t depth,
                             int64_t* node_N, int64_t* n_particles, int64_t* act_N, double* act_V,
                             int32_t n_actions);
/* Whole-tree dump in allocation order (for structural invariants). */
int32_t pomcp_dump_tree(pomcp_handle* h, int64_t cap_nodes, int64_t* n_nodes, int64_t* node_N,
                        int64_t* act_N, double* act_V, int32_t* child /* [node][action][obs], -1 none; NULL ok */);
/* The environment side of the outer simulation loop (POMDPToolbox.RolloutSimulator / test_solver, SURVEY.md
 * §3.5): generate_sor(pomdp, s, a, rng) and rand(rng, initial_state_distribution(pomdp)) on the same model
 * code the planner uses.  Draws come from Philox(seed; step), independent of the planner's streams. */
int32_t pomcp_generate_sor(pomcp_handle* h, const void* state, int32_t a, uint64_t seed, uint32_t step,
                           void* next_state, uint64_t* obs_bits, double* reward, int32_t* terminal);
int32_t pomcp_initial_state(pomcp_handle* h, uint64_t seed, void* state_out);
/* E independent episodes of the outer simulation loop entirely on the device: POMDPToolbox.RolloutSimulator.simulate
 * as test_solver (test/runtests.jl:19) and est_reward (notebooks/Sanity_Checks.ipynb:47-77, 100 episodes spread
 * over processes) run it -
 *     while disc > eps && !isterminal(s) && step <= max_steps:
 *         a = action(policy, b); (sp, o, r) = generate_sor(pomdp, s, a); total += disc * r; b = update(up, b, a, o)
 * One search tree per episode (the planner must have been created with n_trees >= n_episodes and search_mode
 * batched), all trees searched side by side in one launch per step, environment step and belief update on the
 * device, no host round trip inside an episode.  The belief updater follows the planner's belief_mode:
 * UNWEIGHTED = RootUpdater with the DeadReinvigorator (tree reuse; an episode that depletes ends with status
 * PARTICLE_DEPLETION, where the reference throws), EXACT2 = RootUpdater over the exact two-state updater
 * (src/updater.jl:30-39), EXTERNAL = SIRParticleFilter(model, sir_particles) with a fresh root every step. */
typedef struct pomcp_episode_params {
  int32_t max_steps;      /* RolloutSimulator max_steps */
  int32_t initial_state;  /* < 0: s0 ~ initial_state_distribution; >= 0: that state (two-state problems; Sanity_Checks: false) */
  double sim_eps;         /* RolloutSimulator eps: an episode runs while disc > eps (0: max_steps / terminal states only) */
  uint64_t env_seed;      /* episode e draws its environment from Philox(env_seed + e) */
  int64_t sir_particles;  /* EXTERNAL: particles per belief */
  double initial_p1;      /* EXACT2: initial belief BoolDistribution(initial_p1) */
} pomcp_episode_params;
int32_t pomcp_run_episodes(pomcp_handle* h, const pomcp_episode_params* ep, int32_t n_episodes, double* rewards_out,
                           int32_t* steps_out, int32_t* status_out);
/* Index of the current root in the arrays pomcp_dump_tree returns (0 after a fresh root or a compaction). */
int32_t pomcp_root_node(pomcp_handle* h, int64_t* node_id);
int32_t pomcp_get_counters(pomcp_handle* h, pomcp_counters* out);
int32_t pomcp_reset_counters(pomcp_handle* h);
/* Diagnostics: SM cycles summed over the batched workers since the last reset: [0] descent,
 * [1] child insertion, [2] leaf rollouts, [3] backup + particle pushes, [4] waiting for admission,
 * [5] admitted life time. */
int32_t pomcp_worker_cycles(pomcp_handle* h, uint64_t* out8);

/* Root-parallel multi-GPU search: one tree per rank, root statistics merged by
 * one NCCL all-reduce before the arg-max of src/solver.jl:60-70. */
int32_t pomcp_nccl_unique_id(void* id_out_128_bytes);
int32_t pomcp_comm_init(pomcp_handle* h, const void* nccl_unique_id_128_bytes, int32_t rank, int32_t nranks);
int32_t pomcp_search_rootparallel(pomcp_handle* h, int64_t tree_queries, int32_t* action_out,
                                  int64_t* root_N_out, double* root_V_out, int32_t n_actions);

/* Device-timer access for the bench harness: elapsed milliseconds spent in the
 * named phase by the last call (CUDA events on the handle's stream). */
typedef enum pomcp_phase { POMCP_PHASE_SEARCH = 0, POMCP_PHASE_ROLLOUT = 1, POMCP_PHASE_PF = 2,
                           POMCP_PHASE_REROOT = 3,
                           POMCP_PHASE_MERGE = 4, /* ncclAllReduce + finish of pomcp_search_rootparallel */
                           POMCP_PHASE_COUNT = 5 } pomcp_phase;
int32_t pomcp_last_phase_ms(pomcp_handle* h, int32_t phase, double* ms_out, int64_t* launches_out);
/* Device-resident variants used by the bench's kernel-only timing: run on
 * whatever belief is already in HBM and leave results there. */
/* Generic CUDA-event stopwatch on the handle's stream (8 slots). */
int32_t pomcp_event_record(pomcp_handle* h, int32_t slot);
int32_t pomcp_event_elapsed_ms(pomcp_handle* h, int32_t slot_start, int32_t slot_stop, double* ms_out);
int32_t pomcp_search_async(pomcp_handle* h, int64_t tree_queries);
int32_t pomcp_sync(pomcp_handle* h);
int32_t pomcp_flush_l2(pomcp_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* POMCP_B200_H */
