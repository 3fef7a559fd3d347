/* Plain-C client of the drop-in boundary (include/pomcp_b200.h): what a cgo/ccall/JNI stub binds.
 * One decision + one belief update on RockSample(7,8), no Python, no torch.
 *
 *   gcc -O2 -Iinclude examples/abi_smoke.c -o /tmp/abi_smoke -Lpomcp.jl_b200 -lpomcp_b200 -Wl,-rpath,$PWD/pomcp.jl_b200
 *   /tmp/abi_smoke
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pomcp_b200.h"

#define CHECK(call)                                                                                   \
  do {                                                                                                \
    int32_t rc__ = (call);                                                                            \
    if (rc__ != POMCP_OK) {                                                                           \
      fprintf(stderr, "%s -> %s: %s\n", #call, pomcp_status_string(rc__), pomcp_last_error(h));       \
      return 1;                                                                                       \
    }                                                                                                 \
  } while (0)

int main(void) {
  pomcp_handle* h = NULL;
  pomcp_problem prob;
  pomcp_solver_params sp;
  pomcp_default_problem(POMCP_ROCKSAMPLE, &prob);
  prob.rs_n = 7;
  prob.rs_k = 8;
  const int rx[8] = {4, 5, 1, 2, 1, 0, 4, 0}, ry[8] = {2, 6, 2, 0, 4, 3, 3, 1};
  for (int i = 0; i < 8; ++i) { prob.rs_rock_x[i] = rx[i]; prob.rs_rock_y[i] = ry[i]; }
  pomcp_default_solver_params(&sp); /* src/constructor.jl:8-18 */
  sp.c = 20.0;
  sp.tree_queries = 200000;
  sp.seed = 42;
  sp.belief_mode = POMCP_BELIEF_UNWEIGHTED;
  int32_t rc = pomcp_create(&prob, &sp, 0, &h);
  if (rc != POMCP_OK) {
    fprintf(stderr, "pomcp_create -> %s: %s\n", pomcp_status_string(rc), pomcp_last_error(NULL));
    return 1;
  }
  const int nA = pomcp_num_actions(&prob);
  int32_t a = -1;
  int64_t N[POMCP_MAX_ACTIONS];
  double V[POMCP_MAX_ACTIONS];
  CHECK(pomcp_set_belief_initial(h));                      /* initialize_belief(up, initial_state_distribution(p)) */
  CHECK(pomcp_search(h, -1, &a, N, V, nA));                /* action(policy, b) */
  int64_t total = 0;
  for (int i = 0; i < nA; ++i) total += N[i];
  printf("action %d, root visits over actions %lld (tree_queries - 1 = %lld)\n", a, (long long)total,
         (long long)(sp.tree_queries - 1));
  if (total != sp.tree_queries - 1) return 2;
  /* environment step on the same model, then update(up, b, a, o) with the device-side reinvigorator */
  uint32_t s = 0, s2 = 0;
  CHECK(pomcp_initial_state(h, 7, &s));
  uint64_t o = 0;
  double r = 0.0;
  int32_t term = 0, outcome = -1;
  CHECK(pomcp_generate_sor(h, &s, a, 7, 1, &s2, &o, &r, &term));
  CHECK(pomcp_update_reinvigorated(h, a, o, 99, 1000, &outcome));
  int64_t nb = 0;
  CHECK(pomcp_get_belief_particles(h, NULL, 0, &nb));
  printf("obs %llu reward %.1f -> new belief of %lld particles (outcome %d)\n", (unsigned long long)o, r, (long long)nb,
         outcome);
  pomcp_counters c;
  CHECK(pomcp_get_counters(h, &c));
  printf("rollout steps %lld, nodes %lld, kernel launches %lld\n", (long long)c.rollout_steps, (long long)c.nodes,
         (long long)c.kernel_launches);
  CHECK(pomcp_destroy(h));
  printf("abi smoke ok\n");
  return 0;
}
