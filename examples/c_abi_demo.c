/*
 * c_abi_demo.c - a plain C99 caller of include/mapdn_b200.h: no CUDA headers, no Python, host buffers only.
 * This is the binding any FFI (cgo, JNI, ctypes ...) would reproduce: describe the network and the profile
 * store once, then reset / step whole batches of envs (reference call pattern code_examples.py:34-58).
 *
 *   gcc -std=c99 -O2 -Iinclude examples/c_abi_demo.c -o c_abi_demo -Lmapdn_b200 -lmapdn_b200 -lm
 *
 * Prints one line per step: "step k reward_sum obs_sum terminated" (tests/test_c_abi_demo.py replays the same
 * feeder through the Python binding and compares the numbers).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "mapdn_b200.h"

#define N_BUS 6
#define N_BRANCH 5
#define N_LOAD 5
#define N_SGEN 2
#define ROWS_PER_DAY 480
#define N_ROWS (5 * ROWS_PER_DAY)
#define BATCH 8
#define N_STEPS 10

static const double kPi = 3.14159265358979323846;

static int die(const char* what, mapdn_status st) {
  fprintf(stderr, "%s failed (status %d): %s\n", what, (int)st, mapdn_last_error());
  return 2;
}

int main(void) {
  /* a 6-bus feeder: 0 (slack) - 1 - 2 - 3 with a lateral 2 - 4 - 5; PV at buses 3 and 5 */
  static const int32_t br_from[N_BRANCH] = {0, 1, 2, 2, 4}, br_to[N_BRANCH] = {1, 2, 3, 4, 5};
  static const int32_t bus_zone[N_BUS] = {0, 0, 1, 1, 2, 2};
  static const int32_t load_bus[N_LOAD] = {1, 2, 3, 4, 5};
  static const int32_t sgen_bus[N_SGEN] = {3, 5}, sgen_zone[N_SGEN] = {1, 2};
  double br_r[N_BRANCH], br_x[N_BRANCH];
  for (int k = 0; k < N_BRANCH; ++k) { br_r[k] = 0.01 * (k + 1); br_x[k] = 0.02 * (k + 1); }

  mapdn_net_desc net = {0};
  net.n_bus = N_BUS; net.n_branch = N_BRANCH; net.n_load = N_LOAD; net.n_sgen = N_SGEN;
  net.base_mva = 10.0; net.slack_bus = 0; net.slack_vm = 1.0; net.slack_va_deg = 0.0; net.vm_init = 1.0;
  net.br_from = br_from; net.br_to = br_to; net.br_r = br_r; net.br_x = br_x;
  net.bus_zone = bus_zone; net.load_bus = load_bus; net.sgen_bus = sgen_bus; net.sgen_zone = sgen_zone;

  /* five days of 3-minute rows (the three CSVs of the reference, already scaled) */
  double* pv = malloc(sizeof(double) * N_ROWS * N_SGEN);
  double* lp = malloc(sizeof(double) * N_ROWS * N_LOAD);
  double* lq = malloc(sizeof(double) * N_ROWS * N_LOAD);
  if (!pv || !lp || !lq) return 1;
  for (int t = 0; t < N_ROWS; ++t) {
    const double sun = sin(kPi * (t % ROWS_PER_DAY) / ROWS_PER_DAY);
    for (int g = 0; g < N_SGEN; ++g) pv[t * N_SGEN + g] = 0.3 * (sun > 0.0 ? sun : 0.0) * (1.0 + 0.1 * g);
    for (int l = 0; l < N_LOAD; ++l) {
      lp[t * N_LOAD + l] = 0.2 + 0.05 * l + 0.05 * cos(2.0 * kPi * t / ROWS_PER_DAY);
      lq[t * N_LOAD + l] = 0.3 * lp[t * N_LOAD + l];
    }
  }
  double pv_std[N_SGEN], s_max[N_SGEN], lp_std[N_LOAD], lq_std[N_LOAD];
  for (int g = 0; g < N_SGEN; ++g) { pv_std[g] = 0.001; s_max[g] = 1.2 * 0.3 * (1.0 + 0.1 * g); }
  for (int l = 0; l < N_LOAD; ++l) { lp_std[l] = 0.002; lq_std[l] = 0.0005; }
  mapdn_profile_desc prof = {0};
  prof.n_rows = N_ROWS; prof.steps_per_hour = 20; prof.n_days = 4;
  prof.pv = pv; prof.load_p = lp; prof.load_q = lq;
  prof.pv_std = pv_std; prof.load_p_std = lp_std; prof.load_q_std = lq_std; prof.s_max = s_max;

  mapdn_cfg cfg = {0};
  cfg.batch = BATCH; cfg.barrier = MAPDN_BARRIER_BOWL; cfg.voltage_weight = 1.0; cfg.q_weight = 0.1;
  cfg.v_upper = 1.05; cfg.v_lower = 0.95; cfg.episode_limit = 240;
  cfg.action_low = -0.8; cfg.action_high = 0.8; cfg.reset_action = 1; cfg.seed = 2024;

  if (mapdn_abi_version() != MAPDN_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }
  mapdn_env* env = NULL;
  mapdn_status st = mapdn_create(&net, &prof, &cfg, /*device=*/0, &env);
  if (st != MAPDN_OK) return die("mapdn_create", st);
  mapdn_dims d;
  if ((st = mapdn_get_dims(env, &d)) != MAPDN_OK) return die("mapdn_get_dims", st);
  printf("dims n_agents %d obs_dim %d state_dim %d lanes_per_env %d\n", d.n_agents, d.obs_dim, d.state_dim,
         d.lanes_per_env);

  const size_t n_obs = (size_t)BATCH * d.n_agents * d.obs_dim;
  double* actions = malloc(sizeof(double) * BATCH * N_SGEN);
  double* reward = malloc(sizeof(double) * BATCH);
  double* info = malloc(sizeof(double) * BATCH * MAPDN_N_INFO);
  double* obs = malloc(sizeof(double) * n_obs);
  uint8_t* term = malloc(BATCH);
  if (!actions || !reward || !info || !obs || !term) return 1;

  if ((st = mapdn_reset(env, NULL, NULL, /*add_noise=*/1, NULL, NULL, /*converged_dev=*/NULL, NULL)) != MAPDN_OK) return die("mapdn_reset", st);
  for (int k = 0; k < N_STEPS; ++k) {
    for (int e = 0; e < BATCH; ++e)
      for (int g = 0; g < N_SGEN; ++g) actions[e * N_SGEN + g] = -0.8 + 1.6 * ((e * 7 + g * 3 + k) % 11) / 10.0;
    if ((st = mapdn_step_host(env, actions, 1, reward, term, info, obs, NULL)) != MAPDN_OK) return die("mapdn_step_host", st);
    double rs = 0.0, os = 0.0;
    int nt = 0;
    for (int e = 0; e < BATCH; ++e) { rs += reward[e]; nt += term[e]; }
    for (size_t i = 0; i < n_obs; ++i) os += obs[i];
    printf("step %d %.17g %.17g %d\n", k, rs, os, nt);
  }
  printf("launches %lld\n", (long long)mapdn_launch_count(env));
  mapdn_destroy(env);
  free(pv); free(lp); free(lq); free(actions); free(reward); free(info); free(obs); free(term);
  return 0;
}
