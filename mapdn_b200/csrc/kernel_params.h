// Kernel-side view of one mapdn_env handle: pointers into HBM + scalars, passed by value.
#pragma once
#include <stdint.h>

namespace mapdn {

// doubles of per-env shared memory per node (see DESIGN.md "shared-memory layout")
constexpr int kNodeArrays = 22;

enum NodeArr {            // per-env, per-node arrays (index * n_pad)
  A_VM = 0, A_VA, A_E, A_F, A_PS, A_QS, A_AUP, A_BUP, A_ADN, A_BDN,
  A_D0, A_D1, A_D2, A_D3, A_R0, A_R1, A_S0, A_S1, A_S2, A_S3, A_T0, A_T1
};

struct HotLayout {        // byte offsets inside the hot static blob (staged into smem per CTA)
  int gu, bu, gd, bd, gii, bii;        // double [n]: Y[i,parent], Y[parent,i], Y[i,i]
  int parent, cstart, eorder, elev, dlev;  // uint16: [n], [n+1], [n-1], [n_elev+1], [n_dlev+1]
  int bytes;                           // total, multiple of 16
};

struct Params {
  // ---- sizes ----
  int n, n_pad, n_load, n_sgen, n_sgen_pad, n_line, n_elev, n_dlev, obs_dim, state_dim;
  int nb;                 // envs processed by this launch
  int env_stride;         // doubles of smem per env
  HotLayout hot_layout;
  const unsigned char* hot;
  // ---- cold static (global, read through the read-only path) ----
  const int* bus_of_node; const int* node_of_bus;
  const int* lptr; const int* lidx; const double* lscale;     // node -> loads (CSR)
  const int* sptr; const int* sidx; const double* sscale;     // node -> sgens (CSR)
  const int* line_f; const int* line_t; const double* line_c; // lines: nodes, 4 loss coefficients
  const int* zptr; const int* znode;                          // agent -> zone bus slots (node ids)
  const int* zsg_ptr; const int* zsg_idx;                     // zone slot -> sgens sitting on it
  const double* s_max; const double* pv_std; const double* lp_std; const double* lq_std;
  // ---- profile store ----
  const double* prof_pv; const double* prof_lp; const double* prof_lq;
  long long n_rows; int steps_per_hour; int n_day_choices;
  // ---- scalars ----
  double base_mva, inv_base, vm_init, e0, f0, vm0, va0, tol;
  int max_iter;
  int barrier; double voltage_weight, q_weight, line_weight; int use_line_weight;
  double v_upper, v_lower; int episode_limit; double action_low, action_high; int reset_action;
  unsigned long long seed; long long env_id_offset;
  // ---- env state (global, [B, *]) ----
  double* cur_pl; double* cur_ql; double* cur_pv; double* cur_q;
  double* res_vm; double* res_va; double* res_p; double* res_q; double* res_pl;
  int* steps; double* sum_rewards; long long* start_row; unsigned* episode;
  // ---- launch io ----
  const double* in_pl; const double* in_ql; const double* in_pv; const double* in_q;  // SOLVE
  const double* actions;                                                             // STEP
  const int* start_dhi; const unsigned char* mask; int add_noise;                    // RESET
  double* out_vm; double* out_va; double* out_p; double* out_q; double* out_pl;
  int* out_iters; unsigned char* out_conv;
  double* reward; unsigned char* term; double* info; double* obs; double* state;
};

}  // namespace mapdn
