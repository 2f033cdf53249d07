// Fused batched environment kernel for sm_100a: action clip -> bus injections -> polar
// Newton-Raphson on the (radial) Ybus with a zero-fill 2x2-block elimination -> reward / info ->
// next profile row (+ noise) -> zone-masked observation gather. One launch per env step.
//
// Work decomposition: G lanes of a warp own one env instance (32/G envs per warp); the env's
// Newton state lives in shared memory (22 doubles per bus); the network's admittances and the
// elimination schedule (identical for all envs) are staged once per CTA with a TMA bulk copy.
// The Newton loop never touches HBM.
//
// Replaces, per env: reference voltage_control_env.py step :178-211, _take_action :548-566,
// _clip_reactive_power :568-572, pp.runpp (pandapower 2.7.0 newtonpf),
// _calc_reward :574-623, voltage_barrier/*.py, _set_demand_and_pv :491-513, get_obs :232-316,
// reset/manual_reset :96-176.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernel_params.h"
#include "philox.cuh"

namespace mapdn {

enum Mode { MODE_SOLVE = 0, MODE_STEP = 1, MODE_RESET = 2 };
constexpr int kMaxResetAttempts = 16;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double nanmax(double a, double b) { return (b > a || b != b) ? b : a; }

template <int G> __device__ __forceinline__ double group_nanmax(double v) {
#pragma unroll
  for (int m = G / 2; m >= 1; m >>= 1) v = nanmax(v, __shfl_xor_sync(kFull, v, m));
  return v;
}
template <int G> __device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int m = G / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(kFull, v, m);
  return v;
}
template <int G> __device__ __forceinline__ double group_max(double v) {
#pragma unroll
  for (int m = G / 2; m >= 1; m >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, m));
  return v;
}

// ---- voltage barriers: reference voltage_barrier/{l1,l2,bowl,bump,courant_beltrami}.py ----
__device__ __forceinline__ double barrier_fn(int kind, double v) {
  switch (kind) {
    case 0: return fabs(v - 1.0);                                   // l1.py:5-8
    case 1: { const double d = v - 1.0; return 2.0 * d * d; }       // l2.py:5-8
    case 2: {                                                       // bowl.py:5-13
      const double d = fabs(v - 1.0);
      if (d > 0.05) return 2.0 * d - 0.095;
      const double dv = v - 1.0;
      const double pdf = 1.0 / sqrt(2.0 * 3.14159265358979323846 * 0.1 * 0.1) *
                         exp(-0.5 * (dv * dv) / (0.1 * 0.1));
      return -0.01 * pdf + 0.04;
    }
    case 3: {                                                       // bump.py:5-13
      if (fabs(v) < 1.0) { const double v2 = v * v; return exp(-1.0 / (1.0 - v2 * v2)); }
      if (v > 1.0 && v < 3.0) { const double w = v - 2.0, w2 = w * w; return exp(-1.0 / (1.0 - w2 * w2)); }
      return 0.0;
    }
    default: {                                                      // courant_beltrami.py:5-8
      const double hi = fmax(0.0, v - 1.05), lo = fmax(0.0, 0.95 - v);
      return hi * hi + lo * lo;
    }
  }
}

// ---- TMA bulk copy of the hot static blob into shared memory (one thread issues) ----
__device__ __forceinline__ void stage_hot_static(unsigned char* smem_dst, const unsigned char* gsrc,
                                                 int bytes, uint64_t* bar) {
  const uint32_t bar_a = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  const uint32_t dst_a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
        "l"(gsrc), "r"(bytes), "r"(bar_a)
        : "memory");
  }
  // every thread waits for phase 0 of the barrier
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar_a)
        : "memory");
  }
}

// Views of the staged static blob and of one env's shared-memory slab.
struct Hot {
  const double *gu, *bu, *gd, *bd, *gii, *bii;
  const uint16_t *parent, *cstart, *eorder, *elev, *dlev;
};
struct Slab {
  double* a[kNodeArrays];
  double* pv;   // sgen.p_mw   [n_sgen]
  double* q;    // sgen.q_mvar [n_sgen]
};

// ------------------------------------------------------------------------------------------
// Newton-Raphson (pandapower newtonpf) for one env per G-lane group.
// Unknowns (dtheta_i, dV_i/V_i) per PQ bus; J's 2x2 blocks are built from per-edge terms
//   a_ik = ViVk(G_ik sin t_ik - B_ik cos t_ik),  b_ik = ViVk(G_ik cos t_ik + B_ik sin t_ik)
// and eliminated leaf-to-root (no fill on a tree), then back-substituted root-to-leaf.
// Returns converged; `iters` = number of linear solves (pandapower's iteration count).
// ------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ bool nr_solve(const Params& p, const Hot& h, const Slab& s, int gl, bool skip, int& iters) {
  const int n = p.n;
  double* vm = s.a[A_VM]; double* va = s.a[A_VA]; double* e = s.a[A_E]; double* f = s.a[A_F];
  double* aup = s.a[A_AUP]; double* bup = s.a[A_BUP]; double* adn = s.a[A_ADN]; double* bdn = s.a[A_BDN];
  double* d0 = s.a[A_D0]; double* d1 = s.a[A_D1]; double* d2 = s.a[A_D2]; double* d3 = s.a[A_D3];
  double* r0 = s.a[A_R0]; double* r1 = s.a[A_R1];
  double* s0 = s.a[A_S0]; double* s1 = s.a[A_S1]; double* s2 = s.a[A_S2]; double* s3 = s.a[A_S3];
  double* t0 = s.a[A_T0]; double* t1 = s.a[A_T1];
  const double* ps = s.a[A_PS]; const double* qs = s.a[A_QS];

  // flat start (pandapower init="auto"): |V| = vm_init, angle 0 on PQ buses; slack fixed
  for (int i = gl; i < n; i += G) {
    const bool root = (i == 0);
    vm[i] = root ? p.vm0 : p.vm_init;
    va[i] = root ? p.va0 : 0.0;
    e[i] = root ? p.e0 : p.vm_init;
    f[i] = root ? p.f0 : 0.0;
  }
  __syncwarp();

  bool done = skip;      // this group's env converged (idle groups never hold the warp back)
  int it = 0;
  iters = 0;
  while (true) {
    // --- per-edge terms (row i / col parent and row parent / col i) ---
    for (int i = 1 + gl; i < n; i += G) {
      const int pa = h.parent[i];
      const double ei = e[i], fi = f[i], ep = e[pa], fp = f[pa];
      const double cc = ei * ep + fi * fp;        // ViVp cos(ti - tp)
      const double ss = fi * ep - ei * fp;        // ViVp sin(ti - tp)
      const double gu = h.gu[i], bu = h.bu[i], gd = h.gd[i], bd = h.bd[i];
      aup[i] = gu * ss - bu * cc;
      bup[i] = gu * cc + bu * ss;
      adn[i] = -gd * ss - bd * cc;
      bdn[i] = gd * cc - bd * ss;
    }
    __syncwarp();
    // --- mismatch F = S_calc - S_spec and diagonal Jacobian blocks ---
    double nrm = 0.0;
    for (int i = 1 + gl; i < n; i += G) {
      double sa = 0.0, sb = 0.0;
      for (int c = h.cstart[i], ce = h.cstart[i + 1]; c < ce; ++c) { sa += adn[c]; sb += bdn[c]; }
      const double vv = e[i] * e[i] + f[i] * f[i];
      const double gv = h.gii[i] * vv, bv = h.bii[i] * vv;
      const double P = gv + bup[i] + sb;
      const double Q = -bv + aup[i] + sa;
      const double Fp = P - ps[i], Fq = Q - qs[i];
      d0[i] = -Q - bv;   // dP/dtheta
      d1[i] = P + gv;    // dP/dV * V
      d2[i] = P - gv;    // dQ/dtheta
      d3[i] = Q - bv;    // dQ/dV * V
      r0[i] = -Fp;
      r1[i] = -Fq;
      nrm = nanmax(nrm, nanmax(fabs(Fp), fabs(Fq)));
    }
    nrm = group_nanmax<G>(nrm);
    if (!done) {
      if (nrm < p.tol) { done = true; iters = it; }
    }
    const bool all_done = __all_sync(kFull, done);
    if (all_done || it >= p.max_iter) break;
    ++it;
    __syncwarp();
    // --- forward elimination, leaves first (levels of equal height) ---
    for (int lev = 0; lev < p.n_elev; ++lev) {
      for (int idx = h.elev[lev] + gl, ie = h.elev[lev + 1]; idx < ie; idx += G) {
        const int i = h.eorder[idx];
        double D0 = d0[i], D1 = d1[i], D2 = d2[i], D3 = d3[i], R0 = r0[i], R1 = r1[i];
        for (int c = h.cstart[i], ce = h.cstart[i + 1]; c < ce; ++c) {
          D0 -= s0[c]; D1 -= s1[c]; D2 -= s2[c]; D3 -= s3[c];
          R0 -= t0[c]; R1 -= t1[c];
        }
        const double idet = 1.0 / (D0 * D3 - D1 * D2);
        const double i00 = D3 * idet, i01 = -D1 * idet, i10 = -D2 * idet, i11 = D0 * idet;
        const double c0 = i00 * R0 + i01 * R1, c1 = i10 * R0 + i11 * R1;
        r0[i] = c0; r1[i] = c1;                       // D^-1 r  (becomes dx in the back sweep)
        if (h.parent[i] != 0) {
          const double au = aup[i], bu = bup[i];      // J[i,parent] = [[au, bu], [-bu, au]]
          const double m00 = i00 * au - i01 * bu, m01 = i00 * bu + i01 * au;
          const double m10 = i10 * au - i11 * bu, m11 = i10 * bu + i11 * au;
          d0[i] = m00; d1[i] = m01; d2[i] = m10; d3[i] = m11;   // D^-1 J[i,parent]
          const double ad = adn[i], bd = bdn[i];      // J[parent,i] = [[ad, bd], [-bd, ad]]
          s0[i] = ad * m00 + bd * m10;  s1[i] = ad * m01 + bd * m11;
          s2[i] = -bd * m00 + ad * m10; s3[i] = -bd * m01 + ad * m11;
          t0[i] = ad * c0 + bd * c1;    t1[i] = -bd * c0 + ad * c1;
        }
      }
      __syncwarp();
    }
    // --- back substitution by depth (BFS numbering: a depth level is a contiguous range);
    //     fused with the state update and V = Vm * exp(j*theta) ---
    for (int d = 1; d < p.n_dlev; ++d) {
      for (int i = h.dlev[d] + gl, ie = h.dlev[d + 1]; i < ie; i += G) {
        double x0 = r0[i], x1 = r1[i];
        const int pa = h.parent[i];
        if (pa != 0) {
          const double xp0 = r0[pa], xp1 = r1[pa];
          x0 -= d0[i] * xp0 + d1[i] * xp1;
          x1 -= d2[i] * xp0 + d3[i] * xp1;
          r0[i] = x0; r1[i] = x1;
        }
        if (!done) {
          const double th = va[i] + x0;
          const double v = vm[i] + vm[i] * x1;
          double sn, cs;
          sincos(th, &sn, &cs);
          va[i] = th; vm[i] = v; e[i] = v * cs; f[i] = v * sn;
        }
      }
      __syncwarp();
    }
  }
  return done;
}

// sgen.q_mvar from an action: reference _clip_reactive_power :568-572
__device__ __forceinline__ double clip_q(double a, double pv, double smax) {
  return sqrt(smax * smax - pv * pv) * a;
}

// Per-unit edge terms + e,f from (vm, va) - used when the previous solution is reloaded from HBM
// (divergence branch, reference :188-196).
template <int G>
__device__ __forceinline__ void recompute_edges(const Params& p, const Hot& h, const Slab& s, int gl) {
  const int n = p.n;
  for (int i = gl; i < n; i += G) {
    double sn, cs;
    sincos(s.a[A_VA][i], &sn, &cs);
    s.a[A_E][i] = s.a[A_VM][i] * cs;
    s.a[A_F][i] = s.a[A_VM][i] * sn;
  }
  __syncwarp();
  for (int i = 1 + gl; i < n; i += G) {
    const int pa = h.parent[i];
    const double ei = s.a[A_E][i], fi = s.a[A_F][i], ep = s.a[A_E][pa], fp = s.a[A_F][pa];
    const double cc = ei * ep + fi * fp, ss = fi * ep - ei * fp;
    s.a[A_ADN][i] = -h.gd[i] * ss - h.bd[i] * cc;
    s.a[A_BDN][i] = h.gd[i] * cc - h.bd[i] * ss;
  }
  __syncwarp();
}

// One observation entry (reference get_obs :232-274):
// obs_i = [P_zone | Q_zone | pv_i | q_i | vm_zone | va_zone (rad)], zero padded.
// P_zone = res_bus.p_mw + sum of sgen.p_mw of the zone's sgens on that bus (the 15/03/24 fix).
template <class BusP, class BusQ, class BusVm, class BusVa, class SgP, class SgQ>
__device__ __forceinline__ double obs_entry(const Params& p, int agent, int k, BusP busp, BusQ busq, BusVm busvm,
                                            BusVa busva, SgP sgp, SgQ sgq) {
  const int z0 = __ldg(p.zptr + agent), nz = __ldg(p.zptr + agent + 1) - z0;
  if (k < 2 * nz) {
    const bool isq = k >= nz;
    const int slot = z0 + (isq ? k - nz : k);
    const int node = __ldg(p.znode + slot);
    double v = isq ? busq(node) : busp(node);
    for (int j = __ldg(p.zsg_ptr + slot), je = __ldg(p.zsg_ptr + slot + 1); j < je; ++j) {
      const int sg = __ldg(p.zsg_idx + j);
      v += isq ? sgq(sg) : sgp(sg);
    }
    return v;
  }
  if (k == 2 * nz) return sgp(agent);
  if (k == 2 * nz + 1) return sgq(agent);
  if (k < 3 * nz + 2) return busvm(__ldg(p.znode + z0 + k - 2 * nz - 2));
  if (k < 4 * nz + 2) return busva(__ldg(p.znode + z0 + k - 3 * nz - 2));
  return 0.0;
}

template <int G, int MODE>
__global__ void __launch_bounds__(128) env_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t stage_bar;
  stage_hot_static(smem_raw, p.hot, p.hot_layout.bytes, &stage_bar);

  const HotLayout& hl = p.hot_layout;
  Hot h;
  h.gu = reinterpret_cast<const double*>(smem_raw + hl.gu);
  h.bu = reinterpret_cast<const double*>(smem_raw + hl.bu);
  h.gd = reinterpret_cast<const double*>(smem_raw + hl.gd);
  h.bd = reinterpret_cast<const double*>(smem_raw + hl.bd);
  h.gii = reinterpret_cast<const double*>(smem_raw + hl.gii);
  h.bii = reinterpret_cast<const double*>(smem_raw + hl.bii);
  h.parent = reinterpret_cast<const uint16_t*>(smem_raw + hl.parent);
  h.cstart = reinterpret_cast<const uint16_t*>(smem_raw + hl.cstart);
  h.eorder = reinterpret_cast<const uint16_t*>(smem_raw + hl.eorder);
  h.elev = reinterpret_cast<const uint16_t*>(smem_raw + hl.elev);
  h.dlev = reinterpret_cast<const uint16_t*>(smem_raw + hl.dlev);

  const int gl = threadIdx.x % G;
  const int gidx = threadIdx.x / G;
  const int epb = blockDim.x / G;
  Slab s;
  {
    double* base = reinterpret_cast<double*>(smem_raw + hl.bytes) + static_cast<size_t>(gidx) * p.env_stride;
#pragma unroll
    for (int a = 0; a < kNodeArrays; ++a) s.a[a] = base + a * p.n_pad;
    s.pv = base + kNodeArrays * p.n_pad;
    s.q = s.pv + p.n_sgen_pad;
  }
  const int n = p.n, nl = p.n_load, ng = p.n_sgen;
  const uint32_t k0 = static_cast<uint32_t>(p.seed), k1 = static_cast<uint32_t>(p.seed >> 32);

  for (int base = blockIdx.x * epb; base < p.nb; base += gridDim.x * epb) {
    int env = base + gidx;
    bool valid = env < p.nb;
    if (!valid) env = p.nb - 1;
    if (MODE == MODE_RESET && p.mask != nullptr && !p.mask[env]) valid = false;
    const size_t eL = static_cast<size_t>(env) * nl, eG = static_cast<size_t>(env) * ng,
                 eN = static_cast<size_t>(env) * n;

    RngKey key{k0, k1, static_cast<uint32_t>(p.env_id_offset + env), 0u};
    long long start = 0;
    int steps_old = 0;
    if (MODE == MODE_STEP) {
      key.c3base = p.episode[env] * 8u;
      start = p.start_row[env];
      steps_old = p.steps[env];
    }
    if (MODE == MODE_RESET) key.c3base = (p.episode[env] + 1u) * 8u;

    bool conv = false;
    int iters = 0;
    int attempt = 0;
    bool solved = false;
    for (int round = 0; round < (MODE == MODE_RESET ? kMaxResetAttempts : 1); ++round) {
      // ---------------- prologue: element values -> sgen p/q and bus injections ----------------
      if (MODE == MODE_RESET) {
        int day, hour, interval;
        if (p.start_dhi != nullptr) {                               // manual_reset :137-152
          day = p.start_dhi[env * 3 + 0]; hour = p.start_dhi[env * 3 + 1]; interval = p.start_dhi[env * 3 + 2];
        } else {                                                    // :111-113, :381-398
          const u32x4 r = philox4x32_10(0u, static_cast<uint32_t>(attempt), key.env, key.c3base + kStreamTime, k0, k1);
          hour = static_cast<int>(__umulhi(r.x, 24u));
          day = static_cast<int>(__umulhi(r.y, static_cast<uint32_t>(p.n_day_choices)));
          interval = static_cast<int>(__umulhi(r.z, static_cast<uint32_t>(p.steps_per_hour)));
        }
        start = interval + static_cast<long long>(hour) * p.steps_per_hour +
                static_cast<long long>(day) * 24 * p.steps_per_hour;          // :445
      }
      long long row = 0;
      if (MODE == MODE_RESET) { row = start + 1; if (row > p.n_rows - 1) row = p.n_rows - 1; }   // t = steps = 1
      const uint32_t c1 = kResetFlag | static_cast<uint32_t>(attempt);
      for (int j = gl; j < ng; j += G) {
        double pv, q;
        if (MODE == MODE_SOLVE) {
          pv = p.in_pv[eG + j]; q = p.in_q[eG + j];
        } else if (MODE == MODE_STEP) {
          pv = p.cur_pv[eG + j];
          q = clip_q(p.actions[eG + j], pv, __ldg(p.s_max + j));             // :553
        } else {
          pv = __ldg(p.prof_pv + row * ng + j);
          if (p.add_noise) pv += __ldg(p.pv_std + j) * half_normal(key, c1, j);   // :498
          q = 0.0;
          if (p.reset_action) {                                               // :120-122, :334-338
            const u32x4 r = philox4x32_10(j, static_cast<uint32_t>(attempt), key.env, key.c3base + kStreamAction, k0, k1);
            const double a = p.action_low + (p.action_high - p.action_low) * u53(r.x, r.y);
            q = clip_q(a, pv, __ldg(p.s_max + j));
          }
          if (valid) p.cur_pv[eG + j] = pv;
        }
        s.pv[j] = pv; s.q[j] = q;
      }
      __syncwarp();
      for (int i = gl; i < n; i += G) {             // A.1: PD/QD per bus, Sbus = -(PD + jQD)/baseMVA
        double pd = 0.0, qd = 0.0;
        for (int t = __ldg(p.lptr + i), te = __ldg(p.lptr + i + 1); t < te; ++t) {
          const int l = __ldg(p.lidx + t);
          double pl, ql;
          if (MODE == MODE_SOLVE) { pl = p.in_pl[eL + l]; ql = p.in_ql[eL + l]; }
          else if (MODE == MODE_STEP) { pl = p.cur_pl[eL + l]; ql = p.cur_ql[eL + l]; }
          else {
            pl = __ldg(p.prof_lp + row * nl + l);
            ql = __ldg(p.prof_lq + row * nl + l);
            if (p.add_noise) {                                                   // :503, :508
              pl += __ldg(p.lp_std + l) * half_normal(key, c1, ng + l);
              ql += __ldg(p.lq_std + l) * half_normal(key, c1, ng + nl + l);
            }
            if (valid) { p.cur_pl[eL + l] = pl; p.cur_ql[eL + l] = ql; }
          }
          const double sc = __ldg(p.lscale + l);
          pd += pl * sc; qd += ql * sc;
        }
        for (int t = __ldg(p.sptr + i), te = __ldg(p.sptr + i + 1); t < te; ++t) {
          const int g = __ldg(p.sidx + t);
          const double sc = __ldg(p.sscale + g);
          pd -= s.pv[g] * sc; qd -= s.q[g] * sc;
        }
        s.a[A_PS][i] = -pd * p.inv_base;
        s.a[A_QS][i] = -qd * p.inv_base;
      }
      __syncwarp();

      // ---------------- Newton-Raphson ----------------
      conv = nr_solve<G>(p, h, s, gl, !valid, iters);
      if (MODE != MODE_RESET) break;
      if (conv) solved = true;
      if (__all_sync(kFull, solved)) break;
      if (!solved) ++attempt;                       // re-draw this env (reference retry loop :108-133)
      __syncwarp();
    }

    // ---------------- epilogue ----------------
    // divergence branch (reference :188-196): fall back to the previous solution kept in HBM
    if (MODE == MODE_STEP && !conv) {
      for (int i = gl; i < n; i += G) {
        const int b = __ldg(p.bus_of_node + i);
        s.a[A_VM][i] = p.res_vm[eN + b];
        s.a[A_VA][i] = p.res_va[eN + b];
        s.a[A_PS][i] = -p.res_p[eN + b] * p.inv_base;
        s.a[A_QS][i] = -p.res_q[eN + b] * p.inv_base;
      }
    }
    __syncwarp();
    if (MODE == MODE_STEP && !__all_sync(kFull, conv)) {
      // (warp-uniform branch) rebuild e,f and the edge terms for the groups that reloaded; the
      // converged groups recompute identical values
      recompute_edges<G>(p, h, s, gl);
    }
    const bool write_res = valid && (MODE != MODE_STEP || conv);

    // slack injection (pfsoln): S0 = V0 conj(Ybus[0,:] V)
    double P0, Q0;
    {
      double sa = 0.0, sb = 0.0;
      for (int c = h.cstart[0], ce = h.cstart[1]; c < ce; ++c) { sa += s.a[A_ADN][c]; sb += s.a[A_BDN][c]; }
      const double vv = p.vm0 * p.vm0;
      P0 = h.gii[0] * vv + sb;
      Q0 = -h.bii[0] * vv + sa;
    }
    constexpr double kRad2Deg = 57.295779513082320876798;
    // per-bus results + voltage statistics (reference _calc_reward :584-596, :610)
    double cnt_lo = 0, cnt_hi = 0, sum_dev = 0, sum_v = 0, max_drop = 0, max_rise = 0, sum_bar = 0;
    const double v_ref = 0.5 * (p.v_lower + p.v_upper);
    for (int b = gl; b < n; b += G) {
      const int i = __ldg(p.node_of_bus + b);
      const double v = s.a[A_VM][i], th = s.a[A_VA][i];
      const double pb = (i == 0) ? -P0 * p.base_mva : -s.a[A_PS][i] * p.base_mva;
      const double qb = (i == 0) ? -Q0 * p.base_mva : -s.a[A_QS][i] * p.base_mva;
      if (MODE == MODE_SOLVE) {
        if (valid) {
          if (p.out_vm) p.out_vm[eN + b] = v;
          if (p.out_va) p.out_va[eN + b] = th * kRad2Deg;
          if (p.out_p) p.out_p[eN + b] = pb;
          if (p.out_q) p.out_q[eN + b] = qb;
        }
      } else {
        if (write_res) {
          p.res_vm[eN + b] = v; p.res_va[eN + b] = th; p.res_p[eN + b] = pb; p.res_q[eN + b] = qb;
        }
        if (MODE == MODE_STEP) {
          cnt_lo += (v < p.v_lower) ? 1.0 : 0.0;
          cnt_hi += (v > p.v_upper) ? 1.0 : 0.0;
          sum_dev += fabs(v - v_ref);
          sum_v += v;
          max_drop = fmax(max_drop, (v < p.v_lower) ? (p.v_lower - v) : 0.0);
          max_rise = fmax(max_rise, (v > p.v_upper) ? (v - p.v_upper) : 0.0);
          sum_bar += barrier_fn(p.barrier, v);
        }
      }
    }
    // line losses: res_line.pl_mw = Re(Sf + St), 4 static coefficients per line
    double sum_pl = 0.0;
    {
      const size_t ePL = static_cast<size_t>(env) * p.n_line;
      for (int k = gl; k < p.n_line; k += G) {
        const int nf = __ldg(p.line_f + k), nt = __ldg(p.line_t + k);
        const double ef = s.a[A_E][nf], ff = s.a[A_F][nf], et = s.a[A_E][nt], ft = s.a[A_F][nt];
        const double cc = ef * et + ff * ft, ss = ff * et - ef * ft;
        const double* c = p.line_c + 4 * k;
        const double pl = __ldg(c) * (ef * ef + ff * ff) + __ldg(c + 1) * (et * et + ft * ft) +
                          __ldg(c + 2) * cc + __ldg(c + 3) * ss;
        sum_pl += pl;
        if (MODE == MODE_SOLVE) { if (valid && p.out_pl) p.out_pl[ePL + k] = pl; }
        else if (write_res) p.res_pl[ePL + k] = pl;
      }
    }
    if (MODE == MODE_SOLVE) {
      if (valid && gl == 0) {
        if (p.out_iters) p.out_iters[env] = conv ? iters : p.max_iter;
        if (p.out_conv) p.out_conv[env] = conv ? 1 : 0;
      }
      __syncwarp();
      continue;
    }

    if (MODE == MODE_STEP) {
      // q terms: res_sgen.q_mvar = sgen.q_mvar * scaling (:604-605); on divergence the reward uses
      // the previous q, info["q_loss"] the attempted one (:189-196)
      double sum_q_eff = 0.0, sum_q_try = 0.0;
      for (int j = gl; j < ng; j += G) {
        const double q_try = s.q[j];
        const double q_eff = conv ? q_try : p.cur_q[eG + j];
        sum_q_try += fabs(q_try);
        sum_q_eff += fabs(q_eff * __ldg(p.sscale + j));
        s.q[j] = q_eff;                                  // sgen.q_mvar after the step (rolled back on failure)
        if (write_res) p.cur_q[eG + j] = q_eff;
      }
      cnt_lo = group_sum<G>(cnt_lo); cnt_hi = group_sum<G>(cnt_hi);
      sum_dev = group_sum<G>(sum_dev); sum_v = group_sum<G>(sum_v); sum_bar = group_sum<G>(sum_bar);
      max_drop = group_max<G>(max_drop); max_rise = group_max<G>(max_rise);
      sum_pl = group_sum<G>(sum_pl); sum_q_eff = group_sum<G>(sum_q_eff); sum_q_try = group_sum<G>(sum_q_try);
      const double inv_n = 1.0 / n;
      const double pct = (cnt_lo + cnt_hi) * inv_n;
      const double q_loss = sum_q_eff / ng;
      const double v_loss = sum_bar * inv_n * p.voltage_weight;
      const double loss = p.use_line_weight ? (sum_pl / p.n_line) * p.line_weight + v_loss
                                            : q_loss * p.q_weight + v_loss;     // :612-618
      double reward = -loss;
      if (!conv) reward -= 200.0;                                               // :192
      const int steps_new = steps_old + 1;                                      // :202
      if (valid && gl == 0) {
        p.reward[env] = reward;
        p.term[env] = (steps_new >= p.episode_limit || !conv) ? 1 : 0;          // :204
        p.steps[env] = steps_new;
        p.sum_rewards[env] += reward;                                           // :203
        if (p.info) {
          double* o = p.info + static_cast<size_t>(env) * 11;
          o[0] = pct; o[1] = cnt_lo * inv_n; o[2] = cnt_hi * inv_n;
          o[3] = (!conv || pct > 1e-3) ? 0.0 : 1.0;                             // :589, :195
          o[4] = sum_dev * inv_n; o[5] = sum_v * inv_n; o[6] = max_drop; o[7] = max_rise;
          o[8] = sum_pl; o[9] = conv ? q_loss : sum_q_try / ng; o[10] = conv ? 0.0 : 1.0;
        }
      }
      // next profile row (reference _set_demand_and_pv :491-513): t = self.steps before the increment
      long long nrow = start + steps_old;
      if (nrow > p.n_rows - 1) nrow = p.n_rows - 1;
      const uint32_t c1 = static_cast<uint32_t>(steps_old);
      for (int j = gl; j < ng; j += G) {
        double pv = __ldg(p.prof_pv + nrow * ng + j);
        if (p.add_noise) pv += __ldg(p.pv_std + j) * half_normal(key, c1, j);
        s.pv[j] = pv;
        if (valid) p.cur_pv[eG + j] = pv;
      }
      for (int l = gl; l < nl; l += G) {
        double pl = __ldg(p.prof_lp + nrow * nl + l), ql = __ldg(p.prof_lq + nrow * nl + l);
        if (p.add_noise) {
          pl += __ldg(p.lp_std + l) * half_normal(key, c1, ng + l);
          ql += __ldg(p.lq_std + l) * half_normal(key, c1, ng + nl + l);
        }
        if (valid) { p.cur_pl[eL + l] = pl; p.cur_ql[eL + l] = ql; }
      }
    } else {  // MODE_RESET
      for (int j = gl; j < ng; j += G) if (valid) p.cur_q[eG + j] = s.q[j];
      if (valid && gl == 0) {
        p.steps[env] = 1;                         // :100
        p.sum_rewards[env] = 0.0;                 // :101
        p.start_row[env] = start;
        p.episode[env] = p.episode[env] + 1u;
      }
    }
    __syncwarp();
    // observations of the new state (reference get_obs :232-316)
    if (p.obs != nullptr) {
      double* o = p.obs + static_cast<size_t>(env) * ng * p.obs_dim;
      const int tot = ng * p.obs_dim;
      auto busp = [&](int i) { return (i == 0) ? -P0 * p.base_mva : -s.a[A_PS][i] * p.base_mva; };
      auto busq = [&](int i) { return (i == 0) ? -Q0 * p.base_mva : -s.a[A_QS][i] * p.base_mva; };
      auto busvm = [&](int i) { return s.a[A_VM][i]; };
      auto busva = [&](int i) { return s.a[A_VA][i]; };
      auto sgp = [&](int j) { return s.pv[j]; };
      auto sgq = [&](int j) { return s.q[j]; };
      for (int idx = gl; idx < tot; idx += G) {
        const int a = idx / p.obs_dim, k = idx - a * p.obs_dim;
        const double v = obs_entry(p, a, k, busp, busq, busvm, busva, sgp, sgq);
        if (valid) o[idx] = v;
      }
    }
    if (MODE == MODE_RESET && p.state != nullptr) {
      // get_state (:213-230): [P_bus | Q_bus | pv | q | vm | va(deg)]
      double* o = p.state + static_cast<size_t>(env) * p.state_dim;
      for (int idx = gl; idx < p.state_dim; idx += G) {
        double v;
        if (idx < 2 * n) {
          const int b = idx < n ? idx : idx - n;
          const int i = __ldg(p.node_of_bus + b);
          v = idx < n ? ((i == 0) ? -P0 * p.base_mva : -s.a[A_PS][i] * p.base_mva)
                      : ((i == 0) ? -Q0 * p.base_mva : -s.a[A_QS][i] * p.base_mva);
        } else if (idx < 2 * n + ng) v = s.pv[idx - 2 * n];
        else if (idx < 2 * n + 2 * ng) v = s.q[idx - 2 * n - ng];
        else if (idx < 3 * n + 2 * ng) v = s.a[A_VM][__ldg(p.node_of_bus + idx - 2 * n - 2 * ng)];
        else v = s.a[A_VA][__ldg(p.node_of_bus + idx - 3 * n - 2 * ng)] * kRad2Deg;
        if (valid) o[idx] = v;
      }
    }
    __syncwarp();
  }
}

}  // namespace mapdn
