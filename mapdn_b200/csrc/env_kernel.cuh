// Fused batched environment kernel for sm_100a: action clip -> bus injections -> polar
// Newton-Raphson on the (radial) Ybus with a zero-fill 2x2-block elimination -> reward / info ->
// next profile row (+ noise) -> zone-masked observation gather. One launch per env step.
//
// Work decomposition: G lanes of a warp own one env instance (32/G envs per warp); the env's
// Newton state lives in shared memory (9 double2 per PQ bus, 128-bit accesses); the network's
// admittances, the elimination schedule, the element->bus maps and the observation program
// (identical for all envs) are staged once per CTA with a TMA bulk copy. The Newton loop never
// touches HBM and is written branch-light: every "missing child" points at an all-zero slot.
//
// The linear solve works on the forest of PQ buses (the slack bus is not an unknown), each tree
// re-rooted at its centre so that the leaf->root elimination has half the depth of the feeder.
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
constexpr double kRad2Deg = 57.295779513082320876798;

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

// Reciprocal without the library's special-case branch: MUFU.RCP64H seed + two Newton steps
// (|rel err| ~ 1 ulp). A zero / denormal pivot yields inf/NaN, which the solver reports as
// "not converged" - the same outcome pandapower reaches through a singular-matrix warning.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
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
  uint32_t done = 0;
  while (!done) {   // every thread waits for phase 0 of the barrier
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
  const double2 *yup, *ydn, *yii, *ysl;
  const uint64_t *ndesc, *edesc;
  const uint16_t *enode, *elev, *dlev, *lptr, *lidx, *sptr, *sidx, *xptr, *xidx, *node_of_bus, *obs_off;
};
struct Slab {
  double2* a[kNodeArrays2];
  double* base;   // the slab as a flat double array (obs program offsets index this)
  double* pv;     // sgen.p_mw   [n_sgen]
  double* q;      // sgen.q_mvar [n_sgen]
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
  const int npq = p.npq;
  double2* VV = s.a[A_VV]; double2* EF = s.a[A_EF]; const double2* SP = s.a[A_SP];
  double2* UP = s.a[A_UP]; double2* DN = s.a[A_DN]; double2* T = s.a[A_T];
  double2* D01 = s.a[A_D01]; double2* D23 = s.a[A_D23]; double2* R = s.a[A_R];

  // flat start (pandapower init="auto"): |V| = vm_init, angle 0 on every PQ bus; sentinel slots
  for (int i = gl; i <= npq; i += G) {
    const bool sl = (i == npq);
    VV[i] = sl ? make_double2(p.vm0, p.va0) : make_double2(p.vm_init, 0.0);
    EF[i] = sl ? make_double2(p.e0, p.f0) : make_double2(p.vm_init, 0.0);
    UP[i] = make_double2(0.0, 0.0);      // roots keep zeros here; slot npq is the "no child" slot
    DN[i] = make_double2(0.0, 0.0);
    T[i] = make_double2(0.0, 0.0);
  }
  __syncwarp();

  bool done = skip;      // this group's env converged (idle groups never hold the warp back)
  int it = 0;
  iters = 0;
  const double2 v0 = make_double2(p.e0, p.f0);
  while (true) {
    // --- per-edge terms of (i, parent): row i / col parent and row parent / col i ---
#pragma unroll 2
    for (int i = gl; i < npq; i += G) {
      const uint32_t pa = static_cast<uint32_t>(h.ndesc[i]) & 0xFFFFu;
      if (pa != kNone) {
        const double2 vi = EF[i], vp = EF[pa];
        const double cc = vi.x * vp.x + vi.y * vp.y;       // ViVp cos(ti - tp)
        const double ss = vi.y * vp.x - vi.x * vp.y;       // ViVp sin(ti - tp)
        const double2 yu = h.yup[i], yd = h.ydn[i];
        UP[i] = make_double2(yu.x * ss - yu.y * cc, yu.x * cc + yu.y * ss);
        DN[i] = make_double2(-yd.x * ss - yd.y * cc, yd.x * cc - yd.y * ss);
      }
    }
    __syncwarp();
    // --- mismatch F = S_calc - S_spec and diagonal Jacobian blocks ---
    double nrm = 0.0;
#pragma unroll 2
    for (int i = gl; i < npq; i += G) {
      const uint64_t nd = h.ndesc[i];
      const int c0 = static_cast<int>((nd >> 16) & 0xFFFFu), c1 = static_cast<int>((nd >> 32) & 0xFFFFu);
      const int nx = static_cast<int>(nd >> 48);
      const double2 vi = EF[i];
      const double2 ys = h.ysl[i];                       // zero unless the bus is adjacent to the slack
      const double cs0 = vi.x * v0.x + vi.y * v0.y, sn0 = vi.y * v0.x - vi.x * v0.y;
      const double2 u = UP[i], a0 = DN[c0], a1 = DN[c1];  // roots: UP = 0; missing children: zero slot
      double sa = ys.x * sn0 - ys.y * cs0 + u.x + a0.x + a1.x;
      double sb = ys.x * cs0 + ys.y * sn0 + u.y + a0.y + a1.y;
#pragma unroll 1
      for (int c = c1 + 1; c <= c1 + nx; ++c) { const double2 d = DN[c]; sa += d.x; sb += d.y; }
      const double vv = vi.x * vi.x + vi.y * vi.y;
      const double2 yi = h.yii[i];
      const double gv = yi.x * vv, bv = yi.y * vv;
      const double P = gv + sb, Q = sa - bv;
      const double2 sp = SP[i];
      const double Fp = P - sp.x, Fq = Q - sp.y;
      D01[i] = make_double2(-Q - bv, P + gv);     // dP/dtheta, dP/dV * V
      D23[i] = make_double2(P - gv, Q - bv);      // dQ/dtheta, dQ/dV * V
      R[i] = make_double2(-Fp, -Fq);
      nrm = nanmax(nrm, nanmax(fabs(Fp), fabs(Fq)));
    }
    nrm = group_nanmax<G>(nrm);
    if (!done && nrm < p.tol) { done = true; iters = it; }
    if (__all_sync(kFull, done) || it >= p.max_iter) break;
    ++it;
    __syncwarp();
    // --- forward elimination, leaves first (levels of equal height) ---
    for (int lev = 0; lev < p.n_lev; ++lev) {
#pragma unroll 1
      for (int idx = h.elev[lev] + gl, ie = h.elev[lev + 1]; idx < ie; idx += G) {
        const int i = h.enode[idx];
        const uint64_t ed = h.edesc[idx];
        const int c0 = static_cast<int>((ed >> 16) & 0xFFFFu), c1 = static_cast<int>((ed >> 32) & 0xFFFFu);
        const int nx = static_cast<int>(ed >> 48);
        double2 d01 = D01[i], d23 = D23[i], r = R[i];
        {   // children's Schur updates (aliased arrays); a missing child reads the zero slot
          const double2 p01 = UP[c0], p23 = DN[c0], pt = T[c0], q01 = UP[c1], q23 = DN[c1], qt = T[c1];
          d01.x -= p01.x + q01.x; d01.y -= p01.y + q01.y;
          d23.x -= p23.x + q23.x; d23.y -= p23.y + q23.y;
          r.x -= pt.x + qt.x; r.y -= pt.y + qt.y;
        }
#pragma unroll 1
        for (int c = c1 + 1; c <= c1 + nx; ++c) {
          const double2 s01 = UP[c], s23 = DN[c], t = T[c];
          d01.x -= s01.x; d01.y -= s01.y; d23.x -= s23.x; d23.y -= s23.y;
          r.x -= t.x; r.y -= t.y;
        }
        const double2 u = UP[i], d = DN[i];            // J[i,p] = [[a,b],[-b,a]](u), J[p,i] likewise (d); roots: 0
        // adjugate form: everything that does not need 1/det runs beside the reciprocal
        const double idet = fast_rcp(d01.x * d23.y - d01.y * d23.x);
        const double ca0 = d23.y * r.x - d01.y * r.y, ca1 = d01.x * r.y - d23.x * r.x;   // adj(D) r
        const double ma00 = d23.y * u.x + d01.y * u.y, ma01 = d23.y * u.y - d01.y * u.x; // adj(D) J[i,p]
        const double ma10 = -d23.x * u.x - d01.x * u.y, ma11 = d01.x * u.x - d23.x * u.y;
        const double sa00 = d.x * ma00 + d.y * ma10, sa01 = d.x * ma01 + d.y * ma11;     // J[p,i] adj(D) J[i,p]
        const double sa10 = d.x * ma10 - d.y * ma00, sa11 = d.x * ma11 - d.y * ma01;
        const double ta0 = d.x * ca0 + d.y * ca1, ta1 = d.x * ca1 - d.y * ca0;           // J[p,i] adj(D) r
        R[i] = make_double2(ca0 * idet, ca1 * idet);           // D^-1 r  (becomes dx in the back sweep)
        D01[i] = make_double2(ma00 * idet, ma01 * idet);       // D^-1 J[i,p]
        D23[i] = make_double2(ma10 * idet, ma11 * idet);
        UP[i] = make_double2(sa00 * idet, sa01 * idet);        // Schur update for the parent (zero at roots)
        DN[i] = make_double2(sa10 * idet, sa11 * idet);
        T[i] = make_double2(ta0 * idet, ta1 * idet);
      }
      __syncwarp();
    }
    // --- back substitution by depth (a depth level is a contiguous node range) ---
    for (int d = 1; d < p.n_lev; ++d) {
#pragma unroll 1
      for (int i = h.dlev[d] + gl, ie = h.dlev[d + 1]; i < ie; i += G) {
        const uint32_t pa = static_cast<uint32_t>(h.ndesc[i]) & 0xFFFFu;   // depth >= 1: always has a parent
        const double2 xp = R[pa], m01 = D01[i], m23 = D23[i];
        double2 x = R[i];
        x.x -= m01.x * xp.x + m01.y * xp.y;
        x.y -= m23.x * xp.x + m23.y * xp.y;
        R[i] = x;
      }
      __syncwarp();
    }
    // --- update (theta += dtheta, V += V * dV/V) and V = Vm exp(j theta) ---
#pragma unroll 2
    for (int i = gl; i < npq; i += G) {
      if (!done) {
        const double2 x = R[i];
        double2 v = VV[i];
        v.y += x.x;
        v.x += v.x * x.y;
        double sn, cs;
        sincos(v.y, &sn, &cs);
        VV[i] = v;
        EF[i] = make_double2(v.x * cs, v.x * sn);
      }
    }
    __syncwarp();
  }
  return done;
}

// sgen.q_mvar from an action: reference _clip_reactive_power :568-572
__device__ __forceinline__ double clip_q(double a, double pv, double smax) {
  return sqrt(smax * smax - pv * pv) * a;
}

template <int G, int MODE>
__global__ void __launch_bounds__(128) env_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t stage_bar;
  stage_hot_static(smem_raw, p.hot, p.hot_layout.bytes, &stage_bar);

  const HotLayout& hl = p.hot_layout;
  Hot h;
  h.yup = reinterpret_cast<const double2*>(smem_raw + hl.yup);
  h.ydn = reinterpret_cast<const double2*>(smem_raw + hl.ydn);
  h.yii = reinterpret_cast<const double2*>(smem_raw + hl.yii);
  h.ysl = reinterpret_cast<const double2*>(smem_raw + hl.ysl);
  h.ndesc = reinterpret_cast<const uint64_t*>(smem_raw + hl.ndesc);
  h.edesc = reinterpret_cast<const uint64_t*>(smem_raw + hl.edesc);
  h.enode = reinterpret_cast<const uint16_t*>(smem_raw + hl.enode);
  h.elev = reinterpret_cast<const uint16_t*>(smem_raw + hl.elev);
  h.dlev = reinterpret_cast<const uint16_t*>(smem_raw + hl.dlev);
  h.lptr = reinterpret_cast<const uint16_t*>(smem_raw + hl.lptr);
  h.lidx = reinterpret_cast<const uint16_t*>(smem_raw + hl.lidx);
  h.sptr = reinterpret_cast<const uint16_t*>(smem_raw + hl.sptr);
  h.sidx = reinterpret_cast<const uint16_t*>(smem_raw + hl.sidx);
  h.xptr = reinterpret_cast<const uint16_t*>(smem_raw + hl.xptr);
  h.xidx = reinterpret_cast<const uint16_t*>(smem_raw + hl.xidx);
  h.node_of_bus = reinterpret_cast<const uint16_t*>(smem_raw + hl.node_of_bus);
  h.obs_off = reinterpret_cast<const uint16_t*>(smem_raw + hl.obs_off);

  const int gl = threadIdx.x % G;
  const int gidx = threadIdx.x / G;
  const int epb = blockDim.x / G;
  const int npq = p.npq, n = p.n_bus, nl = p.n_load, ng = p.n_sgen, na = npq + 1;
  Slab s;
  {
    double2* base = reinterpret_cast<double2*>(smem_raw + hl.bytes) + static_cast<size_t>(gidx) * p.env_stride2;
#pragma unroll
    for (int a = 0; a < kNodeArrays2; ++a) s.a[a] = base + a * na;
    s.base = reinterpret_cast<double*>(base);
    s.pv = reinterpret_cast<double*>(base + p.pvq_off2);
    s.q = s.pv + ng;
  }
  double* stage_pl = reinterpret_cast<double*>(s.a[A_UP]);   // scratch of the prologue: scaled load p / q
  double* stage_ql = stage_pl + nl;
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
      // (1) coalesced, independent loads of the env's element values into shared memory
#pragma unroll 2
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
#pragma unroll 4
      for (int l = gl; l < nl; l += G) {
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
        stage_pl[l] = pl * sc; stage_ql[l] = ql * sc;
      }
      __syncwarp();
      // (2) A.1: PD/QD per bus, Sbus = -(PD + jQD)/baseMVA
      for (int i = gl; i < npq; i += G) {
        double pd = 0.0, qd = 0.0;
#pragma unroll 1
        for (int t = h.lptr[i], te = h.lptr[i + 1]; t < te; ++t) { const int l = h.lidx[t]; pd += stage_pl[l]; qd += stage_ql[l]; }
#pragma unroll 1
        for (int t = h.sptr[i], te = h.sptr[i + 1]; t < te; ++t) {
          const int g = h.sidx[t];
          const double sc = __ldg(p.sscale + g);
          pd -= s.pv[g] * sc; qd -= s.q[g] * sc;
        }
        s.a[A_SP][i] = make_double2(-pd * p.inv_base, -qd * p.inv_base);
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
    double2* VV = s.a[A_VV]; double2* EF = s.a[A_EF]; double2* SP = s.a[A_SP];
    double2* BP = s.a[A_BP]; double2* OP = s.a[A_OP];
    // divergence branch (reference :188-196): fall back to the previous solution kept in HBM
    if (MODE == MODE_STEP && !conv) {
      for (int i = gl; i < npq; i += G) {
        const int b = __ldg(p.bus_of_node + i);
        const double vm = p.res_vm[eN + b], va = p.res_va[eN + b];
        double sn, cs;
        sincos(va, &sn, &cs);
        VV[i] = make_double2(vm, va);
        EF[i] = make_double2(vm * cs, vm * sn);
        SP[i] = make_double2(-p.res_p[eN + b] * p.inv_base, -p.res_q[eN + b] * p.inv_base);
      }
    }
    __syncwarp();
    const bool write_res = valid && (MODE != MODE_STEP || conv);

    // slack injection (pfsoln): S0 = V0 conj(Ybus[0,:] V) -> sentinel slot of SP
    {
      const double vv = p.vm0 * p.vm0;
      double P0 = p.ysl_g0 * vv, Q0 = -p.ysl_b0 * vv;
      for (int k = 0; k < p.n_slack_adj; ++k) {
        const int i = __ldg(p.sl_node + k);
        const double g = __ldg(p.sl_y + 2 * k), b = __ldg(p.sl_y + 2 * k + 1);
        const double2 vi = EF[i];
        const double cc = p.e0 * vi.x + p.f0 * vi.y, ss = p.f0 * vi.x - p.e0 * vi.y;   // V0 Vi cos/sin(t0 - ti)
        P0 += g * cc + b * ss;
        Q0 += g * ss - b * cc;
      }
      if (gl == 0) SP[npq] = make_double2(P0, Q0);
    }
    // sgen.q_mvar after the step: the clipped action, or the previous q on divergence (:189-196)
    double sum_q_eff = 0.0, sum_q_try = 0.0;
    if (MODE == MODE_STEP) {
      for (int j = gl; j < ng; j += G) {
        const double q_try = s.q[j];
        const double q_eff = conv ? q_try : p.cur_q[eG + j];
        sum_q_try += fabs(q_try);
        sum_q_eff += fabs(q_eff * __ldg(p.sscale + j));     // res_sgen.q_mvar = q * scaling (:604-605)
        s.q[j] = q_eff;
        if (write_res) p.cur_q[eG + j] = q_eff;
      }
    }
    // next profile row (reference _set_demand_and_pv :491-513): t = self.steps before the increment.
    // Elements [pv | load_p | load_q] are drawn in Box-Muller pairs (2m, 2m+1).
    if (MODE == MODE_STEP) {
      long long nrow = start + steps_old;
      if (nrow > p.n_rows - 1) nrow = p.n_rows - 1;
      const uint32_t c1 = static_cast<uint32_t>(steps_old);
      const int n_elem = ng + 2 * nl;
#pragma unroll 2
      for (int m = gl; 2 * m < n_elem; m += G) {
        double z[2] = {0.0, 0.0};
        if (p.add_noise) half_normal_pair(key, c1, m, z[0], z[1]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int el = 2 * m + u;
          if (el >= n_elem) break;
          if (el < ng) {
            const double pv = __ldg(p.prof_pv + nrow * ng + el) + __ldg(p.pv_std + el) * z[u];
            s.pv[el] = pv;
            if (valid) p.cur_pv[eG + el] = pv;
          } else if (el < ng + nl) {
            const int l = el - ng;
            const double pl = __ldg(p.prof_lp + nrow * nl + l) + __ldg(p.lp_std + l) * z[u];
            if (valid) p.cur_pl[eL + l] = pl;
          } else {
            const int l = el - ng - nl;
            const double ql = __ldg(p.prof_lq + nrow * nl + l) + __ldg(p.lq_std + l) * z[u];
            if (valid) p.cur_ql[eL + l] = ql;
          }
        }
      }
    }
    __syncwarp();
    // res_bus columns per node (BP) and the "demand" columns of get_obs (OP = BP + sgens of the bus's own zone)
    for (int i = gl; i <= npq; i += G) {
      const double2 sp = SP[i];
      const double2 bp = make_double2(-sp.x * p.base_mva, -sp.y * p.base_mva);
      double2 op = bp;
#pragma unroll 1
      for (int t = h.xptr[i], te = h.xptr[i + 1]; t < te; ++t) { const int g = h.xidx[t]; op.x += s.pv[g]; op.y += s.q[g]; }
      BP[i] = bp; OP[i] = op;
    }
    __syncwarp();
    // per-bus results + voltage statistics (reference _calc_reward :584-596, :610)
    double cnt_lo = 0, cnt_hi = 0, sum_dev = 0, sum_v = 0, max_drop = 0, max_rise = 0, sum_bar = 0;
    const double v_ref = 0.5 * (p.v_lower + p.v_upper);
#pragma unroll 2
    for (int b = gl; b < n; b += G) {
      const int i = h.node_of_bus[b];
      const double2 vv = VV[i], bp = BP[i];
      const double v = vv.x, th = vv.y;
      if (MODE == MODE_SOLVE) {
        if (valid) {
          if (p.out_vm) p.out_vm[eN + b] = v;
          if (p.out_va) p.out_va[eN + b] = th * kRad2Deg;
          if (p.out_p) p.out_p[eN + b] = bp.x;
          if (p.out_q) p.out_q[eN + b] = bp.y;
        }
      } else {
        if (write_res) {
          p.res_vm[eN + b] = v; p.res_va[eN + b] = th; p.res_p[eN + b] = bp.x; p.res_q[eN + b] = bp.y;
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
#pragma unroll 2
      for (int k = gl; k < p.n_line; k += G) {
        const int nf = __ldg(p.line_f + k), nt = __ldg(p.line_t + k);
        const double2 vf = EF[nf], vt = EF[nt];
        const double cc = vf.x * vt.x + vf.y * vt.y, ss = vf.y * vt.x - vf.x * vt.y;
        const double* c = p.line_c + 4 * k;
        const double pl = __ldg(c) * (vf.x * vf.x + vf.y * vf.y) + __ldg(c + 1) * (vt.x * vt.x + vt.y * vt.y) +
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
    } else {  // MODE_RESET
      for (int j = gl; j < ng; j += G) if (valid) p.cur_q[eG + j] = s.q[j];
      if (valid && gl == 0) {
        p.steps[env] = 1;                         // :100
        p.sum_rewards[env] = 0.0;                 // :101
        p.start_row[env] = start;
        p.episode[env] = p.episode[env] + 1u;
      }
    }
    // observations of the new state (reference get_obs :232-316): a pure gather - the program maps
    // every entry to a double inside this env's slab (or to a constant-zero slot for the padding)
    if (p.obs != nullptr && valid) {
      double* o = p.obs + static_cast<size_t>(env) * ng * p.obs_dim;
      const int tot = ng * p.obs_dim;
#pragma unroll 4
      for (int idx = gl; idx < tot; idx += G) o[idx] = s.base[h.obs_off[idx]];
    }
    if (MODE == MODE_RESET && p.state != nullptr) {
      // get_state (:213-230): [P_bus | Q_bus | pv | q | vm | va(deg)]
      double* o = p.state + static_cast<size_t>(env) * p.state_dim;
      for (int idx = gl; idx < p.state_dim; idx += G) {
        double v;
        if (idx < n) v = BP[h.node_of_bus[idx]].x;
        else if (idx < 2 * n) v = BP[h.node_of_bus[idx - n]].y;
        else if (idx < 2 * n + ng) v = s.pv[idx - 2 * n];
        else if (idx < 2 * n + 2 * ng) v = s.q[idx - 2 * n - ng];
        else if (idx < 3 * n + 2 * ng) v = VV[h.node_of_bus[idx - 2 * n - 2 * ng]].x;
        else v = VV[h.node_of_bus[idx - 3 * n - 2 * ng]].y * kRad2Deg;
        if (valid) o[idx] = v;
      }
    }
    __syncwarp();
  }
}

}  // namespace mapdn
