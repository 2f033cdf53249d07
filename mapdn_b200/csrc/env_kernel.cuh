// Fused batched environment kernel for sm_100a: action clip -> bus injections -> polar
// Newton-Raphson on the (radial) Ybus with a zero-fill 2x2-block elimination -> reward / info ->
// next profile row (+ noise) -> zone-masked observation gather. One launch per env step.
//
// Work decomposition: a group of G threads owns one env instance - a sub-warp slice (G <= 32: 32/G envs
// per warp in lock-step) or 2-4 whole warps (G = 64 / 128, large feeders); the env's
// Newton state lives in shared memory (one 144 B record per PQ bus, 128-bit accesses); the network's
// admittances, the elimination schedule, the element->bus maps and the observation program
// (identical for all envs) are staged once per CTA with a TMA bulk copy. The Newton loop never
// touches HBM and is written branch-light: every "missing child" points at an all-zero slot.
// In MODE_STEP one extra (helper) warp per CTA draws the next profile rows + noise of the CTA's
// envs concurrently with the Newton iteration (warp specialisation, two named barriers).
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

// Profile builds (MAPDN_PROFILE_BUILD=1 python -m mapdn_b200.build): clock64 totals per phase of warp 0 of block 0,
// printed by launch_env_kernel (scripts/phase_prof.py). Compiled out otherwise.
#ifdef MAPDN_PROFILE
#define PROF_DECL long long _pt = clock64(); long long _acc[16] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0};
#define PROF(k) { long long _t = clock64(); _acc[k] += _t - _pt; _pt = _t; }
// inside one forward-sweep step (slots 12..15; they overlap slot 5 "elim"): [12] loop top -> children subtracted (load
// wait), [13] -> results stored (arithmetic), [14] step executions, [15] linear solves
#define PROF_STEP_BEGIN long long _s0 = clock64();
#define PROF_STEP_MID long long _s1 = clock64(); _acc[12] += _s1 - _s0;
#define PROF_STEP_END { _acc[13] += clock64() - _s1; _acc[14] += 1; }
#define PROF_COUNT(k) { _acc[k] += 1; }
#else
#define PROF_DECL
#define PROF(k)
#define PROF_STEP_BEGIN
#define PROF_STEP_MID
#define PROF_STEP_END
#define PROF_COUNT(k)
#endif

enum Mode { MODE_SOLVE = 0, MODE_STEP = 1, MODE_RESET = 2 };
constexpr int kMaxResetAttempts = 16;
constexpr unsigned kFull = 0xffffffffu;
constexpr double kRad2Deg = 57.295779513082320876798;

__device__ __forceinline__ double nanmax(double a, double b) { return (b > a || b != b) ? b : a; }

// ---- group primitives. A group = the G threads that own one env: a sub-warp slice (G <= 32, all groups of
//      a warp run in lock-step, synchronised with __syncwarp) or G/32 whole warps (G = 64 / 128, large feeders:
//      synchronised with a named barrier per group; ids 3.. - 0 is __syncthreads, 1-2 the helper warp's) ----
template <int G> __device__ __forceinline__ void grp_sync(int gidx) {
  if constexpr (G <= 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(3 + gidx), "r"(G) : "memory");
}
// true iff `pred` holds on every thread of the group (includes a group barrier when G > 32)
template <int G> __device__ __forceinline__ bool grp_all(int gidx, bool pred) {
  if constexpr (G <= 32) {
    const unsigned ok = __ballot_sync(kFull, pred);
    const unsigned gmask = (G == 32) ? kFull : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    return (ok & gmask) == gmask;
  } else {
    unsigned r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 q, %1, 0;\n\t"
        "bar.red.and.pred p, %2, %3, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(r)
        : "r"(static_cast<unsigned>(pred)), "r"(3 + gidx), "r"(G)
        : "memory");
    return r != 0;
  }
}
// loop-exit test of the lock-step sub-warp groups: every env handled by this warp is done (G <= 32);
// for multi-warp groups the env's own flag decides (it is group-uniform)
template <int G> __device__ __forceinline__ bool grp_exit(bool done) {
  if constexpr (G <= 32) return __all_sync(kFull, done);
  else return done;
}
template <int G> __device__ __forceinline__ double warp_part_sum(double v) {
#pragma unroll
  for (int m = (G < 32 ? G : 32) / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(kFull, v, m);
  return v;
}
template <int G> __device__ __forceinline__ double warp_part_max(double v) {
#pragma unroll
  for (int m = (G < 32 ? G : 32) / 2; m >= 1; m >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, m));
  return v;
}
// sum / max over the group of NV values at once; `scratch` = >= NV * G/32 doubles of the env's slab (G > 32 only)
template <int G, int NV> __device__ __forceinline__ void grp_reduce(int gidx, int gl, double (&v)[NV], const bool (&is_max)[NV],
                                                                    double* scratch) {
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = is_max[k] ? warp_part_max<G>(v[k]) : warp_part_sum<G>(v[k]);
  if constexpr (G > 32) {
    constexpr int W = G / 32;
    grp_sync<G>(gidx);                                   // scratch is free
    if ((gl & 31) == 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) scratch[(gl >> 5) * NV + k] = v[k];
    }
    grp_sync<G>(gidx);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double a = scratch[k];
#pragma unroll
      for (int w = 1; w < W; ++w) a = is_max[k] ? fmax(a, scratch[w * NV + k]) : a + scratch[w * NV + k];
      v[k] = a;
    }
  }
}

// Reciprocal without the library's special-case branch: MUFU.RCP64H seed (~2^-23 rel. error) + two
// Newton steps (-> ~1 ulp). A zero / denormal pivot yields inf/NaN, which the solver reports as
// "not converged" - the same outcome pandapower reaches through a singular-matrix warning.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// sin/cos for the bus angles: distribution-feeder angles are a few degrees, so the common path is a pair of
// Taylor polynomials on |x| <= pi/4 (truncation < 1e-19, i.e. exact to rounding; two independent FMA chains,
// no range reduction); larger angles fall back to the library routine.
__device__ __forceinline__ void sincos_angle(double x, double* sn, double* cs) {
  if (fabs(x) <= 0.78539816339744830962) {
    const double z = x * x;
    double ps = 1.0 / 355687428096000.0;                 // 1/17!
    ps = fma(ps, z, -1.0 / 1307674368000.0);             // -1/15!
    ps = fma(ps, z, 1.0 / 6227020800.0);                 // 1/13!
    ps = fma(ps, z, -1.0 / 39916800.0);                  // -1/11!
    ps = fma(ps, z, 1.0 / 362880.0);                     // 1/9!
    ps = fma(ps, z, -1.0 / 5040.0);                      // -1/7!
    ps = fma(ps, z, 1.0 / 120.0);                        // 1/5!
    ps = fma(ps, z, -1.0 / 6.0);                         // -1/3!
    double pc = 1.0 / 20922789888000.0;                  // 1/16!
    pc = fma(pc, z, -1.0 / 87178291200.0);               // -1/14!
    pc = fma(pc, z, 1.0 / 479001600.0);                  // 1/12!
    pc = fma(pc, z, -1.0 / 3628800.0);                   // -1/10!
    pc = fma(pc, z, 1.0 / 40320.0);                      // 1/8!
    pc = fma(pc, z, -1.0 / 720.0);                       // -1/6!
    pc = fma(pc, z, 1.0 / 24.0);                         // 1/4!
    pc = fma(pc, z, -0.5);                               // -1/2!
    *sn = fma(ps * z, x, x);
    *cs = fma(pc, z, 1.0);
  } else {
    sincos(x, sn, cs);
  }
}

// exp(x) for x in [-0.125, 0] (the only range the bowl barrier needs): degree-9 Taylor, |err| < 3e-15
__device__ __forceinline__ double exp_small(double x) {
  double r = 1.0 / 362880.0;
  r = fma(r, x, 1.0 / 40320.0); r = fma(r, x, 1.0 / 5040.0); r = fma(r, x, 1.0 / 720.0);
  r = fma(r, x, 1.0 / 120.0); r = fma(r, x, 1.0 / 24.0); r = fma(r, x, 1.0 / 6.0);
  r = fma(r, x, 0.5); r = fma(r, x, 1.0); r = fma(r, x, 1.0);
  return r;
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
      // normal pdf N(v; 1, 0.1); |dv| <= 0.05 here, so the exponent lies in [-0.125, 0]
      const double pdf = 3.9894228040143267794 * exp_small(-50.0 * (dv * dv));
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
__device__ __forceinline__ void stage_hot_issue(unsigned char* smem_dst, const unsigned char* gsrc,
                                                int bytes, uint64_t* bar) {
  const uint32_t bar_a = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  const uint32_t dst_a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
        "l"(gsrc), "r"(bytes), "r"(bar_a)
        : "memory");
  }
  __syncthreads();      // the barrier is initialised before anybody waits on it
}
__device__ __forceinline__ void stage_hot_wait(uint64_t* bar) {
  const uint32_t bar_a = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  uint32_t done = 0;
  while (!done) {       // every thread waits for phase 0 of the barrier
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar_a)
        : "memory");
  }
}

// Named barriers between the solver warps and the helper warp of a CTA (producer / consumer)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Views of the staged static blob and of one env's shared-memory slab.
struct Hot {
  const double2 *yup, *ydn, *yii, *ysl;
  const uint64_t *ndesc, *esched, *bsched;
  const uint16_t *lptr, *lidx, *sptr, *sidx, *xptr, *xidx, *node_of_bus, *obs_off, *line_nodes;
  const double* line_c;
  const uint16_t *nbr_ptr, *nbr_idx;   // meshed nets (dense solver) only
  const double2* nbr_y;
};
struct Slab {
  double2* nodes;   // (npq + 2) records of kNodeArrays2 double2 (sentinel + trash at the end)
  double* base;     // the slab as a flat double array (obs program offsets index this)
  double* pv;       // sgen.p_mw   [n_sgen]
  double* q;        // sgen.q_mvar [n_sgen]
  double* scratch;  // [n_sgen + 2 n_load]
  __device__ __forceinline__ double2* node(int i) const { return nodes + i * kNodeArrays2; }
};

// ------------------------------------------------------------------------------------------
// Newton-Raphson (pandapower newtonpf) for one env per G-lane group.
// Unknowns (dtheta_i, dV_i/V_i) per PQ bus; J's 2x2 blocks are built from per-edge terms
//   a_ik = ViVk(G_ik sin t_ik - B_ik cos t_ik),  b_ik = ViVk(G_ik cos t_ik + B_ik sin t_ik)
// and eliminated leaf-to-root (no fill on a tree), then back-substituted root-to-leaf.
// Returns converged; `iters` = number of linear solves (pandapower's iteration count).
// ------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ bool nr_solve(const Params& p, const Hot& h, const Slab& s, int gidx, int gl, bool skip, int& iters
#ifdef MAPDN_PROFILE
    , long long* _acc, long long& _pt
#endif
) {
  const int npq = p.npq;

  // flat start (pandapower init="auto"): |V| = vm_init, angle 0 on every PQ bus; sentinel record
  for (int i = gl; i <= npq + 1; i += G) {
    const bool sl = (i == npq);
    double2* nd = s.node(i);
    nd[A_VV] = sl ? make_double2(p.vm0, p.va0) : make_double2(p.vm_init, 0.0);
    nd[A_EF] = sl ? make_double2(p.e0, p.f0) : make_double2(p.vm_init, 0.0);
    nd[A_UP] = make_double2(0.0, 0.0);      // record npq is the "no child" slot
    nd[A_DN] = make_double2(0.0, 0.0);
    nd[A_T] = make_double2(0.0, 0.0);
    nd[A_R] = make_double2(0.0, 0.0);       // record npq is also the "no parent" slot of the back sweep
    if (i >= npq) { nd[A_D01] = make_double2(1.0, 0.0); nd[A_D23] = make_double2(0.0, 1.0); }
  }
  grp_sync<G>(gidx);

  bool done = skip;      // this group's env converged (idle groups never hold the warp back)
  int it = 0;
  iters = 0;
  const double2 v0 = make_double2(p.e0, p.f0);
  while (true) {
    PROF(2)
    // --- per-edge terms of (i, parent): row i / col parent and row parent / col i ---
#pragma unroll 4
    for (int i = gl; i < npq; i += G) {
      const int pa = static_cast<int>(static_cast<uint32_t>(h.ndesc[i]) & 0xFFFFu);   // roots: sentinel, Y = 0
      double2* nd = s.node(i);
      const double2 vi = nd[A_EF], vp = s.node(pa)[A_EF];
      const double2 yu = h.yup[i], yd = h.ydn[i];
      const double cc = vi.x * vp.x + vi.y * vp.y;       // ViVp cos(ti - tp)
      const double ss = vi.y * vp.x - vi.x * vp.y;       // ViVp sin(ti - tp)
      nd[A_UP] = make_double2(yu.x * ss - yu.y * cc, yu.x * cc + yu.y * ss);
      nd[A_DN] = make_double2(-yd.x * ss - yd.y * cc, yd.x * cc - yd.y * ss);
    }
    grp_sync<G>(gidx);
    PROF(3)
    // --- mismatch F = S_calc - S_spec and diagonal Jacobian blocks ---
    double nrm = 0.0;
#pragma unroll 4
    for (int i = gl; i < npq; i += G) {
      const uint64_t ndc = h.ndesc[i];
      const int c0 = static_cast<int>((ndc >> 16) & 0xFFFFu), c1 = static_cast<int>((ndc >> 32) & 0xFFFFu);
      const int nx = static_cast<int>(ndc >> 48);
      double2* nd = s.node(i);
      const double2 vi = nd[A_EF];
      const double2 u = nd[A_UP], a0 = s.node(c0)[A_DN], a1 = s.node(c1)[A_DN];  // roots: UP = 0; no child: zero slot
      const double2 sp = nd[A_SP];
      const double2 ys = h.ysl[i];                       // zero unless the bus is adjacent to the slack
      const double2 yi = h.yii[i];
      const double cs0 = vi.x * v0.x + vi.y * v0.y, sn0 = vi.y * v0.x - vi.x * v0.y;
      double sa = ys.x * sn0 - ys.y * cs0 + u.x + a0.x + a1.x;
      double sb = ys.x * cs0 + ys.y * sn0 + u.y + a0.y + a1.y;
      if (p.has_extra_children) {          // warp-uniform: only nets with a bus of degree > 3
#pragma unroll 1
        for (int c = c1 + 1; c <= c1 + nx; ++c) { const double2 d = s.node(c)[A_DN]; sa += d.x; sb += d.y; }
      }
      const double vv = vi.x * vi.x + vi.y * vi.y;
      const double gv = yi.x * vv, bv = yi.y * vv;
      const double P = gv + sb, Q = sa - bv;
      const double Fp = P - sp.x, Fq = Q - sp.y;
      nd[A_D01] = make_double2(-Q - bv, P + gv);     // dP/dtheta, dP/dV * V
      nd[A_D23] = make_double2(P - gv, Q - bv);      // dQ/dtheta, dQ/dV * V
      nd[A_R] = make_double2(-Fp, -Fq);
      nrm = nanmax(nrm, nanmax(fabs(Fp), fabs(Fq)));
    }
    {   // ||F||inf < tol for the whole env <=> every thread of the group is below tol (NaN-safe)
      const bool ok = grp_all<G>(gidx, nrm < p.tol);
      if (!done && ok) { done = true; iters = it; }
    }
    if (grp_exit<G>(done) || it >= p.max_iter) break;
    ++it;
    grp_sync<G>(gidx);
    PROF(4)
    // --- forward elimination, leaves first: a flat schedule of steps (a level of the elimination forest,
    //     split when it is wider than the group), one entry per (step, lane). Lanes follow chains of the
    //     forest: when a bus was eliminated by the same lane in the previous step, its Schur update reaches
    //     the parent in registers (no shared-memory round trip on the critical path); the node's own blocks
    //     are prefetched one step ahead. Idle lanes work on the trash record: no divergent branches. ---
    {
      struct Own { double2 d01, d23, r, u, d; };
      auto load_own = [&](uint64_t e) {
        const double2* nd = s.node(static_cast<int>(e & 0xFFFFu));
        Own o; o.d01 = nd[A_D01]; o.d23 = nd[A_D23]; o.r = nd[A_R]; o.u = nd[A_UP]; o.d = nd[A_DN];
        return o;
      };
      uint64_t ed = h.esched[gl];
      uint64_t ed_next = h.esched[max(0, min(1, p.n_esteps - 1)) * G + gl];
      Own own = load_own(ed);
      double2 s01 = make_double2(0.0, 0.0), s23 = s01, tt = s01;     // Schur update produced by this lane's last step
      for (int st = 0; st < p.n_esteps; ++st) {
        PROF_STEP_BEGIN
        const uint64_t ed_next2 = h.esched[min(st + 2, p.n_esteps - 1) * G + gl];   // independent of the data
#ifndef MAPDN_EXP_CHILD_LOADS_FIRST
        const Own own_next = load_own(ed_next);          // not touched before its own step
#endif
        const int i = static_cast<int>(ed & 0xFFFFu);
        const int c0 = static_cast<int>((ed >> 16) & 0xFFFFu), c1 = static_cast<int>((ed >> 32) & 0xFFFFu);
        const unsigned fl = static_cast<unsigned>(ed >> 48);
        double2* nd = s.node(i);
        double2 d01 = own.d01, d23 = own.d23, r = own.r;
        const double2 u = own.u, d = own.d;   // J[i,p] = [[a,b],[-b,a]](u), J[p,i] likewise (d); zero at roots
        grp_sync<G>(gidx);                                    // the previous step's Schur updates are visible
        // child 0: this lane's registers (chain) or shared memory (a leaf reads the all-zero sentinel record)
        double2 p01 = s01, p23 = s23, pt = tt;
        if (fl & kEschedLoad0) { const double2* k0 = s.node(c0); p01 = k0[A_UP]; p23 = k0[A_DN]; pt = k0[A_T]; }
        d01.x -= p01.x; d01.y -= p01.y; d23.x -= p23.x; d23.y -= p23.y; r.x -= pt.x; r.y -= pt.y;
#ifdef MAPDN_EXP_CHILD_LOADS_FIRST      // experiment: prefetch the next step's blocks behind the child loads
        const Own own_next = load_own(ed_next);
#endif
        if (fl & kEschedLoad1) {                          // child 1 (branching buses only): always shared memory
          const double2* k1 = s.node(c1);
          const double2 q01 = k1[A_UP], q23 = k1[A_DN], qt = k1[A_T];
          d01.x -= q01.x; d01.y -= q01.y; d23.x -= q23.x; d23.y -= q23.y; r.x -= qt.x; r.y -= qt.y;
        }
        if (p.has_extra_children) {                      // warp-uniform: only nets with a bus of degree > 3
          const int nx = static_cast<int>(fl & 0xFFu);
#pragma unroll 1
          for (int c = c1 + 1; c <= c1 + nx; ++c) {
            const double2* k = s.node(c);
            const double2 x01 = k[A_UP], x23 = k[A_DN], xt = k[A_T];
            d01.x -= x01.x; d01.y -= x01.y; d23.x -= x23.x; d23.y -= x23.y;
            r.x -= xt.x; r.y -= xt.y;
          }
        }
        PROF_STEP_MID
        // adjugate form: everything that does not need 1/det runs beside the reciprocal (shorter chain, +8 flops)
        const double idet = fast_rcp(d01.x * d23.y - d01.y * d23.x);
        const double ca0 = d23.y * r.x - d01.y * r.y, ca1 = d01.x * r.y - d23.x * r.x;   // adj(D) r
        const double ma00 = d23.y * u.x + d01.y * u.y, ma01 = d23.y * u.y - d01.y * u.x; // adj(D) J[i,p]
        const double ma10 = -d23.x * u.x - d01.x * u.y, ma11 = d01.x * u.x - d23.x * u.y;
        const double sa00 = d.x * ma00 + d.y * ma10, sa01 = d.x * ma01 + d.y * ma11;     // J[p,i] adj(D) J[i,p]
        const double sa10 = d.x * ma10 - d.y * ma00, sa11 = d.x * ma11 - d.y * ma01;
        const double ta0 = d.x * ca0 + d.y * ca1, ta1 = d.x * ca1 - d.y * ca0;           // J[p,i] adj(D) r
        const double c0v = ca0 * idet, c1v = ca1 * idet;
        const double m00 = ma00 * idet, m01 = ma01 * idet, m10 = ma10 * idet, m11 = ma11 * idet;
        s01 = make_double2(sa00 * idet, sa01 * idet);
        s23 = make_double2(sa10 * idet, sa11 * idet);
        tt = make_double2(ta0 * idet, ta1 * idet);
        if (fl & kEschedStore) { nd[A_UP] = s01; nd[A_DN] = s23; nd[A_T] = tt; }   // else: it travels in registers
        if (!(fl & kEschedIdle)) {
          nd[A_R] = make_double2(c0v, c1v);                       // D^-1 r  (becomes dx in the back sweep)
          nd[A_D01] = make_double2(m00, m01);                     // D^-1 J[i,p]
          nd[A_D23] = make_double2(m10, m11);
        }
        ed = ed_next; ed_next = ed_next2; own = own_next;
        PROF_STEP_END
      }
      PROF_COUNT(15)
      grp_sync<G>(gidx);                                 // the last step's results are visible to the back sweep
    }
    PROF(5)
    // --- back substitution root -> leaves, same flat-schedule form (roots: dx = D^-1 r already); a lane
    //     that solved the parent in the previous step keeps dx_parent in registers ---
    {
      struct OwnB { double2 m01, m23, x; };
      auto load_own = [&](uint64_t e) {
        const double2* nd = s.node(static_cast<int>(e & 0xFFFFu));
        OwnB o; o.m01 = nd[A_D01]; o.m23 = nd[A_D23]; o.x = nd[A_R];
        return o;
      };
      uint64_t bd = h.bsched[gl];
      uint64_t bd_next = h.bsched[max(0, min(1, p.n_bsteps - 1)) * G + gl];
      OwnB own = load_own(bd);
      double2 xl = make_double2(0.0, 0.0);               // dx of the node this lane solved in the previous step
      for (int st = 0; st < p.n_bsteps; ++st) {
        const uint64_t bd_next2 = h.bsched[max(0, min(st + 2, p.n_bsteps - 1)) * G + gl];
#ifndef MAPDN_EXP_CHILD_LOADS_FIRST
        const OwnB own_next = load_own(bd_next);         // D^-1 J and D^-1 r are final since the forward sweep
#endif
        double2* nd = s.node(static_cast<int>(bd & 0xFFFFu));
        grp_sync<G>(gidx);                                    // the previous step's dx are visible
        double2 xp = xl;
        if (!((bd >> 32) & 1u)) xp = s.node(static_cast<int>((bd >> 16) & 0xFFFFu))[A_R];
#ifdef MAPDN_EXP_CHILD_LOADS_FIRST
        const OwnB own_next = load_own(bd_next);         // experiment: behind the parent's dx the step waits for
#endif
        double2 x = own.x;
        x.x -= own.m01.x * xp.x + own.m01.y * xp.y;
        x.y -= own.m23.x * xp.x + own.m23.y * xp.y;
        if (!((bd >> 33) & 1u)) nd[A_R] = x;             // idle lanes (trash record) store nothing
        xl = x;
        bd = bd_next; bd_next = bd_next2; own = own_next;
      }
      grp_sync<G>(gidx);
    }
    PROF(6)
    // --- update (theta += dtheta, V += V * dV/V) and V = Vm exp(j theta) ---
#pragma unroll 4
    for (int i = gl; i < npq; i += G) {
      if (!done) {
        double2* nd = s.node(i);
        const double2 x = nd[A_R];
        double2 v = nd[A_VV];
        v.y += x.x;
        v.x += v.x * x.y;
        double sn, cs;
        sincos_angle(v.y, &sn, &cs);
        nd[A_VV] = v;
        nd[A_EF] = make_double2(v.x * cs, v.x * sn);
      }
    }
    grp_sync<G>(gidx);
  }
  return done;
}

// ------------------------------------------------------------------------------------------
// Fallback for meshed networks: the same Newton-Raphson (same unknowns, same iterates up to rounding),
// but the Jacobian is assembled dense ([2 npq] x [2 npq], row-major, unknown 2i = dtheta_i, 2i+1 = dV_i/V_i)
// in a global-memory workspace and solved by one warp with partial-pivoting LU (what SuperLU does for
// pandapower). O(npq^3) per solve: meant for small meshed feeders, not for the radial hot path.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool nr_solve_dense(const Params& p, const Hot& h, const Slab& s, double* __restrict__ ws,
                                               int lane, bool skip, int& iters) {
  const int npq = p.npq, m = 2 * npq, ld = m + 1;       // column m holds the right-hand side
  for (int i = lane; i <= npq + 1; i += 32) {
    const bool sl = (i == npq);
    double2* nd = s.node(i);
    nd[A_VV] = sl ? make_double2(p.vm0, p.va0) : make_double2(p.vm_init, 0.0);
    nd[A_EF] = sl ? make_double2(p.e0, p.f0) : make_double2(p.vm_init, 0.0);
    nd[A_UP] = make_double2(0.0, 0.0); nd[A_DN] = make_double2(0.0, 0.0); nd[A_T] = make_double2(0.0, 0.0);
    nd[A_R] = make_double2(0.0, 0.0);
  }
  __syncwarp();
  bool done = skip;
  int it = 0;
  iters = 0;
  const double2 v0 = make_double2(p.e0, p.f0);
  while (true) {
    for (int idx = lane; idx < m * ld; idx += 32) ws[idx] = 0.0;
    __syncwarp();
    // mismatch + Jacobian rows of bus i (SURVEY A.4 formulas, unknown dV/V)
    double nrm = 0.0;
    for (int i = lane; i < npq; i += 32) {
      const double2 vi = s.node(i)[A_EF];
      const double2 ys = h.ysl[i], yi = h.yii[i], sp = s.node(i)[A_SP];
      const double cs0 = vi.x * v0.x + vi.y * v0.y, sn0 = vi.y * v0.x - vi.x * v0.y;
      double sa = ys.x * sn0 - ys.y * cs0, sb = ys.x * cs0 + ys.y * sn0;
      double* rp = ws + static_cast<size_t>(2 * i) * ld;
      double* rq = rp + ld;
      for (int t = h.nbr_ptr[i], te = h.nbr_ptr[i + 1]; t < te; ++t) {
        const int j = h.nbr_idx[t];
        const double2 y = h.nbr_y[t], vj = s.node(j)[A_EF];
        const double cc = vi.x * vj.x + vi.y * vj.y, ss = vi.y * vj.x - vi.x * vj.y;
        const double a = y.x * ss - y.y * cc, b = y.x * cc + y.y * ss;
        sa += a; sb += b;
        rp[2 * j] = a; rp[2 * j + 1] = b; rq[2 * j] = -b; rq[2 * j + 1] = a;
      }
      const double vv = vi.x * vi.x + vi.y * vi.y;
      const double gv = yi.x * vv, bv = yi.y * vv;
      const double P = gv + sb, Q = sa - bv;
      const double Fp = P - sp.x, Fq = Q - sp.y;
      rp[2 * i] = -Q - bv; rp[2 * i + 1] = P + gv; rq[2 * i] = P - gv; rq[2 * i + 1] = Q - bv;
      rp[m] = -Fp; rq[m] = -Fq;
      nrm = nanmax(nrm, nanmax(fabs(Fp), fabs(Fq)));
    }
    const bool ok = __all_sync(kFull, nrm < p.tol);
    if (!done && ok) { done = true; iters = it; }
    if (done || it >= p.max_iter) break;
    ++it;
    __syncwarp();
    // LU with partial pivoting, right-looking, rhs carried as column m
    bool singular = false;
    for (int k = 0; k < m; ++k) {
      double best = -1.0; int brow = k;
      for (int r = k + lane; r < m; r += 32) { const double v = fabs(ws[static_cast<size_t>(r) * ld + k]); if (v > best) { best = v; brow = r; } }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const double ob = __shfl_xor_sync(kFull, best, o); const int orow = __shfl_xor_sync(kFull, brow, o);
        if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
      }
      if (!(best > 0.0)) { singular = true; break; }
      if (brow != k)
        for (int c = k + lane; c <= m; c += 32) {
          const double t0 = ws[static_cast<size_t>(k) * ld + c];
          ws[static_cast<size_t>(k) * ld + c] = ws[static_cast<size_t>(brow) * ld + c];
          ws[static_cast<size_t>(brow) * ld + c] = t0;
        }
      __syncwarp();
      const double pinv = 1.0 / ws[static_cast<size_t>(k) * ld + k];
      for (int r = k + 1; r < m; ++r) {
        const double l = ws[static_cast<size_t>(r) * ld + k] * pinv;     // same value in every lane (broadcast load)
        if (l != 0.0)
          for (int c = k + 1 + lane; c <= m; c += 32) ws[static_cast<size_t>(r) * ld + c] -= l * ws[static_cast<size_t>(k) * ld + c];
      }
      __syncwarp();
    }
    if (singular) { done = false; break; }
    // back substitution; dx of bus i goes to its record (R)
    for (int k = m - 1; k >= 0; --k) {
      double acc = 0.0;
      for (int c = k + 1 + lane; c < m; c += 32) acc += ws[static_cast<size_t>(k) * ld + c] * ws[static_cast<size_t>(c) * ld + m];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
      if (lane == 0) ws[static_cast<size_t>(k) * ld + m] = (ws[static_cast<size_t>(k) * ld + m] - acc) / ws[static_cast<size_t>(k) * ld + k];
      __syncwarp();
    }
    for (int i = lane; i < npq; i += 32) {
      double2* nd = s.node(i);
      double2 v = nd[A_VV];
      v.y += ws[static_cast<size_t>(2 * i) * ld + m];
      v.x += v.x * ws[static_cast<size_t>(2 * i + 1) * ld + m];
      double sn, cs;
      sincos_angle(v.y, &sn, &cs);
      nd[A_VV] = v;
      nd[A_EF] = make_double2(v.x * cs, v.x * sn);
    }
    __syncwarp();
  }
  return done;
}

// sgen.q_mvar from an action: reference _clip_reactive_power :568-572
__device__ __forceinline__ double clip_q(double a, double pv, double smax) {
  return sqrt(smax * smax - pv * pv) * a;
}

template <int G, int MODE, bool DENSE = false>
__global__ void __launch_bounds__(G > 32 ? 640 : 384) env_kernel(const __grid_constant__ Params p) {
  static_assert(!DENSE || G == 32, "the dense fallback uses one warp per env");
  static_assert(G == 4 || G == 8 || G == 16 || G == 32 || G == 64 || G == 128, "unsupported group size");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t stage_bar;
  PROF_DECL
  stage_hot_issue(smem_raw, p.hot, p.hot_layout.bytes, &stage_bar);
  bool hot_ready = false;     // the wait is deferred until the first use, behind the prologue's global loads

  const HotLayout& hl = p.hot_layout;
  Hot h;
  h.yup = reinterpret_cast<const double2*>(smem_raw + hl.yup);
  h.ydn = reinterpret_cast<const double2*>(smem_raw + hl.ydn);
  h.yii = reinterpret_cast<const double2*>(smem_raw + hl.yii);
  h.ysl = reinterpret_cast<const double2*>(smem_raw + hl.ysl);
  h.ndesc = reinterpret_cast<const uint64_t*>(smem_raw + hl.ndesc);
  h.esched = reinterpret_cast<const uint64_t*>(smem_raw + hl.esched);
  h.bsched = reinterpret_cast<const uint64_t*>(smem_raw + hl.bsched);
  h.lptr = reinterpret_cast<const uint16_t*>(smem_raw + hl.lptr);
  h.lidx = reinterpret_cast<const uint16_t*>(smem_raw + hl.lidx);
  h.sptr = reinterpret_cast<const uint16_t*>(smem_raw + hl.sptr);
  h.sidx = reinterpret_cast<const uint16_t*>(smem_raw + hl.sidx);
  h.xptr = reinterpret_cast<const uint16_t*>(smem_raw + hl.xptr);
  h.xidx = reinterpret_cast<const uint16_t*>(smem_raw + hl.xidx);
  h.node_of_bus = reinterpret_cast<const uint16_t*>(smem_raw + hl.node_of_bus);
  h.obs_off = reinterpret_cast<const uint16_t*>(smem_raw + hl.obs_off);
  h.line_nodes = reinterpret_cast<const uint16_t*>(smem_raw + hl.line_nodes);
  h.line_c = reinterpret_cast<const double*>(smem_raw + hl.line_c);
  h.nbr_ptr = reinterpret_cast<const uint16_t*>(smem_raw + hl.nbr_ptr);
  h.nbr_idx = reinterpret_cast<const uint16_t*>(smem_raw + hl.nbr_idx);
  h.nbr_y = reinterpret_cast<const double2*>(smem_raw + hl.nbr_y);

  // MODE_STEP: the last warp of the CTA is a helper that draws the next profile rows (+ noise) of the CTA's
  // envs while the solver warps run the Newton iteration (the two only meet at two named barriers)
  const int n_helper_threads = (MODE == MODE_STEP) ? p.helper_threads : 0;
  const int n_solver_threads = blockDim.x - n_helper_threads;
  const bool is_helper = (MODE == MODE_STEP) && threadIdx.x >= n_solver_threads;
  const int hl_bytes = p.hot_layout.bytes;
  const int gl = threadIdx.x % G;
  const int gidx = is_helper ? 0 : threadIdx.x / G;
  const int epb = n_solver_threads / G;
  const int npq = p.npq, n = p.n_bus, nl = p.n_load, ng = p.n_sgen;
  Slab s;
  {
    double2* base = reinterpret_cast<double2*>(smem_raw + hl.bytes) + static_cast<size_t>(gidx) * p.env_stride2;
    s.nodes = base;
    s.base = reinterpret_cast<double*>(base);
    s.pv = reinterpret_cast<double*>(base + p.pvq_off2);
    s.q = s.pv + ng;
    s.scratch = reinterpret_cast<double*>(base + p.scratch_off2);
  }
  double* stage_pl = s.scratch;          // prologue: scaled load p / q
  double* stage_ql = s.scratch + nl;
  double* next_row = s.scratch;          // after the prologue: the helper warp's new sgen.p_mw [n_sgen]
  const uint32_t k0 = static_cast<uint32_t>(p.seed), k1 = static_cast<uint32_t>(p.seed >> 32);

  for (int base = blockIdx.x * epb; base < p.nb; base += gridDim.x * epb) {
    if (is_helper) {
      // ---- helper warp: next profile row of every env of this round (reference _set_demand_and_pv :491-513:
      //      t = self.steps before the increment) + |N(0,1)| * std noise in Box-Muller pairs (2m, 2m+1) over the
      //      elements [pv | load_p | load_q]; new pv goes to the env's scratch (for the obs), loads straight to HBM ----
#ifdef MAPDN_EXP_HELPER_V2
      // experiment: software-pipelined rounds. The per-env scalars of round k+2 and the
      // profile values of round k+1 are in flight while round k's Box-Muller pair is computed; nothing read here is
      // written by the solvers before barrier 2, so the first loads are issued ahead of barrier 1 (only the stores
      // to cur_* have to wait for it).
      {
        const int n_elem = ng + 2 * nl, n_pair = (n_elem + 1) / 2, n_work = epb * n_pair;
        const int hl = threadIdx.x - n_solver_threads, T = n_helper_threads;
        struct HS { int e, m, env, steps; long long nrow; uint32_t ep; bool ok; };
        struct HR { double v[2], sd[2]; };
        auto load_scalars = [&](int w) {
          HS a; a.ok = w < n_work; a.e = 0; a.m = 0; a.env = 0; a.steps = 0; a.nrow = 0; a.ep = 0u;
          if (a.ok) { a.e = w / n_pair; a.m = w - a.e * n_pair; a.env = base + a.e; a.ok = a.env < p.nb; }
          if (a.ok) {
            a.steps = p.steps[a.env];
            a.nrow = p.start_row[a.env] + a.steps;
            if (a.nrow > p.n_rows - 1) a.nrow = p.n_rows - 1;
            a.ep = p.episode[a.env];
          }
          return a;
        };
        auto load_rows = [&](const HS& a) {
          HR r; r.v[0] = r.v[1] = r.sd[0] = r.sd[1] = 0.0;
          if (a.ok) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int el = 2 * a.m + u;
              if (el >= n_elem) break;
              if (el < ng) { r.v[u] = __ldg(p.prof_pv + a.nrow * ng + el); r.sd[u] = __ldg(p.pv_std + el); }
              else if (el < ng + nl) { const int l = el - ng; r.v[u] = __ldg(p.prof_lp + a.nrow * nl + l); r.sd[u] = __ldg(p.lp_std + l); }
              else { const int l = el - ng - nl; r.v[u] = __ldg(p.prof_lq + a.nrow * nl + l); r.sd[u] = __ldg(p.lq_std + l); }
            }
          }
          return r;
        };
        HS s0 = load_scalars(hl), s1 = load_scalars(hl + T);
        HR r0 = load_rows(s0);
        named_bar_sync(1, blockDim.x);                 // the solvers have read the current rows
        for (int w = hl; w < n_work; w += T) {
          const HS s2 = load_scalars(w + 2 * T);
          const HR r1 = load_rows(s1);
          if (s0.ok) {
            RngKey key_h{k0, k1, static_cast<uint32_t>(p.env_id_offset + s0.env), s0.ep * 8u};
            double z[2] = {0.0, 0.0};
            if (p.add_noise) half_normal_pair(key_h, static_cast<uint32_t>(s0.steps), s0.m, z[0], z[1]);
            double* pv_next = reinterpret_cast<double*>(reinterpret_cast<double2*>(smem_raw + hl_bytes) +
                                                         static_cast<size_t>(s0.e) * p.env_stride2 + p.scratch_off2);
            const size_t hL = static_cast<size_t>(s0.env) * nl, hG = static_cast<size_t>(s0.env) * ng;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int el = 2 * s0.m + u;
              if (el >= n_elem) break;
              const double val = r0.v[u] + r0.sd[u] * z[u];
              if (el < ng) { pv_next[el] = val; p.cur_pv[hG + el] = val; }
              else if (el < ng + nl) p.cur_pl[hL + (el - ng)] = val;
              else p.cur_ql[hL + (el - ng - nl)] = val;
            }
          }
          s0 = s1; s1 = s2; r0 = r1;
        }
      }
      named_bar_arrive(2, blockDim.x);                 // new pv of every env is in shared memory
      continue;
#endif
      named_bar_sync(1, blockDim.x);                   // the solvers have read the current rows
      const int n_elem = ng + 2 * nl, n_pair = (n_elem + 1) / 2;
      const int hl = threadIdx.x - n_solver_threads;
      for (int w = hl; w < epb * n_pair; w += n_helper_threads) {
        const int e = w / n_pair, m = w - e * n_pair, env_h = base + e;
        if (env_h >= p.nb) continue;
        const int steps_h = p.steps[env_h];
        long long nrow = p.start_row[env_h] + steps_h;
        if (nrow > p.n_rows - 1) nrow = p.n_rows - 1;
        RngKey key_h{k0, k1, static_cast<uint32_t>(p.env_id_offset + env_h), p.episode[env_h] * 8u};
        double z[2] = {0.0, 0.0};
        if (p.add_noise) half_normal_pair(key_h, static_cast<uint32_t>(steps_h), m, z[0], z[1]);
        double* pv_next = reinterpret_cast<double*>(reinterpret_cast<double2*>(smem_raw + hl_bytes) +
                                                     static_cast<size_t>(e) * p.env_stride2 + p.scratch_off2);
        const size_t hL = static_cast<size_t>(env_h) * nl, hG = static_cast<size_t>(env_h) * ng;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int el = 2 * m + u;
          if (el >= n_elem) break;
          if (el < ng) {
            const double pv = __ldg(p.prof_pv + nrow * ng + el) + __ldg(p.pv_std + el) * z[u];
            pv_next[el] = pv;
            p.cur_pv[hG + el] = pv;
          } else if (el < ng + nl) {
            const int l = el - ng;
            p.cur_pl[hL + l] = __ldg(p.prof_lp + nrow * nl + l) + __ldg(p.lp_std + l) * z[u];
          } else {
            const int l = el - ng - nl;
            p.cur_ql[hL + l] = __ldg(p.prof_lq + nrow * nl + l) + __ldg(p.lq_std + l) * z[u];
          }
        }
      }
      named_bar_arrive(2, blockDim.x);                 // new pv of every env is in shared memory
      continue;
    }
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
#pragma unroll 4
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
      if (!hot_ready) { stage_hot_wait(&stage_bar); hot_ready = true; }
      grp_sync<G>(gidx);
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
        s.node(i)[A_SP] = make_double2(-pd * p.inv_base, -qd * p.inv_base);
      }
      grp_sync<G>(gidx);
      if (MODE == MODE_STEP) named_bar_arrive(1, blockDim.x);   // the current rows are consumed: the helper may overwrite them

      // ---------------- Newton-Raphson ----------------
      PROF(1)
      if (DENSE) {
        double* ws = p.dense_ws + (static_cast<size_t>(blockIdx.x) * epb + gidx) * p.dense_stride;
        conv = nr_solve_dense(p, h, s, ws, gl, !valid, iters);
      } else {
#ifdef MAPDN_PROFILE
        conv = nr_solve<G>(p, h, s, gidx, gl, !valid, iters, _acc, _pt);
#else
        conv = nr_solve<G>(p, h, s, gidx, gl, !valid, iters);
#endif
      }
      PROF(4)
      if (MODE != MODE_RESET) break;
      if (conv) solved = true;
      if (grp_exit<G>(solved)) break;
      if (!solved) ++attempt;                       // re-draw this env (reference retry loop :108-133)
      grp_sync<G>(gidx);
    }

    // ---------------- epilogue ----------------
    // divergence branch (reference :188-196): fall back to the previous solution kept in HBM
    if (MODE == MODE_STEP && !conv) {
      for (int i = gl; i < npq; i += G) {
        const int b = __ldg(p.bus_of_node + i);
        const double vm = p.res_vm[eN + b], va = p.res_va[eN + b];
        double sn, cs;
        sincos(va, &sn, &cs);
        double2* nd = s.node(i);
        nd[A_VV] = make_double2(vm, va);
        nd[A_EF] = make_double2(vm * cs, vm * sn);
        nd[A_SP] = make_double2(-p.res_p[eN + b] * p.inv_base, -p.res_q[eN + b] * p.inv_base);
      }
    }
    grp_sync<G>(gidx);
    const bool write_res = valid && (MODE != MODE_STEP || conv);

    // slack injection (pfsoln): S0 = V0 conj(Ybus[0,:] V) -> sentinel record's SP
    {
      const double vv = p.vm0 * p.vm0;
      double P0 = p.ysl_g0 * vv, Q0 = -p.ysl_b0 * vv;
      for (int k = 0; k < p.n_slack_adj; ++k) {
        const int i = __ldg(p.sl_node + k);
        const double g = __ldg(p.sl_y + 2 * k), b = __ldg(p.sl_y + 2 * k + 1);
        const double2 vi = s.node(i)[A_EF];
        const double cc = p.e0 * vi.x + p.f0 * vi.y, ss = p.f0 * vi.x - p.e0 * vi.y;   // V0 Vi cos/sin(t0 - ti)
        P0 += g * cc + b * ss;
        Q0 += g * ss - b * cc;
      }
      if (gl == 0) s.node(npq)[A_SP] = make_double2(P0, Q0);
    }
    PROF(7)
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
    // the helper warp has drawn the next profile row meanwhile: take over the new sgen.p_mw (Appendix B.3: the
    // observation mixes this solve's bus results with the NEXT row's PV)
    if (MODE == MODE_STEP) {
      named_bar_sync(2, blockDim.x);
      for (int j = gl; j < ng; j += G) s.pv[j] = next_row[j];
    }
    grp_sync<G>(gidx);
    PROF(8)
    // res_bus columns per node (BP) and the "demand" columns of get_obs (OP = BP + sgens of the bus's own zone)
    for (int i = gl; i <= npq; i += G) {
      double2* nd = s.node(i);
      const double2 sp = nd[A_SP];
      const double2 bp = make_double2(-sp.x * p.base_mva, -sp.y * p.base_mva);
      double2 op = bp;
#pragma unroll 1
      for (int t = h.xptr[i], te = h.xptr[i + 1]; t < te; ++t) { const int g = h.xidx[t]; op.x += s.pv[g]; op.y += s.q[g]; }
      nd[A_BP] = bp; nd[A_OP] = op;
    }
    grp_sync<G>(gidx);
    // per-bus results + voltage statistics (reference _calc_reward :584-596, :610)
    double cnt_lo = 0, cnt_hi = 0, sum_dev = 0, sum_v = 0, max_drop = 0, max_rise = 0, sum_bar = 0;
    const double v_ref = 0.5 * (p.v_lower + p.v_upper);
#pragma unroll 4
    for (int b = gl; b < n; b += G) {
      const double2* nd = s.node(h.node_of_bus[b]);
      const double2 vv = nd[A_VV], bp = nd[A_BP];
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
#pragma unroll 4
      for (int k = gl; k < p.n_line; k += G) {
        const double2 vf = s.node(h.line_nodes[2 * k])[A_EF], vt = s.node(h.line_nodes[2 * k + 1])[A_EF];
        const double cc = vf.x * vt.x + vf.y * vt.y, ss = vf.y * vt.x - vf.x * vt.y;
        const double* c = h.line_c + 4 * k;
        const double pl = c[0] * (vf.x * vf.x + vf.y * vf.y) + c[1] * (vt.x * vt.x + vt.y * vt.y) + c[2] * cc + c[3] * ss;
        sum_pl += pl;
        if (MODE == MODE_SOLVE) { if (valid && p.out_pl) p.out_pl[ePL + k] = pl; }
        else if (write_res) p.res_pl[ePL + k] = pl;
      }
    }
    PROF(9)
    if (MODE == MODE_SOLVE) {
      if (valid && gl == 0) {
        if (p.out_iters) p.out_iters[env] = conv ? iters : p.max_iter;
        if (p.out_conv) p.out_conv[env] = conv ? 1 : 0;
      }
      grp_sync<G>(gidx);
      continue;
    }

    if (MODE == MODE_STEP) {
      {
        double rv[10] = {cnt_lo, cnt_hi, sum_dev, sum_v, sum_bar, max_drop, max_rise, sum_pl, sum_q_eff, sum_q_try};
        const bool rmax[10] = {false, false, false, false, false, true, true, false, false, false};
        grp_reduce<G, 10>(gidx, gl, rv, rmax, s.scratch);
        cnt_lo = rv[0]; cnt_hi = rv[1]; sum_dev = rv[2]; sum_v = rv[3]; sum_bar = rv[4];
        max_drop = rv[5]; max_rise = rv[6]; sum_pl = rv[7]; sum_q_eff = rv[8]; sum_q_try = rv[9];
      }
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
    PROF(10)
    // observations of the new state (reference get_obs :232-316): a pure gather - the program maps
    // every entry to a double inside this env's slab (or to a constant-zero slot for the padding)
    if (p.obs != nullptr && valid) {
      double* o = p.obs + static_cast<size_t>(env) * ng * p.obs_dim;
      const int tot = ng * p.obs_dim;
#pragma unroll 4
      for (int idx = gl; idx < tot; idx += G) o[idx] = s.base[h.obs_off[idx]];
    } else if (MODE == MODE_STEP && p.obs32 != nullptr && valid) {
      float* o = p.obs32 + static_cast<size_t>(env) * ng * p.obs_dim;
      const int tot = ng * p.obs_dim;
#pragma unroll 4
      for (int idx = gl; idx < tot; idx += G) o[idx] = static_cast<float>(s.base[h.obs_off[idx]]);
    }
    if (MODE == MODE_RESET && p.state != nullptr) {
      // get_state (:213-230): [P_bus | Q_bus | pv | q | vm | va(deg)] restricted to state_space (cold program)
      double* o = p.state + static_cast<size_t>(env) * p.state_dim;
      for (int idx = gl; idx < p.state_dim; idx += G) {
        const unsigned src = __ldg(p.state_src + idx);
        const int kind = static_cast<int>(src >> 28), ix = static_cast<int>(src & 0x0FFFFFFFu);
        double v = 0.0;
        if (kind == OBS_PBUS) v = s.node(h.node_of_bus[ix])[A_BP].x;
        else if (kind == OBS_QBUS) v = s.node(h.node_of_bus[ix])[A_BP].y;
        else if (kind == OBS_PV) v = s.pv[ix];
        else if (kind == OBS_QSG) v = s.q[ix];
        else if (kind == OBS_VM) v = s.node(h.node_of_bus[ix])[A_VV].x;
        else if (kind == OBS_VA_DEG) v = s.node(h.node_of_bus[ix])[A_VV].y * kRad2Deg;
        if (valid) o[idx] = v;
      }
    }
    grp_sync<G>(gidx);
    PROF(11)
  }
#ifdef MAPDN_PROFILE
  if (p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) for (int k = 0; k < 16; ++k) p.prof[k] = _acc[k];
#endif
  if (!hot_ready) stage_hot_wait(&stage_bar);   // never exit with the bulk copy still in flight
}

}  // namespace mapdn
