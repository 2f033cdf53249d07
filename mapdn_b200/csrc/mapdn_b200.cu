// C-ABI of the batched MAPDN voltage-control env (include/mapdn_b200.h): handle management,
// one-off Ybus assembly + symbolic analysis of the radial network, kernel launches.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mapdn_b200.h"
#include "env_kernel.cuh"

namespace mapdn {

static thread_local std::string g_last_error;

static mapdn_status fail(mapdn_status st, const std::string& msg) {
  g_last_error = msg;
  return st;
}

#define MAPDN_CUDA(expr)                                                                        \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fail(MAPDN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
  } while (0)

// ------------------------------------------------------------------------------------------------
// Ybus assembly on the device (PYPOWER makeYbus). Runs once per handle: the
// topology never changes between env steps (the reference rebuilds Ybus inside every pp.runpp).
// ------------------------------------------------------------------------------------------------
__global__ void ybus_branch_kernel(int n_br, const double* __restrict__ r, const double* __restrict__ x,
                                   const double* __restrict__ b, const double* __restrict__ g,
                                   const double* __restrict__ tap, const double* __restrict__ shift_deg,
                                   const unsigned char* __restrict__ status, double* __restrict__ ybr /*[n_br,8]*/) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_br) return;
  const double st = status[k] ? 1.0 : 0.0;
  const double den = r[k] * r[k] + x[k] * x[k];
  const double ys_g = st * r[k] / den, ys_b = -st * x[k] / den;          // Ys = stat / (r + jx)
  const double ytt_g = ys_g + 0.5 * st * g[k], ytt_b = ys_b + 0.5 * st * b[k];   // Ys + j*Bc/2, Bc = b - jg
  const double t = (tap[k] == 0.0) ? 1.0 : tap[k];
  double sn, cs;
  sincos(shift_deg[k] * (3.14159265358979323846 / 180.0), &sn, &cs);
  const double tr = t * cs, ti = t * sn;                                  // tau = tap * exp(j*shift)
  const double t2 = tr * tr + ti * ti;
  // Yff = Ytt / (tau conj(tau)); Yft = -Ys / conj(tau); Ytf = -Ys / tau
  double* o = ybr + 8 * k;
  o[0] = ytt_g / t2; o[1] = ytt_b / t2;
  // -Ys/conj(tau) = -Ys * tau / |tau|^2
  o[2] = -(ys_g * tr - ys_b * ti) / t2; o[3] = -(ys_g * ti + ys_b * tr) / t2;
  // -Ys/tau = -Ys * conj(tau) / |tau|^2
  o[4] = -(ys_g * tr + ys_b * ti) / t2; o[5] = -(-ys_g * ti + ys_b * tr) / t2;
  o[6] = ytt_g; o[7] = ytt_b;
}

__global__ void ybus_diag_kernel(int n_bus, int n_br, const int* __restrict__ from, const int* __restrict__ to,
                                 const double* __restrict__ ybr, const double* __restrict__ gs,
                                 const double* __restrict__ bs, double inv_base, double* __restrict__ ydiag /*[n_bus,2]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bus) return;
  double g = gs[i] * inv_base, b = bs[i] * inv_base;                      // Ysh = (GS + jBS)/baseMVA
  for (int k = 0; k < n_br; ++k) {
    if (from[k] == i) { g += ybr[8 * k + 0]; b += ybr[8 * k + 1]; }
    if (to[k] == i) { g += ybr[8 * k + 6]; b += ybr[8 * k + 7]; }
  }
  ydiag[2 * i] = g; ydiag[2 * i + 1] = b;
}

// ---- small gather kernels over the env state kept in HBM ----
__global__ void scale_copy_kernel(long long n, const double* __restrict__ in, double scale, double* __restrict__ out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = in[i] * scale;
}
__global__ void int_to_double_kernel(long long n, const int* __restrict__ in, double* __restrict__ out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = static_cast<double>(in[i]);
}
__global__ void i64_to_double_kernel(long long n, const long long* __restrict__ in, double* __restrict__ out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = static_cast<double>(in[i]);
}

// get_obs() from the state in HBM (reference :232-316)
__global__ void get_obs_kernel(const __grid_constant__ Params p) {
  const int tot = p.n_sgen * p.obs_dim;
  const long long gid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (gid >= static_cast<long long>(p.nb) * tot) return;
  const int env = static_cast<int>(gid / tot), idx = static_cast<int>(gid - static_cast<long long>(env) * tot);
  const int a = idx / p.obs_dim, k = idx - a * p.obs_dim;
  const size_t eN = static_cast<size_t>(env) * p.n, eG = static_cast<size_t>(env) * p.n_sgen;
  auto busp = [&](int i) { return p.res_p[eN + __ldg(p.bus_of_node + i)]; };
  auto busq = [&](int i) { return p.res_q[eN + __ldg(p.bus_of_node + i)]; };
  auto busvm = [&](int i) { return p.res_vm[eN + __ldg(p.bus_of_node + i)]; };
  auto busva = [&](int i) { return p.res_va[eN + __ldg(p.bus_of_node + i)]; };
  auto sgp = [&](int j) { return p.cur_pv[eG + j]; };
  auto sgq = [&](int j) { return p.cur_q[eG + j]; };
  p.obs[gid] = obs_entry(p, a, k, busp, busq, busvm, busva, sgp, sgq);
}

// get_state() (reference :213-230): [P_bus | Q_bus | pv | q | vm | va(deg)]
__global__ void get_state_kernel(const __grid_constant__ Params p) {
  const long long gid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (gid >= static_cast<long long>(p.nb) * p.state_dim) return;
  const int env = static_cast<int>(gid / p.state_dim), idx = static_cast<int>(gid - static_cast<long long>(env) * p.state_dim);
  const int n = p.n, ng = p.n_sgen;
  const size_t eN = static_cast<size_t>(env) * n, eG = static_cast<size_t>(env) * ng;
  double v;
  if (idx < n) v = p.res_p[eN + idx];
  else if (idx < 2 * n) v = p.res_q[eN + idx - n];
  else if (idx < 2 * n + ng) v = p.cur_pv[eG + idx - 2 * n];
  else if (idx < 2 * n + 2 * ng) v = p.cur_q[eG + idx - 2 * n - ng];
  else if (idx < 3 * n + 2 * ng) v = p.res_vm[eN + idx - 2 * n - 2 * ng];
  else v = p.res_va[eN + idx - 3 * n - 2 * ng] * 57.295779513082320876798;
  p.state[gid] = v;
}

}  // namespace mapdn

using namespace mapdn;

// ------------------------------------------------------------------------------------------------
struct mapdn_env {
  int device = 0;
  mapdn_cfg cfg{};
  mapdn_dims dims{};
  Params base{};                       // static + state pointers; io fields filled per call
  std::vector<void*> allocs;           // everything cudaMalloc'ed
  int G = 8, threads = 128, epb = 16, smem = 0, max_blocks = 0;
  long long launches = 0;
  // Ybus pieces kept for the test hook
  std::vector<double> ybr, ydiag;
  std::vector<int> br_from, br_to;
  // staging buffers of the *_host entry points
  double *h_actions = nullptr, *h_reward = nullptr, *h_info = nullptr, *h_obs = nullptr;
  unsigned char* h_term = nullptr;
};

namespace {

template <class T>
mapdn_status dev_upload(mapdn_env* e, const std::vector<T>& v, const T** out) {
  void* d = nullptr;
  const size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
  MAPDN_CUDA(cudaMalloc(&d, bytes));
  e->allocs.push_back(d);
  if (!v.empty()) MAPDN_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = static_cast<const T*>(d);
  return MAPDN_OK;
}

template <class T>
mapdn_status dev_alloc(mapdn_env* e, size_t count, T** out) {
  void* d = nullptr;
  const size_t bytes = std::max<size_t>(count * sizeof(T), 16);
  MAPDN_CUDA(cudaMalloc(&d, bytes));
  MAPDN_CUDA(cudaMemset(d, 0, bytes));
  e->allocs.push_back(d);
  *out = static_cast<T*>(d);
  return MAPDN_OK;
}

template <class T>
std::vector<T> vec_or(const T* p, size_t n, T fill) {
  return p ? std::vector<T>(p, p + n) : std::vector<T>(n, fill);
}

using KernelFn = void (*)(const Params);

template <int G>
KernelFn kernel_for_mode(int mode) {
  switch (mode) {
    case MODE_SOLVE: return env_kernel<G, MODE_SOLVE>;
    case MODE_STEP: return env_kernel<G, MODE_STEP>;
    default: return env_kernel<G, MODE_RESET>;
  }
}

KernelFn kernel_for(int G, int mode) {
  switch (G) {
    case 4: return kernel_for_mode<4>(mode);
    case 8: return kernel_for_mode<8>(mode);
    case 16: return kernel_for_mode<16>(mode);
    default: return kernel_for_mode<32>(mode);
  }
}

int env_stride_for(int n, int ng, int G) {
  int stride = kNodeArrays * n + 2 * ng;
  while (stride % 16 != G % 16) ++stride;     // bank-conflict-free interleave of the envs of a warp
  return stride;
}

mapdn_status launch_env_kernel(mapdn_env* e, int mode, Params& p, cudaStream_t st) {
  KernelFn fn = kernel_for(e->G, mode);
  const int needed = (p.nb + e->epb - 1) / e->epb;
  int grid = std::min(needed, std::max(1, e->max_blocks));
  const int rounds = (needed + grid - 1) / grid;
  grid = (needed + rounds - 1) / rounds;       // balance the persistent loop
  fn<<<grid, e->threads, e->smem, st>>>(p);
  MAPDN_CUDA(cudaGetLastError());
  e->launches++;
  return MAPDN_OK;
}

}  // namespace

extern "C" {

int32_t mapdn_abi_version(void) { return MAPDN_ABI_VERSION; }
const char* mapdn_last_error(void) { return g_last_error.c_str(); }

mapdn_status mapdn_destroy(mapdn_env* e) {
  if (!e) return MAPDN_OK;
  cudaSetDevice(e->device);
  for (void* d : e->allocs) cudaFree(d);
  delete e;
  return MAPDN_OK;
}

mapdn_status mapdn_create(const mapdn_net_desc* net, const mapdn_profile_desc* prof, const mapdn_cfg* cfg,
                          int32_t device, mapdn_env** out) {
  if (!net || !cfg || !out) return fail(MAPDN_ERR_INVALID, "null argument");
  *out = nullptr;
  const int n = net->n_bus, nbr = net->n_branch, nl = net->n_load, ng = net->n_sgen;
  if (n < 2 || n > 65000) return fail(MAPDN_ERR_INVALID, "n_bus must be in [2, 65000]");
  if (nbr < 1 || nl < 0 || ng < 1) return fail(MAPDN_ERR_INVALID, "need >=1 branch and >=1 sgen");
  if (cfg->batch < 1) return fail(MAPDN_ERR_INVALID, "batch must be >= 1");
  if (!net->br_from || !net->br_to || !net->br_r || !net->br_x || !net->sgen_bus || !net->sgen_zone ||
      !net->bus_zone || (nl > 0 && !net->load_bus))
    return fail(MAPDN_ERR_INVALID, "missing required network array");
  if (net->slack_bus < 0 || net->slack_bus >= n) return fail(MAPDN_ERR_INVALID, "slack_bus out of range");
  if (!(net->base_mva > 0)) return fail(MAPDN_ERR_INVALID, "base_mva must be positive");
  if (cfg->barrier < 0 || cfg->barrier > 4) return fail(MAPDN_ERR_INVALID, "unknown voltage barrier");
  if (cfg->lanes_per_env != 0 && cfg->lanes_per_env != 4 && cfg->lanes_per_env != 8 &&
      cfg->lanes_per_env != 16 && cfg->lanes_per_env != 32)
    return fail(MAPDN_ERR_INVALID, "lanes_per_env must be 0, 4, 8, 16 or 32");
  for (int k = 0; k < nbr; ++k)
    if (net->br_from[k] < 0 || net->br_from[k] >= n || net->br_to[k] < 0 || net->br_to[k] >= n ||
        net->br_from[k] == net->br_to[k])
      return fail(MAPDN_ERR_INVALID, "branch endpoint out of range");
  for (int l = 0; l < nl; ++l)
    if (net->load_bus[l] < 0 || net->load_bus[l] >= n) return fail(MAPDN_ERR_INVALID, "load_bus out of range");
  for (int j = 0; j < ng; ++j)
    if (net->sgen_bus[j] < 0 || net->sgen_bus[j] >= n) return fail(MAPDN_ERR_INVALID, "sgen_bus out of range");
  if (prof) {
    if (!prof->pv || !prof->load_p || !prof->load_q || !prof->pv_std || !prof->load_p_std ||
        !prof->load_q_std || !prof->s_max || prof->n_rows < 2 || prof->steps_per_hour < 1)
      return fail(MAPDN_ERR_INVALID, "incomplete profile description");
  }
  MAPDN_CUDA(cudaSetDevice(device));
  mapdn_env* e = new (std::nothrow) mapdn_env();
  if (!e) return fail(MAPDN_ERR_NOMEM, "out of host memory");
  e->device = device;
  e->cfg = *cfg;
  mapdn_status st = MAPDN_OK;
  auto bail = [&](mapdn_status s) { mapdn_destroy(e); return s; };
#define TRY(expr) do { st = (expr); if (st != MAPDN_OK) return bail(st); } while (0)
#define TRY_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { \
      fail(MAPDN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); return bail(MAPDN_ERR_CUDA); } } while (0)

  // ---- 1. Ybus on the device ----
  std::vector<double> br_r(net->br_r, net->br_r + nbr), br_x(net->br_x, net->br_x + nbr);
  std::vector<double> br_b = vec_or(net->br_b, nbr, 0.0), br_g = vec_or(net->br_g, nbr, 0.0);
  std::vector<double> br_tap = vec_or(net->br_tap, nbr, 1.0), br_sh = vec_or(net->br_shift_deg, nbr, 0.0);
  std::vector<unsigned char> br_st = vec_or<unsigned char>(net->br_status, nbr, 1);
  std::vector<unsigned char> br_line = vec_or<unsigned char>(net->br_is_line, nbr, 1);
  std::vector<double> gs = vec_or(net->bus_gs_mw, n, 0.0), bs = vec_or(net->bus_bs_mvar, n, 0.0);
  e->br_from.assign(net->br_from, net->br_from + nbr);
  e->br_to.assign(net->br_to, net->br_to + nbr);
  for (int k = 0; k < nbr; ++k)
    if (br_st[k] && br_r[k] == 0.0 && br_x[k] == 0.0) return bail(fail(MAPDN_ERR_INVALID, "zero-impedance branch"));
  {
    const double *d_r, *d_x, *d_b, *d_g, *d_tap, *d_sh, *d_gs, *d_bs;
    const unsigned char* d_st;
    const int *d_f, *d_t;
    double *d_ybr, *d_ydiag;
    TRY(dev_upload(e, br_r, &d_r)); TRY(dev_upload(e, br_x, &d_x)); TRY(dev_upload(e, br_b, &d_b));
    TRY(dev_upload(e, br_g, &d_g)); TRY(dev_upload(e, br_tap, &d_tap)); TRY(dev_upload(e, br_sh, &d_sh));
    TRY(dev_upload(e, br_st, &d_st)); TRY(dev_upload(e, e->br_from, &d_f)); TRY(dev_upload(e, e->br_to, &d_t));
    TRY(dev_upload(e, gs, &d_gs)); TRY(dev_upload(e, bs, &d_bs));
    TRY(dev_alloc(e, static_cast<size_t>(8) * nbr, &d_ybr)); TRY(dev_alloc(e, static_cast<size_t>(2) * n, &d_ydiag));
    ybus_branch_kernel<<<(nbr + 127) / 128, 128>>>(nbr, d_r, d_x, d_b, d_g, d_tap, d_sh, d_st, d_ybr);
    ybus_diag_kernel<<<(n + 127) / 128, 128>>>(n, nbr, d_f, d_t, d_ybr, d_gs, d_bs, 1.0 / net->base_mva, d_ydiag);
    TRY_CUDA(cudaGetLastError());
    e->launches += 2;
    e->ybr.resize(static_cast<size_t>(8) * nbr);
    e->ydiag.resize(static_cast<size_t>(2) * n);
    TRY_CUDA(cudaMemcpy(e->ybr.data(), d_ybr, e->ybr.size() * sizeof(double), cudaMemcpyDeviceToHost));
    TRY_CUDA(cudaMemcpy(e->ydiag.data(), d_ydiag, e->ydiag.size() * sizeof(double), cudaMemcpyDeviceToHost));
  }

  // ---- 2. symbolic analysis: merge parallel branches, BFS tree from the slack bus ----
  struct Pair { double gab = 0, bab = 0, gba = 0, bba = 0; };   // Y[a,b], Y[b,a] with a < b
  std::map<std::pair<int, int>, Pair> pairs;
  for (int k = 0; k < nbr; ++k) {
    if (!br_st[k]) continue;
    const int f = e->br_from[k], t = e->br_to[k];
    const double* y = &e->ybr[8 * static_cast<size_t>(k)];
    Pair& pr = pairs[{std::min(f, t), std::max(f, t)}];
    if (f < t) { pr.gab += y[2]; pr.bab += y[3]; pr.gba += y[4]; pr.bba += y[5]; }   // Ybus[f,t]+=Yft, [t,f]+=Ytf
    else       { pr.gba += y[2]; pr.bba += y[3]; pr.gab += y[4]; pr.bab += y[5]; }
  }
  std::vector<std::vector<int>> adj(n);
  for (auto& kv : pairs) { adj[kv.first.first].push_back(kv.first.second); adj[kv.first.second].push_back(kv.first.first); }
  for (auto& a : adj) std::sort(a.begin(), a.end());
  std::vector<int> order, node_of_bus(n, -1), parent_bus(n, -1), depth(n, 0);
  order.reserve(n);
  order.push_back(net->slack_bus);
  node_of_bus[net->slack_bus] = 0;
  size_t tree_edges = 0;
  for (size_t qh = 0; qh < order.size(); ++qh) {
    const int u = order[qh];
    for (int v : adj[u]) {
      if (node_of_bus[v] >= 0) continue;
      node_of_bus[v] = static_cast<int>(order.size());
      parent_bus[v] = u;
      depth[v] = depth[u] + 1;
      order.push_back(v);
      ++tree_edges;
    }
  }
  if (static_cast<int>(order.size()) != n)
    return bail(fail(MAPDN_ERR_TOPOLOGY, "network is not connected to the slack bus (" +
                                             std::to_string(n - order.size()) + " unreachable buses)"));
  if (pairs.size() != tree_edges)
    return bail(fail(MAPDN_ERR_TOPOLOGY, "network is meshed (" + std::to_string(pairs.size() - tree_edges) +
                                             " loop-closing branches); only radial feeders are supported"));
  std::vector<int> parent(n, 0), nchild(n, 0), height(n, 0);
  for (int i = 1; i < n; ++i) { parent[i] = node_of_bus[parent_bus[order[i]]]; nchild[parent[i]]++; }
  for (int i = n - 1; i >= 1; --i) height[parent[i]] = std::max(height[parent[i]], height[i] + 1);
  std::vector<uint16_t> cstart(n + 1);
  cstart[0] = 1;
  for (int i = 0; i < n; ++i) cstart[i + 1] = static_cast<uint16_t>(cstart[i] + nchild[i]);
  int max_h = 0, max_d = 0;
  for (int i = 1; i < n; ++i) { max_h = std::max(max_h, height[i]); max_d = std::max(max_d, depth[order[i]]); }
  const int n_elev = max_h + 1, n_dlev = max_d + 1;
  std::vector<uint16_t> eorder, elev(n_elev + 1), dlev(n_dlev + 1);
  for (int hgt = 0; hgt <= max_h; ++hgt) {
    elev[hgt] = static_cast<uint16_t>(eorder.size());
    for (int i = 1; i < n; ++i) if (height[i] == hgt) eorder.push_back(static_cast<uint16_t>(i));
  }
  elev[n_elev] = static_cast<uint16_t>(eorder.size());
  {
    int d = 0;
    dlev[0] = 0;
    for (int i = 0; i < n; ++i) while (depth[order[i]] > d) dlev[++d] = static_cast<uint16_t>(i);
    dlev[n_dlev] = static_cast<uint16_t>(n);
  }
  int max_width = 0;
  for (int l = 0; l < n_elev; ++l) max_width = std::max(max_width, elev[l + 1] - elev[l]);

  // ---- 3. hot static blob ----
  HotLayout hl{};
  {
    int off = 0;
    auto take = [&](int bytes) { int o = off; off += (bytes + 15) / 16 * 16; return o; };
    hl.gu = take(8 * n); hl.bu = take(8 * n); hl.gd = take(8 * n); hl.bd = take(8 * n);
    hl.gii = take(8 * n); hl.bii = take(8 * n);
    hl.parent = take(2 * n); hl.cstart = take(2 * (n + 1)); hl.eorder = take(2 * std::max(1, n - 1));
    hl.elev = take(2 * (n_elev + 1)); hl.dlev = take(2 * (n_dlev + 1));
    hl.bytes = off;
  }
  std::vector<unsigned char> hot(hl.bytes, 0);
  {
    double* gu = reinterpret_cast<double*>(hot.data() + hl.gu); double* bu = reinterpret_cast<double*>(hot.data() + hl.bu);
    double* gd = reinterpret_cast<double*>(hot.data() + hl.gd); double* bd = reinterpret_cast<double*>(hot.data() + hl.bd);
    double* gii = reinterpret_cast<double*>(hot.data() + hl.gii); double* bii = reinterpret_cast<double*>(hot.data() + hl.bii);
    uint16_t* par = reinterpret_cast<uint16_t*>(hot.data() + hl.parent);
    for (int i = 0; i < n; ++i) {
      const int b = order[i];
      gii[i] = e->ydiag[2 * b]; bii[i] = e->ydiag[2 * b + 1];
      par[i] = static_cast<uint16_t>(parent[i]);
      if (i == 0) continue;
      const int pb = parent_bus[b];
      const Pair& pr = pairs[{std::min(b, pb), std::max(b, pb)}];
      if (b < pb) { gu[i] = pr.gab; bu[i] = pr.bab; gd[i] = pr.gba; bd[i] = pr.bba; }   // Y[i,parent], Y[parent,i]
      else        { gu[i] = pr.gba; bu[i] = pr.bba; gd[i] = pr.gab; bd[i] = pr.bab; }
    }
    std::memcpy(hot.data() + hl.cstart, cstart.data(), 2 * (n + 1));
    std::memcpy(hot.data() + hl.eorder, eorder.data(), 2 * eorder.size());
    std::memcpy(hot.data() + hl.elev, elev.data(), 2 * elev.size());
    std::memcpy(hot.data() + hl.dlev, dlev.data(), 2 * dlev.size());
  }

  // ---- 4. cold tables ----
  std::vector<int> lptr(n + 1, 0), lidx(nl), sptr(n + 1, 0), sidx(ng);
  {
    std::vector<std::vector<int>> ln(n), sn(n);
    for (int l = 0; l < nl; ++l) ln[node_of_bus[net->load_bus[l]]].push_back(l);
    for (int j = 0; j < ng; ++j) sn[node_of_bus[net->sgen_bus[j]]].push_back(j);
    int a = 0, b = 0;
    for (int i = 0; i < n; ++i) {
      lptr[i] = a; sptr[i] = b;
      for (int l : ln[i]) lidx[a++] = l;
      for (int j : sn[i]) sidx[b++] = j;
    }
    lptr[n] = a; sptr[n] = b;
  }
  std::vector<double> lscale = vec_or(net->load_scaling, nl, 1.0), sscale = vec_or(net->sgen_scaling, ng, 1.0);
  std::vector<int> line_f, line_t;
  std::vector<double> line_c;
  for (int k = 0; k < nbr; ++k) {
    if (!br_line[k]) continue;
    const double* y = &e->ybr[8 * static_cast<size_t>(k)];
    line_f.push_back(node_of_bus[e->br_from[k]]);
    line_t.push_back(node_of_bus[e->br_to[k]]);
    // pl = Re(Sf + St) = Gff|Vf|^2 + Gtt|Vt|^2 + (Gft+Gtf) Re(Vf Vt*) + (Bft-Btf) Im(Vf Vt*)   [x baseMVA]
    line_c.push_back(y[0] * net->base_mva); line_c.push_back(y[6] * net->base_mva);
    line_c.push_back((y[2] + y[4]) * net->base_mva); line_c.push_back((y[3] - y[5]) * net->base_mva);
  }
  const int n_line = static_cast<int>(line_f.size());
  std::vector<int> zptr(ng + 1, 0), znode, zsg_ptr, zsg_idx;
  int obs_dim = 0;
  for (int a = 0; a < ng; ++a) {
    zptr[a] = static_cast<int>(znode.size());
    for (int b = 0; b < n; ++b) {
      if (net->bus_zone[b] != net->sgen_zone[a]) continue;
      zsg_ptr.push_back(static_cast<int>(zsg_idx.size()));
      for (int j = 0; j < ng; ++j)
        if (net->sgen_zone[j] == net->sgen_zone[a] && net->sgen_bus[j] == b) zsg_idx.push_back(j);
      znode.push_back(node_of_bus[b]);
    }
    const int nz = static_cast<int>(znode.size()) - zptr[a];
    obs_dim = std::max(obs_dim, 4 * nz + 2);
  }
  zptr[ng] = static_cast<int>(znode.size());
  zsg_ptr.push_back(static_cast<int>(zsg_idx.size()));

  // ---- 5. launch geometry ----
  cudaDeviceProp dp{};
  TRY_CUDA(cudaGetDeviceProperties(&dp, device));
  int G = cfg->lanes_per_env;
  if (G == 0) G = (n <= 96) ? 8 : (n <= 200 ? 16 : 32);
  const int stride = env_stride_for(n, ng, G);
  const size_t max_smem = dp.sharedMemPerBlockOptin;
  int warps = 4;
  auto smem_for = [&](int w) { return static_cast<size_t>(hl.bytes) + static_cast<size_t>(w) * (32 / G) * stride * 8; };
  while (warps > 1 && smem_for(warps) > max_smem) --warps;
  if (smem_for(warps) > max_smem)
    return bail(fail(MAPDN_ERR_UNSUPPORTED, "network too large for the shared-memory resident solver (" +
                                                std::to_string(smem_for(warps)) + " B needed)"));
  e->G = G; e->threads = 32 * warps; e->epb = warps * (32 / G); e->smem = static_cast<int>(smem_for(warps));
  for (int mode = 0; mode < 3; ++mode) {
    KernelFn fn = kernel_for(G, mode);
    TRY_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, e->smem));
  }
  {
    int per_sm = 0;
    TRY_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel_for(G, MODE_STEP), e->threads, e->smem));
    e->max_blocks = std::max(1, per_sm) * dp.multiProcessorCount;
  }

  // ---- 6. upload + env state ----
  Params& P = e->base;
  P.n = n; P.n_pad = n; P.n_load = nl; P.n_sgen = ng; P.n_sgen_pad = ng; P.n_line = n_line;
  P.n_elev = n_elev; P.n_dlev = n_dlev; P.obs_dim = obs_dim; P.state_dim = 4 * n + 2 * ng;
  P.nb = cfg->batch; P.env_stride = stride; P.hot_layout = hl;
  std::vector<int> bus_of_node(order.begin(), order.end());
  TRY(dev_upload(e, hot, &P.hot));
  TRY(dev_upload(e, bus_of_node, &P.bus_of_node)); TRY(dev_upload(e, node_of_bus, &P.node_of_bus));
  TRY(dev_upload(e, lptr, &P.lptr)); TRY(dev_upload(e, lidx, &P.lidx)); TRY(dev_upload(e, lscale, &P.lscale));
  TRY(dev_upload(e, sptr, &P.sptr)); TRY(dev_upload(e, sidx, &P.sidx)); TRY(dev_upload(e, sscale, &P.sscale));
  TRY(dev_upload(e, line_f, &P.line_f)); TRY(dev_upload(e, line_t, &P.line_t)); TRY(dev_upload(e, line_c, &P.line_c));
  TRY(dev_upload(e, zptr, &P.zptr)); TRY(dev_upload(e, znode, &P.znode));
  TRY(dev_upload(e, zsg_ptr, &P.zsg_ptr)); TRY(dev_upload(e, zsg_idx, &P.zsg_idx));
  const size_t B = static_cast<size_t>(cfg->batch);
  if (prof) {
    const size_t T = static_cast<size_t>(prof->n_rows);
    TRY(dev_upload(e, std::vector<double>(prof->s_max, prof->s_max + ng), &P.s_max));
    TRY(dev_upload(e, std::vector<double>(prof->pv_std, prof->pv_std + ng), &P.pv_std));
    TRY(dev_upload(e, std::vector<double>(prof->load_p_std, prof->load_p_std + nl), &P.lp_std));
    TRY(dev_upload(e, std::vector<double>(prof->load_q_std, prof->load_q_std + nl), &P.lq_std));
    double *dpv, *dlp, *dlq;
    TRY(dev_alloc(e, T * ng, &dpv)); TRY(dev_alloc(e, T * nl, &dlp)); TRY(dev_alloc(e, T * nl, &dlq));
    TRY_CUDA(cudaMemcpy(dpv, prof->pv, T * ng * sizeof(double), cudaMemcpyHostToDevice));
    if (nl) {
      TRY_CUDA(cudaMemcpy(dlp, prof->load_p, T * nl * sizeof(double), cudaMemcpyHostToDevice));
      TRY_CUDA(cudaMemcpy(dlq, prof->load_q, T * nl * sizeof(double), cudaMemcpyHostToDevice));
    }
    P.prof_pv = dpv; P.prof_lp = dlp; P.prof_lq = dlq;
    P.n_rows = prof->n_rows; P.steps_per_hour = prof->steps_per_hour;
    const int episode_days = cfg->episode_limit / (24 * prof->steps_per_hour) + 1;     // reference :397
    P.n_day_choices = prof->n_days - episode_days;                                     // reference :398
    if (P.n_day_choices < 1) return bail(fail(MAPDN_ERR_INVALID, "profile store shorter than one episode"));
    // the latest sampled window must fit: start_max + episode_limit rows
    const long long last = (static_cast<long long>(P.n_day_choices - 1) * 24 + 23) * prof->steps_per_hour +
                           (prof->steps_per_hour - 1) + cfg->episode_limit;
    if (last > prof->n_rows - 1) return bail(fail(MAPDN_ERR_INVALID, "profile store shorter than n_days claims"));
  }
  P.base_mva = net->base_mva; P.inv_base = 1.0 / net->base_mva;
  P.vm_init = (net->vm_init > 0) ? net->vm_init : net->slack_vm;
  P.vm0 = net->slack_vm; P.va0 = net->slack_va_deg * (3.14159265358979323846 / 180.0);
  P.e0 = P.vm0 * std::cos(P.va0); P.f0 = P.vm0 * std::sin(P.va0);
  P.tol = (cfg->tol > 0) ? cfg->tol : 1e-8;
  P.max_iter = (cfg->max_iter > 0) ? cfg->max_iter : 10;
  P.barrier = cfg->barrier; P.voltage_weight = cfg->voltage_weight; P.q_weight = cfg->q_weight;
  P.line_weight = cfg->line_weight; P.use_line_weight = cfg->use_line_weight;
  P.v_upper = cfg->v_upper; P.v_lower = cfg->v_lower; P.episode_limit = cfg->episode_limit;
  P.action_low = cfg->action_low; P.action_high = cfg->action_high; P.reset_action = cfg->reset_action;
  P.seed = cfg->seed; P.env_id_offset = cfg->env_id_offset;
  TRY(dev_alloc(e, B * nl, &P.cur_pl)); TRY(dev_alloc(e, B * nl, &P.cur_ql));
  TRY(dev_alloc(e, B * ng, &P.cur_pv)); TRY(dev_alloc(e, B * ng, &P.cur_q));
  TRY(dev_alloc(e, B * n, &P.res_vm)); TRY(dev_alloc(e, B * n, &P.res_va));
  TRY(dev_alloc(e, B * n, &P.res_p)); TRY(dev_alloc(e, B * n, &P.res_q));
  TRY(dev_alloc(e, B * std::max(1, n_line), &P.res_pl));
  TRY(dev_alloc(e, B, &P.steps)); TRY(dev_alloc(e, B, &P.sum_rewards));
  TRY(dev_alloc(e, B, &P.start_row)); TRY(dev_alloc(e, B, &P.episode));
  // staging for the *_host entry points
  TRY(dev_alloc(e, B * ng, &e->h_actions)); TRY(dev_alloc(e, B, &e->h_reward)); TRY(dev_alloc(e, B, &e->h_term));
  TRY(dev_alloc(e, B * MAPDN_N_INFO, &e->h_info)); TRY(dev_alloc(e, B * ng * obs_dim, &e->h_obs));

  mapdn_dims& d = e->dims;
  d.batch = cfg->batch; d.n_bus = n; d.n_branch = nbr; d.n_line = n_line; d.n_load = nl; d.n_sgen = ng;
  d.n_agents = ng; d.n_actions = 1; d.obs_dim = obs_dim; d.state_dim = P.state_dim; d.n_info = MAPDN_N_INFO;
  d.lanes_per_env = G; d.envs_per_block = e->epb; d.smem_bytes = e->smem; d.n_levels = n_elev;
  // read p_load,q_load,p_pv,a ; write vm,va ; write obs ; reward+done+11 info
  d.algorithmic_bytes_per_env_step = 8LL * (2 * nl + 2 * ng) + 8LL * 2 * n + 8LL * ng * obs_dim + 8LL * 13;
  (void)max_width;
  *out = e;
  return MAPDN_OK;
#undef TRY
#undef TRY_CUDA
}

mapdn_status mapdn_get_dims(const mapdn_env* e, mapdn_dims* out) {
  if (!e || !out) return fail(MAPDN_ERR_INVALID, "null argument");
  *out = e->dims;
  return MAPDN_OK;
}

int64_t mapdn_launch_count(const mapdn_env* e) { return e ? e->launches : 0; }

mapdn_status mapdn_reset(mapdn_env* e, const int32_t* start_dhi_dev, const uint8_t* mask_dev, int32_t add_noise,
                         double* obs_dev, double* state_dev, void* stream) {
  if (!e) return fail(MAPDN_ERR_INVALID, "null handle");
  if (!e->base.prof_pv) return fail(MAPDN_ERR_INVALID, "handle was created without a profile store");
  MAPDN_CUDA(cudaSetDevice(e->device));
  Params p = e->base;
  p.start_dhi = start_dhi_dev; p.mask = mask_dev; p.add_noise = add_noise; p.obs = obs_dev; p.state = state_dev;
  return launch_env_kernel(e, MODE_RESET, p, static_cast<cudaStream_t>(stream));
}

mapdn_status mapdn_step(mapdn_env* e, const double* actions_dev, int32_t add_noise, double* reward_dev,
                        uint8_t* terminated_dev, double* info_dev, double* obs_dev, void* stream) {
  if (!e || !actions_dev || !reward_dev || !terminated_dev) return fail(MAPDN_ERR_INVALID, "null argument");
  if (!e->base.prof_pv) return fail(MAPDN_ERR_INVALID, "handle was created without a profile store");
  MAPDN_CUDA(cudaSetDevice(e->device));
  Params p = e->base;
  p.actions = actions_dev; p.add_noise = add_noise; p.reward = reward_dev; p.term = terminated_dev;
  p.info = info_dev; p.obs = obs_dev;
  return launch_env_kernel(e, MODE_STEP, p, static_cast<cudaStream_t>(stream));
}

mapdn_status mapdn_step_host(mapdn_env* e, const double* actions_host, int32_t add_noise, double* reward_host,
                             uint8_t* terminated_host, double* info_host, double* obs_host, void* stream) {
  if (!e || !actions_host || !reward_host || !terminated_host) return fail(MAPDN_ERR_INVALID, "null argument");
  MAPDN_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t B = e->dims.batch, ng = e->dims.n_sgen, od = e->dims.obs_dim;
  MAPDN_CUDA(cudaMemcpyAsync(e->h_actions, actions_host, B * ng * sizeof(double), cudaMemcpyHostToDevice, st));
  mapdn_status s = mapdn_step(e, e->h_actions, add_noise, e->h_reward, e->h_term, info_host ? e->h_info : nullptr,
                              obs_host ? e->h_obs : nullptr, stream);
  if (s != MAPDN_OK) return s;
  MAPDN_CUDA(cudaMemcpyAsync(reward_host, e->h_reward, B * sizeof(double), cudaMemcpyDeviceToHost, st));
  MAPDN_CUDA(cudaMemcpyAsync(terminated_host, e->h_term, B, cudaMemcpyDeviceToHost, st));
  if (info_host) MAPDN_CUDA(cudaMemcpyAsync(info_host, e->h_info, B * MAPDN_N_INFO * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (obs_host) MAPDN_CUDA(cudaMemcpyAsync(obs_host, e->h_obs, B * ng * od * sizeof(double), cudaMemcpyDeviceToHost, st));
  MAPDN_CUDA(cudaStreamSynchronize(st));
  return MAPDN_OK;
}

mapdn_status mapdn_get_obs(mapdn_env* e, double* obs_dev, void* stream) {
  if (!e || !obs_dev) return fail(MAPDN_ERR_INVALID, "null argument");
  MAPDN_CUDA(cudaSetDevice(e->device));
  Params p = e->base;
  p.obs = obs_dev;
  const long long tot = static_cast<long long>(p.nb) * p.n_sgen * p.obs_dim;
  get_obs_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MAPDN_CUDA(cudaGetLastError());
  e->launches++;
  return MAPDN_OK;
}

mapdn_status mapdn_get_state(mapdn_env* e, double* state_dev, void* stream) {
  if (!e || !state_dev) return fail(MAPDN_ERR_INVALID, "null argument");
  MAPDN_CUDA(cudaSetDevice(e->device));
  Params p = e->base;
  p.state = state_dev;
  const long long tot = static_cast<long long>(p.nb) * p.state_dim;
  get_state_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MAPDN_CUDA(cudaGetLastError());
  e->launches++;
  return MAPDN_OK;
}

mapdn_status mapdn_get_field(mapdn_env* e, int32_t field, double* out_dev, void* stream) {
  if (!e || !out_dev) return fail(MAPDN_ERR_INVALID, "null argument");
  MAPDN_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Params& p = e->base;
  const long long B = p.nb;
  const double* src = nullptr;
  long long cnt = 0;
  double scale = 1.0;
  switch (field) {
    case MAPDN_FIELD_VM: src = p.res_vm; cnt = B * p.n; break;
    case MAPDN_FIELD_VA_DEG: src = p.res_va; cnt = B * p.n; scale = 57.295779513082320876798; break;
    case MAPDN_FIELD_P_BUS: src = p.res_p; cnt = B * p.n; break;
    case MAPDN_FIELD_Q_BUS: src = p.res_q; cnt = B * p.n; break;
    case MAPDN_FIELD_P_SGEN: src = p.cur_pv; cnt = B * p.n_sgen; break;
    case MAPDN_FIELD_Q_SGEN: src = p.cur_q; cnt = B * p.n_sgen; break;
    case MAPDN_FIELD_LINE_LOSS: src = p.res_pl; cnt = B * p.n_line; break;
    case MAPDN_FIELD_P_LOAD: src = p.cur_pl; cnt = B * p.n_load; break;
    case MAPDN_FIELD_Q_LOAD: src = p.cur_ql; cnt = B * p.n_load; break;
    case MAPDN_FIELD_SUM_REWARDS: src = p.sum_rewards; cnt = B; break;
    case MAPDN_FIELD_STEPS:
      int_to_double_kernel<<<static_cast<unsigned>((B + 255) / 256), 256, 0, st>>>(B, p.steps, out_dev);
      MAPDN_CUDA(cudaGetLastError()); e->launches++;
      return MAPDN_OK;
    case MAPDN_FIELD_START_ROW:
      i64_to_double_kernel<<<static_cast<unsigned>((B + 255) / 256), 256, 0, st>>>(B, p.start_row, out_dev);
      MAPDN_CUDA(cudaGetLastError()); e->launches++;
      return MAPDN_OK;
    default: return fail(MAPDN_ERR_INVALID, "unknown field");
  }
  if (cnt == 0) return MAPDN_OK;
  scale_copy_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, st>>>(cnt, src, scale, out_dev);
  MAPDN_CUDA(cudaGetLastError());
  e->launches++;
  return MAPDN_OK;
}

mapdn_status mapdn_solve(mapdn_env* e, int32_t nb, const double* p_load, const double* q_load, const double* p_sgen,
                         const double* q_sgen, double* vm, double* va_deg, double* p_bus, double* q_bus, double* pl,
                         int32_t* iters, uint8_t* converged, void* stream) {
  if (!e || nb < 1 || !p_sgen || !q_sgen || (e->dims.n_load > 0 && (!p_load || !q_load)))
    return fail(MAPDN_ERR_INVALID, "null argument");
  MAPDN_CUDA(cudaSetDevice(e->device));
  Params p = e->base;
  p.nb = nb;
  p.in_pl = p_load; p.in_ql = q_load; p.in_pv = p_sgen; p.in_q = q_sgen;
  p.out_vm = vm; p.out_va = va_deg; p.out_p = p_bus; p.out_q = q_bus; p.out_pl = pl;
  p.out_iters = iters; p.out_conv = converged;
  return launch_env_kernel(e, MODE_SOLVE, p, static_cast<cudaStream_t>(stream));
}

mapdn_status mapdn_get_ybus_dense(mapdn_env* e, double* g_host, double* b_host) {
  if (!e || !g_host || !b_host) return fail(MAPDN_ERR_INVALID, "null argument");
  const int n = e->dims.n_bus;
  std::fill(g_host, g_host + static_cast<size_t>(n) * n, 0.0);
  std::fill(b_host, b_host + static_cast<size_t>(n) * n, 0.0);
  for (int i = 0; i < n; ++i) { g_host[static_cast<size_t>(i) * n + i] = e->ydiag[2 * i]; b_host[static_cast<size_t>(i) * n + i] = e->ydiag[2 * i + 1]; }
  for (size_t k = 0; k < e->br_from.size(); ++k) {
    const size_t f = e->br_from[k], t = e->br_to[k];
    g_host[f * n + t] += e->ybr[8 * k + 2]; b_host[f * n + t] += e->ybr[8 * k + 3];
    g_host[t * n + f] += e->ybr[8 * k + 4]; b_host[t * n + f] += e->ybr[8 * k + 5];
  }
  return MAPDN_OK;
}

}  // extern "C"
