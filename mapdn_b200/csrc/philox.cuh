// Counter-based RNG of the batched env: Philox4x32-10 (Salmon et al., SC'11), keyed by
// (seed, global env id, episode, step, element). Replaces the process-global np.random stream
// of the reference (voltage_control_env.py:49,384,389,398,498,503,508,337) - same
// distributions, reproducible per env and independent of the multi-GPU sharding.
// NumPy mirror used by the parity tests: oracle/philox_ref.py.
#pragma once
#include <stdint.h>

namespace mapdn {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kStreamNoise = 0, kStreamTime = 1, kStreamAction = 2;
constexpr uint32_t kResetFlag = 0x80000000u;

struct u32x4 { uint32_t x, y, z, w; };

__device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
    const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += kPhiloxW0; k1 += kPhiloxW1;
  }
  return {c0, c1, c2, c3};
}

// (0,1) double from two words: ((hi<<32|lo)>>11 + 0.5) * 2^-53
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  const uint64_t bits = ((static_cast<uint64_t>(hi) << 32) | lo) >> 11;
  return (static_cast<double>(bits) + 0.5) * (1.0 / 9007199254740992.0);
}

struct RngKey { uint32_t k0, k1, env, c3base; };   // c3base = episode * 8

// |N(0,1)| noise of the profile rows (reference: np.abs(np.random.randn()), :498,503,508).
// One Philox call yields the Box-Muller pair of elements (2m, 2m+1): |r cos(2 pi u2)|, |r sin(2 pi u2)|,
// drawn at counter c1 (= steps, or kResetFlag|attempt).
__device__ __forceinline__ void half_normal_pair(const RngKey& k, uint32_t c1, uint32_t pair, double& z0, double& z1) {
  const u32x4 r = philox4x32_10(pair, c1, k.env, k.c3base + kStreamNoise, k.k0, k.k1);
  const double u1 = u53(r.x, r.y), u2 = u53(r.z, r.w);
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincos(6.283185307179586476925 * u2, &sn, &cs);
  z0 = fabs(rad * cs);
  z1 = fabs(rad * sn);
}
__device__ __forceinline__ double half_normal(const RngKey& k, uint32_t c1, uint32_t elem) {
  double z0, z1;
  half_normal_pair(k, c1, elem >> 1, z0, z1);
  return (elem & 1u) ? z1 : z0;
}

}  // namespace mapdn
