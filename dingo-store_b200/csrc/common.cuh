// common.cuh — shared device/host utilities of libb200vs (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace b200vs {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};

#define B200VS_CUDA(expr)                                                                       \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      throw ::b200vs::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) +    \
                                " (" __FILE__ ":" + std::to_string(__LINE__) + ")");           \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Order-preserving float <-> uint32 key.  Selection minimises (key, id) lexicographically.
// L2: key = ord(dist).  IP/cosine: key = ord(-ip) (exact negation), so "smaller key = better" always.
// -0 is canonicalised to +0 first so that -0 == +0 like the CPU float compare.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
  f = __fadd_rn(f, 0.0f);
  uint32_t u = __float_as_uint(f);
#else
  f = f + 0.0f;
  uint32_t u;
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t u) {
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
constexpr uint32_t KEY_SENTINEL_D = 0xFFFFFFFFu;
constexpr long long KEY_SENTINEL_ID = 0x7FFFFFFFFFFFFFFFLL;

__device__ __forceinline__ bool key_less(uint32_t ad, long long aid, uint32_t bd, long long bid) {
  return ad < bd || (ad == bd && aid < bid);
}

// ---------------------------------------------------------------------------------------------
// Exact FP32 distance in the reference's AVX-512 evaluation order
// (src/simd/distances_avx512.cc:48-143): 16 lane partials over i mod 16 with un-fused multiply and add,
// fold hi8+lo8, 8-wide tail, fold hi4+lo4, 4-wide tail, masked 1..3 tail, (m0+m1)+(m2+m3).
//
// Cooperative over a QUAD of 4 consecutive threads: thread t (0..3) of the quad owns AVX lanes
// 4t..4t+3 and streams its 16-byte slice of every 64-byte chunk with one 128-bit load.
// x: the database row (global), q: the query (shared or global), both 16-byte aligned when d%4==0;
// a scalar path covers other d.  The result is valid in ALL four threads of the quad.
// `qmask` must name exactly the 4 lanes of the calling quad group(s) that are active together —
// callers keep whole warps converged and pass 0xffffffff.
// ---------------------------------------------------------------------------------------------
template <bool L2>
__device__ __forceinline__ void acc4(float4& a, const float4 x, const float4 y) {
  if (L2) {
    float t;
    t = __fsub_rn(x.x, y.x); a.x = __fadd_rn(a.x, __fmul_rn(t, t));
    t = __fsub_rn(x.y, y.y); a.y = __fadd_rn(a.y, __fmul_rn(t, t));
    t = __fsub_rn(x.z, y.z); a.z = __fadd_rn(a.z, __fmul_rn(t, t));
    t = __fsub_rn(x.w, y.w); a.w = __fadd_rn(a.w, __fmul_rn(t, t));
  } else {
    a.x = __fadd_rn(a.x, __fmul_rn(x.x, y.x));
    a.y = __fadd_rn(a.y, __fmul_rn(x.y, y.y));
    a.z = __fadd_rn(a.z, __fmul_rn(x.z, y.z));
    a.w = __fadd_rn(a.w, __fmul_rn(x.w, y.w));
  }
}

template <bool L2>
__device__ __forceinline__ float term1(float x, float y) {
  if (L2) { float t = __fsub_rn(x, y); return __fmul_rn(t, t); }
  return __fmul_rn(x, y);
}

__device__ __forceinline__ float4 ld_f4(const float* p, bool vec) {
  if (vec) return *reinterpret_cast<const float4*>(p);
  return make_float4(p[0], p[1], p[2], p[3]);
}

template <bool L2>
__device__ __forceinline__ float quad_distance(const float* __restrict__ x, const float* __restrict__ q, int d,
                                               int t /*lane in quad 0..3*/, bool vec /*16B-aligned rows*/) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  const int nfull = d >> 4;
  const float* xp = x + 4 * t;
  const float* qp = q + 4 * t;
  if (vec) {  // 16-byte aligned rows: no branch inside the loop, so the unrolled body issues its 128-bit loads back to back
#pragma unroll 8
    for (int c = 0; c < nfull; ++c) {
      const float4 xv = *reinterpret_cast<const float4*>(xp);
      const float4 qv = *reinterpret_cast<const float4*>(qp);
      acc4<L2>(a, xv, qv);
      xp += 16; qp += 16;
    }
  } else {
#pragma unroll 4
    for (int c = 0; c < nfull; ++c) {
      const float4 xv = make_float4(xp[0], xp[1], xp[2], xp[3]);
      const float4 qv = make_float4(qp[0], qp[1], qp[2], qp[3]);
      acc4<L2>(a, xv, qv);
      xp += 16; qp += 16;
    }
  }
  int rem = d & 15;
  int base = nfull << 4;
  // msum1[j] = acc[8+j] + acc[j]  (threads 0,1 hold j = 0..7)
  float4 o;
  o.x = __shfl_down_sync(0xffffffffu, a.x, 2, 4);
  o.y = __shfl_down_sync(0xffffffffu, a.y, 2, 4);
  o.z = __shfl_down_sync(0xffffffffu, a.z, 2, 4);
  o.w = __shfl_down_sync(0xffffffffu, a.w, 2, 4);
  float4 m1 = make_float4(__fadd_rn(o.x, a.x), __fadd_rn(o.y, a.y), __fadd_rn(o.z, a.z), __fadd_rn(o.w, a.w));
  if (rem >= 8) {
    if (t < 2) {
      const float* xx = x + base + 4 * t;
      const float* qq = q + base + 4 * t;
      m1.x = __fadd_rn(m1.x, term1<L2>(xx[0], qq[0]));
      m1.y = __fadd_rn(m1.y, term1<L2>(xx[1], qq[1]));
      m1.z = __fadd_rn(m1.z, term1<L2>(xx[2], qq[2]));
      m1.w = __fadd_rn(m1.w, term1<L2>(xx[3], qq[3]));
    }
    base += 8; rem -= 8;
  }
  // msum2[j] = msum1[4+j] + msum1[j]  (thread 0 holds j = 0..3)
  o.x = __shfl_down_sync(0xffffffffu, m1.x, 1, 4);
  o.y = __shfl_down_sync(0xffffffffu, m1.y, 1, 4);
  o.z = __shfl_down_sync(0xffffffffu, m1.z, 1, 4);
  o.w = __shfl_down_sync(0xffffffffu, m1.w, 1, 4);
  float4 m2 = make_float4(__fadd_rn(o.x, m1.x), __fadd_rn(o.y, m1.y), __fadd_rn(o.z, m1.z), __fadd_rn(o.w, m1.w));
  if (rem >= 4) {
    if (t == 0) {
      const float* xx = x + base;
      const float* qq = q + base;
      m2.x = __fadd_rn(m2.x, term1<L2>(xx[0], qq[0]));
      m2.y = __fadd_rn(m2.y, term1<L2>(xx[1], qq[1]));
      m2.z = __fadd_rn(m2.z, term1<L2>(xx[2], qq[2]));
      m2.w = __fadd_rn(m2.w, term1<L2>(xx[3], qq[3]));
    }
    base += 4; rem -= 4;
  }
  if (rem > 0) {  // masked_read zero-fills: the add of (+0 or 0*0) is still performed on every lane
    if (t == 0) {
      const float* xx = x + base;
      const float* qq = q + base;
      const float x0 = xx[0], q0 = qq[0];
      const float x1 = rem > 1 ? xx[1] : 0.f, q1 = rem > 1 ? qq[1] : 0.f;
      const float x2 = rem > 2 ? xx[2] : 0.f, q2 = rem > 2 ? qq[2] : 0.f;
      m2.x = __fadd_rn(m2.x, term1<L2>(x0, q0));
      m2.y = __fadd_rn(m2.y, term1<L2>(x1, q1));
      m2.z = __fadd_rn(m2.z, term1<L2>(x2, q2));
      m2.w = __fadd_rn(m2.w, term1<L2>(0.f, 0.f));
    }
  }
  const float h0 = __fadd_rn(m2.x, m2.y);
  const float h1 = __fadd_rn(m2.z, m2.w);
  const float r = __fadd_rn(h0, h1);
  return __shfl_sync(0xffffffffu, r, 0, 4);
}

// ceil-div / pow2 helpers
__host__ __device__ __forceinline__ int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace b200vs
