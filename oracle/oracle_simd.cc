// oracle_simd.cc — CPU ORACLE (test infrastructure): distance arithmetic.
//
// Restates, in portable C++, the exact floating-point evaluation order of the reference's
// AVX-512 kernels, which is what `fvec_hook` installs on every AVX-512 host
// (src/simd/hook.cc:69-84) and therefore what faiss / hnswlib call through
// set_fvec_L2sqr_hook / set_fvec_inner_product_hook (src/vector/vector_index.cc:153-185):
//
//   fvec_inner_product_avx512   src/simd/distances_avx512.cc:48-92
//   fvec_L2sqr_avx512           src/simd/distances_avx512.cc:94-143
//
// Order (both kernels):  16 lane partials over i mod 16 (un-fused multiply then add: the
// reference builds that file with -mavx512f -mavx512dq -mavx512bw, CMakeLists.txt:614-616, no
// -mfma, so GCC emits vmulps+vaddps);  fold hi8+lo8;  optional 8-wide tail;  fold hi4+lo4;
// optional 4-wide tail;  optional masked 1..3 tail;  two hadds: (m0+m1)+(m2+m3).
//
// Must be compiled with -ffp-contract=off (Makefile does).  Pinned bit-for-bit against
// oracle/_ref (the reference's own objects) by tests/test_oracle_simd.py.
#include "oracle_common.h"

namespace {

template <bool L2>
#if defined(__x86_64__)
__attribute__((target_clones("avx512f", "avx2", "default")))
#endif
float avx512_order(const float* x, const float* y, size_t d) {
  float acc[16];
  for (int l = 0; l < 16; ++l) acc[l] = 0.0f;
  while (d >= 16) {
    for (int l = 0; l < 16; ++l) {
      if (L2) { const float t = x[l] - y[l]; acc[l] = acc[l] + t * t; }
      else    { acc[l] = acc[l] + x[l] * y[l]; }
    }
    x += 16; y += 16; d -= 16;
  }
  float m1[8];
  for (int l = 0; l < 8; ++l) m1[l] = acc[8 + l] + acc[l];   // msum1 = hi; msum1 += lo
  if (d >= 8) {
    for (int l = 0; l < 8; ++l) {
      if (L2) { const float t = x[l] - y[l]; m1[l] = m1[l] + t * t; }
      else    { m1[l] = m1[l] + x[l] * y[l]; }
    }
    x += 8; y += 8; d -= 8;
  }
  float m2[4];
  for (int l = 0; l < 4; ++l) m2[l] = m1[4 + l] + m1[l];
  if (d >= 4) {
    for (int l = 0; l < 4; ++l) {
      if (L2) { const float t = x[l] - y[l]; m2[l] = m2[l] + t * t; }
      else    { m2[l] = m2[l] + x[l] * y[l]; }
    }
    x += 4; y += 4; d -= 4;
  }
  if (d > 0) {  // masked_read zero-fills the missing lanes: the add of +0*+0 is still performed
    float bx[4] = {0, 0, 0, 0}, by[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < d; ++i) { bx[i] = x[i]; by[i] = y[i]; }
    for (int l = 0; l < 4; ++l) {
      if (L2) { const float t = bx[l] - by[l]; m2[l] = m2[l] + t * t; }
      else    { m2[l] = m2[l] + bx[l] * by[l]; }
    }
  }
  const float h0 = m2[0] + m2[1];   // _mm_hadd_ps(msum2, msum2)
  const float h1 = m2[2] + m2[3];
  return h0 + h1;                   // second hadd, lane 0
}

}  // namespace

extern "C" {

float oracle_fvec_L2sqr(const float* x, const float* y, size_t d) { return avx512_order<true>(x, y, d); }
float oracle_fvec_inner_product(const float* x, const float* y, size_t d) { return avx512_order<false>(x, y, d); }

// src/simd/distances_ref.cc:23-31
float oracle_fvec_L2sqr_seq(const float* x, const float* y, size_t d) {
  float res = 0;
  for (size_t i = 0; i < d; i++) { const float tmp = x[i] - y[i]; res += tmp * tmp; }
  return res;
}
// src/simd/distances_ref.cc:51-56
float oracle_fvec_inner_product_seq(const float* x, const float* y, size_t d) {
  float res = 0;
  for (size_t i = 0; i < d; i++) res += x[i] * y[i];
  return res;
}

// VectorIndexUtils::NormalizeVectorForFaiss, src/vector/vector_index_utils.cc:480-491.
// The squared norm there is faiss::fvec_norm_L2sqr (faiss fork not vendored; its summation order is
// unpinned).  ORACLE CHOICE: the hooked inner-product order, norm = <x,x>; 1-ulp differences in the
// norm are far inside the 1e-4 distance budget.  The divide is a true IEEE division by sqrt(norm).
void oracle_normalize_faiss(float* x, int32_t d) {
  static const float kFloatAccuracy = 0.00001;
  float n2 = oracle_fvec_inner_product(x, x, (size_t)d);
  if (n2 > 0 && std::abs(1.0f - n2) > kFloatAccuracy) {
    float n = std::sqrt(n2);
    for (int32_t i = 0; i < d; i++) x[i] = x[i] / n;
  }
}

// VectorIndexUtils::NormalizeVectorForHnsw, src/vector/vector_index_utils.cc:493-500:
// sequential scalar sum, inv = 1/(sqrtf(norm)+1e-30f), multiply.
void oracle_normalize_hnsw(const float* data, int32_t d, float* out) {
  float norm = 0.0f;
  for (int i = 0; i < d; i++) norm += data[i] * data[i];
  norm = 1.0f / (sqrtf(norm) + 1e-30f);
  for (int i = 0; i < d; i++) out[i] = data[i] * norm;
}

const char* oracle_version(void) { return "b200vs-oracle 1 (dingo-store dc8c439c restatement)"; }

}  // extern "C"
