// ref_shim.cc — extern "C" window onto the REFERENCE's own distance kernels.
// Compiled together with /root/reference/src/simd/*.cc (sources stay where they lie; nothing is
// copied) into oracle/_ref/libdingo_simd_ref.so by oracle/Makefile.  Test infrastructure only.
#include <cstddef>
#include <string>

#include "simd/distances_avx.h"
#include "simd/distances_avx512.h"
#include "simd/distances_ref.h"
#include "simd/distances_sse.h"
#include "simd/hook.h"

extern "C" {
// what fvec_hook() installed on THIS host (src/simd/hook.cc:69-124) — i.e. what faiss/hnswlib would call
float ref_fvec_L2sqr(const float* x, const float* y, size_t d) { return dingodb::fvec_L2sqr(x, y, d); }
float ref_fvec_inner_product(const float* x, const float* y, size_t d) { return dingodb::fvec_inner_product(x, y, d); }
float ref_fvec_norm_L2sqr(const float* x, size_t d) { return dingodb::fvec_norm_L2sqr(x, d); }
// explicit variants
float ref_fvec_L2sqr_avx512(const float* x, const float* y, size_t d) { return dingodb::fvec_L2sqr_avx512(x, y, d); }
float ref_fvec_inner_product_avx512(const float* x, const float* y, size_t d) { return dingodb::fvec_inner_product_avx512(x, y, d); }
float ref_fvec_L2sqr_avx(const float* x, const float* y, size_t d) { return dingodb::fvec_L2sqr_avx(x, y, d); }
float ref_fvec_inner_product_avx(const float* x, const float* y, size_t d) { return dingodb::fvec_inner_product_avx(x, y, d); }
float ref_fvec_L2sqr_sse(const float* x, const float* y, size_t d) { return dingodb::fvec_L2sqr_sse(x, y, d); }
float ref_fvec_inner_product_sse(const float* x, const float* y, size_t d) { return dingodb::fvec_inner_product_sse(x, y, d); }
float ref_fvec_L2sqr_ref(const float* x, const float* y, size_t d) { return dingodb::fvec_L2sqr_ref(x, y, d); }
float ref_fvec_inner_product_ref(const float* x, const float* y, size_t d) { return dingodb::fvec_inner_product_ref(x, y, d); }
int ref_cpu_support_avx512() { return dingodb::cpu_support_avx512() ? 1 : 0; }
const char* ref_simd_type() {
  static std::string s;
  dingodb::fvec_hook_info(s);
  return s.c_str();
}
}
