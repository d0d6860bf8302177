// filter.cuh — device form of the reference FilterFunctors (src/vector/vector_index.h:67-146).
#pragma once
#include "common.cuh"

namespace b200vs {

// Device form of the reference FilterFunctors (src/vector/vector_index.h:67-146), all ANDed.
struct FilterDev {
  int has_range;
  int negate;
  long long rmin, rmax;
  const long long* sorted_ids;  // device, ascending
  long long n_ids;
};

__device__ __forceinline__ bool filter_pass(const FilterDev& f, long long id) {
  if (f.has_range && !(id >= f.rmin && id < f.rmax)) return false;  // RangeFilterFunctor::Check
  if (f.sorted_ids) {                                               // SortFilterFunctor::IsExist
    long long lo = 0, hi = f.n_ids - 1;
    bool exist = false;
    while (lo <= hi) {
      long long mid = (lo + hi) >> 1;
      long long v = f.sorted_ids[mid];
      if (v == id) { exist = true; break; }
      if (id < v) hi = mid - 1; else lo = mid + 1;
    }
    if (f.negate ? exist : !exist) return false;
  }
  return true;
}

}  // namespace b200vs
