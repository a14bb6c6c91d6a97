// Shared by the host-side (no device work) entry points: a range-parallel loop on host threads and NumPy's pairwise
// summation, so that sums computed here are the float64 VALUES np.sum returns for the same contiguous array.
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace wcx {

// fn(begin, end) over [0, total) in chunks of `grain`, chunk c on thread c % nt
template <class F>
void host_parallel(int64_t total, int64_t grain, int32_t threads, F&& fn) {
  if (total <= 0) return;
  const int64_t chunks = std::max<int64_t>(1, (total + grain - 1) / grain);
  const int32_t nt = (int32_t)std::max<int64_t>(1, std::min<int64_t>(threads, chunks));
  if (nt == 1) { fn(0, total); return; }
  std::vector<std::thread> pool;
  pool.reserve(nt);
  for (int32_t t = 0; t < nt; ++t)
    pool.emplace_back([=, &fn]() {
      for (int64_t c = t; c < chunks; c += nt) fn(c * grain, std::min(total, (c + 1) * grain));
    });
  for (auto& th : pool) th.join();
}

// numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum: fewer than 8 elements are added in order; up to 128
// elements go through eight running sums combined as ((0+1)+(2+3))+((4+5)+(6+7)) followed by a sequential tail; longer
// arrays are split at the multiple of eight at or below the middle.
inline double numpy_pairwise_sum(const double* a, int64_t n) {
  if (n < 8) {
    double res = 0.;
    for (int64_t i = 0; i < n; ++i) res += a[i];
    return res;
  }
  if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int64_t i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  return numpy_pairwise_sum(a, n2) + numpy_pairwise_sum(a + n2, n - n2);
}

}  // namespace wcx
