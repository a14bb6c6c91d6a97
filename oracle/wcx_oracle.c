/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the WisecondorX `get_reference` hot loops.
 *
 * Used by tests/ (parity checker at sizes the NumPy oracle is too slow for), by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference leg.  It is NOT
 * a product path: nothing under wisecondorx_b200/ links or loads it.
 *
 * Pinned: tests/test_oracle_golden.py checks every entry point bit-for-bit against the golden
 * vectors generated from the live reference (tests/golden/get_reference.npz).
 *
 * Citations are relative to /root/reference/src/wisecondorx/.
 *   - distance:   newref_tools.py:260  np.sum(np.power(chr_data - row, 2), 1)
 *                 three roundings per term (subtract, square, add) and NumPy's pairwise
 *                 summation order over the contiguous sample axis;
 *                 compile with -ffp-contract=off so no FMA is formed.
 *   - selection:  newref_tools.py:261-277  strict `<` against the running maximum and
 *                 bisect_right insertion into a sorted list of ref_size entries
 *                 initialised to (-1, 1e10).
 *   - null ratio: newref_tools.py:210-224  log2(col[b] / median(col[idx[b,:]])), the gather
 *                 indexes the FULL column with chromosome-excluded positions, -1 wraps.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* NumPy pairwise summation (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum). */
static double pairwise_sum(const double *a, ptrdiff_t n) {
  if (n < 8) {
    double res = 0.0;
    for (ptrdiff_t i = 0; i < n; i++) res += a[i];
    return res;
  } else if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    ptrdiff_t i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; j++) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
  } else {
    ptrdiff_t n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
  }
}

double wcxo_sqdist(const double *a, const double *b, int32_t S, double *scratch) {
  for (int32_t s = 0; s < S; s++) {
    double t = b[s] - a[s]; /* chr_data - row */
    scratch[s] = t * t;     /* np.power(x, 2) == x*x bitwise */
  }
  return pairwise_sum(scratch, S);
}

/* bisect.bisect_right on an ascending array */
static int bisect_right(const double *v, int n, double x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (x < v[mid]) hi = mid; else lo = mid + 1;
  }
  return lo;
}

/* indexes/distances of get_reference for rows [row_begin,row_end).
 * X: [N,S] row-major; per/cum: masked bins per chromosome and cumulative (C entries).
 * Gonosomal mode (C > 22): chromosomes other than index 22/23 get placeholder rows (0, 1.0)
 * (newref_tools.py:186-191). */
int wcxo_topk(const double *X, int64_t N, int32_t S, const int64_t *per, const int64_t *cum,
              int32_t C, int64_t row_begin, int64_t row_end, int32_t k, int32_t *idx,
              double *dist, int32_t nthreads) {
  if (nthreads <= 0) nthreads = 1;
  int err = 0;
#pragma omp parallel num_threads(nthreads)
  {
    double *scratch = (double *)malloc(sizeof(double) * (size_t)(S > 0 ? S : 1));
    if (!scratch) {
#pragma omp atomic write
      err = 1;
    }
#pragma omp for schedule(dynamic, 1)
    for (int64_t r = row_begin; r < row_end; r++) {
      if (!scratch) continue;
      int32_t *oi = idx + (r - row_begin) * (int64_t)k;
      double *od = dist + (r - row_begin) * (int64_t)k;
      int c = 0;
      while (c < C && cum[c] <= r) c++;
      if (C > 22 && c != 22 && c != 23) {
        for (int t = 0; t < k; t++) { oi[t] = 0; od[t] = 1.0; }
        continue;
      }
      int64_t cs = cum[c] - per[c], ce = cum[c];
      for (int t = 0; t < k; t++) { oi[t] = -1; od[t] = 1e10; }
      double cur_max = 1e10;
      const double *row = X + r * (int64_t)S;
      int64_t pos = 0;
      for (int64_t j = 0; j < N; j++) {
        if (j >= cs && j < ce) continue;
        double d = wcxo_sqdist(row, X + j * (int64_t)S, S, scratch);
        if (d < cur_max) {
          int p = bisect_right(od, k, d);
          /* pop the last, insert at p */
          memmove(od + p + 1, od + p, sizeof(double) * (size_t)(k - 1 - p));
          memmove(oi + p + 1, oi + p, sizeof(int32_t) * (size_t)(k - 1 - p));
          od[p] = d;
          oi[p] = (int32_t)pos;
          cur_max = od[k - 1];
        }
        pos++;
      }
    }
    free(scratch);
  }
  return err;
}

static int cmp_double(const void *a, const void *b) {
  double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}

/* np.median of k values: NaN if any NaN, else mean of the two middle order statistics
 * (or the middle one for odd k). */
static double median_k(double *v, int k) {
  for (int i = 0; i < k; i++) if (isnan(v[i])) return NAN;
  qsort(v, (size_t)k, sizeof(double), cmp_double);
  if (k % 2) return v[k / 2];
  return (v[k / 2 - 1] + v[k / 2]) / 2.0; /* np.mean of two: (a+b)/2 */
}

/* out[r - row_begin, m] = log2(X[r, ids[m]] / median(X[wrap(idx[r - row_begin, :]), ids[m]])) */
int wcxo_null_ratios(const double *X, int64_t N, int32_t S, const int32_t *idx,
                     int64_t row_begin, int64_t row_end, int32_t k, const int32_t *ids,
                     int32_t M, double *out, int32_t nthreads) {
  if (nthreads <= 0) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
  {
    double *v = (double *)malloc(sizeof(double) * (size_t)(k > 0 ? k : 1));
#pragma omp for schedule(dynamic, 16)
    for (int64_t r = row_begin; r < row_end; r++) {
      const int32_t *ri = idx + (r - row_begin) * (int64_t)k;
      for (int m = 0; m < M; m++) {
        int32_t s = ids[m];
        for (int t = 0; t < k; t++) {
          int64_t g = ri[t];
          if (g < 0) g += N;
          v[t] = X[g * (int64_t)S + s];
        }
        double med = median_k(v, k);
        out[(r - row_begin) * (int64_t)M + m] = log2(X[r * (int64_t)S + s] / med);
      }
    }
    free(v);
  }
  return 0;
}

int wcxo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
