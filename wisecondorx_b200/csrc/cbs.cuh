// Host interface of the CUDA circular binary segmentation (cbs.cu).
#pragma once
#include "wcx_common.cuh"

namespace wcx {

struct CbsStats {
  int64_t rounds;           // host-driven recursion rounds
  int64_t segments_tested;  // calls of the change-point finder
  int64_t perm_tests;       // segments that needed the permutation test
  int64_t t_tests;          // edge t-tests that needed permutations
  int64_t permutations;     // permutations evaluated
  int64_t launches;         // kernel launches
};

struct CbsWorkspace;
CbsWorkspace* cbs_workspace_create();
void cbs_workspace_destroy(CbsWorkspace* ws);

// sequential stopping boundary of the permutation tests (DNAcopy getbdry table, see cbs.cu); n = 0 switches it off
void cbs_set_boundary(CbsWorkspace* ws, const int32_t* sbdry, int32_t n);

// y, w: concatenated NA-free series (host); off[nseries + 1]; ends_out capacity = off[nseries]
int cbs_segment(CbsWorkspace* ws, const double* y, const double* w, const int64_t* off, int32_t nseries,
                const int32_t* series_ids, double alpha, int32_t nperm, int32_t kmax, int32_t nmin, int32_t min_width,
                uint32_t seed, int32_t* ends_out, int32_t* nseg_out, CbsStats* stats, cudaStream_t st);

}  // namespace wcx
