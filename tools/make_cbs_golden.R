#!/usr/bin/env Rscript
# Golden vectors for the CUDA circular binary segmentation from the REAL DNAcopy (Bioconductor, the reference pins
# bioconductor-dnacopy ==1.76 in conda.yml).  R is not part of the build image, so this script is for a site that has
# R + DNAcopy + jsonlite:
#
#     Rscript tools/make_cbs_golden.R tests/golden/cbs_dnacopy.json
#
# Every case is segmented exactly the way the reference's include/CBS.R does it (CBS.R:70-73):
#     CNA(genomdat, chrom, maploc, data.type = "logratio", sampleid = "X") ; segment(CNA.object, alpha, verbose = 1,
#     weights)   -- all other arguments at the package defaults.
# tests/test_cbs_dnacopy.py picks the file up when it exists and compares breakpoints (loc.end) of the NumPy oracle
# (oracle/cbs_oracle.py) and of the GPU path with DNAcopy's.  The permutation stream of R cannot be reproduced, so
# the cases are built with margins (SNR >= 1.5 sigma per change over >= 8 bins, or pure noise) where the decision does
# not hang on single permutations; the test reports concordance per case and requires identical breakpoints there.
suppressMessages({library(DNAcopy); library(jsonlite)})
args <- commandArgs(trailingOnly = TRUE)
out <- if (length(args) >= 1) args[1] else "cbs_dnacopy.json"
set.seed(20260101)
mk <- function(n, cps, means, sd = 0.05, weighted = TRUE, na_runs = list()) {
  y <- rnorm(n, 0, sd)
  b <- c(0, cps, n)
  for (i in seq_along(means)) y[(b[i] + 1):b[i + 1]] <- y[(b[i] + 1):b[i + 1]] + means[i]
  w <- if (weighted) runif(n, 0.5, 2.0) else rep(1, n)
  for (r in na_runs) y[r[1]:r[2]] <- NA
  list(y = y, w = w)
}
cases <- list(
  noise_small      = mk(150, c(), c(0)),
  noise_large      = mk(3000, c(), c(0)),
  one_change_small = mk(180, c(90), c(0, 0.12)),
  one_change_large = mk(2500, c(1400), c(0, 0.10)),
  two_change_arc   = mk(1200, c(500, 560), c(0, 0.25, 0)),
  short_arc        = mk(800, c(300, 312), c(0, 0.4, 0)),
  many_changes     = mk(5000, c(700, 1500, 1540, 3000, 4200), c(0, 0.15, -0.2, 0.05, -0.1, 0.1)),
  unweighted       = mk(1000, c(400), c(0, -0.15), weighted = FALSE),
  with_na_runs     = mk(2000, c(900), c(0, 0.2), na_runs = list(c(100, 140), c(880, 905))),
  edge_change      = mk(600, c(6), c(0.5, 0)),
  chr1_15kb_like   = mk(16598, c(4000, 4400, 12000), c(0, 0.58, 0, -0.3), sd = 0.12)
)
res <- list()
for (nm in names(cases)) {
  cs <- cases[[nm]]
  keep <- !is.na(cs$y)
  x <- seq_along(cs$y)
  set.seed(1)
  cna <- CNA(cs$y, rep(1, length(cs$y)), x, data.type = "logratio", sampleid = "X")
  seg <- segment(cna, alpha = 1e-4, verbose = 0, weights = cs$w)$output
  res[[nm]] <- list(y = ifelse(is.na(cs$y), 0, cs$y), w = cs$w, loc_start = seg$loc.start, loc_end = seg$loc.end,
                    seg_mean = seg$seg.mean, num_mark = seg$num.mark, alpha = 1e-4, seed = 1)
}
writeLines(toJSON(list(dnacopy_version = as.character(packageVersion("DNAcopy")), cases = res), digits = 17, auto_unbox = TRUE), out)
cat("wrote", out, "\n")
