// ransac.cuh — the pieces of OpenCV's RANSACPointSetRegistrator that both robust estimators on the path share
// (solvePnPRansac, sfm.py:67; findEssentialMat, sfm.py:307): the RNG index stream and the iteration-budget update.
#pragma once
#include <float.h>
#include <math.h>

namespace {

// ------------------------------------------------------------------ RNG subset stream
// cv::RNG (multiply-with-carry, state 2^64-1) as RANSACPointSetRegistrator::getSubset uses it:
// 5 distinct indices per iteration, redraw on duplicates.
__host__ __device__ inline void ransac_subsets(int n, int iters, int* out) {
  unsigned long long state = 0xFFFFFFFFFFFFFFFFull;
  for (int it = 0; it < iters; ++it) {
    int* s = out + 5 * it;
    for (int i = 0; i < 5; ++i) {
      for (;;) {
        state = (state & 0xFFFFFFFFull) * 4164903690ull + (state >> 32);
        int j = (int)((unsigned int)state % (unsigned int)n);
        bool dup = false;
        for (int k = 0; k < i; ++k) dup |= (s[k] == j);
        if (!dup) { s[i] = j; break; }
      }
    }
  }
}

__host__ __device__ inline int update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = fmax(p, 0.0); p = fmin(p, 1.0);
  ep = fmax(ep, 0.0); ep = fmin(ep, 1.0);
  double num = fmax(1.0 - p, DBL_MIN);
  double denom = 1.0 - pow(1.0 - ep, (double)model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}


}  // namespace
