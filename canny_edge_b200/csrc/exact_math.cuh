// exact_math.cuh — device arithmetic that must be bit-exact, shared by front.cu and selftest.cu.
#pragma once
#include <cuda_runtime.h>

namespace cb {

// Correctly rounded a/b given y = RN(1/b) (host computed): two Markstein correction steps.  After the
// first, q is faithful; the second then yields RN(a/b) (Markstein 1990, Thm. on q' = q + r*y with
// exact residual r = a - b*q).  5 issue slots instead of the ~10 of the MUFU.RCP/FCHK sequence.
// tests/test_gpu_division.py checks it against __fdiv_rn over every float numerator in range for the
// count tables of the benchmarked sigmas.
__device__ __forceinline__ float div_exact(float a, float b, float y) {
    float q = __fmul_rn(a, y);
    float r = __fmaf_rn(-b, q, a);
    q = __fmaf_rn(r, y, q);
    r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, y, q);
}

// floor(sqrt(n)), 0 <= n < 2^24 here (n <= 2*1020^2): MUFU.SQRT estimate + integer fix-up == (int)sqrt((double)n)
__device__ __forceinline__ int isqrt_floor(int n) {
    int m = (int)sqrtf((float)n);  // compiled without fast-math: correctly rounded; the fix-up keeps it exact regardless
    if (m * m > n) --m;
    if ((m + 1) * (m + 1) <= n) ++m;
    return m;
}


}  // namespace cb
