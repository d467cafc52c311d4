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

// floor(sqrt(n)), 0 <= n < 2^24 here (n <= 2*1020^2): ONE MUFU.SQRT estimate + integer fix-up == (int)sqrt((double)n) of
// src/utils.cpp:212.  sqrt.approx.f32 is within 2^-22 relative, i.e. < 0.001 absolute below 2^12, so the truncated estimate is
// off by at most one and a single correction in each direction makes it exact (tests/test_gpu_parity.py::test_isqrt_table_device
// checks every n of the domain).
__device__ __forceinline__ int isqrt_floor(int n) {
    float s;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(s) : "f"((float)n));
    int m = (int)s;
    if (m * m > n) --m;
    if ((m + 1) * (m + 1) <= n) ++m;
    return m;
}


}  // namespace cb
