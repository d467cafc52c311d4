// selftest.cu — device-side table dumps used by the GPU tests to check the exact-arithmetic
// building blocks of front.cu over their WHOLE domain (not just on images).
#include <math.h>

#include "canny_math.h"
#include "exact_math.cuh"
#include "internal.h"

namespace cb {

__global__ void direction_table_kernel(int gmax, int16_t* __restrict__ out) {
    const int n = 2 * gmax + 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * n; i += (long long)gridDim.x * blockDim.x) {
        const int gy = (int)(i / n) - gmax, gx = (int)(i % n) - gmax;
        out[i] = (int16_t)dir_code_to_angle(direction_code<int>(gx, gy));
    }
}
__global__ void isqrt_table_kernel(int n_max, int32_t* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i <= n_max; i += (long long)gridDim.x * blockDim.x)
        out[i] = isqrt_floor((int)i);
}
// every float a in {0} u [2^-100, 256*b] for every entry b of the count table.  (Below 2^-100 the quotient
// nears the subnormal range where the residuals stop being exact; the front kernel switches to IEEE division
// when sigma is so small that such sums can occur — see FrontParams::ieee_div.)
__global__ void div_check_kernel(const float* __restrict__ cnt, const float* __restrict__ rcp, int n_tab,
                                 unsigned long long* __restrict__ mismatches) {
    const int t = blockIdx.y;
    if (t >= n_tab) return;
    const float b = cnt[t], y = rcp[t];
    if (!(b > 0.f)) return;
    const unsigned first = __float_as_uint(7.888609052210118e-31f);  // 2^-100
    const unsigned last = __float_as_uint(256.0f * b);
    unsigned long long bad = 0;
    for (unsigned long long i = first + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i <= last;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float a = __uint_as_float((unsigned)i);
        const float q = div_exact(a, b, y), ref = __fdiv_rn(a, b);
        if (__float_as_uint(q) != __float_as_uint(ref)) ++bad;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && __float_as_uint(div_exact(0.f, b, y)) != 0u) ++bad;
    if (bad) atomicAdd(mismatches, bad);
}

// every float mantissa a in [1, 2) (the quotient's rounding does not depend on the exponent away from the subnormal
// range): does ONE Markstein correction already give RN(a / b)?
__global__ void div3_check_kernel(float b, float y, unsigned long long* __restrict__ mismatches) {
    unsigned long long bad = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (1u << 23); i += gridDim.x * blockDim.x) {
        const float a = __uint_as_float(0x3F800000u | i);
        const float q0 = __fmul_rn(a, y);
        const float r = __fmaf_rn(-b, q0, a);
        const float q = __fmaf_rn(r, y, q0);
        if (__float_as_uint(q) != __float_as_uint(__fdiv_rn(a, b))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// same question for the ONE-instruction form q = fma(a, c, a) with c = RN(1/b - 1): the interior count is 1 +- an ulp or two, so
// a/b = a + a*(1/b - 1) and the fused multiply-add rounds that once.  Whether this single rounding always lands on RN(a/b)
// depends on b; it is decided by trying every mantissa.
__global__ void div1_check_kernel(float b, float c, unsigned long long* __restrict__ mismatches) {
    unsigned long long bad = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (1u << 23); i += gridDim.x * blockDim.x) {
        const float a = __uint_as_float(0x3F800000u | i);
        if (__float_as_uint(__fmaf_rn(a, c, a)) != __float_as_uint(__fdiv_rn(a, b))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// *mode = 1 (q = fma(a, c, a)), 3 (one Markstein correction) or 5 (two): the cheapest form that equals IEEE division by b for
// every float mantissa.  For the one-instruction form the three floats around 1/b - 1 are tried (*c is the nearest on entry and
// the one that works on return): which of them makes the single rounding land on RN(a/b) everywhere depends on b.
int check_div_mode_device(b200_ctx* ctx, float b, float y, float* c, int* mode) {
    *mode = 5;
    CB_TRY(ensure_ws(ctx->ws_misc, 256));
    unsigned long long* d = reinterpret_cast<unsigned long long*>(ctx->ws_misc.ptr);
    unsigned long long h[4] = {1, 1, 1, 1};
    const float cand[3] = {*c, nextafterf(*c, INFINITY), nextafterf(*c, -INFINITY)};
    CB_CUDA(cudaMemsetAsync(d, 0, 4 * sizeof(*d), ctx->stream));
    for (int k = 0; k < 3; ++k) div1_check_kernel<<<1184, 256, 0, ctx->stream>>>(b, cand[k], d + k);
    div3_check_kernel<<<1184, 256, 0, ctx->stream>>>(b, y, d + 3);
    CB_CUDA(cudaGetLastError());
    ctx->launches += 4;
    CB_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 3; ++k)
        if (h[k] == 0) { *mode = 1; *c = cand[k]; return B200_OK; }
    *mode = (h[3] == 0) ? 3 : 5;
    return B200_OK;
}

}  // namespace cb

using namespace cb;

extern "C" {

int b200_direction_table_host(int gmax, int16_t* out) {
    if (!out || gmax < 0 || gmax > 20000) { set_error("bad argument"); return B200_ERR_INVALID_ARG; }
    const int n = 2 * gmax + 1;
    for (int gy = -gmax; gy <= gmax; ++gy)
        for (int gx = -gmax; gx <= gmax; ++gx)
            out[(size_t)(gy + gmax) * n + (gx + gmax)] = (int16_t)dir_code_to_angle(direction_code<long long>(gx, gy));
    return B200_OK;
}

int b200_direction_table_device(b200_ctx* ctx, int gmax, int16_t* out_host) {
    if (!ctx || !out_host || gmax < 0 || gmax > 1020) { set_error("bad argument"); return B200_ERR_INVALID_ARG; }
    CB_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)(2 * gmax + 1) * (2 * gmax + 1);
    CB_TRY(ensure_ws(ctx->ws_planes, n * 2));
    int16_t* d = reinterpret_cast<int16_t*>(ctx->ws_planes.ptr);
    direction_table_kernel<<<1024, 256, 0, ctx->stream>>>(gmax, d);
    CB_CUDA(cudaGetLastError());
    ctx->launches++;
    CB_CUDA(cudaMemcpyAsync(out_host, d, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

int b200_isqrt_table_device(b200_ctx* ctx, int n_max, int32_t* out_host) {
    if (!ctx || !out_host || n_max < 0 || n_max > (1 << 24)) { set_error("bad argument"); return B200_ERR_INVALID_ARG; }
    CB_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)n_max + 1;
    CB_TRY(ensure_ws(ctx->ws_planes, n * 4));
    int32_t* d = reinterpret_cast<int32_t*>(ctx->ws_planes.ptr);
    isqrt_table_kernel<<<1024, 256, 0, ctx->stream>>>(n_max, d);
    CB_CUDA(cudaGetLastError());
    ctx->launches++;
    CB_CUDA(cudaMemcpyAsync(out_host, d, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

int b200_division_check_device(b200_ctx* ctx, float sigma, unsigned long long* mismatches) {
    if (!ctx || !mismatches) { set_error("bad argument"); return B200_ERR_INVALID_ARG; }
    CB_CUDA(cudaSetDevice(ctx->device));
    CB_TRY(prepare_gauss(ctx, sigma));
    const int n1 = ctx->gauss.radius + 1, n_tab = n1 * n1;
    CB_TRY(ensure_ws(ctx->ws_misc, 256));
    unsigned long long* d = reinterpret_cast<unsigned long long*>(ctx->ws_misc.ptr);
    CB_CUDA(cudaMemsetAsync(d, 0, sizeof(*d), ctx->stream));
    dim3 grid(592, n_tab);
    div_check_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->gauss.d_count, ctx->gauss.d_count + n_tab, n_tab, d);
    CB_CUDA(cudaGetLastError());
    ctx->launches++;
    CB_CUDA(cudaMemcpyAsync(mismatches, d, sizeof(*d), cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

int b200_division_mode_device(b200_ctx* ctx, float sigma, int* mode) {
    if (!ctx || !mode) { set_error("bad argument"); return B200_ERR_INVALID_ARG; }
    CB_CUDA(cudaSetDevice(ctx->device));
    CB_TRY(prepare_gauss(ctx, sigma));
    *mode = ctx->gauss.div_mode;
    return B200_OK;
}

}  // extern "C"
