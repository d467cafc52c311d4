// Stand-alone probe: how fast does ONE warp (and 2, 4, 8 warps per scheduler) get through the packed-FP32 blur run of the front
// kernels?  The run is the column pass's: 64-bit shared-memory loads, FMUL2.FTZ products shared by symmetric taps, FADD2 sums.
// Prints cycles per warp-level instruction of the run for each warps-per-SM setting.
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I../../canny_edge_b200/csrc blur_probe.cu -o blur_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "front_packed.cuh"
using namespace cb::pk;

template <int R, int S, bool SCALAR>
__global__ void probe(const float* w, float* out, long long* cycles, int iters) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 132 * (S + 2 * R); i += blockDim.x) sm[i] = 1.0f + (i % 7);
    __syncthreads();
    u64 ws2[R + 1];
    float ws[R + 1];
#pragma unroll
    for (int j = 0; j <= R; ++j) { ws[j] = w[R + j]; ws2[j] = pack2(ws[j], ws[j]); }
    const float* tcol = sm + 2 * (threadIdx.x & 63);
    u64 total = 0;
    float totals = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        asm volatile("" ::: "memory");   // the shared-memory lines may have changed: reload and recompute every iteration
        if (!SCALAR) {
            blur_run2<R, S>(ws2, [&](int i) { return *reinterpret_cast<const u64*>(tcol + i * 132); },
                            [&](int o, u64 sum) { total = add2(total, sum); });
        } else {
            float acc[S];
#pragma unroll
            for (int i = 0; i < S + 2 * R; ++i) {
                const float x = tcol[i * 132];
                float q[R + 1];
#pragma unroll
                for (int j = 0; j <= R; ++j) q[j] = __fmul_rn(x, ws[j]);
#pragma unroll
                for (int t = 0; t <= 2 * R; ++t) {
                    const int o = i - t;
                    if (o >= 0 && o < S) { const int j = t < R ? R - t : t - R; acc[o] = (t == 0) ? q[j] : __fadd_rn(acc[o], q[j]); }
                }
                if (i >= 2 * R) totals += acc[i - 2 * R];
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float a, b;
    unpack2(total, a, b);
    if (a + b + totals == 12345.f) out[0] = a;
}

template <int R, int S, bool SCALAR>
void run(const char* name, int warps, const float* dw, float* dout, long long* dcyc, int sms) {
    const int iters = 200;
    const size_t smem = 132 * (S + 2 * R) * 4;
    probe<R, S, SCALAR><<<sms, 32 * warps, smem>>>(dw, dout, dcyc, iters);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, dcyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += h[i];
    avg /= sms;
    // instructions of one run per warp: loads + products + sums + the emit add
    const int loads = S + 2 * R, prods = (R + 1) * (S + 2 * R) - R * (R + 1), sums = 2 * R * S, emits = S;
    const double instr = (double)(loads + prods + sums + emits) * iters;
    printf("%-28s warps/SM %2d (per scheduler %d): %8.0f clk per block, %.2f clk per warp-instruction, SM issues %.2f blur instr/clk\n", name, warps,
           warps / 4, avg, avg / instr, instr * warps / avg);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    float hw[11] = {0.0005f, 0.005f, 0.03f, 0.1f, 0.22f, 0.285f, 0.22f, 0.1f, 0.03f, 0.005f, 0.0005f};
    float *dw, *dout;
    long long* dcyc;
    cudaMalloc(&dw, sizeof(hw)); cudaMalloc(&dout, 64); cudaMalloc(&dcyc, sizeof(long long) * 256);
    cudaMemcpy(dw, hw, sizeof(hw), cudaMemcpyHostToDevice);
    for (int warps : {4, 8, 16, 32}) run<5, 18, false>("packed S=18", warps, dw, dout, dcyc, sms);
    for (int warps : {4, 8, 16}) run<5, 34, false>("packed S=34", warps, dw, dout, dcyc, sms);
    for (int warps : {4, 8, 16, 32}) run<5, 34, true>("scalar S=34", warps, dw, dout, dcyc, sms);
    return 0;
}
