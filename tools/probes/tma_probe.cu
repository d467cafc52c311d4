// Stand-alone probe: one CTA loads a u8 box with cp.async.bulk.tensor.{2d,3d} and copies it out.
// Build: nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a tma_probe.cu -o tma_probe -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, int box_bytes, uint8_t* out, int* status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint32_t bar = smem_u32(smem + 65536);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(smem_u32(smem)), "l"(&tmap), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(smem)), "l"(&tmap), "r"(bar), "r"(x), "r"(y) : "memory");
    }
    int ok = 0;
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(0) : "memory");
        if (done) { ok = 1; break; }
    }
    if (threadIdx.x == 0) *status = ok;
    if (ok) for (int i = threadIdx.x; i < box_bytes; i += blockDim.x) out[i] = smem[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    int only = argc > 1 ? atoi(argv[1]) : -1;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { printf("no entry point\n"); return 1; }
    EncodeFn enc = (EncodeFn)sym;
    const int W = 256, H = 256, F = 2;
    std::vector<uint8_t> h((size_t)W * H * F);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 8));
    uint8_t *d, *dout; int* dstat;
    cudaMalloc(&d, h.size()); cudaMalloc(&dout, 65536); cudaMalloc(&dstat, 4);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
    cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
    struct Case { int rank, bw, bh, x, y, z; };
    Case cases[] = {{3, 144, 32, -4, -4, 0}, {3, 144, 32, 16, 8, 1}, {3, 128, 32, 0, 0, 0}, {3, 256, 32, 0, 0, 0}, {2, 144, 32, -4, -4, 0},
                    {2, 128, 32, 0, 0, 0}, {3, 176, 32, 100, 240, 1}, {3, 16, 32, 0, 0, 0}, {3, 64, 8, 0, 0, 0}};
    int idx = -1;
    for (auto& c : cases) {
        ++idx; if (only >= 0 && idx != only) continue;
        CUtensorMap tm;
        cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F};
        cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
        cuuint32_t box[3] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, c.rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("rank %d box %dx%d at (%d,%d,%d): encode=%d ", c.rank, c.bw, c.bh, c.x, c.y, c.z, (int)r);
        if (r != CUDA_SUCCESS) { printf("\n"); continue; }
        cudaMemset(dout, 0xEE, 65536); cudaMemset(dstat, 0xFF, 4);
        const int bytes = c.bw * c.bh;
        if (c.rank == 3) probe<3><<<1, 128, 66560>>>(tm, c.x, c.y, c.z, bytes, dout, dstat);
        else probe<2><<<1, 128, 66560>>>(tm, c.x, c.y, c.z, bytes, dout, dstat);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel error: %s\n", cudaGetErrorString(e)); return 2; }
        int st; std::vector<uint8_t> o(bytes);
        cudaMemcpy(&st, dstat, 4, cudaMemcpyDeviceToHost); cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < c.bh; ++r2) for (int cc = 0; cc < c.bw; ++cc) {
            int gx = c.x + cc, gy = c.y + r2;
            uint8_t want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[(size_t)(c.rank == 3 ? c.z : 0) * W * H + (size_t)gy * W + gx] : 0;
            bad += o[(size_t)r2 * c.bw + cc] != want;
        }
        printf("completed=%d mismatches=%d\n", st, bad);
    }
    return 0;
}
