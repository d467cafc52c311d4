// synth.cu — procedural frames generated directly in HBM (same integer hash as the host generator in
// canny_math.h, so the CPU oracle and the GPU see identical bytes without shipping GiBs), plus the
// device-side edge counter the bench uses as its per-step result.
#include "canny_math.h"
#include "internal.h"

namespace cb {

__global__ void synth_kernel(uint8_t* __restrict__ d, int n_frames, int row0, int rows, int width, int kind,
                             uint64_t seed, int first_frame) {
    const long long per = (long long)rows * width;
    const long long total = per * n_frames;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i / per);
        const long long rem = i - (long long)f * per;
        const int r = (int)(rem / width), c = (int)(rem - (long long)r * width);
        d[i] = synth_pixel(kind, seed, first_frame + f, c, row0 + r);
    }
}

__global__ void count255_kernel(const uint8_t* __restrict__ d, size_t n, unsigned long long* __restrict__ out) {
    unsigned long long local = 0;
    const size_t n16 = ((reinterpret_cast<uintptr_t>(d) & 15) == 0) ? (n >> 4) : 0;
    const uint4* v = reinterpret_cast<const uint4*>(d);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 q = v[i];
        // bytes are 0 or 255: popcount of bit 7 of every byte
        local += __popc(q.x & 0x80808080u) + __popc(q.y & 0x80808080u) + __popc(q.z & 0x80808080u) + __popc(q.w & 0x80808080u);
    }
    for (size_t i = (n16 << 4) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        local += (d[i] == 255);
    for (int off = 16; off; off >>= 1) local += __shfl_down_sync(0xffffffffu, local, off);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

// Position-dependent checksum of an edge map: sum over edge pixels of mix64(global pixel index), mod 2^64.  A sum, so the checksums
// of the row bands of one image (each computed on its own GPU with its global offset) add up to the checksum of the whole image.
__global__ void hash255_kernel(const uint8_t* __restrict__ d, size_t n, unsigned long long offset, unsigned long long* __restrict__ out) {
    unsigned long long local = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (d[i] == 255) local += mix64(offset + i);
    for (int off = 16; off; off >>= 1) local += __shfl_down_sync(0xffffffffu, local, off);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}
int launch_hash255(b200_ctx* ctx, cudaStream_t st, const uint8_t* d, size_t n, unsigned long long offset, unsigned long long* d_out) {
    CB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), st));
    size_t blocks = (n + 1023) / 1024;
    const size_t cap = 32 * (size_t)(ctx->sm_count > 0 ? ctx->sm_count : 148);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    hash255_kernel<<<(int)blocks, 256, 0, st>>>(d, n, offset, d_out);
    CB_CUDA(cudaGetLastError());
    ctx->launches++;
    return B200_OK;
}

int launch_synth(b200_ctx* ctx, cudaStream_t st, uint8_t* d, int n_frames, int row0, int rows, int width, int kind,
                 uint64_t seed, int first_frame) {
    const long long total = (long long)rows * width * n_frames;
    long long blocks = (total + 255) / 256;
    const long long cap = 64LL * (ctx->sm_count > 0 ? ctx->sm_count : 148);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    synth_kernel<<<(int)blocks, 256, 0, st>>>(d, n_frames, row0, rows, width, kind, seed, first_frame);
    CB_CUDA(cudaGetLastError());
    ctx->launches++;
    return B200_OK;
}

int launch_count255(b200_ctx* ctx, cudaStream_t st, const uint8_t* d, size_t n, unsigned long long* d_count) {
    CB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
    size_t blocks = ((n >> 4) + 255) / 256;
    const size_t cap = 16 * (size_t)(ctx->sm_count > 0 ? ctx->sm_count : 148);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    count255_kernel<<<(int)blocks, 256, 0, st>>>(d, n, d_count);
    CB_CUDA(cudaGetLastError());
    ctx->launches++;
    return B200_OK;
}

}  // namespace cb
