// Stand-alone probe: issue rate of the instruction kinds the front kernel is made of, on one B200.
// Every kernel runs N_ITER iterations of UNROLL independent dependency chains per thread; the figure
// printed is warp-instructions per clock per SM (4 = every SMSP issues every cycle).
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a pipe_probe.cu -o pipe_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int N_ITER = 4096;
constexpr int CH = 8;  // independent chains per thread

#define PROBE_KERNEL(name, DECL, BODY, SINK)                                              \
    __global__ void __launch_bounds__(256) name(float* out, long long* cycles, float seed) { \
        DECL;                                                                             \
        __syncthreads();                                                                  \
        long long t0 = clock64();                                                         \
        for (int it = 0; it < N_ITER; ++it) {                                             \
            _Pragma("unroll") for (int c = 0; c < CH; ++c) { BODY; }                      \
        }                                                                                 \
        long long t1 = clock64();                                                         \
        __syncthreads();                                                                  \
        SINK;                                                                             \
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;                               \
    }

#define DECL_F32                                         \
    float a[CH], b = seed, w = seed * 0.5f;              \
    for (int c = 0; c < CH; ++c) a[c] = seed + c + threadIdx.x
#define SINK_F32                                         \
    float s = 0;                                         \
    for (int c = 0; c < CH; ++c) s += a[c];              \
    if (s == 123.456f) out[threadIdx.x] = s

#define DECL_F32X2                                                     \
    unsigned long long a[CH], b, w;                                    \
    {                                                                  \
        float2 t = make_float2(seed, seed * 0.5f);                     \
        b = *reinterpret_cast<unsigned long long*>(&t);                \
        w = b + 12345;                                                 \
    }                                                                  \
    for (int c = 0; c < CH; ++c) a[c] = b + c + threadIdx.x
#define SINK_F32X2                                        \
    unsigned long long s = 0;                            \
    for (int c = 0; c < CH; ++c) s += a[c];              \
    if (s == 123456ull) out[threadIdx.x] = (float)s

#define DECL_I32                                         \
    int a[CH], b = (int)seed, w = (int)seed * 3;         \
    for (int c = 0; c < CH; ++c) a[c] = (int)seed + c + threadIdx.x
#define SINK_I32                                         \
    int s = 0;                                           \
    for (int c = 0; c < CH; ++c) s += a[c];              \
    if (s == 123456) out[threadIdx.x] = (float)s

PROBE_KERNEL(k_fmul, DECL_F32, asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b)), SINK_F32)
PROBE_KERNEL(k_fadd, DECL_F32, asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b)), SINK_F32)
PROBE_KERNEL(k_ffma, DECL_F32, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(b), "f"(w)), SINK_F32)
PROBE_KERNEL(k_fmul_fadd, DECL_F32,
             asm volatile("mul.rn.f32 %0, %0, %1;\n\tadd.rn.f32 %0, %0, %2;" : "+f"(a[c]) : "f"(b), "f"(w)), SINK_F32)
PROBE_KERNEL(k_fmul2, DECL_F32X2, asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[c]) : "l"(b)), SINK_F32X2)
PROBE_KERNEL(k_fadd2, DECL_F32X2, asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[c]) : "l"(b)), SINK_F32X2)
PROBE_KERNEL(k_ffma2, DECL_F32X2, asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[c]) : "l"(b), "l"(w)), SINK_F32X2)
PROBE_KERNEL(k_fmul2_fadd2, DECL_F32X2,
             asm volatile("mul.rn.f32x2 %0, %0, %1;\n\tadd.rn.f32x2 %0, %0, %2;" : "+l"(a[c]) : "l"(b), "l"(w)), SINK_F32X2)
PROBE_KERNEL(k_imad, DECL_I32, asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(b), "r"(w)), SINK_I32)
PROBE_KERNEL(k_iadd, DECL_I32, asm volatile("add.s32 %0, %0, %1;" : "+r"(a[c]) : "r"(b)), SINK_I32)
PROBE_KERNEL(k_lop3, DECL_I32, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b), "r"(w)), SINK_I32)
PROBE_KERNEL(k_prmt, DECL_I32, asm volatile("prmt.b32 %0, %0, %1, 0x7650;" : "+r"(a[c]) : "r"(b)), SINK_I32)
PROBE_KERNEL(k_i2f, DECL_I32,
             { float t; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(t) : "r"(a[c])); a[c] = __float_as_int(t); }, SINK_I32)
PROBE_KERNEL(k_f2i, DECL_I32,
             { asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(a[c]) : "f"(__int_as_float(a[c]))); }, SINK_I32)
PROBE_KERNEL(k_sqrt, DECL_F32, asm volatile("sqrt.approx.f32 %0, %0;" : "+f"(a[c])), SINK_F32)
PROBE_KERNEL(k_hfma2, DECL_I32, asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(b), "r"(w)), SINK_I32)
PROBE_KERNEL(k_hadd2, DECL_I32, asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(a[c]) : "r"(b)), SINK_I32)
PROBE_KERNEL(k_vabsdiff, DECL_I32, asm volatile("max.s16x2 %0, %0, %1;" : "+r"(a[c]) : "r"(b)), SINK_I32)
// mixes that model the blur inner loop: one product feeding two sums (scalar and packed)
PROBE_KERNEL(k_mix_1mul_2add, DECL_F32,
             { float q; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(q) : "f"(a[c]), "f"(b));
               asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(q));
               asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[(c + 1) % CH]) : "f"(q)); }, SINK_F32)
PROBE_KERNEL(k_mix2_1mul_2add, DECL_F32X2,
             { unsigned long long q; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(q) : "l"(a[c]), "l"(b));
               asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[c]) : "l"(q));
               asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[(c + 1) % CH]) : "l"(q)); }, SINK_F32X2)

PROBE_KERNEL(k_mix2_fma0_2add, DECL_F32X2,
             { unsigned long long q; unsigned long long nz = 0x8000000080000000ull;
               asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q) : "l"(a[c]), "l"(b), "l"(nz));
               asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[c]) : "l"(q));
               asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[(c + 1) % CH]) : "l"(q)); }, SINK_F32X2)
PROBE_KERNEL(k_fadd2_rz, DECL_F32X2, asm volatile("add.rz.f32x2 %0, %0, %1;" : "+l"(a[c]) : "l"(b)), SINK_F32X2)
PROBE_KERNEL(k_fadd_rz, DECL_F32, asm volatile("add.rz.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b)), SINK_F32)
PROBE_KERNEL(k_shfl, DECL_I32, asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(a[c])), SINK_I32)
PROBE_KERNEL(k_fsetp_sel, DECL_F32,
             asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(a[c]) : "f"(b), "f"(w)), SINK_F32)
PROBE_KERNEL(k_f2fp, DECL_I32,
             asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(a[c]) : "f"(__int_as_float(a[c])), "f"(__int_as_float(b))), SINK_I32)
PROBE_KERNEL(k_frnd, DECL_F32, asm volatile("cvt.rzi.f32.f32 %0, %0;" : "+f"(a[c])), SINK_F32)
PROBE_KERNEL(k_fmnmx, DECL_F32, asm volatile("max.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b)), SINK_F32)
PROBE_KERNEL(k_vote, DECL_I32,
             asm volatile("{.reg .pred p; setp.gt.s32 p, %0, %1; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "+r"(a[c]) : "r"(b)), SINK_I32)

// the form the blur uses: products by FMUL2, accumulation by FFMA2(acc, one, q) with a RUN-TIME one (ptxas folds a literal 1.0
// back into an add and then contracts it with the multiply)
__global__ void __launch_bounds__(256) k_mix2_mul_fmaone(float* out, long long* cycles, float seed) {
    unsigned long long a[CH], b, one;
    { float2 t = make_float2(seed, seed * 0.5f); b = *reinterpret_cast<unsigned long long*>(&t);
      float2 o = make_float2(seed, seed); one = *reinterpret_cast<unsigned long long*>(&o); }
    for (int c = 0; c < CH; ++c) a[c] = b + c + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < N_ITER; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            unsigned long long q;
            asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(q) : "l"(a[c]), "l"(b));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[c]) : "l"(one), "l"(q));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[(c + 1) % CH]) : "l"(one), "l"(q));
        }
    }
    long long t1 = clock64();
    __syncthreads();
    unsigned long long s = 0;
    for (int c = 0; c < CH; ++c) s += a[c];
    if (s == 123456ull) out[threadIdx.x] = (float)s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
// shared-memory loads: 32-bit and 128-bit, conflict free
__global__ void __launch_bounds__(256) k_lds32(float* out, long long* cycles, float seed) {
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += 256) sm[i] = seed + i;
    __syncthreads();
    float acc[CH] = {0};
    long long t0 = clock64();
    for (int it = 0; it < N_ITER; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(&sm[(threadIdx.x + 32 * c + it) & 4095])));
            acc[c] += v;
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int c = 0; c < CH; ++c) s += acc[c];
    if (s == 123.456f) out[threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
__global__ void __launch_bounds__(256) k_lds128(float* out, long long* cycles, float seed) {
    __shared__ float4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = make_float4(seed, i, 0, 1);
    __syncthreads();
    float acc[CH] = {0};
    long long t0 = clock64();
    for (int it = 0; it < N_ITER; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "r"((unsigned)__cvta_generic_to_shared(&sm[(threadIdx.x + 32 * c + it) & 2047])));
            acc[c] += v.x + v.w;
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int c = 0; c < CH; ++c) s += acc[c];
    if (s == 123.456f) out[threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

typedef void (*Kern)(float*, long long*, float);
struct Entry { const char* name; Kern k; int instr_per_body; };

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    float* out;
    long long* cyc;
    const int ctas_per_sm = 4;
    const int grid = sms * ctas_per_sm;
    cudaMalloc(&out, 4096);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    long long* h = new long long[grid];
    Entry es[] = {
        {"FMUL", k_fmul, 1}, {"FADD", k_fadd, 1}, {"FFMA", k_ffma, 1}, {"FMUL+FADD (dependent)", k_fmul_fadd, 2},
        {"FMUL2 (mul.f32x2)", k_fmul2, 1}, {"FADD2 (add.f32x2)", k_fadd2, 1}, {"FFMA2 (fma.f32x2)", k_ffma2, 1},
        {"FMUL2+FADD2 (dependent; check SASS for FFMA2 fusion)", k_fmul2_fadd2, 2},
        {"IMAD", k_imad, 1}, {"IADD", k_iadd, 1}, {"LOP3", k_lop3, 1}, {"PRMT", k_prmt, 1}, {"I2F", k_i2f, 1}, {"F2I", k_f2i, 1},
        {"MUFU.SQRT", k_sqrt, 1}, {"HFMA2", k_hfma2, 1}, {"HADD2", k_hadd2, 1}, {"VIMNMX.S16x2", k_vabsdiff, 1},
        {"mix 1 FMUL + 2 FADD", k_mix_1mul_2add, 3}, {"mix 1 FMUL2 + 2 FADD2", k_mix2_1mul_2add, 3},
        {"mix 1 FFMA2(x,w,-0) + 2 FADD2", k_mix2_fma0_2add, 3}, {"mix 1 FMUL2 + 2 FFMA2(acc,one,q)", k_mix2_mul_fmaone, 3}, {"FADD2.RZ", k_fadd2_rz, 1}, {"FADD.RZ", k_fadd_rz, 1},
        {"SHFL", k_shfl, 1}, {"FSETP+SEL", k_fsetp_sel, 2}, {"F2FP.F16.F32.PACK", k_f2fp, 1}, {"FRND.TRUNC", k_frnd, 1},
        {"FMNMX", k_fmnmx, 1}, {"ISETP+VOTE", k_vote, 2},
        {"LDS.32", k_lds32, 1}, {"LDS.128", k_lds128, 1},
    };
    printf("device %s, %d SMs, %d CTAs x 256 threads per SM\n", prop.name, sms, ctas_per_sm);
    for (auto& e : es) {
        e.k<<<grid, 256>>>(out, cyc, 1.0f);
        cudaEvent_t ev0, ev1; cudaEventCreate(&ev0); cudaEventCreate(&ev1);
        cudaEventRecord(ev0);
        for (int rep = 0; rep < 4; ++rep) e.k<<<grid, 256>>>(out, cyc, 1.0f);
        cudaEventRecord(ev1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, ev0, ev1); ms /= 4;
        if (err != cudaSuccess) { printf("%s: %s\n", e.name, cudaGetErrorString(err)); return 1; }
        cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < grid; ++i) avg += (double)h[i];
        avg /= grid;
        const double warp_instr = (double)N_ITER * CH * e.instr_per_body * 8 /*warps per CTA*/ * ctas_per_sm;
        printf("%-55s %.3f warp-instr/clk64/SM (%.0f clk64)  | event: %.1f us -> %.3f warp-instr/ns/SM = %.3f per clk @1.965GHz, clk64 rate %.3f GHz\n",
               e.name, warp_instr / avg, avg, ms * 1e3, warp_instr / (ms * 1e6), warp_instr / (ms * 1e6) / 1.965, avg / (ms * 1e6));
    }
    return 0;
}
