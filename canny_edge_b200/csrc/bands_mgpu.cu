// bands_mgpu.cu — ONE image split into row bands over several B200s, driven from C (BASELINE config 5).
//
// The reference has no counterpart: it runs one frame on one GPU and finishes hysteresis on the CPU (src/cuda.cu:392-450, :436).
// b200_band_front / _boundary_export / _finalize (band.cu) are the per-band building blocks; this file adds what a C or C++ caller
// in the position of src/main.cpp:128 needs to use them across GPUs without any Python or torch: the two exchange steps.
//
//   step 1, halo rows    window/2 + 2 input rows from each neighbour.
//   step 2, label merge  every band's boundary-row records (label, flags) reach every rank, which then unions the records that
//                        touch across a boundary and finalises its own band (band.cu).
//
// Two transports, same results:
//   P2P  (default when every rank can map its peers' buffers: CUDA IPC between processes, plain pointers inside one process)
//        * halo rows are PULLED by the copy engines over NVLink (cudaMemcpyAsync from the neighbour's mapped band buffer on a copy
//          stream) while the front kernel already works on the band's interior rows, which need no halo; the rows next to the
//          band edges follow in two small launches once the copies have landed;
//        * boundary records travel SPARSE: only candidate pixels (typically < 2 % of a boundary row) are listed, and the kernel
//          that builds the dense cross-band table reads every peer's list directly from that peer's memory (NVLink loads), so there
//          is no collective at all on the data path: readiness is a step counter each producer stores into its consumers' flag
//          words (st.release.sys / ld.acquire.sys);
//   NCCL (fallback, or B200_BANDS_TRANSPORT=nccl): ncclSend/ncclRecv of the halo rows in one group and one ncclAllGather of the
//        dense records — what canny_edge_b200/sharded.py did with torch.distributed in round 1.
// NCCL is resolved at run time (dlopen of the libnccl.so.2 already in the process, e.g. torch's): the library has no link-time
// dependency on it and single-GPU users never load it.  In P2P mode NCCL only bootstraps (all-gather of the IPC handles).
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only; every call goes through the table below
#include <string.h>
#include <unistd.h>

#include <vector>

#include "internal.h"

namespace cb {

// ---- NCCL through dlopen ----------------------------------------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclCommCount) CommCount = nullptr;
    decltype(&ncclCommUserRank) CommUserRank = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
    if (g_nccl.handle) return B200_OK;
    // the copy already in the process first (a communicator handed in by the caller belongs to THAT instance)
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) { set_error("cannot load libnccl.so.2: %s", dlerror()); return B200_ERR_UNSUPPORTED; }
#define CB_NCCL_SYM(name)                                                                        \
    g_nccl.name = reinterpret_cast<decltype(g_nccl.name)>(dlsym(h, "nccl" #name));                \
    if (!g_nccl.name) { set_error("libnccl has no symbol nccl" #name); return B200_ERR_UNSUPPORTED; }
    CB_NCCL_SYM(GetUniqueId) CB_NCCL_SYM(CommInitRank) CB_NCCL_SYM(CommDestroy) CB_NCCL_SYM(CommCount) CB_NCCL_SYM(CommUserRank)
    CB_NCCL_SYM(AllGather) CB_NCCL_SYM(Send) CB_NCCL_SYM(Recv) CB_NCCL_SYM(GroupStart) CB_NCCL_SYM(GroupEnd) CB_NCCL_SYM(GetErrorString)
    CB_NCCL_SYM(GetVersion)
#undef CB_NCCL_SYM
    g_nccl.handle = h;
    return B200_OK;
}
#define CB_NCCL(expr)                                                                                        \
    do {                                                                                                     \
        ncclResult_t r_ = (expr);                                                                            \
        if (r_ != ncclSuccess) {                                                                             \
            cb::set_error("%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(r_), __FILE__, __LINE__);    \
            return B200_ERR_CUDA;                                                                            \
        }                                                                                                    \
    } while (0)

// ---- what a band shares with its peers ----------------------------------------------------------------------------------------------
constexpr int kMaxBands = 64;
struct SparseRec { int32_t index, label, flags, pad; };   // record `index` of the band (band.cu: record_pixel), its label and flags
struct BandShared {                     // one allocation per band, mapped by every peer
    unsigned int halo_ready[2];         // step number stored by the UPPER ([0]) / LOWER ([1]) neighbour: "my rows of this step are in place"
    unsigned int rec_ready[kMaxBands];  // step number stored by band b: "my sparse records of this step are complete"
    unsigned int err;                   // set by a wait that gave up (a peer never signalled)
    unsigned int count[2];              // entries in recs[parity] (parity = step & 1: a band may be one step ahead of a reader)
    unsigned int pad[64 - 2 - 1 - 2];
    // SparseRec recs[2][S] follows
};
static_assert(sizeof(BandShared) == (kMaxBands + 64) * 4, "layout");
__host__ __device__ inline size_t shared_bytes(int S) { return sizeof(BandShared) + 2 * (size_t)S * sizeof(SparseRec); }
__device__ __forceinline__ SparseRec* shared_recs(BandShared* s, int parity, int S) {
    return reinterpret_cast<SparseRec*>(s + 1) + (size_t)parity * S;
}

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Waits (bounded: a few seconds) until *flag >= step.  A peer that never signals becomes an error code, not a hung GPU.
__device__ __forceinline__ bool wait_step(const unsigned int* flag, unsigned int step) {
    for (long long spin = 0; spin < (1LL << 24); ++spin) {
        if ((int)(ld_acquire_sys(flag) - step) >= 0) return true;
        __nanosleep(200);
    }
    return false;
}

// "my band rows of step `step` are in place": one system-scope store into each neighbour's flag word
__global__ void band_signal_halo_kernel(unsigned int* up_flag, unsigned int* down_flag, unsigned int step) {
    __threadfence_system();
    if (up_flag) st_release_sys(up_flag, step);
    if (down_flag) st_release_sys(down_flag, step);
}
__global__ void band_wait_halo_kernel(BandShared* mine, int need_up, int need_down, unsigned int step) {
    bool ok = true;
    if (need_up) ok = wait_step(&mine->halo_ready[0], step) && ok;
    if (need_down) ok = wait_step(&mine->halo_ready[1], step) && ok;
    if (!ok) mine->err = 1;
}

// dense records of this band -> sparse list in the shared block (candidates only), then "ready" into every peer's flag word
__global__ void band_compact_kernel(const b200_band_record* __restrict__ rec, int S, BandShared* mine, int parity) {
    SparseRec* out = shared_recs(mine, parity, S);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = i < S && (rec[i].flags & 2) != 0;
    const unsigned vote = __ballot_sync(0xffffffffu, have);
    if (!vote) return;
    const int lane = threadIdx.x & 31;
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(&mine->count[parity], (unsigned int)__popc(vote));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (have) {
        const b200_band_record r = rec[i];
        out[base + __popc(vote & ((1u << lane) - 1u))] = SparseRec{i, r.label, r.flags, 0};
    }
}
__global__ void band_signal_records_kernel(BandShared* const* peers, int n_bands, int me, unsigned int step) {
    const int b = threadIdx.x;
    if (b >= n_bands) return;
    __threadfence_system();
    st_release_sys(&peers[b]->rec_ready[me], step);
}
// block b: wait for band b's records of this step, then scatter them into the dense table all[b*S + index] (zeroed beforehand:
// flags 0 = "not a candidate").  Peer memory is read through NVLink with L1-bypassing loads.  The block that handles the band of the
// NEXT parity's previous use also clears nothing: counts are reset by their owner (band_reset_count_kernel) two steps later.
__global__ void band_gather_kernel(BandShared* const* peers, int S, int me, unsigned int step, b200_band_record* __restrict__ all) {
    const int b = blockIdx.x;
    BandShared* mine = peers[me];
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = (b == me) ? 1 : (wait_step(&mine->rec_ready[b], step) ? 1 : 0);
    __syncthreads();
    if (!s_ok) { if (threadIdx.x == 0) mine->err = 2; return; }
    BandShared* src = peers[b];
    const int parity = (int)(step & 1u);
    const unsigned int n = __ldcv(&src->count[parity]);
    const int4* recs = reinterpret_cast<const int4*>(shared_recs(src, parity, S));
    for (unsigned int i = threadIdx.x; i < n && i < (unsigned int)S; i += blockDim.x) {
        const int4 r = __ldcv(recs + i);          // {index, label, flags, pad}
        if (r.x >= 0 && r.x < S) all[(size_t)b * S + r.x] = b200_band_record{r.y, r.z};
    }
}
__global__ void band_reset_count_kernel(BandShared* mine, int parity) { mine->count[parity] = 0; }

}  // namespace cb

using namespace cb;

struct b200_bands {
    b200_ctx* ctx = nullptr;
    int rank = 0, world = 1, height = 0, width = 0;
    int row0 = 0, rows = 0, halo = 0, above = 0, below = 0;
    float sigma = 0.f;
    int lo = 0, hi = 0;
    int transport = 0;                   // 0 none (one band), 1 P2P, 2 NCCL
    bool split_front = true;             // P2P: interior rows first, edge rows after the halo copies
    ncclComm_t comm = nullptr;
    bool own_comm = false;
    int S = 0;                           // records per band
    uint8_t* buf = nullptr;              // [above + rows + below][width]
    BandShared* shared = nullptr;
    b200_band_record* rec = nullptr;     // S
    b200_band_record* all = nullptr;     // world * S
    // peers
    uint8_t* up_buf = nullptr;           // neighbours' band buffers (first byte of their allocation), mapped here
    uint8_t* down_buf = nullptr;
    std::vector<BandShared*> peer_shared;   // [world], own entry = shared
    BandShared** d_peer_shared = nullptr;   // device copy of the table
    std::vector<void*> ipc_opened;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_halo = nullptr, ev_t[8] = {};
    bool timing = false;
    unsigned int step = 0;
    int phase = 0;                       // 0 idle, 1 after begin, 2 after front
    uint8_t* cur_edges = nullptr;
};

namespace {

struct Geo { int row0, rows, above, below; };
Geo band_geo(int height, int world, int rank, int halo) {
    const int base = height / world, rem = height % world;
    Geo g;
    g.row0 = rank * base + (rank < rem ? rank : rem);
    g.rows = base + (rank < rem ? 1 : 0);
    g.above = halo < g.row0 ? halo : g.row0;
    const int after = height - (g.row0 + g.rows);
    g.below = halo < after ? halo : after;
    return g;
}

int bands_alloc(b200_bands* b) {
    CB_CUDA(cudaSetDevice(b->ctx->device));
    const size_t buf_bytes = (size_t)(b->above + b->rows + b->below) * b->width;
    CB_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->buf), buf_bytes));
    CB_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->shared), shared_bytes(b->S)));
    CB_CUDA(cudaMemset(b->shared, 0, shared_bytes(b->S)));
    CB_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->rec), (size_t)b->S * sizeof(b200_band_record)));
    CB_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->all), (size_t)b->world * b->S * sizeof(b200_band_record)));
    CB_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->d_peer_shared), sizeof(BandShared*) * (size_t)b->world));
    CB_CUDA(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
    CB_CUDA(cudaEventCreateWithFlags(&b->ev_start, cudaEventDisableTiming));
    CB_CUDA(cudaEventCreateWithFlags(&b->ev_halo, cudaEventDisableTiming));
    for (auto& e : b->ev_t) CB_CUDA(cudaEventCreate(&e));
    CB_CUDA(cudaDeviceSynchronize());
    return B200_OK;
}

int bands_init_common(b200_bands* b, b200_ctx* ctx, int rank, int world, int height, int width, float sigma, int lo, int hi) {
    if (!ctx) { set_error("b200_bands needs an explicit context"); return B200_ERR_INVALID_ARG; }
    if (world < 1 || world > kMaxBands || rank < 0 || rank >= world) { set_error("bad rank %d of %d (at most %d bands)", rank, world, kMaxBands); return B200_ERR_INVALID_ARG; }
    if (height < 2 || width < 2) { set_error("height and width must be >= 2"); return B200_ERR_INVALID_ARG; }
    if (!thresholds_supported(lo, hi)) { set_error("thresholds minVal=%d > 255 >= maxVal=%d are not reproduced (src/utils.cpp:327-340)", lo, hi); return B200_ERR_UNSUPPORTED; }
    CB_CUDA(cudaSetDevice(ctx->device));
    CB_TRY(prepare_gauss(ctx, sigma));
    b->ctx = ctx; b->rank = rank; b->world = world; b->height = height; b->width = width;
    b->sigma = sigma; b->lo = lo; b->hi = hi;
    b->halo = ctx->gauss.radius + 2;
    if (world > 1 && height / world < (b->halo > 2 ? b->halo : 2)) {
        set_error("bands of %d rows are shorter than the %d-row halo; use fewer bands", height / world, b->halo);
        return B200_ERR_INVALID_ARG;
    }
    const Geo g = band_geo(height, world, rank, b->halo);
    b->row0 = g.row0; b->rows = g.rows; b->above = g.above; b->below = g.below;
    if ((long long)b->rows * width >= (1LL << 31)) { set_error("band exceeds int indexing"); return B200_ERR_UNSUPPORTED; }
    b->S = b200_band_record_count(width);
    b->peer_shared.assign((size_t)world, nullptr);
    if (const char* e = getenv("B200_BANDS_SPLIT")) b->split_front = e[0] != '0';
    return bands_alloc(b);
}

int bands_publish_peers(b200_bands* b) {
    CB_CUDA(cudaSetDevice(b->ctx->device));
    CB_CUDA(cudaMemcpy(b->d_peer_shared, b->peer_shared.data(), sizeof(BandShared*) * (size_t)b->world, cudaMemcpyHostToDevice));
    return B200_OK;
}

struct IpcBlob {   // what every rank tells every other rank at creation
    cudaIpcMemHandle_t buf, shared;
    int device;
    int pid;
    char host[56];
};

// NCCL all-gather of a small host blob (bootstrap only)
int nccl_allgather_host(b200_bands* b, const void* mine, size_t bytes, std::vector<unsigned char>& all) {
    unsigned char* d = nullptr;
    CB_CUDA(cudaMalloc(reinterpret_cast<void**>(&d), bytes * (size_t)(b->world + 1)));
    CB_CUDA(cudaMemcpy(d, mine, bytes, cudaMemcpyHostToDevice));
    cudaStream_t st = b->copy_stream;
    CB_NCCL(g_nccl.AllGather(d, d + bytes, bytes, ncclUint8, b->comm, st));
    CB_CUDA(cudaStreamSynchronize(st));
    all.resize(bytes * (size_t)b->world);
    CB_CUDA(cudaMemcpy(all.data(), d + bytes, all.size(), cudaMemcpyDeviceToHost));
    CB_CUDA(cudaFree(d));
    return B200_OK;
}

BandGeom geom_of(const b200_bands* b, uint8_t* d_edges) {
    return BandGeom{b->buf, b->above, b->below, b->rows, b->row0, b->height, b->width, b->lo, b->hi, d_edges};
}

void mark(b200_bands* b, int i) { if (b->timing) cudaEventRecord(b->ev_t[i], b->ctx->stream); }

}  // namespace

extern "C" {

int b200_bands_unique_id(void* id_out) {
    if (!id_out) { set_error("null pointer"); return B200_ERR_INVALID_ARG; }
    CB_TRY(load_nccl());
    ncclUniqueId id;
    CB_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == B200_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
    memcpy(id_out, &id, sizeof(id));
    return B200_OK;
}

int b200_bands_destroy(b200_bands* b) {
    if (!b) return B200_OK;
    if (b->ctx) cudaSetDevice(b->ctx->device);
    cudaDeviceSynchronize();
    if (b->comm && b->world > 1 && g_nccl.handle) {
        // nobody unmaps or frees while a peer may still be reading: a last (tiny) collective is the barrier
        unsigned char* d = nullptr;
        if (cudaMalloc(reinterpret_cast<void**>(&d), (size_t)b->world + 1) == cudaSuccess) {
            if (g_nccl.AllGather(d, d + 1, 1, ncclUint8, b->comm, b->copy_stream) == ncclSuccess) cudaStreamSynchronize(b->copy_stream);
            cudaFree(d);
        }
    }
    for (void* p : b->ipc_opened) cudaIpcCloseMemHandle(p);
    if (b->own_comm && b->comm) g_nccl.CommDestroy(b->comm);
    cudaFree(b->buf); cudaFree(b->shared); cudaFree(b->rec); cudaFree(b->all); cudaFree(b->d_peer_shared);
    if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
    if (b->ev_start) cudaEventDestroy(b->ev_start);
    if (b->ev_halo) cudaEventDestroy(b->ev_halo);
    for (auto& e : b->ev_t) if (e) cudaEventDestroy(e);
    cudaGetLastError();
    delete b;
    return B200_OK;
}

int b200_bands_create(b200_ctx* ctx, void* nccl_comm, const void* unique_id, int rank, int world, int height, int width,
                      float sigma, int min_val, int max_val, b200_bands** out) {
    if (!out) { set_error("null out pointer"); return B200_ERR_INVALID_ARG; }
    *out = nullptr;
    b200_bands* b = new b200_bands();
    auto fail = [&](int rc) { b200_bands_destroy(b); return rc; };
    int rc = bands_init_common(b, ctx, rank, world, height, width, sigma, min_val, max_val);
    if (rc != B200_OK) return fail(rc);
    b->peer_shared[(size_t)rank] = b->shared;
    if (world == 1) {
        rc = bands_publish_peers(b);
        if (rc != B200_OK) return fail(rc);
        *out = b;
        return B200_OK;
    }
    rc = load_nccl();
    if (rc != B200_OK) return fail(rc);
    auto init = [&]() -> int {
        if (nccl_comm) {
            b->comm = reinterpret_cast<ncclComm_t>(nccl_comm);
            int n = 0, r = -1;
            CB_NCCL(g_nccl.CommCount(b->comm, &n));
            CB_NCCL(g_nccl.CommUserRank(b->comm, &r));
            if (n != world || r != rank) { set_error("communicator is rank %d of %d, expected %d of %d", r, n, rank, world); return B200_ERR_INVALID_ARG; }
        } else {
            if (!unique_id) { set_error("either an ncclComm_t or a unique id (b200_bands_unique_id on rank 0, broadcast by the caller) is needed"); return B200_ERR_INVALID_ARG; }
            ncclUniqueId id;
            memcpy(&id, unique_id, sizeof(id));
            CB_NCCL(g_nccl.CommInitRank(&b->comm, world, id, rank));
            b->own_comm = true;
        }
        // ---- try P2P: exchange IPC handles, map the neighbours' band buffers and everybody's shared block ----
        const char* want = getenv("B200_BANDS_TRANSPORT");
        const bool force_nccl = want && (want[0] == 'n' || want[0] == 'N');
        IpcBlob mine;
        memset(&mine, 0, sizeof(mine));
        int ok = force_nccl ? 0 : 1;
        if (ok && (cudaIpcGetMemHandle(&mine.buf, b->buf) != cudaSuccess || cudaIpcGetMemHandle(&mine.shared, b->shared) != cudaSuccess)) {
            cudaGetLastError();
            ok = 0;
        }
        mine.device = ok ? ctx->device : -1;     // -1: "I cannot do P2P"
        mine.pid = (int)getpid();
        gethostname(mine.host, sizeof(mine.host) - 1);
        std::vector<unsigned char> blob;
        CB_TRY(nccl_allgather_host(b, &mine, sizeof(mine), blob));
        const IpcBlob* peers = reinterpret_cast<const IpcBlob*>(blob.data());
        for (int r = 0; r < world; ++r) ok = ok && peers[r].device >= 0 && strcmp(peers[r].host, mine.host) == 0 && (r == rank || peers[r].pid != mine.pid);
        if (ok) {
            for (int r = 0; r < world && ok; ++r) {
                if (r == rank) continue;
                void* p = nullptr;
                if (cudaIpcOpenMemHandle(&p, peers[r].shared, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
                b->ipc_opened.push_back(p);
                b->peer_shared[(size_t)r] = reinterpret_cast<BandShared*>(p);
                if (r == rank - 1 || r == rank + 1) {
                    void* q = nullptr;
                    if (cudaIpcOpenMemHandle(&q, peers[r].buf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
                    b->ipc_opened.push_back(q);
                    (r == rank - 1 ? b->up_buf : b->down_buf) = reinterpret_cast<uint8_t*>(q);
                }
            }
        }
        // everybody must agree: one more tiny all-gather of the outcome
        std::vector<unsigned char> votes;
        const unsigned char my_vote = ok ? 1 : 0;
        CB_TRY(nccl_allgather_host(b, &my_vote, 1, votes));
        for (unsigned char v : votes) ok = ok && v;
        if (ok) {
            b->transport = 1;
            CB_TRY(bands_publish_peers(b));
        } else {
            for (void* p : b->ipc_opened) cudaIpcCloseMemHandle(p);
            b->ipc_opened.clear();
            b->up_buf = b->down_buf = nullptr;
            b->transport = 2;
        }
        return B200_OK;
    };
    rc = init();
    if (rc != B200_OK) return fail(rc);
    *out = b;
    return B200_OK;
}

int b200_bands_create_group(b200_ctx* const* ctxs, int n_bands, int height, int width, float sigma, int min_val, int max_val,
                            b200_bands** out) {
    if (!ctxs || !out || n_bands < 1) { set_error("bad argument to b200_bands_create_group"); return B200_ERR_INVALID_ARG; }
    for (int i = 0; i < n_bands; ++i) out[i] = nullptr;
    auto fail = [&](int rc) { for (int i = 0; i < n_bands; ++i) { b200_bands_destroy(out[i]); out[i] = nullptr; } return rc; };
    for (int i = 0; i < n_bands; ++i) {
        out[i] = new b200_bands();
        const int rc = bands_init_common(out[i], ctxs[i], i, n_bands, height, width, sigma, min_val, max_val);
        if (rc != B200_OK) return fail(rc);
    }
    // one process: peers are plain pointers; bands on different devices need peer access switched on
    for (int i = 0; i < n_bands; ++i) {
        b200_bands* b = out[i];
        for (int j = 0; j < n_bands; ++j) {
            b->peer_shared[(size_t)j] = out[j]->shared;
            const int di = b->ctx->device, dj = out[j]->ctx->device;
            if (di != dj) {
                cudaSetDevice(di);
                const cudaError_t e = cudaDeviceEnablePeerAccess(dj, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    set_error("no peer access from device %d to %d: %s", di, dj, cudaGetErrorString(e));
                    cudaGetLastError();
                    return fail(B200_ERR_UNSUPPORTED);
                }
                cudaGetLastError();
            }
        }
        if (i > 0) b->up_buf = out[i - 1]->buf;
        if (i + 1 < n_bands) b->down_buf = out[i + 1]->buf;
        b->transport = n_bands > 1 ? 1 : 0;
        const int rc = bands_publish_peers(b);
        if (rc != B200_OK) return fail(rc);
    }
    return B200_OK;
}

int b200_bands_info(const b200_bands* b, int* row0, int* rows, int* halo_rows, int* transport) {
    if (!b) { set_error("null handle"); return B200_ERR_INVALID_ARG; }
    if (row0) *row0 = b->row0;
    if (rows) *rows = b->rows;
    if (halo_rows) *halo_rows = b->halo;
    if (transport) *transport = b->transport;
    return B200_OK;
}

int b200_bands_input(b200_bands* b, uint8_t** d_band_rows) {
    if (!b || !d_band_rows) { set_error("null pointer"); return B200_ERR_INVALID_ARG; }
    *d_band_rows = b->buf + (size_t)b->above * b->width;
    return B200_OK;
}

int b200_bands_set_timing(b200_bands* b, int on) {
    if (!b) { set_error("null handle"); return B200_ERR_INVALID_ARG; }
    b->timing = on != 0;
    return B200_OK;
}

// ---- the three phases of a step (b200_bands_run chains them; a group in ONE process issues each phase for all its bands before
// the next, so that every "ready" store is queued before the waits that need it whatever the hardware queue mapping) -------------
int b200_bands_begin(b200_bands* b, uint8_t* d_edges) {
    if (!b || !d_edges) { set_error("null pointer"); return B200_ERR_INVALID_ARG; }
    if (b->phase != 0) { set_error("b200_bands_begin: the previous step was not finished"); return B200_ERR_INVALID_ARG; }
    CB_CUDA(cudaSetDevice(b->ctx->device));
    b->step++;
    b->cur_edges = d_edges;
    cudaStream_t st = b->ctx->stream;
    mark(b, 0);
    if (b->transport == 1) {
        unsigned int* up = b->rank > 0 ? &b->peer_shared[(size_t)b->rank - 1]->halo_ready[1] : nullptr;
        unsigned int* down = b->rank + 1 < b->world ? &b->peer_shared[(size_t)b->rank + 1]->halo_ready[0] : nullptr;
        band_signal_halo_kernel<<<1, 1, 0, st>>>(up, down, b->step);
        // this step's record list starts empty (the parity buffer was last read two steps ago)
        band_reset_count_kernel<<<1, 1, 0, st>>>(b->shared, (int)(b->step & 1u));
        CB_CUDA(cudaGetLastError());
        b->ctx->launches += 2;
    }
    b->phase = 1;
    return B200_OK;
}

int b200_bands_front(b200_bands* b) {
    if (!b) { set_error("null handle"); return B200_ERR_INVALID_ARG; }
    if (b->phase != 1) { set_error("b200_bands_front without b200_bands_begin"); return B200_ERR_INVALID_ARG; }
    CB_CUDA(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->stream;
    const BandGeom g = geom_of(b, b->cur_edges);
    CB_TRY(band_prepare(b->ctx, g, b->sigma));
    const size_t W = (size_t)b->width;
    const bool has_up = b->rank > 0, has_down = b->rank + 1 < b->world;
    if (b->transport == 1) {
        // halo rows: pulled by the copy engines while the interior rows are computed
        CB_CUDA(cudaEventRecord(b->ev_start, st));
        CB_CUDA(cudaStreamWaitEvent(b->copy_stream, b->ev_start, 0));
        band_wait_halo_kernel<<<1, 1, 0, b->copy_stream>>>(b->shared, has_up ? 1 : 0, has_down ? 1 : 0, b->step);
        CB_CUDA(cudaGetLastError());
        b->ctx->launches++;
        if (has_up) {
            const Geo n = band_geo(b->height, b->world, b->rank - 1, b->halo);
            CB_CUDA(cudaMemcpyAsync(b->buf, b->up_buf + (size_t)(n.above + n.rows - b->halo) * W, (size_t)b->halo * W, cudaMemcpyDefault, b->copy_stream));
        }
        if (has_down) {
            const Geo n = band_geo(b->height, b->world, b->rank + 1, b->halo);
            CB_CUDA(cudaMemcpyAsync(b->buf + (size_t)(b->above + b->rows) * W, b->down_buf + (size_t)n.above * W, (size_t)b->halo * W, cudaMemcpyDefault, b->copy_stream));
        }
        CB_CUDA(cudaEventRecord(b->ev_halo, b->copy_stream));
        // rows whose stencil stays inside the band: [row0 + e_top, row0 + rows - e_bot)
        const int edge = 64 - (2 * b->ctx->gauss.radius + 4) > b->halo ? 64 - (2 * b->ctx->gauss.radius + 4) : b->halo;   // one slab of the front kernel
        int e_top = has_up ? edge : 0, e_bot = has_down ? edge : 0;
        if (!b->split_front || e_top + e_bot >= b->rows) { e_top = has_up ? b->rows : 0; e_bot = 0; if (!has_up && has_down) e_bot = b->rows; }
        const int interior = b->rows - e_top - e_bot;
        if (interior > 0) CB_TRY(band_front_rows(b->ctx, st, g, b->row0 + e_top, interior));
        mark(b, 1);
        CB_CUDA(cudaStreamWaitEvent(st, b->ev_halo, 0));
        mark(b, 2);
        if (e_top > 0) CB_TRY(band_front_rows(b->ctx, st, g, b->row0, e_top));
        if (e_bot > 0) CB_TRY(band_front_rows(b->ctx, st, g, b->row0 + b->rows - e_bot, e_bot));
    } else {
        if (b->transport == 2) {
            CB_NCCL(g_nccl.GroupStart());
            uint8_t* own = b->buf + (size_t)b->above * W;
            if (has_up) {
                CB_NCCL(g_nccl.Send(own, (size_t)b->halo * W, ncclUint8, b->rank - 1, b->comm, st));
                CB_NCCL(g_nccl.Recv(b->buf, (size_t)b->above * W, ncclUint8, b->rank - 1, b->comm, st));
            }
            if (has_down) {
                CB_NCCL(g_nccl.Send(own + (size_t)(b->rows - b->halo) * W, (size_t)b->halo * W, ncclUint8, b->rank + 1, b->comm, st));
                CB_NCCL(g_nccl.Recv(own + (size_t)b->rows * W, (size_t)b->below * W, ncclUint8, b->rank + 1, b->comm, st));
            }
            CB_NCCL(g_nccl.GroupEnd());
        }
        mark(b, 1);
        mark(b, 2);
        CB_TRY(band_front_rows(b->ctx, st, g, b->row0, b->rows));
    }
    CB_TRY(band_label(b->ctx, st, g));
    mark(b, 3);
    CB_TRY(b200_band_boundary_export(b->ctx, b->rows, b->width, b->rec));
    if (b->transport == 1) {
        const int blocks = (b->S + 255) / 256;
        band_compact_kernel<<<blocks, 256, 0, st>>>(b->rec, b->S, b->shared, (int)(b->step & 1u));
        band_signal_records_kernel<<<1, kMaxBands, 0, st>>>(b->d_peer_shared, b->world, b->rank, b->step);
        CB_CUDA(cudaGetLastError());
        b->ctx->launches += 2;
    }
    mark(b, 4);
    b->phase = 2;
    return B200_OK;
}

int b200_bands_finish(b200_bands* b) {
    if (!b) { set_error("null handle"); return B200_ERR_INVALID_ARG; }
    if (b->phase != 2) { set_error("b200_bands_finish without b200_bands_front"); return B200_ERR_INVALID_ARG; }
    CB_CUDA(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->stream;
    const b200_band_record* all = b->rec;
    if (b->transport == 1) {
        CB_CUDA(cudaMemsetAsync(b->all, 0, (size_t)b->world * b->S * sizeof(b200_band_record), st));
        band_gather_kernel<<<b->world, 256, 0, st>>>(b->d_peer_shared, b->S, b->rank, b->step, b->all);
        CB_CUDA(cudaGetLastError());
        b->ctx->launches++;
        all = b->all;
    } else if (b->transport == 2) {
        CB_NCCL(g_nccl.AllGather(b->rec, b->all, (size_t)b->S * sizeof(b200_band_record), ncclUint8, b->comm, st));
        all = b->all;
    }
    mark(b, 5);
    CB_TRY(b200_band_finalize(b->ctx, all, b->world, b->rank, b->rows, b->width, b->cur_edges));
    mark(b, 6);
    b->phase = 0;
    return B200_OK;
}

int b200_bands_run(b200_bands* b, uint8_t* d_edges) {
    CB_TRY(b200_bands_begin(b, d_edges));
    CB_TRY(b200_bands_front(b));
    return b200_bands_finish(b);
}

int b200_bands_run_group(b200_bands* const* bands, int n_bands, uint8_t* const* d_edges) {
    if (!bands || !d_edges || n_bands < 1) { set_error("bad argument to b200_bands_run_group"); return B200_ERR_INVALID_ARG; }
    for (int i = 0; i < n_bands; ++i) CB_TRY(b200_bands_begin(bands[i], d_edges[i]));
    for (int i = 0; i < n_bands; ++i) CB_TRY(b200_bands_front(bands[i]));
    for (int i = 0; i < n_bands; ++i) CB_TRY(b200_bands_finish(bands[i]));
    return B200_OK;
}

int b200_bands_check(b200_bands* b) {
    if (!b) { set_error("null handle"); return B200_ERR_INVALID_ARG; }
    CB_CUDA(cudaSetDevice(b->ctx->device));
    CB_CUDA(cudaStreamSynchronize(b->ctx->stream));
    CB_CUDA(cudaStreamSynchronize(b->copy_stream));
    unsigned int err = 0;
    CB_CUDA(cudaMemcpy(&err, &b->shared->err, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) { set_error("band %d: a peer never signalled (%s)", b->rank, err == 1 ? "halo rows" : "boundary records"); return B200_ERR_CUDA; }
    return B200_OK;
}

int b200_bands_stage_ms(b200_bands* b, float* ms6) {
    if (!b || !ms6) { set_error("null pointer"); return B200_ERR_INVALID_ARG; }
    CB_CUDA(cudaSetDevice(b->ctx->device));
    CB_CUDA(cudaStreamSynchronize(b->ctx->stream));
    for (int i = 0; i < 6; ++i) {
        ms6[i] = 0.f;
        if (cudaEventElapsedTime(&ms6[i], b->ev_t[i], b->ev_t[i + 1]) != cudaSuccess) { cudaGetLastError(); ms6[i] = -1.f; }
    }
    return B200_OK;
}

}  // extern "C"
