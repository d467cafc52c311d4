// CPU emulation of the tile-local hysteresis linking (canny_edge_b200/csrc/local_link.cuh) + the border-only global link +
// resolve, phase by phase, "threads" run one after another.  Test infrastructure: built by tests/test_local_link_cpu.py with
// g++ and compared with the oracle's hysteresis.  It exercises the very functions the CUDA kernel calls (they are
// __host__ __device__ inline), so the bit arithmetic and the decomposition "local inside a tile, global only for border pixels"
// are checked without a GPU; barriers and atomics are what it cannot check.
#include <stdint.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "local_link.cuh"

namespace {

using namespace cb::ll;

int gfind(std::vector<int>& parent, int x) {
    while (x >= 0) {
        const int p = parent[(size_t)x];
        if (p == x || p < 0) return p < 0 ? p : x;
        x = p;
    }
    return x;
}
void gunion(std::vector<int>& parent, int a, int b) {
    a = gfind(parent, a);
    b = gfind(parent, b);
    if (a == b) return;
    if (a < b) std::swap(a, b);
    parent[(size_t)a] = b;
}

}  // namespace

// cls: H x W bytes, 0 / 1 (weak) / 255 (strong).  out: 0 / 255.  slab_rows: class rows per tile (<= 64), row0: global row of
// cls row 0 (the (0,1)/(1,0) rule applies to GLOBAL rows 0 and 1).  Returns the number of weak pixels the global link visited.
// n_threads > 1: the init / link / emit phases of every tile run on that many host threads with real atomics (a join between
// phases stands in for the kernel's barriers), so the lock-free union-find is exercised under true concurrency as well.
extern "C" long long ll_emulate_mt(const uint8_t* cls, int H, int W, int slab_rows, int row0, uint8_t* out, int n_threads);
extern "C" long long ll_emulate(const uint8_t* cls, int H, int W, int slab_rows, int row0, uint8_t* out) {
    return ll_emulate_mt(cls, H, W, slab_rows, row0, out, 1);
}
extern "C" long long ll_emulate_mt(const uint8_t* cls, int H, int W, int slab_rows, int row0, uint8_t* out, int n_threads) {
    std::vector<int> parent((size_t)H * W, -12345);   // poison: slots of non-weak pixels must never be followed
    std::vector<int> all_list, border_list;
    std::vector<uint32_t> weak(kRows * kWords), strong(kRows * kWords);
    std::vector<int> lab(kRows * kPitch);
    auto amin = [](int* addr, int v) {
        int old = __atomic_load_n(addr, __ATOMIC_RELAXED);
        while (v < old && !__atomic_compare_exchange_n(addr, &old, v, false, __ATOMIC_ACQ_REL, __ATOMIC_RELAXED)) {}
        return old;
    };
    // runs fn(r, k) for all 256 words of a tile, on n_threads threads (interleaved assignment, like lanes of different warps)
    auto for_words = [&](auto fn) {
        if (n_threads <= 1) {
            for (int i = 0; i < kRows * kWords; ++i) fn(i / kWords, i % kWords);
            return;
        }
        std::vector<std::thread> ts;
        for (int t = 0; t < n_threads; ++t)
            ts.emplace_back([&, t] { for (int i = t; i < kRows * kWords; i += n_threads) fn(i / kWords, i % kWords); });
        for (auto& th : ts) th.join();
    };
    for (int y0 = 0; y0 < H; y0 += slab_rows) {
        const int rows = std::min(slab_rows, H - y0);
        for (int x0 = 0; x0 < W; x0 += kCols) {
            const int cols = std::min(kCols, W - x0);
            std::fill(weak.begin(), weak.end(), 0u);
            std::fill(strong.begin(), strong.end(), 0u);
            std::fill(lab.begin(), lab.end(), -777);
            for (int r = 0; r < rows; ++r)
                for (int c = 0; c < cols; ++c) {
                    const int v = cls[(size_t)(y0 + r) * W + x0 + c];
                    if (v == 1) weak[r * kWords + (c >> 5)] |= 1u << (c & 31);
                    if (v == 255) strong[r * kWords + (c >> 5)] |= 1u << (c & 31);
                }
            for_words([&](int r, int k) { init_word(weak.data(), strong.data(), lab.data(), r, k); });
            for_words([&](int r, int k) { link_word(weak.data(), lab.data(), r, k, /*skip_01_10=*/(row0 + y0 + r == 0) && x0 == 0, amin); });
            const int gbase = y0 * W + x0;
            for (int r = 0; r < kRows; ++r)
                for (int k = 0; k < kWords; ++k) {
                    uint32_t w = weak[r * kWords + k];
                    uint32_t b = border_bits(w, r, k, 0, rows - 1, cols - 1);
                    while (w) {
                        const int c = 32 * k + ctz32(w);
                        const int g = gbase + r * W + c;
                        parent[(size_t)g] = global_parent(lab.data(), r * kPitch + c, gbase, W);
                        all_list.push_back(g);
                        if (b & (w & (0u - w))) border_list.push_back(g);
                        w &= w - 1;
                    }
                }
        }
    }
    // the global link kernel (hysteresis.cu: ccl_sparse_link_kernel), restricted to the border list
    for (int g : border_list) {
        const int y = g / W, x = g - y * W;
        const bool has_n = y > 0, has_s = y + 1 < H, has_w = x > 0, has_e = x + 1 < W;
        const bool q01 = (row0 + y == 0) && x == 1, q10 = (row0 + y == 1) && x == 0;
        auto c = [&](int dy, int dx) -> int { return cls[(size_t)(y + dy) * W + x + dx]; };
        const int c_nw = (has_n && has_w) ? c(-1, -1) : 0, c_n = has_n ? c(-1, 0) : 0, c_ne = (has_n && has_e && !q10) ? c(-1, 1) : 0;
        const int c_w = has_w ? c(0, -1) : 0, c_e = has_e ? c(0, 1) : 0;
        const int c_sw = (has_s && has_w && !q01) ? c(1, -1) : 0, c_s = has_s ? c(1, 0) : 0, c_se = (has_s && has_e) ? c(1, 1) : 0;
        if (((c_nw | c_n | c_ne | c_w | c_e | c_sw | c_s | c_se) & 0x80) != 0) gunion(parent, g, -1);
        if (c_e == 1) gunion(parent, g, g + 1);
        if (c_s == 1) {
            gunion(parent, g, g + W);
        } else {
            if (c_sw == 1) gunion(parent, g, g + W - 1);
            if (c_se == 1) gunion(parent, g, g + W + 1);
        }
    }
    if (row0 == 0 && H >= 2 && W >= 2) {
        const int c01 = cls[1], c10 = cls[W];
        if (c10 == 1 && (c01 == 255 || (c01 == 1 && gfind(parent, 1) == -1))) gunion(parent, W, -1);
    }
    for (size_t i = 0; i < (size_t)H * W; ++i) out[i] = cls[i] == 255 ? 255 : 0;
    for (int g : all_list) out[(size_t)g] = gfind(parent, g) == -1 ? 255 : 0;
    return (long long)border_list.size();
}
