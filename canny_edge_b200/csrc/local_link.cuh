// local_link.cuh — hysteresis linking INSIDE one tile (one slab of class rows x one strip) from two bitmaps.
//
// EXPERIMENTAL (front2.cu's LOCAL_LINK instantiation, off by default: B200_CANNY_LOCAL_LINK=1).  The arithmetic below is
// checked on the CPU against the oracle (tests/test_local_link_cpu.py runs these very functions through a sequential emulation
// of the kernel's phases); the CUDA wiring has not been run on hardware yet.
//
// Idea: the list-driven link kernel (hysteresis.cu) spends one thread, eight class-byte loads and up to four global union-find
// operations on EVERY weak pixel.  But when phase 3b of the front kernel ends, the weak pixels of a tile (64 class rows x 124
// columns) sit in shared memory as a bitmap of 4 words per row, and so can the strong ones.  Linking inside the tile is then
// bit arithmetic on runs plus a union-find over tile-local indices in shared memory; only weak pixels on the tile's border
// (first / last class row of the slab, first / last column of the strip: ~5 % of them) still need the global kernel, which
// finds every neighbour in another tile through the class map exactly as before.
//
// Three phases, each run by one thread per (row, word) and separated by a barrier:
//   init   every run of consecutive weak pixels inside a word gets its lowest pixel as label (SUPER when any pixel of the run
//          touches a strong pixel of this tile), every other pixel of the run points at that head;
//   link   runs are united with the run that continues them in the previous word and with every weak pixel they touch in the
//          row below (src/utils.cpp:360-427's 8-neighbourhood; the one link the reference does not follow, global (0,1) ->
//          (1,0), is left out here too: hysteresis.cu applies it one-way at the very end);
//   emit   every weak pixel looks up its root: the value for its global union-find slot (root's index, or SUPER).
// Labels only ever decrease (atomicMin hooking), so a root is the smallest index of its component, which keeps the global
// forest's "parent <= self" order when local indices are translated to global ones (both are raster orders of the same tile).
#pragma once
#include <stdint.h>

#include "canny_math.h"

namespace cb {
namespace ll {

constexpr int kRows = 64;          // class rows of a slab
constexpr int kWords = 4;          // bitmap words per row (124 columns)
constexpr int kCols = 124;
constexpr int kPitch = 124;        // label array pitch: 64 x 124 int32 = 31 744 B
constexpr int kSuper = -1;

CB_HD int ctz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

CB_HD uint32_t word_at(const uint32_t* m, int r, int k) {
    return (r < 0 || r >= kRows || k < 0 || k >= kWords) ? 0u : m[r * kWords + k];
}
// word k of row r dilated by one pixel to each side, across word borders
CB_HD uint32_t dilate_row(const uint32_t* m, int r, int k) {
    const uint32_t c = word_at(m, r, k);
    return c | (c << 1) | (c >> 1) | (word_at(m, r, k - 1) >> 31) | (word_at(m, r, k + 1) << 31);
}
// lowest run of consecutive set bits of m; *rest = m without it
CB_HD uint32_t lowest_run(uint32_t m, uint32_t* rest) {
    const uint32_t low = m & (0u - m);
    const uint32_t t = m + low;          // the carry runs through the run (wraps to 0 when the run reaches bit 31)
    *rest = m & t;
    return m & ~t;
}

template <typename LabPtr>
CB_HD int find(LabPtr lab, int x) {
    while (x >= 0) {
        const int p = lab[x];
        if (p == x) break;
        x = p;
    }
    return x;  // a root index, or kSuper
}
// AtomicMin: int(int* addr, int value) -> old value
template <typename AtomicMin>
CB_HD void unite(int* lab, int a, int b, AtomicMin amin) {
    while (true) {
        a = find(static_cast<const volatile int*>(lab), a);
        b = find(static_cast<const volatile int*>(lab), b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }   // a > b >= kSuper: a is a real slot
        const int old = amin(lab + a, b);
        if (old == a) return;
        a = old;                                        // re-parented meanwhile: unite what it points to with b
    }
}

// ---- phase "init" for (row r, word k) -------------------------------------------------------------------------------
CB_HD void init_word(const uint32_t* weak, const uint32_t* strong, int* lab, int r, int k) {
    uint32_t m = weak[r * kWords + k];
    if (!m) return;
    const uint32_t near_strong = dilate_row(strong, r - 1, k) | dilate_row(strong, r, k) | dilate_row(strong, r + 1, k);
    const int base = r * kPitch + 32 * k;
    while (m) {
        uint32_t rest;
        uint32_t run = lowest_run(m, &rest);
        m = rest;
        const int head = base + ctz32(run);
        lab[head] = (run & near_strong) ? kSuper : head;
        run &= run - 1;
        while (run) {
            lab[base + ctz32(run)] = head;
            run &= run - 1;
        }
    }
}

// ---- phase "link" for (row r, word k).  skip_01_10: this tile's row r is GLOBAL row 0 and its column 0 is GLOBAL column 0 ----
template <typename AtomicMin>
CB_HD void link_word(const uint32_t* weak, int* lab, int r, int k, bool skip_01_10, AtomicMin amin) {
    uint32_t m = weak[r * kWords + k];
    if (!m) return;
    const int base = r * kPitch + 32 * k;
    const uint32_t below = word_at(weak, r + 1, k);
    const bool below_left = (word_at(weak, r + 1, k - 1) >> 31) != 0;    // weak pixel at column 32k-1 of the row below
    const bool below_right = (word_at(weak, r + 1, k + 1) & 1u) != 0;    // ... at column 32k+32
    const bool left = (word_at(weak, r, k - 1) >> 31) != 0;              // weak pixel at column 32k-1 of this row
    while (m) {
        uint32_t rest;
        const uint32_t run = lowest_run(m, &rest);
        m = rest;
        const int head = base + ctz32(run);
        if ((run & 1u) && left) unite(lab, head, base - 1, amin);        // the run continues the previous word's last run
        if (r + 1 >= kRows) continue;
        uint32_t dil = run | (run << 1) | (run >> 1);
        if (skip_01_10 && k == 0 && !(run & 1u)) dil &= ~1u;             // (0,1) does not reach (1,0) unless through (0,0)
        uint32_t nb = dil & below;
        while (nb) {
            unite(lab, head, base + kPitch + ctz32(nb), amin);
            nb &= nb - 1;
        }
        if ((run & 1u) && below_left) unite(lab, head, base + kPitch - 1, amin);
        if ((run >> 31) && below_right) unite(lab, head, base + kPitch + 32, amin);
    }
}

// ---- phase "emit": value for the global union-find slot of the weak pixel with tile-local index idx ----
//   gbase   launch-relative index of the tile's (row 0, column 0);  width: image width
template <typename LabPtr>
CB_HD int global_parent(LabPtr lab, int idx, int gbase, int width) {
    const int root = find(lab, idx);
    if (root < 0) return root;
    const int rr = root / kPitch, rc = root - rr * kPitch;
    return gbase + rr * width + rc;
}

// weak pixels of word k of row r that the global link kernel still has to visit: the tile's border
//   r_lo, r_hi: first / last class row of this slab;  c_last: last valid column of the strip (123, or less at the image's right edge)
CB_HD uint32_t border_bits(uint32_t w, int r, int k, int r_lo, int r_hi, int c_last) {
    if (r == r_lo || r == r_hi) return w;
    uint32_t mask = 0;
    if (k == 0) mask |= 1u;
    if ((c_last >> 5) == k) mask |= 1u << (c_last & 31);
    return w & mask;
}

}  // namespace ll
}  // namespace cb
