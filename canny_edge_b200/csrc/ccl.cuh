// ccl.cuh — lock-free union-find primitives shared by hysteresis.cu and band.cu.
//
// A forest lives in an int32 array: slot x holds the parent of x, a root holds itself.  Links always
// point to a SMALLER value (atomicMin), so there are no cycles.  Negative values are terminal:
//   kSuper (-1)  the virtual root every component that contains a seed hangs under;
//   <= -2        band.cu's "this root was claimed by boundary record i" code (-2 - i); still a root,
//                not strong.
#pragma once
#include <stdint.h>

namespace cb {

constexpr int kTile = 64;          // tile edge (pixels)
constexpr int kCclThreads = 256;
constexpr int32_t kSuper = -1;     // virtual root of every strong component
constexpr int32_t kNone = INT32_MIN;  // "no such root"

// ---------------------------------------------------------------------------------------------
// shared-memory union-find (labels are tile-local pixel indices, 0..4095)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int s_find(const volatile int* lab, int x) {
    int p = lab[x];
    while (p != x) { x = p; p = lab[x]; }
    return x;
}
__device__ __forceinline__ void s_union(int* lab, int a, int b) {
    while (true) {
        a = s_find(lab, a);
        b = s_find(lab, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }  // a > b: hang a under b
        int old = atomicMin(&lab[a], b);
        if (old == a) return;
        a = old;  // somebody re-parented a meanwhile: keep uniting what it pointed to with b
    }
}

// ---------------------------------------------------------------------------------------------
// global union-find (labels are frame-relative pixel indices; SUPER = -1 is a root without a slot)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int g_find(const int32_t* parent, int x) {
    while (x >= 0) {
        int p = __ldcg(parent + x);  // L2: other CTAs update these slots with atomics
        if (p == x) break;
        x = p;
    }
    return x;
}
__device__ __forceinline__ void g_union(int32_t* parent, int a, int b) {
    while (true) {
        a = g_find(parent, a);
        b = g_find(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }  // a > b >= SUPER; a is a real slot
        int old = atomicMin(parent + a, b);
        if (old == a) return;
        a = old;
    }
}

// find with path halving (the sparse kernels: chains along a long edge would otherwise be walked again and again).
// Values only ever move towards the root, so the racy plain store is benign; negative values are terminal and never stored.
__device__ __forceinline__ int g_find_halve(int32_t* parent, int x) {
    while (x >= 0) {
        const int p = __ldcg(parent + x);
        if (p == x || p < 0) return p < 0 ? p : x;
        const int gp = __ldcg(parent + p);
        if (gp == p) return p;
        if (gp < 0) return gp;
        parent[x] = gp;
        x = gp;
    }
    return x;
}
__device__ __forceinline__ void g_union_halve(int32_t* parent, int a, int b) {
    while (true) {
        a = g_find_halve(parent, a);
        b = g_find_halve(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }  // a > b >= SUPER; a is a real slot
        const int old = atomicMin(parent + a, b);
        if (old == a) return;
        a = old;
    }
}

}  // namespace cb
