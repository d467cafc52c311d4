// canny_math.h — the exact integer arithmetic shared by every kernel (and by the host-side table
// tests through b200_direction_host / b200_isqrt_host).  Everything here is bit-defined: no
// transcendental, no data-dependent rounding.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#else
#define CB_HD inline
#endif

namespace cb {

// Direction codes written by the fused kernels (2 bits) and the angle they stand for
// (reference src/utils.cpp:220-231 writes the angle itself as int16).
enum : int { DIR_0 = 0, DIR_45 = 1, DIR_90 = 2, DIR_135 = 3 };

CB_HD int dir_code_to_angle(int code) { return code * 45; }

// Exact integer form of the reference's binning of atan2(gy,gx) (src/utils.cpp:215-231):
//   theta in [22.5,67.5) u [202.5,247.5) -> 45,  [112.5,157.5) u [292.5,337.5) -> 135,
//   [67.5,112.5) u [247.5,292.5) -> 90, else 0.
// tan(22.5 deg) = sqrt(2)-1 and tan(67.5 deg) = sqrt(2)+1, so with ax=|gx|, ay=|gy|:
//   ay < (sqrt2-1) ax  <=>  (ay+ax)^2 < 2 ax^2   -> 0
//   ay > (sqrt2+1) ax  <=>  ay > ax and (ay-ax)^2 > 2 ax^2 -> 90
// otherwise a diagonal, 45 when gx and gy have the same sign.  The boundaries are irrational so no
// integer pair sits on one; the nearest any pair with |g| <= 1020 comes is 1.8e-5 deg, ten times the
// float spacing of theta, so the reference's float rounding never changes the bin
// (tests/test_direction_table.py checks all 2041^2 pairs against the oracle).
// T must hold 2*(2*gmax)^2: int for |g| <= 1020 (the fused path), long long for raw int16 input.
template <typename T>
CB_HD int direction_code(T gx, T gy) {
    T ax = gx < 0 ? -gx : gx;
    T ay = gy < 0 ? -gy : gy;
    T two_ax2 = 2 * ax * ax;
    T s = ay + ax;
    if (s * s < two_ax2) return DIR_0;
    T d = ay - ax;
    if (ay > ax && d * d > two_ax2) return DIR_90;
    if (ax == 0 && ay == 0) return DIR_0;
    return ((gx > 0) == (gy > 0)) ? DIR_45 : DIR_135;
}

// floor(sqrt(n)) for 0 <= n < 2^31, == (int)sqrt((double)n) of src/utils.cpp:212.
// Host version (tests); the device version in the kernels uses MUFU + the same integer fix-up.
inline int isqrt_floor_host(int n) {
    if (n <= 0) return 0;
    int m = 0;
    for (int bit = 1 << 15; bit; bit >>= 1) {
        int t = m | bit;
        if ((long long)t * t <= n) m = t;
    }
    return m;
}

// ---------------------------------------------------------------------------------------------
// Procedural frames: pure 64-bit integer hashing so host and device produce identical bytes.
// ---------------------------------------------------------------------------------------------
CB_HD uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
CB_HD uint64_t hash4(uint64_t seed, uint64_t f, uint64_t a, uint64_t b) {
    return mix64(mix64(mix64(mix64(seed) ^ f) ^ a) ^ b);
}

enum : int { SYNTH_SHAPES = 0, SYNTH_NOISE = 1, SYNTH_CONST = 2 };

// "shapes": bilinear value noise on a 256-px lattice (background 64..191) + one random disc per
// 64x64 cell (3x3 cells visited; radius 6..37, grey delta -96..+95) + per-pixel noise -8..+7.
CB_HD uint8_t synth_pixel(int kind, uint64_t seed, int frame, int x, int y) {
    if (kind == SYNTH_CONST) return 128;
    if (kind == SYNTH_NOISE) return (uint8_t)(hash4(seed ^ 0x7015Eull, (uint64_t)frame, (uint64_t)x, (uint64_t)y) & 255);
    int X = x >> 8, Y = y >> 8, fx = x & 255, fy = y & 255;
    int v00 = (int)(hash4(seed, frame, X, Y) & 127), v10 = (int)(hash4(seed, frame, X + 1, Y) & 127);
    int v01 = (int)(hash4(seed, frame, X, Y + 1) & 127), v11 = (int)(hash4(seed, frame, X + 1, Y + 1) & 127);
    int top = v00 * (256 - fx) + v10 * fx, bot = v01 * (256 - fx) + v11 * fx;
    int v = 64 + ((top * (256 - fy) + bot * fy) >> 16);
    int cx = x >> 6, cy = y >> 6;
    for (int dy = -1; dy <= 1; dy++) {
        for (int dx = -1; dx <= 1; dx++) {
            int ux = cx + dx, uy = cy + dy;
            if (ux < 0 || uy < 0) continue;
            uint64_t h = hash4(seed ^ 0xD15Cull, frame, ux, uy);
            int ox = (ux << 6) + (int)(h & 63), oy = (uy << 6) + (int)((h >> 6) & 63);
            int r = 6 + (int)((h >> 12) & 31), d = (int)((h >> 20) % 192) - 96;
            int ddx = x - ox, ddy = y - oy;
            if (ddx * ddx + ddy * ddy <= r * r) v += d;
        }
    }
    v += (int)(hash4(seed ^ 0xA015Eull, frame, x, y) & 15) - 8;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

}  // namespace cb
