/*
 * TEST INFRASTRUCTURE ONLY — CPU oracle for the Canny hot path.
 *
 * A plain-C restatement of the reference's CPU algorithm (/root/reference/src/utils.cpp), one
 * function per reference stage, each citing the lines it follows.  It exists so the CUDA path can be
 * checked on a box where /root/reference is absent.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product (canny_edge_b200/) never
 * does, and has no CPU fallback.
 *
 * Parity is PINNED: tests/test_oracle.py checks this file against (1) every value-pinning vector of the
 * reference's own tests/utils/test_utils.cpp, (2) golden fixtures under tests/golden/ generated from
 * the compiled, unmodified reference (oracle/_ref/libcanny_ref.so, recipe in oracle/Makefile), and
 * (3) when oracle/_ref is present, the compiled reference itself on random inputs.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (NO -march=native / -ffast-math: an FMA-contracted blur
 * differs from the reference by +-1 in a few pixels per Mpix).
 *
 * All arrays are row-major, pitch == width, as the reference's Mat.data note says (utils.cpp:12-15).
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORACLE_EDGE 255 /* utils.h:5 */
#define ORACLE_NOEDGE 0 /* utils.h:6 */

/* utils.cpp:78 — window = 1 + 2*ceil(3*sigma); the product and ceil are evaluated in float
 * (C++ picks the float overloads), the sum converts to int by truncation. */
int oracle_window(float sigma) {
    float t = 3 * sigma;
    return (int)(1 + 2 * ceilf(t));
}

/* utils.cpp:77-95 createGaussianKernel.  w needs oracle_window(sigma) floats. */
int oracle_gaussian_kernel(float sigma, float* w, int* window) {
    int n = oracle_window(sigma);
    int mid = n / 2;
    float total = 0.0f;
    for (int i = 0; i < n; i++) {
        float x = (float)(i - mid);
        /* utils.cpp:87: exp() on a float argument is expf in C++; sqrt(6.2831853) is double, so the
         * denominator and the division are double and the quotient narrows to float. */
        float e = expf(-((x * x) / (2 * sigma * sigma)));
        float p = (float)((double)e / (sqrt(6.2831853) * (double)sigma));
        w[i] = p;
        total += p;
    }
    for (int i = 0; i < n; i++) w[i] /= total; /* utils.cpp:92-94 */
    *window = n;
    return 0;
}

/* utils.cpp:26-68 gaussian: row pass u8->f32 (37-49), column pass f32->i16 with truncation (52-64).
 * Taps outside the image are skipped and the weight sum renormalises. */
int oracle_gaussian(const uint8_t* img, float sigma, int h, int w, int16_t* out) {
    int n = oracle_window(sigma);
    float* k = (float*)malloc(sizeof(float) * (size_t)n);
    float* tmp = (float*)malloc(sizeof(float) * (size_t)h * w);
    if (!k || !tmp) { free(k); free(tmp); return -1; }
    int win;
    oracle_gaussian_kernel(sigma, k, &win);
    int mid = n / 2;

    for (int r = 0; r < h; r++) {
        const uint8_t* row = img + (size_t)r * w;
        for (int c = 0; c < w; c++) {
            float acc = 0, wsum = 0;
            for (int t = -mid; t <= mid; t++) {
                int cc = c + t;
                if (cc >= 0 && cc < w) {
                    acc += ((float)row[cc]) * k[mid + t];
                    wsum += k[mid + t];
                }
            }
            tmp[(size_t)r * w + c] = acc / wsum;
        }
    }
    for (int c = 0; c < w; c++) {
        for (int r = 0; r < h; r++) {
            float acc = 0, wsum = 0;
            for (int t = -mid; t <= mid; t++) {
                int rr = r + t;
                if (rr >= 0 && rr < h) {
                    acc += tmp[(size_t)rr * w + c] * k[mid + t];
                    wsum += k[mid + t];
                }
            }
            out[(size_t)r * w + c] = (int16_t)(acc / wsum);
        }
    }
    free(tmp);
    free(k);
    return 0;
}

/* utils.cpp:106-187 calculateXYGradient.  gx replicates the border horizontally and drops the
 * missing row vertically (114-149); gy = (row below) - (row above), replicating vertically and
 * dropping the missing column (155-186).  Needs h,w >= 2 (the reference reads out of bounds below
 * that).  Results are stored to int16 exactly as the reference's `short` arrays are. */
int oracle_xy_gradient(const int16_t* b, int h, int w, int16_t* gx, int16_t* gy) {
    if (h < 2 || w < 2) return -2;
    for (int r = 0; r < h; r++) {
        for (int c = 0; c < w; c++) {
            size_t i = (size_t)r * w + c;
            int lft = c > 0 ? c - 1 : c, rgt = c < w - 1 ? c + 1 : c;
            int v = 2 * b[(size_t)r * w + rgt] - 2 * b[(size_t)r * w + lft];
            if (r != h - 1) v += b[(size_t)(r + 1) * w + rgt] - b[(size_t)(r + 1) * w + lft];
            if (r != 0) v += b[(size_t)(r - 1) * w + rgt] - b[(size_t)(r - 1) * w + lft];
            gx[i] = (int16_t)v;

            int up = r > 0 ? r - 1 : r, dn = r < h - 1 ? r + 1 : r;
            int u = 2 * b[(size_t)dn * w + c] - 2 * b[(size_t)up * w + c];
            if (c != w - 1) u += b[(size_t)dn * w + c + 1] - b[(size_t)up * w + c + 1];
            if (c != 0) u += b[(size_t)dn * w + c - 1] - b[(size_t)up * w + c - 1];
            gy[i] = (int16_t)u;
        }
    }
    return 0;
}

/* utils.cpp:215-231: the per-pixel angle binning on its own (also used for whole-domain tables). */
static int16_t oracle_angle_of(int gx, int gy) {
    float th = (float)atan2((double)gy, (double)gx);
    th = (float)((double)th * (180 / 3.1415926535));
    if (th < 0) th = 360 + th;
    if ((th >= 22.5 && th < 67.5) || (th >= 202.5 && th < 247.5)) return 45;
    if ((th >= 112.5 && th < 157.5) || (th >= 292.5 && th < 337.5)) return 135;
    if ((th >= 67.5 && th < 112.5) || (th >= 247.5 && th < 292.5)) return 90;
    return 0;
}

/* angle for every (gx,gy) in [-gmax,gmax]^2, row-major by gy then gx: out[(gy+gmax)*(2gmax+1) + gx+gmax] */
int oracle_angle_table(int gmax, int16_t* out) {
    int n = 2 * gmax + 1;
    for (int gy = -gmax; gy <= gmax; gy++)
        for (int gx = -gmax; gx <= gmax; gx++) out[(size_t)(gy + gmax) * n + (gx + gmax)] = oracle_angle_of(gx, gy);
    return 0;
}

/* (int)sqrt((double)n) for n = 0..n_max (utils.cpp:212) */
int oracle_isqrt_table(int n_max, int32_t* out) {
    for (int n = 0; n <= n_max; n++) out[n] = (int)sqrt((double)n);
    return 0;
}

/* utils.cpp:201-236 sobelOperator: magnitude = (int)sqrt(double(gx^2+gy^2)) (212); angle from a double
 * atan2 narrowed to float, scaled by the double 180/PI with PI = 3.1415926535 (utils.h:4), wrapped to
 * [0,360) in float, then binned to 0/45/90/135 (215-231).  Unlike the reference it does not free its
 * input. */
int oracle_sobel(const int16_t* b, int h, int w, int16_t* mag, int16_t* ang) {
    if (h < 2 || w < 2) return -2;
    size_t n = (size_t)h * w;
    int16_t* gx = (int16_t*)malloc(sizeof(int16_t) * n);
    int16_t* gy = (int16_t*)malloc(sizeof(int16_t) * n);
    if (!gx || !gy) { free(gx); free(gy); return -1; }
    oracle_xy_gradient(b, h, w, gx, gy);
    for (size_t i = 0; i < n; i++) {
        int sq = gx[i] * gx[i] + gy[i] * gy[i];
        mag[i] = (int16_t)(int)sqrt((double)sq);
        int16_t a = oracle_angle_of(gx[i], gy[i]);
        ang[i] = a;
    }
    free(gx);
    free(gy);
    return 0;
}

/* utils.cpp:248-308 nonmaximalSuppression.  Neighbour pair per angle: 0 -> left/right (253-265);
 * 45 -> up-right / down-left (266-278); 90 -> up/down (279-291); 135 -> up-left / down-right
 * (292-304).  A neighbour outside the image is ignored; ties suppress (<=).  An angle that is none of
 * the four leaves the output element unwritten in the reference (uninitialised new[]); the oracle
 * writes 0 there — the stage never produces such an angle. */
int oracle_nonmaximal(const int16_t* mag, const int16_t* ang, int h, int w, int16_t* out) {
    for (int r = 0; r < h; r++) {
        for (int c = 0; c < w; c++) {
            size_t i = (size_t)r * w + c;
            int dr, dc;
            switch (ang[i]) {
                case 0: dr = 0; dc = 1; break;
                case 45: dr = -1; dc = 1; break;
                case 90: dr = 1; dc = 0; break;
                case 135: dr = 1; dc = 1; break;
                default: out[i] = 0; continue;
            }
            int keep = 1;
            for (int s = -1; s <= 1; s += 2) {
                int rr = r + s * dr, cc = c + s * dc;
                if (rr < 0 || rr >= h || cc < 0 || cc >= w) continue;
                if (mag[i] <= mag[(size_t)rr * w + cc]) keep = 0;
            }
            out[i] = keep ? mag[i] : ORACLE_NOEDGE;
        }
    }
    return 0;
}

/* utils.cpp:360-427 findEdgePixels: breadth-first flood from `start` through pixels >= lo, marking
 * them EDGE.  The neighbour guards are restated literally, including the `current - width > 0` test on
 * the two upper diagonals (378, 399): for current == width (row 1, column 0) it is false, so that
 * pixel never reaches (row 0, column 1) — the one missing directed link.  The start pixel is not marked
 * visited (361-365). */
int oracle_find_edge_pixels(int16_t* e, uint8_t* visited, int start, int lo, int hi, int h, int w) {
    (void)hi;
    if (visited[start]) return 0;
    int total = h * w;
    /* every pixel is enqueued at most once, plus possibly `start` a second time */
    int* q = (int*)malloc(sizeof(int) * ((size_t)total + 2));
    if (!q) return -1;
    int head = 0, tail = 0;
    q[tail++] = start;
#define ORACLE_TRY(idx)                              \
    do {                                             \
        int j_ = (idx);                              \
        if (e[j_] >= lo && !visited[j_]) {           \
            q[tail++] = j_;                          \
            visited[j_] = 1;                         \
        }                                            \
    } while (0)
    while (head < tail) {
        int cur = q[head++];
        e[cur] = ORACLE_EDGE;
        int col = cur % w;
        if (col > 0) {
            if (cur + w < total) ORACLE_TRY(cur + w - 1);
            if (cur - w > 0) ORACLE_TRY(cur - w - 1);
            ORACLE_TRY(cur - 1);
        }
        if (col < w - 1) {
            if (cur + w < total) ORACLE_TRY(cur + w + 1);
            if (cur - w > 0) ORACLE_TRY(cur - w + 1);
            ORACLE_TRY(cur + 1);
        }
        if (cur + w < total) ORACLE_TRY(cur + w);
        if (cur - w >= 0) ORACLE_TRY(cur - w);
    }
#undef ORACLE_TRY
    free(q);
    return 0;
}

/* utils.cpp:322-342 hysteresis, in place: scan 1 zeroes < lo and floods from every pixel >= hi
 * (327-334); scan 2 zeroes everything still < hi (336-340). */
int oracle_hysteresis(int16_t* e, int h, int w, int lo, int hi) {
    size_t n = (size_t)h * w;
    uint8_t* visited = (uint8_t*)calloc(n ? n : 1, 1);
    if (!visited) return -1;
    for (size_t i = 0; i < n; i++) {
        if (e[i] < lo) e[i] = ORACLE_NOEDGE;
        else if (e[i] >= hi) oracle_find_edge_pixels(e, visited, (int)i, lo, hi, h, w);
    }
    for (size_t i = 0; i < n; i++)
        if (e[i] < hi) e[i] = ORACLE_NOEDGE;
    free(visited);
    return 0;
}

/* utils.cpp:429-492 canny: gaussian -> sobelOperator -> nonmaximalSuppression -> hysteresis
 * (438,452,465,478); display and printing left out.  `edges` gets the int16 0/255 map; the optional
 * planes receive each stage's output (what `steps` would show).  Returns seconds spent in the four
 * stages, or a negative status. */
double oracle_canny(const uint8_t* img, float sigma, int lo, int hi, int h, int w, int16_t* edges,
                    int16_t* blur_out, int16_t* mag_out, int16_t* ang_out, int16_t* nms_out) {
    if (h < 2 || w < 2) return -2.0;
    size_t n = (size_t)h * w;
    int16_t* blur = (int16_t*)malloc(sizeof(int16_t) * n);
    int16_t* mag = (int16_t*)malloc(sizeof(int16_t) * n);
    int16_t* ang = (int16_t*)malloc(sizeof(int16_t) * n);
    int16_t* nms = (int16_t*)malloc(sizeof(int16_t) * n);
    double secs = -1.0;
    if (blur && mag && ang && nms) {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        oracle_gaussian(img, sigma, h, w, blur);
        oracle_sobel(blur, h, w, mag, ang);
        oracle_nonmaximal(mag, ang, h, w, nms);
        if (nms_out) memcpy(nms_out, nms, sizeof(int16_t) * n);
        oracle_hysteresis(nms, h, w, lo, hi);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        secs = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        if (edges) memcpy(edges, nms, sizeof(int16_t) * n);
        if (blur_out) memcpy(blur_out, blur, sizeof(int16_t) * n);
        if (mag_out) memcpy(mag_out, mag, sizeof(int16_t) * n);
        if (ang_out) memcpy(ang_out, ang, sizeof(int16_t) * n);
    }
    free(blur);
    free(mag);
    free(ang);
    free(nms);
    return secs;
}

/* ------------------------------------------------------------------------------------------------
 * Workload generator (NOT part of the reference): the bench's procedural frames, so that bench.py's
 * CPU legs (`--impl reference`, cpu_baseline) build their inputs without loading the product library.
 * Same pure-integer hash as canny_edge_b200/csrc/canny_math.h::synth_pixel;
 * tests/test_oracle.py checks the two generators byte for byte.
 * kind: 0 "shapes", 1 uniform noise, 2 constant 128.
 * ---------------------------------------------------------------------------------------------- */
static uint64_t o_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static uint64_t o_hash4(uint64_t seed, uint64_t f, uint64_t a, uint64_t b) {
    return o_mix64(o_mix64(o_mix64(o_mix64(seed) ^ f) ^ a) ^ b);
}
static uint8_t o_synth_pixel(int kind, uint64_t seed, int frame, int x, int y) {
    if (kind == 2) return 128;
    if (kind == 1) return (uint8_t)(o_hash4(seed ^ 0x7015Eull, (uint64_t)frame, (uint64_t)x, (uint64_t)y) & 255);
    int X = x >> 8, Y = y >> 8, fx = x & 255, fy = y & 255;
    int v00 = (int)(o_hash4(seed, frame, X, Y) & 127), v10 = (int)(o_hash4(seed, frame, X + 1, Y) & 127);
    int v01 = (int)(o_hash4(seed, frame, X, Y + 1) & 127), v11 = (int)(o_hash4(seed, frame, X + 1, Y + 1) & 127);
    int top = v00 * (256 - fx) + v10 * fx, bot = v01 * (256 - fx) + v11 * fx;
    int v = 64 + ((top * (256 - fy) + bot * fy) >> 16);
    int cx = x >> 6, cy = y >> 6;
    for (int dy = -1; dy <= 1; dy++) {
        for (int dx = -1; dx <= 1; dx++) {
            int ux = cx + dx, uy = cy + dy;
            if (ux < 0 || uy < 0) continue;
            uint64_t h = o_hash4(seed ^ 0xD15Cull, frame, ux, uy);
            int ox = (ux << 6) + (int)(h & 63), oy = (uy << 6) + (int)((h >> 6) & 63);
            int r = 6 + (int)((h >> 12) & 31), d = (int)((h >> 20) % 192) - 96;
            int ddx = x - ox, ddy = y - oy;
            if (ddx * ddx + ddy * ddy <= r * r) v += d;
        }
    }
    v += (int)(o_hash4(seed ^ 0xA015Eull, frame, x, y) & 15) - 8;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
int oracle_synth_rows(uint8_t* out, int row0, int rows, int width, int kind, uint64_t seed, int frame) {
    if (!out || rows < 0 || width <= 0) return -1;
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < width; c++) out[(size_t)r * width + c] = o_synth_pixel(kind, seed, frame, c, row0 + r);
    return 0;
}
