// Per-pixel arithmetic of tools.augment_transform (reference utils/tools.py:452-594) -- the OpenCV calls of
// perspective_shift / channel_shift restated operation by operation, shared by the sm_100a kernel (augment.cu)
// and by a host twin that the tests run on the CPU against OpenCV itself (oracle/augment_host.cpp).
//
//   cv2.warpPerspective(img f32, M, INTER_LINEAR <- flags=INTER_AREA, BORDER_REFLECT_101)   tools.py:581
//       imgproc/imgwarp.cpp WarpPerspectiveInvoker: double precision, per block of bw0 columns the numerators and
//       the denominator at the block's first column plus M[.] * x1 inside it; X = round_half_even((X0 + M0 x1) * 32 / W);
//       remap: 1/32-pixel coordinates, float table (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy fx -- every product and
//       partial sum is exact for 8-bit valued inputs, so the order of the four terms does not matter;
//   cv2.warpPerspective(mask, M, INTER_NEAREST, BORDER_REFLECT_101)                           tools.py:582
//   crop [30 : T-30]^2, cv2.resize(f32, (T,T), INTER_AREA) enlarging = two-tap "area" linear kernel, rows then
//       columns, each product and sum rounded to f32 on its own (no FMA)                      tools.py:585-586
//   cv2.resize(mask f32, INTER_NEAREST): source index min(floor(dx * (T-60)/T), T-61)         tools.py:587-588
//   channel_shift: uint8(clip(int16(img) + shift, 0, 255))                                     tools.py:550-555
// Every floating-point operation goes through an explicitly rounded primitive so that neither nvcc (-fmad) nor the
// host compiler can contract a multiply and an add.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PYLC_HD __host__ __device__ __forceinline__
#else
#define PYLC_HD inline
#endif

namespace pylc_aug {

#if defined(__CUDA_ARCH__)
PYLC_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
PYLC_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
PYLC_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
PYLC_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
PYLC_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
PYLC_HD int round_even(double v) { return __double2int_rn(v); }
#else
// volatile keeps the host compiler from fusing or re-associating across the primitives
PYLC_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
PYLC_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
PYLC_HD double ddiv(double a, double b) { volatile double r = a / b; return r; }
PYLC_HD float fmul(float a, float b) { volatile float r = a * b; return r; }
PYLC_HD float fadd(float a, float b) { volatile float r = a + b; return r; }
PYLC_HD int round_even(double v) { return (int)lrint(v); }      // default rounding mode: to nearest, ties to even
#endif

constexpr int kCrop = 30;            // tools.py:584,587

// saturate_cast<short>: warpPerspective hands remap its integer coordinates as 16-bit values (this also bounds the
// reflection loop below when a denominator crosses zero and the coordinate saturates)
PYLC_HD int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

PYLC_HD int reflect101(int p, int n) {
    if ((unsigned)p < (unsigned)n) return p;
    if (n == 1) return 0;
    do {
        p = p < 0 ? -p : 2 * (n - 1) - p;
    } while ((unsigned)p >= (unsigned)n);
    return p;
}

// columns per block of WarpPerspectiveInvoker for a w x h destination
PYLC_HD int warp_block_width(int w, int h) {
    const int bh0 = h < 16 ? h : 16;
    const int bw0 = 1024 / bh0;
    return bw0 < w ? bw0 : w;
}

// fixed-point source coordinates of destination pixel (x, y): tab = 32 (bilinear) or 1 (nearest)
PYLC_HD void warp_xy(const double *M, int x, int y, int bw0, double tab, int &X, int &Y) {
    const int xb = x / bw0 * bw0;
    const double x0 = (double)xb, x1 = (double)(x - xb), yy = (double)y;
    const double X0 = dadd(dadd(dmul(M[0], x0), dmul(M[1], yy)), M[2]);
    const double Y0 = dadd(dadd(dmul(M[3], x0), dmul(M[4], yy)), M[5]);
    const double W0 = dadd(dadd(dmul(M[6], x0), dmul(M[7], yy)), M[8]);
    double W = dadd(W0, dmul(M[6], x1));
    W = W != 0.0 ? ddiv(tab, W) : 0.0;
    double fX = dmul(dadd(X0, dmul(M[0], x1)), W);
    double fY = dmul(dadd(Y0, dmul(M[3], x1)), W);
    fX = fX < -2147483648.0 ? -2147483648.0 : (fX > 2147483647.0 ? 2147483647.0 : fX);
    fY = fY < -2147483648.0 ? -2147483648.0 : (fY > 2147483647.0 ? 2147483647.0 : fY);
    X = round_even(fX);
    Y = round_even(fY);
}

// taps of the warped image at destination (x, y): four source offsets (row * T + col) and their float weights
struct WarpTaps {
    int o00, o01, o10, o11;
    float w00, w01, w10, w11;
};
PYLC_HD WarpTaps warp_taps(const double *M, int x, int y, int T, int bw0) {
    int X, Y;
    warp_xy(M, x, y, bw0, 32.0, X, Y);
    const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);   // arithmetic shifts: floor for negative coordinates
    const float fx = (float)(X & 31) * 0.03125f, fy = (float)(Y & 31) * 0.03125f;
    const float gx = fadd(1.0f, -fx), gy = fadd(1.0f, -fy);
    const int x0 = reflect101(sx, T), x1 = reflect101(sx + 1, T), y0 = reflect101(sy, T), y1 = reflect101(sy + 1, T);
    WarpTaps t;
    t.o00 = y0 * T + x0, t.o01 = y0 * T + x1, t.o10 = y1 * T + x0, t.o11 = y1 * T + x1;
    t.w00 = fmul(gy, gx), t.w01 = fmul(gy, fx), t.w10 = fmul(fy, gx), t.w11 = fmul(fy, fx);
    return t;
}
PYLC_HD float warp_value(const uint8_t *plane, const WarpTaps &t) {
    return fadd(fadd(fadd(fmul((float)plane[t.o00], t.w00), fmul((float)plane[t.o01], t.w01)), fmul((float)plane[t.o10], t.w10)),
                fmul((float)plane[t.o11], t.w11));
}
PYLC_HD int warp_nearest_offset(const double *M, int x, int y, int T, int bw0) {
    int X, Y;
    warp_xy(M, x, y, bw0, 1.0, X, Y);
    return reflect101(sat_short(Y), T) * T + reflect101(sat_short(X), T);
}

// enlarging INTER_AREA tap of destination index d: first source sample s (second = min(s + 1, ssize - 1)) and the
// second sample's weight f
PYLC_HD void area_up_tap(int d, int ssize, int dsize, int &s, float &f) {
    const double scale = ddiv((double)ssize, (double)dsize), inv = ddiv((double)dsize, (double)ssize);
    s = (int)floor(dmul((double)d, scale));
    f = (float)dadd((double)(d + 1), -dmul((double)(s + 1), inv));
    f = f <= 0.0f ? 0.0f : fadd(f, -floorf(f));
    if (s >= ssize - 1) {
        f = 0.0f;
        s = ssize - 1;
    }
}

// One output pixel (dx, dy) of one augmented copy: ch image bytes (planes T*T apart) and the mask byte.
PYLC_HD void augment_pixel(const uint8_t *img, const uint8_t *mask, int ch, int T, const double *M, int shift, int dx, int dy,
                           uint8_t *out_img, uint8_t *out_mask) {
    const int S = T - 2 * kCrop, bw0 = warp_block_width(T, T);
    int sx, sy;
    float fx, fy;
    area_up_tap(dx, S, T, sx, fx);
    area_up_tap(dy, S, T, sy, fy);
    const int sx1 = sx + 1 < S ? sx + 1 : S - 1, sy1 = sy + 1 < S ? sy + 1 : S - 1;
    const float a0 = fadd(1.0f, -fx), b0 = fadd(1.0f, -fy);
    const WarpTaps t00 = warp_taps(M, sx + kCrop, sy + kCrop, T, bw0), t01 = warp_taps(M, sx1 + kCrop, sy + kCrop, T, bw0);
    const WarpTaps t10 = warp_taps(M, sx + kCrop, sy1 + kCrop, T, bw0), t11 = warp_taps(M, sx1 + kCrop, sy1 + kCrop, T, bw0);
    const size_t TT = (size_t)T * T, o = (size_t)dy * T + dx;
    for (int c = 0; c < ch; ++c) {
        const uint8_t *p = img + c * TT;
        const float r0 = fadd(fmul(warp_value(p, t00), a0), fmul(warp_value(p, t01), fx));
        const float r1 = fadd(fmul(warp_value(p, t10), a0), fmul(warp_value(p, t11), fx));
        const float v = fadd(fmul(r0, b0), fmul(r1, fy));
        int b = (int)(int16_t)v + shift;                     // np.int16(img) truncates; + shift; clip; uint8
        b = b < 0 ? 0 : (b > 255 ? 255 : b);
        out_img[c * TT + o] = (uint8_t)b;
    }
    // mask: nearest resize of the cropped nearest warp
    const double scale = ddiv((double)S, (double)T);
    int mx = (int)floor(dmul((double)dx, scale)), my = (int)floor(dmul((double)dy, scale));
    mx = mx < S - 1 ? mx : S - 1;
    my = my < S - 1 ? my : S - 1;
    out_mask[o] = mask[warp_nearest_offset(M, mx + kCrop, my + kCrop, T, bw0)];
}

}  // namespace pylc_aug
