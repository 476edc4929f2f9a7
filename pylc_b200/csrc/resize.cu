// Test-time fit resize on the device: cv2.resize(img, (w', h'), INTER_AREA) of tools.adjust_to_tile
// (utils/tools.py:189-206), bit-exact with OpenCV's general area filter for u8 images.
//
// OpenCV 4.x (imgproc/resize.cpp, computeResizeAreaTab + resizeArea_<uchar, float>; third-party, not
// vendored in the reference -- requirements.txt:8 only says cv2 >= 3.4) computes, for a
// down-scale by a non-integer factor:
//     tab(d)      : the run of source cells [s0, s0+n) that overlap destination cell d and their
//                   area fractions alpha_k, evaluated in double and stored as float
//     buf(sy, dx) = (((0 + S[sy][s0]*a0) + S[sy][s0+1]*a1) + ...)          float, mul and add rounded separately
//     sum(dy, dx) = beta0*buf(sy0) ; sum += beta_j*buf(sy0+j)               float, mul and add rounded separately
//     D[dy][dx]   = saturate_cast<uchar>(cvRound(sum))                      round-half-even
// The host builds the tables with the same double arithmetic (pylc_area_table); the kernel replays
// the float operations in the same order with __fmul_rn / __fadd_rn (never contracted to FMA), so
// every output byte equals OpenCV's.  Integer-factor down-scales (OpenCV's separate "fast area"
// code) and any up-scale are refused with PYLC_ERR_GEOMETRY: the caller keeps the host cv2 path.
//
// One thread owns one destination column of a band of destination rows and walks down the source
// rows it needs, holding `sum` in registers: no shared memory, no barrier.  Neighbouring threads
// read neighbouring source bytes (L1-coalesced); every source row is read by exactly one band,
// plus the single row two bands may share.
#include <math.h>

#include "common.cuh"

namespace pylc {

constexpr int kAreaRows = 16;  // destination rows per CTA band

struct AreaArgs {
    const uint8_t *src;
    size_t src_pitch;
    uint8_t *dst;
    size_t dst_pitch;
    int dw, dh;
    const int32_t *xs, *xn, *ys, *yn;
    const float *xa, *ya;
};

template <int CN>
__global__ void __launch_bounds__(128) area_resize_kernel(AreaArgs a) {
    const int dx = blockIdx.x * 128 + threadIdx.x;
    if (dx >= a.dw) return;
    const int sx0 = __ldg(a.xs + dx) * CN, nx = __ldg(a.xn + dx);
    float alpha[PYLC_AREA_TAPS];
#pragma unroll
    for (int k = 0; k < PYLC_AREA_TAPS; ++k) alpha[k] = __ldg(a.xa + (size_t)dx * PYLC_AREA_TAPS + k);
    const int dy0 = blockIdx.y * kAreaRows, dy1 = min(a.dh, dy0 + kAreaRows);
    for (int dy = dy0; dy < dy1; ++dy) {
        const int sy0 = __ldg(a.ys + dy), ny = __ldg(a.yn + dy);
        float sum[CN];
#pragma unroll
        for (int c = 0; c < CN; ++c) sum[c] = 0.f;
        for (int j = 0; j < ny; ++j) {
            const uint8_t *srow = a.src + (size_t)(sy0 + j) * a.src_pitch + sx0;
            const float beta = __ldg(a.ya + (size_t)dy * PYLC_AREA_TAPS + j);
            float buf[CN];
#pragma unroll
            for (int c = 0; c < CN; ++c) buf[c] = 0.f;
#pragma unroll
            for (int k = 0; k < PYLC_AREA_TAPS; ++k) {
                if (k < nx) {
#pragma unroll
                    for (int c = 0; c < CN; ++c)
                        buf[c] = __fadd_rn(buf[c], __fmul_rn((float)__ldg(srow + k * CN + c), alpha[k]));
                }
            }
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                const float t = __fmul_rn(beta, buf[c]);
                sum[c] = j == 0 ? t : __fadd_rn(sum[c], t);
            }
        }
        uint8_t *d = a.dst + (size_t)dy * a.dst_pitch + (size_t)dx * CN;
#pragma unroll
        for (int c = 0; c < CN; ++c) d[c] = (uint8_t)min(255, max(0, __float2int_rn(sum[c])));
    }
}

}  // namespace pylc

using namespace pylc;

// HOST: one axis of computeResizeAreaTab, grouped per destination index.
extern "C" int pylc_area_table(int ssize, int dsize, int32_t *start, int32_t *count, float *weights) {
    if (ssize < 1 || dsize < 1 || !start || !count || !weights) return PYLC_ERR_ARG;
    if (dsize > ssize) return PYLC_ERR_GEOMETRY;   // up-scaling: OpenCV switches to a linear filter
    const double inv_scale = (double)dsize / ssize;
    const double scale = 1. / inv_scale;
    for (int dx = 0; dx < dsize; ++dx) {
        const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        const double cell = fmin(scale, ssize - fsx1);
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        sx2 = sx2 < ssize - 1 ? sx2 : ssize - 1;
        sx1 = sx1 < sx2 ? sx1 : sx2;
        int n = 0, s0 = sx1;
        float *w = weights + (size_t)dx * PYLC_AREA_TAPS;
        for (int k = 0; k < PYLC_AREA_TAPS; ++k) w[k] = 0.f;
        if (sx1 - fsx1 > 1e-3) {
            s0 = sx1 - 1;
            w[n++] = (float)((sx1 - fsx1) / cell);
        }
        for (int sx = sx1; sx < sx2; ++sx) {
            if (n >= PYLC_AREA_TAPS) return PYLC_ERR_GEOMETRY;
            w[n++] = (float)(1.0 / cell);
        }
        if (fsx2 - sx2 > 1e-3) {
            if (n >= PYLC_AREA_TAPS) return PYLC_ERR_GEOMETRY;
            w[n++] = (float)(fmin(fmin(fsx2 - sx2, 1.), cell) / cell);
        }
        if (n == 0 || s0 < 0 || s0 + n > ssize) return PYLC_ERR_GEOMETRY;
        start[dx] = s0;
        count[dx] = n;
    }
    return PYLC_OK;
}

// HOST: does cv2.resize(INTER_AREA) from (W,H) to (w,h) take the general area filter this library
// reproduces?  (identity counts: a 1:1 table copies the bytes.)
extern "C" int pylc_area_supported(int W, int H, int w, int h) {
    if (W < 1 || H < 1 || w < 1 || h < 1 || w > W || h > H) return 0;
    if (w == W && h == H) return 1;
    const double sx = 1. / ((double)w / W), sy = 1. / ((double)h / H);
    const double eps = 2.220446049250313e-16;
    const bool fast = fabs(sx - (int)(sx + 0.5)) < eps && fabs(sy - (int)(sy + 0.5)) < eps;   // "is_area_fast"
    if (fast) return 0;
    return (sx < PYLC_AREA_TAPS - 1) && (sy < PYLC_AREA_TAPS - 1);
}

extern "C" int pylc_fit_resize_area_u8(const uint8_t *src, int H, int W, int ch, size_t src_pitch, uint8_t *dst, int h,
                                       int w, size_t dst_pitch, const int32_t *x_start, const int32_t *x_count,
                                       const float *x_weights, const int32_t *y_start, const int32_t *y_count,
                                       const float *y_weights, pylc_stream_t stream) {
    if (!src || !dst || !x_start || !x_count || !x_weights || !y_start || !y_count || !y_weights) return PYLC_ERR_ARG;
    if ((ch != 1 && ch != 3) || H < 1 || W < 1 || h < 1 || w < 1) return PYLC_ERR_ARG;
    if (src_pitch < (size_t)W * ch || dst_pitch < (size_t)w * ch) return PYLC_ERR_ARG;
    if (!pylc_area_supported(W, H, w, h)) return PYLC_ERR_GEOMETRY;
    AreaArgs a;
    a.src = src; a.src_pitch = src_pitch; a.dst = dst; a.dst_pitch = dst_pitch; a.dw = w; a.dh = h;
    a.xs = x_start; a.xn = x_count; a.xa = x_weights; a.ys = y_start; a.yn = y_count; a.ya = y_weights;
    const dim3 grid((unsigned)((w + 127) / 128), (unsigned)((h + kAreaRows - 1) / kAreaRows));
    if (ch == 1) area_resize_kernel<1><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    else area_resize_kernel<3><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    return finish_launch();
}

// Pitched host -> device upload on the copy engine (cudaMemcpy2DAsync): converts a tightly packed
// decoded image into the 16-byte-pitched device layout without a host-side repack.  `src_host`
// should be pinned for the copy to be asynchronous.
extern "C" int pylc_upload_pitched(void *dst, size_t dst_pitch, const void *src_host, size_t src_pitch, size_t width_bytes,
                                   size_t rows, pylc_stream_t stream) {
    if (!dst || !src_host || width_bytes == 0 || rows == 0 || dst_pitch < width_bytes || src_pitch < width_bytes)
        return PYLC_ERR_ARG;
    cudaError_t e = cudaMemcpy2DAsync(dst, dst_pitch, src_host, src_pitch, width_bytes, rows, cudaMemcpyHostToDevice,
                                      (cudaStream_t)stream);
    return e == cudaSuccess ? PYLC_OK : (int)e;
}
