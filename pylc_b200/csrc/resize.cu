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
// A CTA owns a band of destination rows x a strip of destination columns.  It first parks every
// source byte the band needs in shared memory with 16-byte loads issued back to back (one memory
// round trip per CTA instead of one per source row), then one thread per destination column walks
// down the staged rows holding `sum` in registers.  Every source row is read by exactly one band,
// plus the row two bands may share.
#include <math.h>

#include "common.cuh"

namespace pylc {

constexpr int kAreaThreads = 128;
constexpr int kAreaRowBytes = 832;   // staged bytes per source row (52 chunks of 16 B)
constexpr int kAreaMaxRows = 40;     // staged source rows per band

struct AreaArgs {
    const uint8_t *src;
    size_t src_pitch, src_bytes;     // src_bytes: one past the last valid source byte
    uint8_t *dst;
    size_t dst_pitch;
    int dw, dh;
    int cols_cta, rows_cta;          // destination columns / rows per CTA (the host sizes them to the staging buffer)
    int row_stride, stage_rows;      // staging buffer: bytes per staged source row (multiple of 16), rows
    const int32_t *xs, *xn, *ys, *yn;
    const float *xa, *ya;
};

template <int CN, int NX>
__global__ void __launch_bounds__(kAreaThreads) area_resize_kernel(AreaArgs a) {
    extern __shared__ __align__(16) uint8_t s_src[];     // stage_rows x row_stride bytes (sized by the host to the patch)
    const int dx0 = blockIdx.x * a.cols_cta, dx1 = min(a.dw, dx0 + a.cols_cta);
    const int dy0 = blockIdx.y * a.rows_cta, dy1 = min(a.dh, dy0 + a.rows_cta);
    const int x_lo = __ldg(a.xs + dx0), x_hi = __ldg(a.xs + dx1 - 1) + __ldg(a.xn + dx1 - 1);   // source columns [x_lo, x_hi)
    const int s_lo = __ldg(a.ys + dy0), s_hi = __ldg(a.ys + dy1 - 1) + __ldg(a.yn + dy1 - 1);   // source rows    [s_lo, s_hi)
    const int nrows = s_hi - s_lo, row_bytes = (x_hi - x_lo) * CN;

    // ---- stage: row r of the band at s_src[r * kAreaRowBytes + (its global address & 15) ...] ----
    const int cpr = (row_bytes + 15 + 15) / 16;           // chunks per row incl. the alignment slop
    if (nrows > a.stage_rows || cpr * 16 > a.row_stride) __trap();   // the host sizes the patch; never taken
    const uintptr_t src0 = (uintptr_t)a.src, src_end = src0 + a.src_bytes;
    for (int id = threadIdx.x; id < nrows * cpr; id += kAreaThreads) {
        const int r = id / cpr, ck = id - r * cpr;
        const uintptr_t row_addr = src0 + (size_t)(s_lo + r) * a.src_pitch + (size_t)x_lo * CN;
        const uintptr_t g = (row_addr & ~(uintptr_t)15) + 16u * ck;
        uint8_t *d = s_src + r * a.row_stride + 16 * ck;
        if (g >= src0 && g + 16 <= src_end) {
            *reinterpret_cast<uint4 *>(d) = ld_stream16(reinterpret_cast<const void *>(g));
        } else {
            for (int i = 0; i < 16; ++i) d[i] = (g + i >= src0 && g + i < src_end) ? __ldg(reinterpret_cast<const uint8_t *>(g + i)) : 0;
        }
    }
    __syncthreads();

    // ---- filter ---------------------------------------------------------------------------------
    // Running pointers and 32-bit shared-window addresses keep the per-row bookkeeping to a few
    // instructions; NX is the compile-time tap bound (3 for scale factors below 2: the fit-resize case).
    const int dx = dx0 + threadIdx.x;
    if (threadIdx.x >= a.cols_cta || dx >= dx1) return;
    const int nx = __ldg(a.xn + dx);
    float alpha[NX];
#pragma unroll
    for (int k = 0; k < NX; ++k) alpha[k] = __ldg(a.xa + (size_t)dx * PYLC_AREA_TAPS + k);
    const uint32_t sbase = (uint32_t)(__ldg(a.xs + dx) - x_lo) * CN;     // byte offset inside a staged row
    const uint32_t lo0 = (uint32_t)((src0 + (size_t)s_lo * a.src_pitch + (size_t)x_lo * CN) & 15), plo = (uint32_t)(a.src_pitch & 15);
    const int32_t *ys_p = a.ys + dy0, *yn_p = a.yn + dy0;
    const float *ya_p = a.ya + (size_t)dy0 * PYLC_AREA_TAPS;
    uint8_t *d = a.dst + (size_t)dy0 * a.dst_pitch + (size_t)dx * CN;
    // `buf` of the last source row is kept: consecutive destination rows usually share one source row
    // (OpenCV recomputes it, to the same value), which nearly halves the horizontal passes.
    float buf[CN];
    int r_buf = -1;
    for (int dy = dy0; dy < dy1; ++dy, ++ys_p, ++yn_p, ya_p += PYLC_AREA_TAPS, d += a.dst_pitch) {
        const int r0 = __ldg(ys_p) - s_lo, ny = __ldg(yn_p);
        float sum[CN];
#pragma unroll
        for (int c = 0; c < CN; ++c) sum[c] = 0.f;
        for (int j = 0; j < ny; ++j) {
            const int r = r0 + j;
            if (r != r_buf) {
                r_buf = r;
                const uint32_t addr = sbase + (uint32_t)r * (uint32_t)a.row_stride + ((lo0 + (uint32_t)r * plo) & 15u);
#pragma unroll
                for (int c = 0; c < CN; ++c) buf[c] = 0.f;
#pragma unroll
                for (int k = 0; k < NX; ++k) {
                    if (k < nx) {
#pragma unroll
                        for (int c = 0; c < CN; ++c) {
                            const uint32_t b = s_src[addr + k * CN + c];
                            // (float)byte without the conversion pipe: 2^23 + b as a bit pattern, minus 2^23
                            const float px = __uint_as_float(0x4B000000u | b) - 8388608.f;
                            buf[c] = __fadd_rn(buf[c], __fmul_rn(px, alpha[k]));
                        }
                    }
                }
            }
            const float beta = __ldg(ya_p + j);
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                const float t = __fmul_rn(beta, buf[c]);
                sum[c] = j == 0 ? t : __fadd_rn(sum[c], t);
            }
        }
#pragma unroll
        for (int c = 0; c < CN; ++c) d[c] = (uint8_t)min(255, max(0, __float2int_rn(sum[c])));
    }
}


// ---- two-pass form for scale factors below 2 (the fit-resize case: 1.0 <= scale < 2) -------------------
// With at most 3 source cells per destination cell and axis the filter is a fixed 3 x 3 tap structure:
// missing taps have weight 0 in the tables, and adding `0 * x` is an exact no-op in OpenCV's float
// sequence (all terms are >= 0), so the kernel has no tap-count predicates at all.
//   stage  the CTA's source patch -> shared memory (16-byte loads, as above)
//   H      buf[r][e] = ((S*a0) + S*a1) + S*a2 for every staged source row r and every element e = (dx, c)
//          of the CTA's 384 destination bytes per row -- each horizontal sum is computed ONCE (OpenCV
//          recomputes it for every destination row that uses the source row, to the same value).
//          `S * a` is one FFMA on the bit pattern 2^23 + S with addend -(2^23 * a): the exact product
//          rounded once, i.e. what __fmul_rn((float)S, a) gives, without the int -> float conversion.
//   V      a thread walks its three elements down the destination rows with a rolling window of three
//          horizontal sums in registers (consecutive destination rows start 1 or 2 source rows apart), so a
//          destination byte costs ~1.2 shared-memory loads, 3 FMUL + 2 FADD, one saturating convert, one store.
constexpr int kArea2Elems = 384;     // destination bytes per row and CTA: 128 RGB columns or 384 gray columns
constexpr int kArea2NE = kArea2Elems / kAreaThreads;

template <int CN>
__global__ void __launch_bounds__(kAreaThreads) area_resize_lt2_kernel(AreaArgs a) {
    extern __shared__ __align__(16) uint8_t s_dyn2[];
    // layout: [stage_rows + 2][kArea2Elems] float sums | [rows_cta] uint4 {r0, beta0, beta1, beta2} | staged bytes
    float *s_buf = reinterpret_cast<float *>(s_dyn2);
    uint4 *s_row = reinterpret_cast<uint4 *>(s_dyn2 + (size_t)(a.stage_rows + 2) * kArea2Elems * 4);
    uint8_t *s_src = reinterpret_cast<uint8_t *>(s_row + a.rows_cta);
    const int dx0 = blockIdx.x * a.cols_cta, dx1 = min(a.dw, dx0 + a.cols_cta);
    const int dy0 = blockIdx.y * a.rows_cta, dy1 = min(a.dh, dy0 + a.rows_cta);
    const int x_lo = __ldg(a.xs + dx0), x_hi = __ldg(a.xs + dx1 - 1) + __ldg(a.xn + dx1 - 1);
    const int s_lo = __ldg(a.ys + dy0), s_hi = __ldg(a.ys + dy1 - 1) + __ldg(a.yn + dy1 - 1);
    const int nrows = s_hi - s_lo, row_bytes = (x_hi - x_lo) * CN;
    const int cpr = (row_bytes + 15 + 15) / 16;
    if (nrows > a.stage_rows || cpr * 16 > a.row_stride) __trap();   // the host sizes the patch; never taken
    const uintptr_t src0 = (uintptr_t)a.src, src_end = src0 + a.src_bytes;
    {
        int r = threadIdx.x / cpr, ck = threadIdx.x - r * cpr;
        const int dr = kAreaThreads / cpr, dck = kAreaThreads - dr * cpr;
        for (; r < nrows; r += dr, ck += dck) {
            if (ck >= cpr) {
                ck -= cpr;
                if (++r >= nrows) break;
            }
            const uintptr_t row_addr = src0 + (size_t)(s_lo + r) * a.src_pitch + (size_t)x_lo * CN;
            const uintptr_t g = (row_addr & ~(uintptr_t)15) + 16u * ck;
            uint8_t *d = s_src + r * a.row_stride + 16 * ck;
            if (g >= src0 && g + 16 <= src_end) {
                *reinterpret_cast<uint4 *>(d) = ld_stream16(reinterpret_cast<const void *>(g));
            } else {
                for (int i = 0; i < 16; ++i) d[i] = (g + i >= src0 && g + i < src_end) ? __ldg(reinterpret_cast<const uint8_t *>(g + i)) : 0;
            }
        }
    }
    for (int i = threadIdx.x; i < dy1 - dy0; i += kAreaThreads) {
        const float *w = a.ya + (size_t)(dy0 + i) * PYLC_AREA_TAPS;
        s_row[i] = make_uint4((uint32_t)(__ldg(a.ys + dy0 + i) - s_lo), __float_as_uint(__ldg(w)), __float_as_uint(__ldg(w + 1)),
                              __float_as_uint(__ldg(w + 2)));
    }
    // per-element column tables (registers): start byte inside a staged row, three weights and their addends
    uint32_t sb[kArea2NE];
    float al[kArea2NE][3], ad[kArea2NE][3];
    bool on[kArea2NE];
#pragma unroll
    for (int i = 0; i < kArea2NE; ++i) {
        const int e = threadIdx.x + i * kAreaThreads;
        const int dxl = e / CN, c = e - dxl * CN, dx = dx0 + dxl;
        on[i] = dxl < a.cols_cta && dx < dx1;
        const int dxs = on[i] ? dx : dx0;
        sb[i] = (uint32_t)(__ldg(a.xs + dxs) - x_lo) * CN + c;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            al[i][k] = __ldg(a.xa + (size_t)dxs * PYLC_AREA_TAPS + k);
            ad[i][k] = -8388608.f * al[i][k];          // exact: a power-of-two multiple
        }
    }
    __syncthreads();

    // ---- H: one horizontal sum per (staged row, element); rows `nrows`, `nrows + 1` are finite don't-cares for the
    //      0-weight vertical taps of the last destination rows (a 1:1 resize has ONE live tap per row)
    const uint32_t lo0 = (uint32_t)((src0 + (size_t)s_lo * a.src_pitch + (size_t)x_lo * CN) & 15), plo = (uint32_t)(a.src_pitch & 15);
    const uint32_t src_s = (uint32_t)__cvta_generic_to_shared(s_src), buf_s = (uint32_t)__cvta_generic_to_shared(s_buf);
#pragma unroll 2
    for (int r = 0; r <= nrows + 1; ++r) {
        const uint32_t row = src_s + (uint32_t)r * (uint32_t)a.row_stride + ((lo0 + (uint32_t)r * plo) & 15u);
#pragma unroll
        for (int i = 0; i < kArea2NE; ++i) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                uint32_t b;
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(row + sb[i] + (uint32_t)(k * CN)));
                const float prod = __fmaf_rn(__uint_as_float(0x4B000000u | b), al[i][k], ad[i][k]);   // == __fmul_rn((float)b, alpha)
                acc = k == 0 ? prod : __fadd_rn(acc, prod);
            }
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(buf_s + (uint32_t)(r * kArea2Elems + threadIdx.x + i * kAreaThreads) * 4u), "f"(acc) : "memory");
        }
    }
    __syncthreads();

    // ---- V: rolling three-row window per element ---------------------------------------------------------
    float w0[kArea2NE], w1[kArea2NE], w2[kArea2NE];
    int r_prev = -4;
    uint8_t *d = a.dst + (size_t)dy0 * a.dst_pitch + (size_t)dx0 * CN + threadIdx.x;
    const uint32_t row_s = (uint32_t)__cvta_generic_to_shared(s_row);
    for (int i = 0; i < dy1 - dy0; ++i, d += a.dst_pitch) {
        uint4 rw;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rw.x), "=r"(rw.y), "=r"(rw.z), "=r"(rw.w) : "r"(row_s + 16u * i));
        const int r0 = (int)rw.x, adv = r0 - r_prev;
        r_prev = r0;
        const float b0 = __uint_as_float(rw.y), b1 = __uint_as_float(rw.z), b2 = __uint_as_float(rw.w);
        const uint32_t base = buf_s + (uint32_t)(r0 * kArea2Elems + threadIdx.x) * 4u;
#pragma unroll
        for (int e = 0; e < kArea2NE; ++e) {
            const uint32_t p = base + (uint32_t)(e * kAreaThreads) * 4u;
            if (adv == 1) {
                w0[e] = w1[e];
                w1[e] = w2[e];
            } else if (adv == 2) {
                w0[e] = w2[e];
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w1[e]) : "r"(p + kArea2Elems * 4u));
            } else {
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w0[e]) : "r"(p));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w1[e]) : "r"(p + kArea2Elems * 4u));
            }
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w2[e]) : "r"(p + 2u * kArea2Elems * 4u));
            float sum = __fmul_rn(b0, w0[e]);
            sum = __fadd_rn(sum, __fmul_rn(b1, w1[e]));
            sum = __fadd_rn(sum, __fmul_rn(b2, w2[e]));
            uint32_t o;
            asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(o) : "f"(sum));      // saturate_cast<uchar>(cvRound(sum))
            if (on[e]) d[e * kAreaThreads] = (uint8_t)o;
        }
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
    a.src = src; a.src_pitch = src_pitch; a.src_bytes = (size_t)(H - 1) * src_pitch + (size_t)W * ch;
    a.dst = dst; a.dst_pitch = dst_pitch; a.dw = w; a.dh = h;
    a.xs = x_start; a.xn = x_count; a.xa = x_weights; a.ys = y_start; a.yn = y_count; a.ya = y_weights;
    const double sx = (double)W / w, sy = (double)H / h;
    if (sx < 2.0 && sy < 2.0) {
        // two-pass form: 384 destination bytes per row and CTA, up to 16 destination rows
        a.cols_cta = kArea2Elems / ch;
        a.rows_cta = 16;
        a.row_stride = (((int)ceil(a.cols_cta * sx) + 2) * ch + 30 + 15) & ~15;
        a.stage_rows = (int)ceil(a.rows_cta * sy) + 2;
        const size_t smem = (size_t)(a.stage_rows + 2) * kArea2Elems * 4 + (size_t)a.rows_cta * 16 + (size_t)(a.stage_rows + 2) * a.row_stride + 64;   // + slack: 0-weight taps read a few bytes past the last row
        const dim3 grid((unsigned)((w + a.cols_cta - 1) / a.cols_cta), (unsigned)((h + a.rows_cta - 1) / a.rows_cta));
        cudaError_t e = ch == 1 ? cudaFuncSetAttribute(area_resize_lt2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                : cudaFuncSetAttribute(area_resize_lt2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        if (ch == 1) area_resize_lt2_kernel<1><<<grid, kAreaThreads, smem, (cudaStream_t)stream>>>(a);
        else area_resize_lt2_kernel<3><<<grid, kAreaThreads, smem, (cudaStream_t)stream>>>(a);
        return finish_launch();
    }
    // size the CTA's destination patch so that its source footprint fits the staging buffer:
    // a run of n destination cells covers at most n*scale + 2 source cells
    int cols = (int)floor(((kAreaRowBytes - 30) / ch - 2) / sx);
    int rows = (int)floor((kAreaMaxRows - 2) / sy);
    a.cols_cta = cols > kAreaThreads ? kAreaThreads : (cols < 1 ? 1 : cols);
    a.rows_cta = rows > 16 ? 16 : (rows < 1 ? 1 : rows);
    // ... and the staging buffer to the patch, so small footprints leave room for many resident CTAs
    a.row_stride = (((int)ceil(a.cols_cta * sx) + 2) * ch + 30 + 15) & ~15;
    a.stage_rows = (int)ceil(a.rows_cta * sy) + 2;
    if (a.row_stride > kAreaRowBytes) a.row_stride = kAreaRowBytes;
    if (a.stage_rows > kAreaMaxRows) a.stage_rows = kAreaMaxRows;
    const size_t smem = (size_t)a.row_stride * a.stage_rows;
    const dim3 grid((unsigned)((w + a.cols_cta - 1) / a.cols_cta), (unsigned)((h + a.rows_cta - 1) / a.rows_cta));
    const bool few = sx < 2.0;     // a destination cell overlaps at most floor(scale) + 2 source cells
    if (ch == 1) {
        if (few) area_resize_kernel<1, 3><<<grid, kAreaThreads, smem, (cudaStream_t)stream>>>(a);
        else area_resize_kernel<1, PYLC_AREA_TAPS><<<grid, kAreaThreads, smem, (cudaStream_t)stream>>>(a);
    } else {
        if (few) area_resize_kernel<3, 3><<<grid, kAreaThreads, smem, (cudaStream_t)stream>>>(a);
        else area_resize_kernel<3, PYLC_AREA_TAPS><<<grid, kAreaThreads, smem, (cudaStream_t)stream>>>(a);
    }
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
