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
#include "tma.cuh"

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


// ---- TMA + packed-f32x2 form for scale factors below 1.9 (sm_100a) -----------------------------------------
// The fit-resize proper: what tools.adjust_to_tile asks for on 16-byte pitched device images.
//   * A CTA owns 512 destination BYTES of a row (elements e = (dx, c); 128 threads x 4 consecutive elements, so
//     a thread's store is one 32-bit word and a warp's a 128-byte line) and `rb` destination rows.
//   * Its source patch arrives by cp.async.bulk.tensor.2d loads of 8-row boxes of the [H][pitch/4] source
//     tensor, every box on its own mbarrier, requested by one thread two boxes ahead of the row walk.  The patch
//     origin comes from the closed form of computeResizeAreaTab (area_start, a dozen double operations), not
//     from a dependent table load, so the copy engine starts ~1 us into the kernel.  No thread executes a
//     staging instruction or a bounds check; out-of-image parts are zero-filled and only ever meet zero weights.
//   * A thread walks DOWN the source rows: H(r) = the horizontal sums of its four elements on source row r.  Its
//     twelve tap bytes come from 3-4 aligned shared-memory words and byte permutes (see `words` below), and S*a is
//     one FFMA on the bit pattern 2^23+S, two elements per instruction: fma.rn.f32x2 / add.rn.f32x2 -- SASS
//     FFMA2 / FADD2 -- round each half exactly like the scalar FMUL / FADD of OpenCV's sequence.  The last three
//     H rows live in registers; the row loop is unrolled by three so the ring rotates by renaming, without
//     moves.  Destination row i is emitted when source row ys[i] + 2 has been summed (destination rows start on
//     strictly increasing source rows for any down-scale): V = ((b0*H0) + b1*H1) + b2*H2 with the products as
//     FFMA2 against an opaque zero, so the assembler cannot contract a product into the following add (ptxas
//     does fuse mul.rn.f32x2 + add.rn.f32x2, which would round once instead of twice).
//   Measured per-CTA timeline (3000x2000 colour, 720 CTAs in one wave): copy engine started 0.9 us after kernel
//   entry, tables in registers at 1.3 us, first source row summed at 2.7 us (median), last CTA done at 17.6 us;
//   the row walk issues at ~64 % of the SM's slots -- the kernel is bound by its ~75 instructions per source row
//   and quad, not by DRAM (18 MB read, 12 MB written).
//   Every horizontal sum is computed exactly once per CTA and never leaves registers.
constexpr int kA3Elems = 512, kA3BoxRows = 8, kA3MaxBoxes = 16, kA3Ahead = 2;

struct Area3Args {
    uint8_t *dst;
    size_t dst_pitch;
    int dst_row_bytes, dh;           // w * ch, h
    int W, H, w, h;                  // source / destination sizes in pixels
    double scale_x, scale_y;         // 1. / ((double)w / W), 1. / ((double)h / H): computeResizeAreaTab's `scale`, evaluated on the host
    int rb, bw, nbox;                // destination rows per CTA; staged bytes per source row (multiple of 16); boxes per CTA
    const int32_t *xs, *ys;
    const float *xa, *ya;
    float zero;                      // 0.0f the compiler cannot see
    int org_mask;                    // patch origin = first source byte & org_mask (16-byte aligned box start in global memory)
    bool dst_words;                  // destination rows are 4-byte aligned
};

// start[d] of pylc_area_table in closed form: the same double operations in the same order, every one rounded on
// its own (the intrinsics keep the device compiler from contracting d * scale + scale into an FMA, which the host
// build never does), so the host table and the device agree bit for bit.  `scale` = 1. / ((double)dsize / ssize).
__device__ __forceinline__ int area_start(int d, double scale, int ssize) {
    const double fsx1 = __dmul_rn((double)d, scale), fsx2 = __dadd_rn(fsx1, scale);
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    sx2 = sx2 < ssize - 1 ? sx2 : ssize - 1;
    sx1 = sx1 < sx2 ? sx1 : sx2;
    return (__dsub_rn((double)sx1, fsx1) > 1e-3) ? sx1 - 1 : sx1;
}

__device__ __forceinline__ unsigned long long f2_pack(uint32_t lo, uint32_t hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t b;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(addr));
    return b;
}
__device__ __forceinline__ uint32_t cvt_u8_sat(float v) {      // saturate_cast<uchar>(cvRound(v))
    uint32_t o;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(o) : "f"(v));
    return o;
}

#ifdef PYLC_CF_DEBUG
__device__ unsigned long long g_rs_dbg[8 * 2048];
__device__ __forceinline__ unsigned long long rs_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define RDBG(slot) do { const int b_ = blockIdx.y * gridDim.x + blockIdx.x; if (threadIdx.x == 0 && b_ < 2048) g_rs_dbg[b_ * 8 + (slot)] = rs_gtime(); } while (0)
#else
#define RDBG(slot) do { } while (0)
#endif

template <int CN>
__global__ void __launch_bounds__(kAreaThreads, 6) area_resize_x2_kernel(const __grid_constant__ CUtensorMap tm_src, const Area3Args a) {
    extern __shared__ __align__(128) uint8_t s_dyn3[];        // nbox boxes of 8 rows x bw bytes | rb x uint4 {r0, beta0, beta1, beta2}
    __shared__ __align__(8) unsigned long long s_bar[kA3MaxBoxes];
    const int tid = threadIdx.x;
    RDBG(0);
    const int e0 = blockIdx.x * kA3Elems, dy0 = blockIdx.y * a.rb;
    const int nd = min(a.rb, a.dh - dy0);
    const uint32_t src_s = smem_u32(s_dyn3), bar0 = smem_u32(s_bar);
    const uint32_t row_s = src_s + (uint32_t)a.nbox * kA3BoxRows * (uint32_t)a.bw;

    // table loads first (one round trip to L2 / DRAM, consumed after the copy engine has been started)
    int xs_v[4];
    uint32_t alw[4][3], cc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e = min(e0 + 4 * tid + i, a.dst_row_bytes - 1);
        const int dx = e / CN;
        cc[i] = (uint32_t)(e - dx * CN);
        xs_v[i] = __ldg(a.xs + dx);
#pragma unroll
        for (int k = 0; k < 3; ++k) alw[i][k] = __float_as_uint(__ldg(a.xa + (size_t)dx * PYLC_AREA_TAPS + k));
    }
    const int x_tab = __ldg(a.xs + e0 / CN), y_tab = __ldg(a.ys + dy0);      // only for the consistency check below
    int ys_v = 0;
    uint32_t yw[3] = {0u, 0u, 0u};
    if (tid < nd) {                  // rb <= 64 < kAreaThreads: one destination row per thread
        const float *wv = a.ya + (size_t)(dy0 + tid) * PYLC_AREA_TAPS;
        ys_v = __ldg(a.ys + dy0 + tid);
        yw[0] = __float_as_uint(__ldg(wv)), yw[1] = __float_as_uint(__ldg(wv + 1)), yw[2] = __float_as_uint(__ldg(wv + 2));
    }
    // patch origin (every thread: a dozen double operations): source byte of the block's first element, source row of
    // its first destination row
    const int x_lo = area_start(e0 / CN, a.scale_x, a.W), y_lo = area_start(dy0, a.scale_y, a.H);
    const int byte0 = x_lo * CN, s_lo = y_lo;
    // Boxes are requested in the order they are consumed, kA3Ahead boxes ahead of the row walk: with every box of
    // every CTA requested at kernel entry, first boxes queued behind other CTAs' last ones and some CTAs saw their
    // first row 7 us into a 17 us kernel.
    auto issue_box = [&](int j) {
        mbar_arrive_expect_tx(bar0 + 8u * j, (uint32_t)(kA3BoxRows * a.bw));
        tma_load_2d(src_s + (uint32_t)(j * kA3BoxRows * a.bw), &tm_src, (byte0 & a.org_mask) >> 2, s_lo + j * kA3BoxRows, bar0 + 8u * j);
    };
    if (tid == 0) {
        tma_prefetch_desc(&tm_src);
        for (int j = 0; j < a.nbox; ++j) mbar_init(bar0 + 8u * j, 1);
        mbar_fence_init();
        for (int j = 0; j < kA3Ahead && j < a.nbox; ++j) issue_box(j);
    }
    RDBG(1);
    uint32_t sb[4];
    unsigned long long al2[2][3], ad2[2][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) sb[i] = (uint32_t)((xs_v[i] * CN + (int)cc[i]) - (byte0 & a.org_mask));
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            al2[p][k] = f2_pack(alw[2 * p][k], alw[2 * p + 1][k]);
            ad2[p][k] = f2_pack(__float_as_uint(-8388608.f * __uint_as_float(alw[2 * p][k])),       // exact: power-of-two multiples
                                __float_as_uint(-8388608.f * __uint_as_float(alw[2 * p + 1][k])));
        }
    if (tid < nd) sts128(row_s + 16u * tid, make_uint4((uint32_t)(ys_v - y_lo), yw[0], yw[1], yw[2]));
    // the tables must be the ones area_start restates (pylc_area_table's): the patch was fetched for them
    if (tid == 0 && (x_tab != x_lo || y_tab != y_lo)) __trap();
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (sb[i] + 2u * CN + 16u > (uint32_t)a.bw) __trap(); // the host sizes the patch (incl. the word form's 16-byte window); never taken

    // Keep the loop invariants in registers: left alone, the compiler re-derives the addends (an FMUL each), the
    // shared-memory base (S2R + LEA) and the parameters inside every horizontal sum to save registers.
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int k = 0; k < 3; ++k) asm volatile("" : "+l"(al2[p][k]), "+l"(ad2[p][k]));
    uint32_t src_b = src_s, bar_b = bar0, bw = (uint32_t)a.bw;
    asm volatile("" : "+r"(src_b), "+r"(bar_b), "+r"(bw));
    const unsigned long long z2 = f2_pack(__float_as_uint(a.zero), __float_as_uint(a.zero));
    // Word form of the tap fetch.  The twelve bytes of a quad lie within 16 bytes of its first tap (elements of one
    // pixel are adjacent, a pixel boundary inside the quad adds 0 or CN bytes; gray: consecutive columns start 1 or 2
    // bytes apart), so 3-4 aligned words per source row replace 12 byte loads -- the byte loads, two wavefronts each
    // because a warp's taps span more than 128 bytes, kept the LSU pipe busier than the issue slots.  The words are
    // funnel-shifted to the quad's first byte (thread-constant amount), tap k's four bytes d_i + k*CN are picked by
    // ONE byte permute with a thread-constant selector (nibble i = d_i), and a second permute with a literal selector
    // drops each byte into the 2^23 + S bit pattern.  Needs d_3 + 2 <= 7 (gray) / d_3 <= 6 (colour): voted per CTA.
    const uint32_t d1 = sb[1] - sb[0], d2 = sb[2] - sb[0], d3 = sb[3] - sb[0];
    const bool quad_ok = d1 <= d2 && d2 <= d3 && d3 <= (CN == 1 ? 5u : 6u);
    const bool words = __syncthreads_and(quad_ok) != 0;          // (also the barrier: mbarriers initialised, row table written)
    RDBG(2);
    uint32_t sel = d1 << 4 | d2 << 8 | d3 << 12, wofs = sb[0] & ~3u, wsh = (sb[0] & 3u) * 8u;
    asm volatile("" : "+r"(sel), "+r"(wofs), "+r"(wsh));
    auto hsum = [&](int r, unsigned long long &o0, unsigned long long &o1) {
        if ((r & (kA3BoxRows - 1)) == 0) {
            const int jn = r / kA3BoxRows + kA3Ahead;
            if (tid == 0 && jn < a.nbox) issue_box(jn);
            mbar_wait(bar_b + (uint32_t)r, 0);       // 8 bytes per barrier, 8 rows per box
        }
        const uint32_t row = src_b + (uint32_t)r * bw;
        uint32_t p[4][3];
        if (words) {
            uint32_t t[3];
            const uint32_t wa = row + wofs;
            if (CN == 1) {
                const uint32_t w0 = lds32(wa), w1 = lds32(wa + 4), w2 = lds32(wa + 8);
                const uint32_t a0 = __funnelshift_r(w0, w1, wsh), a1 = __funnelshift_r(w1, w2, wsh);
#pragma unroll
                for (int k = 0; k < 3; ++k) t[k] = __byte_perm(a0, a1, sel + 0x1111u * k);
            } else {
                const uint32_t w0 = lds32(wa), w1 = lds32(wa + 4), w2 = lds32(wa + 8), w3 = lds32(wa + 12);
                const uint32_t a0 = __funnelshift_r(w0, w1, wsh), a1 = __funnelshift_r(w1, w2, wsh), a2 = __funnelshift_r(w2, w3, wsh),
                               a3 = __funnelshift_r(w3, w3, wsh);
                t[0] = __byte_perm(a0, a1, sel);
                t[1] = __byte_perm(__funnelshift_r(a0, a1, 24), __funnelshift_r(a1, a2, 24), sel);
                t[2] = __byte_perm(__funnelshift_r(a1, a2, 16), __funnelshift_r(a2, a3, 16), sel);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) p[i][k] = __byte_perm(t[k], 0x4B000000u, 0x7440u + i);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) p[i][k] = 0x4B000000u | lds_u8(row + sb[i] + (uint32_t)(k * CN));
        }
        o0 = f2_fma(f2_pack(p[0][0], p[1][0]), al2[0][0], ad2[0][0]);
        o1 = f2_fma(f2_pack(p[2][0], p[3][0]), al2[1][0], ad2[1][0]);
#pragma unroll
        for (int k = 1; k < 3; ++k) {
            o0 = f2_add(o0, f2_fma(f2_pack(p[0][k], p[1][k]), al2[0][k], ad2[0][k]));
            o1 = f2_add(o1, f2_fma(f2_pack(p[2][k], p[3][k]), al2[1][k], ad2[1][k]));
        }
    };
    const int n_valid = max(0, min(4, a.dst_row_bytes - (e0 + 4 * tid)));
    uint8_t *d = a.dst + (size_t)dy0 * a.dst_pitch + (size_t)(e0 + 4 * tid);
    uint4 rw = lds128(row_s);
    int i = 0;
    // emits destination row i from the H rows (h0: source row r0, h1: r0 + 1, h2: r0 + 2); true when it was the last one
    auto emit = [&](unsigned long long h00, unsigned long long h01, unsigned long long h10, unsigned long long h11, unsigned long long h20,
                    unsigned long long h21) -> bool {
        const unsigned long long b0 = f2_pack(rw.y, rw.y), b1 = f2_pack(rw.z, rw.z), b2 = f2_pack(rw.w, rw.w);
        unsigned long long s0 = f2_fma(h00, b0, z2), s1 = f2_fma(h01, b0, z2);
        s0 = f2_add(s0, f2_fma(h10, b1, z2));
        s1 = f2_add(s1, f2_fma(h11, b1, z2));
        s0 = f2_add(s0, f2_fma(h20, b2, z2));
        s1 = f2_add(s1, f2_fma(h21, b2, z2));
        float v0, v1, v2, v3;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(s0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(v2), "=f"(v3) : "l"(s1));
        const uint32_t o0 = cvt_u8_sat(v0), o1 = cvt_u8_sat(v1), o2 = cvt_u8_sat(v2), o3 = cvt_u8_sat(v3);
        if (n_valid == 4 && a.dst_words) {
            *reinterpret_cast<uint32_t *>(d) = __byte_perm(__byte_perm(o0, o1, 0x0040), __byte_perm(o2, o3, 0x0040), 0x5410);
        } else {
            if (n_valid > 0) d[0] = (uint8_t)o0;
            if (n_valid > 1) d[1] = (uint8_t)o1;
            if (n_valid > 2) d[2] = (uint8_t)o2;
            if (n_valid > 3) d[3] = (uint8_t)o3;
        }
        d += a.dst_pitch;
        if (++i == nd) return true;
        rw = lds128(row_s + 16u * (uint32_t)i);
        return false;
    };
    unsigned long long A0 = 0, A1 = 0, B0 = 0, B1 = 0, C0 = 0, C1 = 0;
    for (int r = 0;; r += 3) {
        hsum(r, A0, A1);
        if (r == 0) RDBG(3);
        if ((int)rw.x + 2 == r && emit(B0, B1, C0, C1, A0, A1)) break;
        hsum(r + 1, B0, B1);
        if ((int)rw.x + 2 == r + 1 && emit(C0, C1, A0, A1, B0, B1)) break;
        hsum(r + 2, C0, C1);
        if ((int)rw.x + 2 == r + 2 && emit(A0, A1, B0, B1, C0, C1)) break;
    }
    RDBG(4);
}

// Returns PYLC_OK / a CUDA error after launching, or -1 when the form does not apply (alignment, scale >= 1.9,
// no tensor-map support): the caller falls through to the per-thread kernels.
static int launch_area_resize_x2(const uint8_t *src, int H, int W, int ch, size_t src_pitch, uint8_t *dst, int h, int w, size_t dst_pitch,
                                 const int32_t *xs, const float *xa, const int32_t *ys, const float *ya, cudaStream_t st) {
    if (((uintptr_t)src % 16) || (src_pitch % 16) || tma_disabled()) return -1;
    const double sx = (double)W / w, sy = (double)H / h;
    if (sx >= 1.9 || sy >= 2.0) return -1;
    Area3Args a;
    a.dst = dst; a.dst_pitch = dst_pitch; a.dst_row_bytes = w * ch; a.dh = h; a.W = W; a.H = H; a.w = w; a.h = h;
    a.xs = xs; a.ys = ys; a.xa = xa; a.ya = ya; a.zero = 0.f;
    a.scale_x = 1. / ((double)w / W);      // pylc_area_table's `scale`, the same expression
    a.scale_y = 1. / ((double)h / H);
    a.dst_words = ((uintptr_t)dst % 4 == 0) && (dst_pitch % 4 == 0);
    // the box of a tensor load must START on a 16-byte boundary of global memory (measured: a 4-byte aligned start
    // faults with "illegal instruction"), so the patch origin is the first source byte rounded down to 16
    a.org_mask = ~15;
    a.bw = (((int)ceil((kA3Elems / ch + 2) * sx) + 2) * ch + 16 + 16 + 16 + 15) & ~15;     // + origin round-down + the word form's window + slack
    // (fuzzed over 10 000 column blocks of 400 random geometries: the largest staged offset reaches this bound minus 16)
    if (a.bw > 1024) return -1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int col_blocks = (a.dst_row_bytes + kA3Elems - 1) / kA3Elems;
    auto kern = ch == 1 ? area_resize_x2_kernel<1> : area_resize_x2_kernel<3>;
    // Destination rows per CTA.  A taller band re-computes fewer boundary rows (a band needs ~rb*sy + 4 source
    // rows) but costs shared memory and resident CTAs; what hurts most is a grid slightly larger than one wave.
    // Model: a wave of n resident CTAs per SM takes (rb*sy + 4) * max(1, n / 6) row times (about six 4-warp CTAs
    // saturate the issue slots); pick the band height with the smallest modelled total.
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    size_t smem = 0;
    double best = 0;
    bool found = false;
    const int cand[5] = {64, 48, 32, 24, 16};
    for (int ci = 0; ci < 5; ++ci) {
        const int rb = cand[ci];
        const int nbox = ((int)ceil(rb * sy) + 4 + kA3BoxRows - 1) / kA3BoxRows;
        const size_t sm = (size_t)nbox * kA3BoxRows * a.bw + (size_t)rb * 16;
        if (nbox > kA3MaxBoxes || sm > 100 * 1024) continue;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kAreaThreads, sm) != cudaSuccess || per_sm < 1) {
            cudaGetLastError();
            continue;
        }
        const long long ctas = (long long)col_blocks * ((h + rb - 1) / rb), slots = (long long)sms * per_sm;
        const double unit = rb * sy + 4.0;
        const long long full = ctas / slots, rem = ctas % slots;
        double t = (double)full * unit * fmax(1.0, per_sm / 6.0);
        if (rem) t += unit * fmax(1.0, (double)((rem + sms - 1) / sms) / 6.0);
        if (!found || t < best) {
            best = t;
            a.rb = rb; a.nbox = nbox; smem = sm;
            found = true;
        }
    }
    if (!found) return -1;
    CUtensorMap tm;
    {
        const uint64_t dims[2] = {(uint64_t)(src_pitch / 4), (uint64_t)H};
        const uint64_t strides[1] = {(uint64_t)src_pitch};
        const uint32_t box[2] = {(uint32_t)(a.bw / 4), (uint32_t)kA3BoxRows};
        if (!tma_encode_u32(&tm, src, 2, dims, strides, box)) return -1;
    }
    const dim3 grid((unsigned)col_blocks, (unsigned)((h + a.rb - 1) / a.rb));
    kern<<<grid, kAreaThreads, smem, st>>>(tm, a);
    return finish_launch();
}

}  // namespace pylc
#ifdef PYLC_CF_DEBUG
extern "C" __attribute__((visibility("default"))) int pylc_debug_read_rs(void *dst, size_t bytes) { return (int)cudaMemcpyFromSymbol(dst, pylc::g_rs_dbg, bytes); }
#endif

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
    {
        const int rc = launch_area_resize_x2(src, H, W, ch, src_pitch, dst, h, w, dst_pitch, x_start, x_weights, y_start, y_weights,
                                             (cudaStream_t)stream);
        if (rc != -1) return rc;
    }
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
