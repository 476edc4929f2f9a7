// Fused stitch + softmax + argmax + colourise (tools.reconstruct, utils/tools.py:239-313).
//
// Closed form of the reference's in-place band merging, S = T/2:
//   output block (ky,kx) of S x S pixels draws on one quadrant of up to four tiles
//     Hrow(i, ty, x) = raw L[i,0] / L[i,nc-1] at the left / right border,
//                      else (softmax(L[i,kx-1]) + softmax(L[i,kx])) / 2
//     M(y, x)        = Hrow(0) / Hrow(nr-1) at the top / bottom border,
//                      else (softmax(Hrow(ky-1)) + softmax(Hrow(ky))) / 2
//   so corners are raw logits, edges one softmax-average, the interior a double softmax.
// Every logit is read exactly once; a CTA owns a slab of rows of one block, so the block kind
// (how many tiles feed it) is uniform per CTA and there is no divergence.
#include "common.cuh"

namespace pylc {

struct StitchArgs {
    const float *logits;            // contiguous [nr*nc, C, T, T] or NULL
    const float *const *batches;    // device array of batch pointers or NULL
    int tiles_per_batch;
    int nr, nc, C, T, S;
    int overlap;                    // 1: S == T/2, 0: S == T
    int h, w, nbx, nby;
    int rows_per_cta, slabs, upr;   // units (PX pixels) per block row
    uint8_t *labels;
    uint8_t *rgb;
    float *stitched;
};

__device__ __forceinline__ const float *tile_base(const StitchArgs &a, int i, int j) {
    const int k = i * a.nc + j;
    const size_t tile_elems = (size_t)a.C * a.T * a.T;
    if (a.logits) return a.logits + (size_t)k * tile_elems;
    const int b = k / a.tiles_per_batch;
    return a.batches[b] + (size_t)(k - b * a.tiles_per_batch) * tile_elems;
}

template <int PX>
struct Vec;
template <>
struct Vec<2> {
    static __device__ __forceinline__ void load(const float *p, float (&o)[2]) {
        const float2 t = ld_stream_f2(p);
        o[0] = t.x;
        o[1] = t.y;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[2]) {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    }
};
template <>
struct Vec<1> {
    static __device__ __forceinline__ void load(const float *p, float (&o)[1]) { o[0] = __ldg(p); }
    static __device__ __forceinline__ void store(float *p, const float (&v)[1]) { *p = v[0]; }
};

// softmax over the class axis, fp32, split in two steps so that the scale by 1/sum can be folded into
// whatever consumes the probabilities:
//   exp_cls     : v[c] <- 2^(v[c]*k - max*k), returns r = 1/sum.  k = log2(e) for logits; for the SUM of
//                 two probability vectors (an average that was never scaled by 1/2) k = log2(e)/2, which
//                 is bit-identical to halving first because scaling by a power of two is exact.
//                 SUBMAX = false skips the max subtraction: inputs that are probabilities (or sums of
//                 two) lie in [0, 2], so 2^(x*k) cannot overflow and softmax is shift-invariant.
//   combine_cls : a[c] <- (a[c]*ra + b[c]*rb) * w as one FMUL + one FFMA per class.
// All classes of a pixel share the same rounded max*k and the same reciprocals, so neither
// approximation can reorder classes; values stay within ~1e-6 relative of torch's CPU softmax.
template <int CMAX, int PX, bool SUBMAX>
__device__ __forceinline__ void exp_cls(float (&v)[CMAX][PX], int C, float k, float (&r)[PX]) {
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        float ms = 0.f;
        if (SUBMAX) {
            float m = v[0][j];
#pragma unroll
            for (int c = 1; c < CMAX; ++c)
                if (c < C) m = fmaxf(m, v[c][j]);
            ms = -m * k;
        }
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) {
                const float e = ex2_approx(SUBMAX ? fmaf(v[c][j], k, ms) : v[c][j] * k);
                v[c][j] = e;
                s += e;
            }
        r[j] = rcp_approx(s);
    }
}

template <int CMAX, int PX>
__device__ __forceinline__ void combine_cls(float (&a)[CMAX][PX], const float (&ra)[PX], const float (&b)[CMAX][PX],
                                            const float (&rb)[PX], int C, float w) {
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        const float wa = ra[j] * w, wb = rb[j] * w;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) a[c][j] = fmaf(a[c][j], wa, b[c][j] * wb);
    }
}

template <int CMAX, int PX>
__device__ __forceinline__ void load_cls(const float *p, size_t cstride, int C, float (&v)[CMAX][PX]) {
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
        if (c < C) Vec<PX>::load(p + c * cstride, v[c]);
}

// Writes one unit: optional f32 map, first-maximum argmax (np.argmax), label bytes, optional RGB.
template <int CMAX, int PX>
__device__ __forceinline__ void emit_unit(const StitchArgs &a, const uint32_t *s_lut, int C, const float (&m)[CMAX][PX], size_t o) {
    if (a.stitched) {
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) Vec<PX>::store(a.stitched + (size_t)c * a.h * a.w + o, m[c]);
    }
    uint32_t lab[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        float best = m[0][j];
        uint32_t bi = 0;
#pragma unroll
        for (int c = 1; c < CMAX; ++c)
            if (c < C && m[c][j] > best) {  // strict: first maximum wins (np.argmax)
                best = m[c][j];
                bi = c;
            }
        lab[j] = bi;
    }
    if (a.labels) {
        if (PX == 2) *reinterpret_cast<uint16_t *>(a.labels + o) = (uint16_t)(lab[0] | (lab[PX - 1] << 8));
        else a.labels[o] = (uint8_t)lab[0];
    }
    if (a.rgb) {
        uint8_t *d = a.rgb + o * 3;
        if (PX == 2) {
            const uint32_t p0 = s_lut[lab[0]], p1 = s_lut[lab[PX - 1]];
            uint16_t *d2 = reinterpret_cast<uint16_t *>(d);
            d2[0] = (uint16_t)p0;
            d2[1] = (uint16_t)((p0 >> 16) | (p1 << 8));
            d2[2] = (uint16_t)(p1 >> 8);
        } else {
            const uint32_t p0 = s_lut[lab[0]];
            d[0] = (uint8_t)p0;
            d[1] = (uint8_t)(p0 >> 8);
            d[2] = (uint8_t)(p0 >> 16);
        }
    }
}

// NH: tiles per strip (1 border / 2 interior columns); NV: strips (1 border / 2 interior rows)
template <int C_T, int CMAX, int PX, int NH, int NV>
__device__ __forceinline__ void stitch_units(const StitchArgs &a, const uint32_t *s_lut, int ky, int kx, int row0, int nrows) {
    const int C = C_T > 0 ? C_T : a.C;
    const int T = a.T, S = a.S;
    const size_t cstride = (size_t)T * T;
    // source tiles: strip s in {0,1} x column q in {0,1}
    int ti[2], ty0[2], tj[2], tx0[2];
    if (!a.overlap) {
        ti[0] = ky; ty0[0] = 0; tj[0] = kx; tx0[0] = 0;
    } else {
        if (NV == 1) {
            ti[0] = ky == 0 ? 0 : a.nr - 1;
            ty0[0] = ky == 0 ? 0 : S;
        } else {
            ti[0] = ky - 1; ty0[0] = S;
            ti[1] = ky;     ty0[1] = 0;
        }
        if (NH == 1) {
            tj[0] = kx == 0 ? 0 : a.nc - 1;
            tx0[0] = kx == 0 ? 0 : S;
        } else {
            tj[0] = kx - 1; tx0[0] = S;
            tj[1] = kx;     tx0[1] = 0;
        }
    }
    const float *base[NV][NH];
#pragma unroll
    for (int s = 0; s < NV; ++s)
#pragma unroll
        for (int q = 0; q < NH; ++q) base[s][q] = tile_base(a, ti[s], tj[q]) + (size_t)ty0[s] * T + tx0[q];

    // a thread's units are kThreads apart: (row, column) advance by constants with a carry
    int row = threadIdx.x / a.upr, col = threadIdx.x - row * a.upr;
    const int d_row = kThreads / a.upr, d_col = kThreads - d_row * a.upr;
    for (; row < nrows; row += d_row, col += d_col) {
        if (col >= a.upr) {
            col -= a.upr;
            if (++row >= nrows) break;
        }
        const int yl = row0 + row;
        const int xl = col * PX;
        const size_t off = (size_t)yl * T + xl;

        float v[NV][NH][CMAX][PX];
#pragma unroll
        for (int s = 0; s < NV; ++s)
#pragma unroll
            for (int q = 0; q < NH; ++q) load_cls<CMAX, PX>(base[s][q] + off, cstride, C, v[s][q]);

        // strips: at interior columns the two tiles' softmaxes are averaged.  When a second softmax
        // follows (NV == 2) the strip keeps the SUM of the two and the 1/2 moves into that softmax's scale.
        if (NH == 2) {
#pragma unroll
            for (int s = 0; s < NV; ++s) {
                float r0[PX], r1[PX];
                exp_cls<CMAX, PX, true>(v[s][0], C, kLog2e, r0);
                exp_cls<CMAX, PX, true>(v[s][1], C, kLog2e, r1);
                combine_cls<CMAX, PX>(v[s][0], r0, v[s][1], r1, C, NV == 2 ? 1.f : 0.5f);
            }
        }
        if (NV == 2) {
            float r0[PX], r1[PX];
            if (NH == 2) {   // inputs: sums of two probability vectors, in [0, 2]
                exp_cls<CMAX, PX, false>(v[0][0], C, 0.5f * kLog2e, r0);
                exp_cls<CMAX, PX, false>(v[NV - 1][0], C, 0.5f * kLog2e, r1);
            } else {         // left / right border: raw logits
                exp_cls<CMAX, PX, true>(v[0][0], C, kLog2e, r0);
                exp_cls<CMAX, PX, true>(v[NV - 1][0], C, kLog2e, r1);
            }
            combine_cls<CMAX, PX>(v[0][0], r0, v[NV - 1][0], r1, C, 0.5f);
        }
        emit_unit<CMAX, PX>(a, s_lut, C, v[0][0], (size_t)(ky * S + yl) * a.w + (size_t)(kx * S + xl));
    }
}

template <int C_T, int CMAX, int PX>
__global__ void __launch_bounds__(kThreads, 2) stitch_kernel(StitchArgs a, const __grid_constant__ ColourLut lut) {
    __shared__ uint32_t s_lut[PYLC_MAX_CLASSES];
    if (threadIdx.x < PYLC_MAX_CLASSES) s_lut[threadIdx.x] = lut.rgb[threadIdx.x];
    __syncthreads();
    int bid = blockIdx.x;
    const int slab = bid % a.slabs;
    bid /= a.slabs;
    // Blocks are walked from the LAST tile row to the first.  In the tiled pipeline the logits were
    // written tile by tile just before this launch and are several times the L2: what is still cached
    // are the most recently written tiles, so reading those first turns the tail of the producer's
    // stores into L2 hits instead of evicting them unread.  (Ordering the blocks by weight -- interior
    // blocks read four quadrants per pixel, corners one -- was measured and makes no difference.)
    bid = a.nbx * a.nby - 1 - bid;
    const int kx = bid % a.nbx, ky = bid / a.nbx;
    const bool two_h = a.overlap && kx > 0 && kx < a.nc;
    const bool two_v = a.overlap && ky > 0 && ky < a.nr;
    if (two_h) {
        if (two_v) stitch_units<C_T, CMAX, PX, 2, 2>(a, s_lut, ky, kx, slab * a.rows_per_cta, a.rows_per_cta);
        else stitch_units<C_T, CMAX, PX, 2, 1>(a, s_lut, ky, kx, slab * a.rows_per_cta, a.rows_per_cta);
    } else {
        if (two_v) stitch_units<C_T, CMAX, PX, 1, 2>(a, s_lut, ky, kx, slab * a.rows_per_cta, a.rows_per_cta);
        else stitch_units<C_T, CMAX, PX, 1, 1>(a, s_lut, ky, kx, slab * a.rows_per_cta, a.rows_per_cta);
    }
}


}  // namespace pylc

using namespace pylc;

extern "C" int pylc_stitch_argmax_colour(const float *logits, const float *const *tile_batches, int tiles_per_batch,
                                         int nr, int nc, int C, int T, int S, const uint8_t *lut_rgb,
                                         uint8_t *labels, uint8_t *rgb, float *stitched, pylc_stream_t stream) {
    if ((!logits && !tile_batches) || nr < 1 || nc < 1 || T < 1 || S < 1) return PYLC_ERR_ARG;
    if (!logits && tiles_per_batch < 1) return PYLC_ERR_ARG;
    if (rgb && !lut_rgb) return PYLC_ERR_ARG;
    if (C < 1 || C > PYLC_MAX_CLASSES) return PYLC_ERR_CLASSES;
    if (!(S == T || 2 * S == T)) return PYLC_ERR_GEOMETRY;  // the reference is only correct for these (tools.py:235-236)
    StitchArgs a;
    a.logits = logits;
    a.batches = tile_batches;
    a.tiles_per_batch = logits ? nr * nc : tiles_per_batch;
    a.nr = nr; a.nc = nc; a.C = C; a.T = T; a.S = S;
    a.overlap = S < T;
    a.nbx = a.overlap ? nc + 1 : nc;
    a.nby = a.overlap ? nr + 1 : nr;
    a.h = a.nby * S;
    a.w = a.nbx * S;
    a.labels = labels; a.rgb = rgb; a.stitched = stitched;
    const bool px2 = (S % 2 == 0) && C <= 12 && ((uintptr_t)logits % 8 == 0) &&
                     (!stitched || (uintptr_t)stitched % 8 == 0) && (!labels || (uintptr_t)labels % 2 == 0) &&
                     (!rgb || (uintptr_t)rgb % 2 == 0);
    const int PX = px2 ? 2 : 1;
    a.upr = S / PX;
    // One CTA per slab of up to eight units per thread.  Measured on B200 (273 / 45 tiles, C = 9): 8 units
    // 0.395 / 0.078 ms, 4 units 0.416 / 0.080, 2 units 0.449 / 0.084, 16 units 0.428 / 0.082.  Persistent
    // variants (round-robin groups with register double-buffering; weight-balanced contiguous runs per
    // CTA) were slower, 0.43-0.67 / 0.085-0.125 ms: neighbouring CTAs walking neighbouring slabs of the
    // same class planes is what keeps the DRAM pages open, and 16 resident warps hide the softmax.
    int rows = S;
    while (rows > 1 && rows * a.upr > kThreads * 8 && rows % 2 == 0) rows /= 2;
    a.rows_per_cta = rows;
    a.slabs = S / rows;
    ColourLut lut;
    if (lut_rgb) build_colour_lut(lut_rgb, C, &lut);
    else for (int i = 0; i < PYLC_MAX_CLASSES; ++i) lut.rgb[i] = 0;
    const unsigned grid = (unsigned)((long long)a.nbx * a.nby * a.slabs);
    cudaStream_t st = (cudaStream_t)stream;
    if (px2) {
        if (C == 9) stitch_kernel<9, 9, 2><<<grid, kThreads, 0, st>>>(a, lut);
        else if (C == 11) stitch_kernel<11, 11, 2><<<grid, kThreads, 0, st>>>(a, lut);
        else if (C <= 4) stitch_kernel<0, 4, 2><<<grid, kThreads, 0, st>>>(a, lut);
        else if (C <= 8) stitch_kernel<0, 8, 2><<<grid, kThreads, 0, st>>>(a, lut);
        else stitch_kernel<0, 12, 2><<<grid, kThreads, 0, st>>>(a, lut);
    } else {
        if (C <= 12) stitch_kernel<0, 12, 1><<<grid, kThreads, 0, st>>>(a, lut);
        else if (C <= 16) stitch_kernel<0, 16, 1><<<grid, kThreads, 0, st>>>(a, lut);
        else stitch_kernel<0, 32, 1><<<grid, kThreads, 0, st>>>(a, lut);
    }
    return finish_launch();
}
