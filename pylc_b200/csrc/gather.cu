// Tile gather kernels: image tiles (+ moments), mask tiles fused with palette encode and per-tile
// class histogram, and the normalising f32 gather.
//
// Work decomposition (all three): the used source area ((nH-1)S+T) x ((nW-1)S+T) is cut into
// S x S blocks.  With m = T/S every block belongs to at most m*m destination tiles, so a CTA
// that owns a horizontal slab of one block reads every source byte exactly once and writes it
// (transformed) into each of those tiles.  A thread moves 16 pixels per step with 128-bit
// loads/stores: 16 B (gray), 48 B (interleaved RGB) in, 16 B per destination plane out.
#include "common.cuh"

namespace pylc {

struct GatherGeom {
    const uint8_t *src;
    size_t pitch;
    int T, S, nH, nW, m;
    int nbx, nby;       // S-blocks across / down
    int rows_per_cta;   // slab height (divides S)
    int slabs;          // S / rows_per_cta
    int gpr;            // 16-pixel groups per block row = S / 16
};

struct TileSpan {
    int r_lo, r_hi, c_lo, c_hi;
};

__device__ __forceinline__ TileSpan tile_span(const GatherGeom &g, int by, int bx) {
    TileSpan t;
    t.r_lo = max(0, by - g.m + 1);
    t.r_hi = min(g.nH - 1, by);
    t.c_lo = max(0, bx - g.m + 1);
    t.c_hi = min(g.nW - 1, bx);
    return t;
}

enum { MODE_GRAY = 0, MODE_RGB = 1, MODE_MASK = 2 };

// De-interleave 4 RGB pixels (3 words) into one word per channel.
__device__ __forceinline__ void deinterleave4(uint32_t a, uint32_t b, uint32_t c, uint32_t &r, uint32_t &g,
                                              uint32_t &bl) {
    r = __byte_perm(__byte_perm(a, b, 0x0630), c, 0x5210);
    g = __byte_perm(__byte_perm(a, b, 0x0741), c, 0x6210);
    bl = __byte_perm(__byte_perm(a, b, 0x0052), c, 0x7410);
}

// ------------------------------------------------------------------------------------------------
// image gather: u8 -> u8 tiles (+ sum, sum of squares per destination tile and channel)
// ------------------------------------------------------------------------------------------------
template <int CH, bool ALIGNED, bool STATS>
__global__ void __launch_bounds__(kThreads) gather_img_kernel(GatherGeom g, uint8_t *__restrict__ dst,
                                                              unsigned long long *__restrict__ stat) {
    __shared__ unsigned long long s_sum[CH * 2];
    int bid = blockIdx.x;
    const int slab = bid % g.slabs;
    bid /= g.slabs;
    const int bx = bid % g.nbx;
    const int by = bid / g.nbx;
    const TileSpan ts = tile_span(g, by, bx);
    if (STATS && threadIdx.x < CH * 2) s_sum[threadIdx.x] = 0;
    if (STATS) __syncthreads();

    uint32_t s1[CH], s2[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) s1[k] = s2[k] = 0;

    const int units = g.rows_per_cta * g.gpr;
    const size_t TT = (size_t)g.T * g.T;
    for (int u = threadIdx.x; u < units; u += kThreads) {
        const int row = u / g.gpr;
        const int grp = u - row * g.gpr;
        const int ly = slab * g.rows_per_cta + row;  // row inside the S-block
        const int lx = grp * 16;
        const int y = by * g.S + ly;
        const int x = bx * g.S + lx;
        const uint8_t *p = g.src + (size_t)y * g.pitch + (size_t)x * CH;
        uint4 o[CH];
        if (CH == 1) {
            o[0] = ld16<ALIGNED>(p);
        } else {
            uint4 q0 = ld16<ALIGNED>(p), q1 = ld16<ALIGNED>(p + 16), q2 = ld16<ALIGNED>(p + 32);
            uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
            uint32_t r[4], gg[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) deinterleave4(w[3 * k], w[3 * k + 1], w[3 * k + 2], r[k], gg[k], b[k]);
            o[0] = make_uint4(r[0], r[1], r[2], r[3]);
            o[CH > 1 ? 1 : 0] = make_uint4(gg[0], gg[1], gg[2], gg[3]);
            o[CH > 2 ? 2 : 0] = make_uint4(b[0], b[1], b[2], b[3]);
        }
        if (STATS) {
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const uint32_t ww[4] = {o[k].x, o[k].y, o[k].z, o[k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s1[k] = __dp4a(ww[j], 0x01010101u, s1[k]);
                    s2[k] = __dp4a(ww[j], ww[j], s2[k]);
                }
            }
        }
        for (int r = ts.r_lo; r <= ts.r_hi; ++r) {
            const int ty = (by - r) * g.S + ly;
            for (int c = ts.c_lo; c <= ts.c_hi; ++c) {
                const int tx = (bx - c) * g.S + lx;
                uint8_t *d = dst + ((size_t)(r * g.nW + c) * CH) * TT + (size_t)ty * g.T + tx;
#pragma unroll
                for (int k = 0; k < CH; ++k) st_stream16(d + k * TT, o[k]);
            }
        }
    }
    if (STATS) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            uint32_t a = __reduce_add_sync(0xFFFFFFFFu, s1[k] & 0xFFFFu) ;
            uint32_t ah = __reduce_add_sync(0xFFFFFFFFu, s1[k] >> 16);
            unsigned long long t1 = (unsigned long long)a + ((unsigned long long)ah << 16);
            uint32_t b = __reduce_add_sync(0xFFFFFFFFu, s2[k] & 0xFFFFu);
            uint32_t bh = __reduce_add_sync(0xFFFFFFFFu, s2[k] >> 16);
            unsigned long long t2 = (unsigned long long)b + ((unsigned long long)bh << 16);
            if ((threadIdx.x & 31) == 0) {
                atomicAdd(&s_sum[k * 2], t1);
                atomicAdd(&s_sum[k * 2 + 1], t2);
            }
        }
        __syncthreads();
        const int nt_c = ts.c_hi - ts.c_lo + 1;
        const int nt = (ts.r_hi - ts.r_lo + 1) * nt_c;
        for (int i = threadIdx.x; i < nt * CH * 2; i += kThreads) {
            const int t = i / (CH * 2), k = i - t * (CH * 2);
            const int r = ts.r_lo + t / nt_c, c = ts.c_lo + t % nt_c;
            atomicAdd(&stat[(size_t)(r * g.nW + c) * CH * 2 + k], s_sum[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// mask gather + palette encode + per-tile histogram
// ------------------------------------------------------------------------------------------------
template <bool ALIGNED, bool WIDE, bool HIST>
__global__ void __launch_bounds__(kThreads)
    gather_mask_kernel(GatherGeom g, const __grid_constant__ PaletteHash ph, int C, uint8_t *__restrict__ dst,
                       long long *__restrict__ px_dist) {
    __shared__ uint32_t s_tab[256];
    __shared__ unsigned s_hist[PYLC_MAX_CLASSES];
    s_tab[threadIdx.x] = ph.tab[threadIdx.x];
    if (threadIdx.x < PYLC_MAX_CLASSES) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t mul = ph.mul;

    int bid = blockIdx.x;
    const int slab = bid % g.slabs;
    bid /= g.slabs;
    const int bx = bid % g.nbx;
    const int by = bid / g.nbx;
    const TileSpan ts = tile_span(g, by, bx);

    ClassCounter<WIDE> cc;
    cc.reset();
    const int units = g.rows_per_cta * g.gpr;
    const size_t TT = (size_t)g.T * g.T;
    int since_flush = 0;
    for (int base = 0; base < units; base += kThreads) {
        if (HIST && ++since_flush > (WIDE ? 15 : 63)) {  // warp-uniform: keeps the packed fields from overflowing
            flush_counter<WIDE>(cc, C, s_hist);
            cc.reset();
            since_flush = 1;
        }
        const int u = base + threadIdx.x;
        if (u >= units) continue;
        const int row = u / g.gpr;
        const int grp = u - row * g.gpr;
        const int ly = slab * g.rows_per_cta + row;
        const int lx = grp * 16;
        const int y = by * g.S + ly;
        const int x = bx * g.S + lx;
        const uint8_t *p = g.src + (size_t)y * g.pitch + (size_t)x * 3;
        uint4 q0 = ld16<ALIGNED>(p), q1 = ld16<ALIGNED>(p + 16), q2 = ld16<ALIGNED>(p + 32);
        const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
        uint32_t ow[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t a = w[3 * k], b = w[3 * k + 1], c = w[3 * k + 2];
            const uint32_t c0 = encode_key(a, s_tab, mul);
            const uint32_t c1 = encode_key(__funnelshift_r(a, b, 24), s_tab, mul);
            const uint32_t c2 = encode_key(__funnelshift_r(b, c, 16), s_tab, mul);
            const uint32_t c3 = encode_key(c >> 8, s_tab, mul);
            if (HIST) {
                cc.add(c0);
                cc.add(c1);
                cc.add(c2);
                cc.add(c3);
            }
            ow[k] = c0 + (c1 << 8) + (c2 << 16) + (c3 << 24);
        }
        if (HIST) cc.end_unit();
        const uint4 o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        for (int r = ts.r_lo; r <= ts.r_hi; ++r) {
            const int ty = (by - r) * g.S + ly;
            for (int c = ts.c_lo; c <= ts.c_hi; ++c) {
                const int tx = (bx - c) * g.S + lx;
                st_stream16(dst + (size_t)(r * g.nW + c) * TT + (size_t)ty * g.T + tx, o);
            }
        }
    }
    if (HIST) {
        flush_counter<WIDE>(cc, C, s_hist);
        __syncthreads();
        const int nt_c = ts.c_hi - ts.c_lo + 1;
        const int nt = (ts.r_hi - ts.r_lo + 1) * nt_c;
        for (int i = threadIdx.x; i < nt * C; i += kThreads) {
            const int t = i / C, k = i - t * C;
            const int r = ts.r_lo + t / nt_c, c = ts.c_lo + t % nt_c;
            if (s_hist[k]) atomicAdd((unsigned long long *)&px_dist[(size_t)(r * g.nW + c) * C + k],
                                     (unsigned long long)s_hist[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// normalising gather: u8 source -> network-ready f32 tiles (models/model.py:416-445, 376-377)
// ------------------------------------------------------------------------------------------------
struct NormParams {
    float mean[3], std[3];
    float post_div;  // 255 (models/model.py:435,445) or 1 (grayscale `default` branch, 431-432)
    int out_ch;
};

// 16 u8 -> 16 f32 through a 256-entry table of exactly rounded (x - mean) / std / post_div.
__device__ __forceinline__ void store_norm16(float *d, uint4 v, const float *lut) {
    const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 f;
        f.x = lut[ww[j] & 0xFF];
        f.y = lut[(ww[j] >> 8) & 0xFF];
        f.z = lut[(ww[j] >> 16) & 0xFF];
        f.w = lut[ww[j] >> 24];
        st_stream_f4(d + 4 * j, f);
    }
}

template <int CH, bool ALIGNED>
__global__ void __launch_bounds__(kThreads) gather_norm_kernel(GatherGeom g, NormParams np, float *__restrict__ dst) {
    // IEEE sub/div/div in the reference's order, so the f32 tiles are bit-equal to
    // ((x - mean) / std) / 255 evaluated by torch on the CPU.
    __shared__ float s_lut[CH][256];
#pragma unroll
    for (int k = 0; k < CH; ++k)
        s_lut[k][threadIdx.x] = __fdiv_rn(__fdiv_rn(__fsub_rn((float)threadIdx.x, np.mean[k]), np.std[k]), np.post_div);
    __syncthreads();
    int bid = blockIdx.x;
    const int slab = bid % g.slabs;
    bid /= g.slabs;
    const int bx = bid % g.nbx;
    const int by = bid / g.nbx;
    const TileSpan ts = tile_span(g, by, bx);
    const int units = g.rows_per_cta * g.gpr;
    const size_t TT = (size_t)g.T * g.T;
    for (int u = threadIdx.x; u < units; u += kThreads) {
        const int row = u / g.gpr;
        const int grp = u - row * g.gpr;
        const int ly = slab * g.rows_per_cta + row;
        const int lx = grp * 16;
        const uint8_t *p = g.src + (size_t)(by * g.S + ly) * g.pitch + (size_t)(bx * g.S + lx) * CH;
        uint4 o[3];
        if (CH == 1) {
            o[0] = ld16<ALIGNED>(p);
        } else {
            uint4 q0 = ld16<ALIGNED>(p), q1 = ld16<ALIGNED>(p + 16), q2 = ld16<ALIGNED>(p + 32);
            uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
            uint32_t r[4], gg[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) deinterleave4(w[3 * k], w[3 * k + 1], w[3 * k + 2], r[k], gg[k], b[k]);
            o[0] = make_uint4(r[0], r[1], r[2], r[3]);
            o[1] = make_uint4(gg[0], gg[1], gg[2], gg[3]);
            o[2] = make_uint4(b[0], b[1], b[2], b[3]);
        }
        for (int r = ts.r_lo; r <= ts.r_hi; ++r) {
            const int ty = (by - r) * g.S + ly;
            for (int c = ts.c_lo; c <= ts.c_hi; ++c) {
                const int tx = (bx - c) * g.S + lx;
                float *d = dst + ((size_t)(r * g.nW + c) * np.out_ch) * TT + (size_t)ty * g.T + tx;
                if (CH == 1) {
                    for (int k = 0; k < np.out_ch; ++k) store_norm16(d + k * TT, o[0], s_lut[0]);
                } else {
#pragma unroll
                    for (int k = 0; k < 3; ++k) store_norm16(d + k * TT, o[k], s_lut[CH == 3 ? k : 0]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int make_geom(const uint8_t *src, int H, int W, int ch, size_t pitch, int T, int S, GatherGeom *g) {
    if (!src || H <= 0 || W <= 0 || T <= 0 || S <= 0) return PYLC_ERR_ARG;
    if (pitch < (size_t)W * ch) return PYLC_ERR_ARG;
    if (T % 16 || S % 16 || T % S || T > 4096) return PYLC_ERR_GEOMETRY;
    int nH, nW;
    pylc_tile_grid(H, W, T, S, &nH, &nW);
    g->src = src;
    g->pitch = pitch;
    g->T = T;
    g->S = S;
    g->nH = nH;
    g->nW = nW;
    g->m = T / S;
    g->nbx = nW > 0 ? nW - 1 + g->m : 0;
    g->nby = nH > 0 ? nH - 1 + g->m : 0;
    g->gpr = S / 16;
    // slab height: at most 8 units per thread, at least 16 rows when the block allows it
    int rows = S;
    while (rows > 1 && (rows * g->gpr > kThreads * 8 || rows > 64) && rows % 2 == 0) rows /= 2;
    g->rows_per_cta = rows;
    g->slabs = S / rows;
    return PYLC_OK;
}

static bool aligned16(const void *p, size_t pitch) { return ((uintptr_t)p % 16 == 0) && (pitch % 16 == 0); }

}  // namespace pylc

using namespace pylc;

extern "C" int pylc_tile_grid(int H, int W, int T, int S, int *nH, int *nW) {
    if (T <= 0 || S <= 0 || !nH || !nW) return PYLC_ERR_ARG;
    *nH = H >= T ? (H - T) / S + 1 : 0;
    *nW = W >= T ? (W - T) / S + 1 : 0;
    return PYLC_OK;
}

extern "C" int pylc_tile_gather_u8(const uint8_t *src, int H, int W, int ch, size_t src_pitch, int T, int S,
                                   uint8_t *dst, uint64_t *stat, pylc_stream_t stream) {
    if (ch != 1 && ch != 3) return PYLC_ERR_ARG;
    GatherGeom g;
    int rc = make_geom(src, H, W, ch, src_pitch, T, S, &g);
    if (rc) return rc;
    const long long ctas = (long long)g.nbx * g.nby * g.slabs;
    if (ctas == 0) return PYLC_OK;  // source smaller than one tile: nothing to write
    if (!dst) return PYLC_ERR_ARG;
    if ((uintptr_t)dst % 16) return PYLC_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const bool al = aligned16(src, src_pitch);
    auto *sp = reinterpret_cast<unsigned long long *>(stat);
#define LAUNCH(CH, AL, ST) gather_img_kernel<CH, AL, ST><<<(unsigned)ctas, kThreads, 0, st>>>(g, dst, sp)
    if (ch == 1) {
        if (al) { if (stat) LAUNCH(1, true, true); else LAUNCH(1, true, false); }
        else    { if (stat) LAUNCH(1, false, true); else LAUNCH(1, false, false); }
    } else {
        if (al) { if (stat) LAUNCH(3, true, true); else LAUNCH(3, true, false); }
        else    { if (stat) LAUNCH(3, false, true); else LAUNCH(3, false, false); }
    }
#undef LAUNCH
    return finish_launch();
}

extern "C" int pylc_mask_gather_encode_hist(const uint8_t *src, int H, int W, size_t src_pitch, int T, int S,
                                            const uint8_t *palette, int C, uint8_t *dst, int64_t *px_dist,
                                            pylc_stream_t stream) {
    if (!palette) return PYLC_ERR_ARG;
    if (C < 1 || C > PYLC_MAX_CLASSES) return PYLC_ERR_CLASSES;
    GatherGeom g;
    int rc = make_geom(src, H, W, 3, src_pitch, T, S, &g);
    if (rc) return rc;
    PaletteHash ph;
    rc = build_palette_hash(palette, C, &ph);
    if (rc) return rc;
    const long long ctas = (long long)g.nbx * g.nby * g.slabs;
    if (ctas == 0) return PYLC_OK;
    if (!dst) return PYLC_ERR_ARG;
    if ((uintptr_t)dst % 16) return PYLC_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const bool al = aligned16(src, src_pitch);
    auto *pd = reinterpret_cast<long long *>(px_dist);
#define LAUNCH(AL, WD, HS) gather_mask_kernel<AL, WD, HS><<<(unsigned)ctas, kThreads, 0, st>>>(g, ph, C, dst, pd)
    if (!px_dist) { if (al) LAUNCH(true, false, false); else LAUNCH(false, false, false); }
    else if (C <= 12) { if (al) LAUNCH(true, false, true); else LAUNCH(false, false, true); }
    else { if (al) LAUNCH(true, true, true); else LAUNCH(false, true, true); }
#undef LAUNCH
    return finish_launch();
}

extern "C" int pylc_tile_gather_norm_f32(const uint8_t *src, int H, int W, int ch, size_t src_pitch, int T, int S,
                                         const float *mean, const float *std, float post_div, int out_ch,
                                         float *dst, pylc_stream_t stream) {
    if (!mean || !std || (ch != 1 && ch != 3)) return PYLC_ERR_ARG;
    if (!((ch == 1 && (out_ch == 1 || out_ch == 3)) || (ch == 3 && out_ch == 3))) return PYLC_ERR_ARG;
    GatherGeom g;
    int rc = make_geom(src, H, W, ch, src_pitch, T, S, &g);
    if (rc) return rc;
    if ((long long)g.nbx * g.nby == 0) return PYLC_OK;
    if (!dst) return PYLC_ERR_ARG;
    if ((uintptr_t)dst % 16) return PYLC_ERR_ALIGN;
    NormParams np;
    for (int k = 0; k < 3; ++k) {
        np.mean[k] = mean[ch == 1 ? 0 : k];
        np.std[k] = std[ch == 1 ? 0 : k];
    }
    np.post_div = post_div;
    np.out_ch = out_ch;
    const long long ctas = (long long)g.nbx * g.nby * g.slabs;
    if (ctas == 0) return PYLC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool al = aligned16(src, src_pitch);
    if (ch == 1) {
        if (al) gather_norm_kernel<1, true><<<(unsigned)ctas, kThreads, 0, st>>>(g, np, dst);
        else gather_norm_kernel<1, false><<<(unsigned)ctas, kThreads, 0, st>>>(g, np, dst);
    } else {
        if (al) gather_norm_kernel<3, true><<<(unsigned)ctas, kThreads, 0, st>>>(g, np, dst);
        else gather_norm_kernel<3, false><<<(unsigned)ctas, kThreads, 0, st>>>(g, np, dst);
    }
    return finish_launch();
}
