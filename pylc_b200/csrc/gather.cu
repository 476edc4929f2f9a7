// Tile gather kernels: image tiles (+ moments), mask tiles fused with palette encode and per-tile
// class histogram, and the normalising f32 gather.
//
// Work decomposition (u8 kernels): the used source area ((nH-1)S+T) x ((nW-1)S+T) is cut into
// S x S blocks.  With m = T/S every block belongs to at most m*m destination tiles, so whoever
// owns a piece of a block reads every source byte exactly once and writes it (transformed) into
// each of those tiles.  A block is cut into ITEMS of 256 units (one per thread; a unit is 16
// consecutive pixels: 16 B gray / 48 B interleaved RGB in, one 16-byte store per destination
// plane out, so a warp-level store is a contiguous 512-byte run of one tile row).
//
// The kernels are persistent: the grid is (#SMs x resident CTAs), CTA k takes the contiguous
// item range [k*n/G, (k+1)*n/G).  That balances to +-1 item however many items there are (no
// partial last wave), pays the palette-table / LUT set-up once per CTA, and -- because consecutive
// items belong to the same block -- lets histogram and moment partials stay in registers / shared
// memory until the block changes.  The next item's loads are issued before the current item is
// processed, so every thread keeps two units in flight.
#include "common.cuh"

#include <stdlib.h>

namespace pylc {

struct GatherGeom {
    const uint8_t *src;
    size_t pitch;
    int T, S, nH, nW, m;
    int nbx, nby;        // S-blocks across / down
    int gpr;             // 16-pixel groups per block row = S / 16
    int rows_item;       // block rows per item (rows_item * gpr <= 256)
    int slabs;           // items per block = S / rows_item
    int items;           // nbx * nby * slabs
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

struct Item {
    int blk, by, bx, slab;
};
__device__ __forceinline__ Item decode_item(const GatherGeom &g, int item) {
    Item it;
    it.slab = item % g.slabs;
    it.blk = item / g.slabs;
    it.bx = it.blk % g.nbx;
    it.by = it.blk / g.nbx;
    return it;
}
// consecutive items: advance without the divisions of decode_item
__device__ __forceinline__ Item next_item(const GatherGeom &g, Item it) {
    if (++it.slab == g.slabs) {
        it.slab = 0;
        ++it.blk;
        if (++it.bx == g.nbx) {
            it.bx = 0;
            ++it.by;
        }
    }
    return it;
}
__device__ __forceinline__ void cta_item_range(int items, int &first, int &last) {
    first = (int)((long long)items * blockIdx.x / gridDim.x);
    last = (int)((long long)items * (blockIdx.x + 1) / gridDim.x);
}

// De-interleave 4 RGB pixels (3 words) into one word per channel.
__device__ __forceinline__ void deinterleave4(uint32_t a, uint32_t b, uint32_t c, uint32_t &r, uint32_t &g,
                                              uint32_t &bl) {
    r = __byte_perm(__byte_perm(a, b, 0x0630), c, 0x5210);
    g = __byte_perm(__byte_perm(a, b, 0x0741), c, 0x6210);
    bl = __byte_perm(__byte_perm(a, b, 0x0052), c, 0x7410);
}

template <int CH>
struct SrcUnit {
    uint4 q[CH == 1 ? 1 : 3];
};
template <int CH, bool ALIGNED>
__device__ __forceinline__ void load_unit(const GatherGeom &g, const Item &it, int row, int grp, SrcUnit<CH> &u) {
    const int y = it.by * g.S + it.slab * g.rows_item + row;
    const int x = it.bx * g.S + grp * 16;
    const uint8_t *p = g.src + (size_t)y * g.pitch + (size_t)x * CH;
    u.q[0] = ld16<ALIGNED>(p);
    if (CH == 3) {
        u.q[CH == 3 ? 1 : 0] = ld16<ALIGNED>(p + 16);
        u.q[CH == 3 ? 2 : 0] = ld16<ALIGNED>(p + 32);
    }
}

// Per-thread cursor over the slabs of one S-block (T/S <= 2, i.e. at most four destination tiles):
// the source unit pointer and the destination unit pointers are computed when the block changes and
// advanced by constant strides from slab to slab, so the steady-state loop carries no multiplies.
struct BlockCursor {
    const uint8_t *src;
    uint8_t *dst[4];
    int n;
};
template <int CH>
__device__ __forceinline__ const uint8_t *unit_src(const GatherGeom &g, const Item &it, int row, int grp) {
    const int y = it.by * g.S + it.slab * g.rows_item + row;
    const int x = it.bx * g.S + grp * 16;
    return g.src + (size_t)y * g.pitch + (size_t)x * CH;
}
template <int CH>
__device__ __forceinline__ void cursor_set(const GatherGeom &g, const Item &it, int row, int grp, uint8_t *dst, size_t tile_bytes,
                                           BlockCursor &c) {
    c.src = unit_src<CH>(g, it, row, grp);
    const TileSpan ts = tile_span(g, it.by, it.bx);
    const int ly = it.slab * g.rows_item + row, lx = grp * 16;
    c.n = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) c.dst[i] = dst;
#pragma unroll
    for (int dr = 0; dr < 2; ++dr)
#pragma unroll
        for (int dc = 0; dc < 2; ++dc) {
            const int r = ts.r_lo + dr, cc = ts.c_lo + dc;
            if (r <= ts.r_hi && cc <= ts.c_hi) {
                uint8_t *d = dst + (size_t)(r * g.nW + cc) * tile_bytes + (size_t)((it.by - r) * g.S + ly) * g.T +
                             (size_t)((it.bx - cc) * g.S + lx);
                // compact into slots 0..n-1 without dynamic register indexing
                if (c.n == 0) c.dst[0] = d;
                else if (c.n == 1) c.dst[1] = d;
                else if (c.n == 2) c.dst[2] = d;
                else c.dst[3] = d;
                ++c.n;
            }
        }
}
template <int CH>
__device__ __forceinline__ void load_ptr(const uint8_t *p, bool aligned, SrcUnit<CH> &u) {
    if (aligned) {
        u.q[0] = ld_stream16(p);
        if (CH == 3) {
            u.q[CH == 3 ? 1 : 0] = ld_stream16(p + 16);
            u.q[CH == 3 ? 2 : 0] = ld_stream16(p + 32);
        }
    } else {
        u.q[0] = ld_bytes16(p);
        if (CH == 3) {
            u.q[CH == 3 ? 1 : 0] = ld_bytes16(p + 16);
            u.q[CH == 3 ? 2 : 0] = ld_bytes16(p + 32);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// image gather: u8 -> u8 tiles (+ sum, sum of squares per destination tile and channel)
// ------------------------------------------------------------------------------------------------
template <int CH, bool STATS>
__device__ __forceinline__ void flush_moments(const GatherGeom &g, int blk, uint32_t (&s1)[CH], uint32_t (&s2)[CH],
                                              unsigned long long *s_sum, unsigned long long *stat) {
    if (!STATS) return;
    // warp partials (16-bit halves keep the 32-lane sums inside u32), then one shared atomic per warp
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const uint32_t a = __reduce_add_sync(0xFFFFFFFFu, s1[k] & 0xFFFFu), ah = __reduce_add_sync(0xFFFFFFFFu, s1[k] >> 16);
        const uint32_t b = __reduce_add_sync(0xFFFFFFFFu, s2[k] & 0xFFFFu), bh = __reduce_add_sync(0xFFFFFFFFu, s2[k] >> 16);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&s_sum[k * 2], (unsigned long long)a + ((unsigned long long)ah << 16));
            atomicAdd(&s_sum[k * 2 + 1], (unsigned long long)b + ((unsigned long long)bh << 16));
        }
        s1[k] = s2[k] = 0;
    }
    __syncthreads();
    const int bx = blk % g.nbx, by = blk / g.nbx;
    const TileSpan ts = tile_span(g, by, bx);
    const int nt_c = ts.c_hi - ts.c_lo + 1;
    const int nt = (ts.r_hi - ts.r_lo + 1) * nt_c;
    for (int i = threadIdx.x; i < nt * CH * 2; i += kThreads) {
        const int t = i / (CH * 2), k = i - t * (CH * 2);
        const int r = ts.r_lo + t / nt_c, c = ts.c_lo + t % nt_c;
        atomicAdd(&stat[(size_t)(r * g.nW + c) * CH * 2 + k], s_sum[k]);
    }
    __syncthreads();
    if (threadIdx.x < CH * 2) s_sum[threadIdx.x] = 0;
    __syncthreads();
}

template <int CH, bool ALIGNED, bool STATS>
__global__ void __launch_bounds__(kThreads) gather_img_kernel(GatherGeom g, uint8_t *__restrict__ dst,
                                                              unsigned long long *__restrict__ stat) {
    __shared__ unsigned long long s_sum[CH * 2];
    if (STATS && threadIdx.x < CH * 2) s_sum[threadIdx.x] = 0;
    if (STATS) __syncthreads();
    int first, last;
    cta_item_range(g.items, first, last);
    if (first >= last) return;
    const int row = threadIdx.x / g.gpr, grp = threadIdx.x - row * g.gpr;
    const bool on = row < g.rows_item;
    const size_t TT = (size_t)g.T * g.T;

    uint32_t s1[CH], s2[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) s1[k] = s2[k] = 0;

    Item it = decode_item(g, first);
    SrcUnit<CH> cur, nxt;
    if (on) load_unit<CH, ALIGNED>(g, it, row, grp, cur);
    int units_since_flush = 0;
    for (int item = first; item < last; ++item) {
        Item it_n = it;
        if (item + 1 < last) {
            it_n = next_item(g, it);
            if (on) load_unit<CH, ALIGNED>(g, it_n, row, grp, nxt);
        }
        if (on) {
            uint4 o[CH];
            if (CH == 1) {
                o[0] = cur.q[0];
            } else {
                const uint4 q0 = cur.q[0], q1 = cur.q[CH == 3 ? 1 : 0], q2 = cur.q[CH == 3 ? 2 : 0];
                const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
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
            const TileSpan ts = tile_span(g, it.by, it.bx);
            const int ly = it.slab * g.rows_item + row, lx = grp * 16;
            for (int r = ts.r_lo; r <= ts.r_hi; ++r) {
                const int ty = (it.by - r) * g.S + ly;
                for (int c = ts.c_lo; c <= ts.c_hi; ++c) {
                    const int tx = (it.bx - c) * g.S + lx;
                    uint8_t *d = dst + ((size_t)(r * g.nW + c) * CH) * TT + (size_t)ty * g.T + tx;
#pragma unroll
                    for (int k = 0; k < CH; ++k) st_stream16(d + k * TT, o[k]);
                }
            }
        }
        // moments are per destination tile: flush when the block changes (u32 partials hold 4096 units)
        if (STATS && (it_n.blk != it.blk || item + 1 == last || ++units_since_flush >= 2048)) {
            flush_moments<CH, STATS>(g, it.blk, s1, s2, s_sum, stat);
            units_since_flush = 0;
        }
        it = it_n;
        cur = nxt;
    }
}

// Staged form for 16-byte aligned sources and T/S <= 2: one small CTA per strip of rows of an S-block.
// The strip's source bytes reach shared memory through coalesced 16-byte cp.async (every load of the
// CTA in flight at once), a thread takes its 16-pixel units from there with conflict-free 16-byte reads,
// the destination tile bases are computed once per CTA, and the moments -- the strip lies in one block --
// are reduced and added to the block's tiles once per CTA.
// Grid: x = strip of the block, y = block column, z = image * nby + block row (a stack of equally sized
// sources is one launch).  A CTA moves only ~8 units per thread, so its fixed cost is kept small: no
// per-thread integer division (chunk -> row by a host-computed reciprocal), the <= 4 destination tiles as a
// branch-free 2 x 2 table.  (The first version spent 60 % of its ~870 instructions per thread on set-up
// and flush and ran at 75 % issue utilisation: ncu, profiles/ncu_r2c_summary.md.)
#ifndef PYLC_IMG_STRIP_KB
#define PYLC_IMG_STRIP_KB 32
#endif
constexpr int kImgStripKB = PYLC_IMG_STRIP_KB;
struct StagedArgs {
    int strip_rows;
    uint32_t magic_cpr, magic_gpr;   // ceil(2^32 / d) for d = 16-byte chunks per strip row / 16-pixel units per row
    size_t img_stride;
};
__device__ __forceinline__ int div_small(int q, int d, uint32_t magic) {   // q / d for q < 2^16
    return d == 1 ? q : (int)__umulhi((uint32_t)q, magic);
}

template <int CH, bool STATS>
__global__ void __launch_bounds__(kThreads) gather_img_staged_kernel(const GatherGeom g, uint8_t *__restrict__ dst,
                                                                    unsigned long long *__restrict__ stat, const StagedArgs sa) {
    __shared__ unsigned long long s_sum[CH * 2];
    extern __shared__ __align__(16) uint8_t s_src[];          // strip_rows rows of S * CH bytes
    const int tid = threadIdx.x;
    if (STATS && tid < CH * 2) s_sum[tid] = 0;
    const int bx = blockIdx.y;
    const int img = blockIdx.z / g.nby, by = blockIdx.z - img * g.nby;
    const int y0 = blockIdx.x * sa.strip_rows;
    const int nrow = min(sa.strip_rows, g.S - y0);
    const int row_bytes = g.S * CH, cpr = row_bytes >> 4;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_src);
    {
        const uint8_t *src0 = g.src + (size_t)img * sa.img_stride + (size_t)(by * g.S + y0) * g.pitch + (size_t)bx * row_bytes;
        const int nchunk = nrow * cpr;
        for (int q = tid; q < nchunk; q += kThreads) {       // the strip is dense in shared memory: chunk q at q * 16
            const int r = div_small(q, cpr, sa.magic_cpr), ck = q - r * cpr;
            cp_async16(sbase + (uint32_t)q * 16u, src0 + (size_t)r * g.pitch + ck * 16);
        }
        cp_async_commit();
    }
    // destination tiles of block (by, bx): tile (by - dr, bx - dc) holds it at rows dr * S, columns dc * S
    const size_t TT = (size_t)g.T * g.T;
    const long long tile0 = (long long)img * g.nH * g.nW;
    uint8_t *tb[4];
    bool ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int dr = i >> 1, dc = i & 1;
        const int r = by - dr, c = bx - dc;
        ok[i] = dr < g.m && dc < g.m && r >= 0 && r < g.nH && c >= 0 && c < g.nW;
        const long long tile = ok[i] ? tile0 + (long long)r * g.nW + c : tile0;
        tb[i] = dst + (size_t)tile * CH * TT + (size_t)(dr * g.S + y0) * g.T + dc * g.S;
    }
    cp_async_wait<0>();
    __syncthreads();

    uint32_t s1[CH], s2[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) s1[k] = s2[k] = 0;
    const int nunit = nrow * g.gpr;
    for (int q = tid; q < nunit; q += kThreads) {             // unit q of the strip: shared-memory offset q * 16 * CH
        const int row = div_small(q, g.gpr, sa.magic_gpr), grp = q - row * g.gpr;
        const uint32_t sa_u = sbase + (uint32_t)q * (16u * CH);
        uint4 o[CH];
        if (CH == 1) {
            o[0] = lds128(sa_u);
        } else {
            const uint4 q0 = lds128(sa_u), q1 = lds128(sa_u + 16), q2 = lds128(sa_u + 32);
            const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
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
        const int off = row * g.T + grp * 16;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (ok[i]) {
#pragma unroll
                for (int k = 0; k < CH; ++k) st_stream16(tb[i] + k * TT + off, o[k]);
            }
    }
    if (STATS) {     // u32 partials hold 4096 units per thread; a strip is at most a few per thread
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const uint32_t a = __reduce_add_sync(0xFFFFFFFFu, s1[k] & 0xFFFFu), ah = __reduce_add_sync(0xFFFFFFFFu, s1[k] >> 16);
            const uint32_t b = __reduce_add_sync(0xFFFFFFFFu, s2[k] & 0xFFFFu), bh = __reduce_add_sync(0xFFFFFFFFu, s2[k] >> 16);
            if ((tid & 31) == 0) {
                atomicAdd(&s_sum[k * 2], (unsigned long long)a + ((unsigned long long)ah << 16));
                atomicAdd(&s_sum[k * 2 + 1], (unsigned long long)b + ((unsigned long long)bh << 16));
            }
        }
        __syncthreads();
        if (tid < 4 * CH * 2) {                                // one thread per (destination tile, channel, moment)
            const int i = tid / (CH * 2), k = tid - i * (CH * 2);
            const int r = by - (i >> 1), c = bx - (i & 1);
            if ((i >> 1) < g.m && (i & 1) < g.m && r >= 0 && r < g.nH && c >= 0 && c < g.nW)
                atomicAdd(&stat[(size_t)(tile0 + (long long)r * g.nW + c) * CH * 2 + k], s_sum[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// mask gather + palette encode + per-tile histogram
// ------------------------------------------------------------------------------------------------
template <class CC>
__device__ __forceinline__ void flush_block_hist(const GatherGeom &g, int blk, CC &cc, int C,
                                                 unsigned *s_hist, long long *px_dist) {
    cc.widen();
    flush_counter(cc, C, s_hist);
    cc.reset();
    __syncthreads();
    const int bx = blk % g.nbx, by = blk / g.nbx;
    const TileSpan ts = tile_span(g, by, bx);
    const int nt_c = ts.c_hi - ts.c_lo + 1;
    const int nt = (ts.r_hi - ts.r_lo + 1) * nt_c;
    for (int i = threadIdx.x; i < nt * C; i += kThreads) {
        const int t = i / C, k = i - t * C;
        const int r = ts.r_lo + t / nt_c, c = ts.c_lo + t % nt_c;
        if (s_hist[k]) atomicAdd((unsigned long long *)&px_dist[(size_t)(r * g.nW + c) * C + k], (unsigned long long)s_hist[k]);
    }
    __syncthreads();
    if (threadIdx.x < PYLC_MAX_CLASSES) s_hist[threadIdx.x] = 0;
    __syncthreads();
}

template <bool ALIGNED, int NG, bool HIST>
__global__ void __launch_bounds__(kThreads)
    gather_mask_kernel(GatherGeom g, const __grid_constant__ PaletteHash ph, int C, uint8_t *__restrict__ dst,
                       long long *__restrict__ px_dist) {
    __shared__ uint32_t s_tab[256];
    __shared__ unsigned s_hist[PYLC_MAX_CLASSES];
    s_tab[threadIdx.x] = ph.tab[threadIdx.x];
    if (threadIdx.x < PYLC_MAX_CLASSES) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t mul = ph.mul;
    int first, last;
    cta_item_range(g.items, first, last);
    if (first >= last) return;
    const int row = threadIdx.x / g.gpr, grp = threadIdx.x - row * g.gpr;
    const bool on = row < g.rows_item;
    const size_t TT = (size_t)g.T * g.T;

    using CC = typename CounterSel<NG>::type;
    CC cc;
    cc.reset();
    int since_flush = 0, since_widen = 0;
    Item it = decode_item(g, first);
    SrcUnit<3> cur, nxt;
    if (on) load_unit<3, ALIGNED>(g, it, row, grp, cur);
    for (int item = first; item < last; ++item) {
        Item it_n = it;
        if (item + 1 < last) {
            it_n = next_item(g, it);
            if (on) load_unit<3, ALIGNED>(g, it_n, row, grp, nxt);
        }
        if (on) {
            const uint4 q0 = cur.q[0], q1 = cur.q[1], q2 = cur.q[2];
            const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
            uint32_t ow[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t a = w[3 * k], b = w[3 * k + 1], c = w[3 * k + 2];
                const uint32_t c0 = encode_key(a, s_tab, mul);
                const uint32_t c1 = encode_key(__funnelshift_r(a, b, 24), s_tab, mul);
                const uint32_t c2 = encode_key(__funnelshift_r(b, c, 16), s_tab, mul);
                const uint32_t c3 = encode_key(c >> 8, s_tab, mul);
                ow[k] = c0 + (c1 << 8) + (c2 << 16) + (c3 << 24);
            }
            if (HIST) {
                cc.add16(ow[0], ow[1], ow[2], ow[3]);
                if (++since_widen == CC::kWidenUnits) {
                    cc.widen();
                    since_widen = 0;
                }
            }
            const uint4 o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            const TileSpan ts = tile_span(g, it.by, it.bx);
            const int ly = it.slab * g.rows_item + row, lx = grp * 16;
            for (int r = ts.r_lo; r <= ts.r_hi; ++r) {
                const int ty = (it.by - r) * g.S + ly;
                for (int c = ts.c_lo; c <= ts.c_hi; ++c) {
                    const int tx = (it.bx - c) * g.S + lx;
                    st_stream16(dst + (size_t)(r * g.nW + c) * TT + (size_t)ty * g.T + tx, o);
                }
            }
        }
        // histograms are per destination tile: flush when the block changes, or before the packed
        // per-thread fields could overflow (CC::kFlushUnits units)
        if (HIST && (it_n.blk != it.blk || item + 1 == last || ++since_flush >= CC::kFlushUnits)) {
            flush_block_hist(g, it.blk, cc, C, s_hist, px_dist);
            since_flush = since_widen = 0;
        }
        it = it_n;
        cur = nxt;
    }
}

// T/S <= 2 version (the reference's two strides: S = T on extraction, S = T/2 on the test path):
// BlockCursor addressing, warp-private histogram flushes (no CTA barrier anywhere in the loop, so
// the eight warps of a CTA drift apart and hide each other's load latency).
template <class CC>
__device__ __forceinline__ void flush_warp_hist(const GatherGeom &g, int blk, CC &cc, int C, long long *px_dist) {
    const int lane = threadIdx.x & 31;
    cc.widen();
    unsigned mine = 0;
#pragma unroll
    for (int c = 0; c < CC::kMaxClasses; ++c) {
        if (c < C) {
            const unsigned v = __reduce_add_sync(0xFFFFFFFFu, cc.count(c));
            if (lane == c) mine = v;
        }
    }
    cc.reset();
    if (lane < C && mine) {
        const int bx = blk % g.nbx, by = blk / g.nbx;
        const TileSpan ts = tile_span(g, by, bx);
        for (int r = ts.r_lo; r <= ts.r_hi; ++r)
            for (int c = ts.c_lo; c <= ts.c_hi; ++c)
                atomicAdd((unsigned long long *)&px_dist[(size_t)(r * g.nW + c) * C + lane], (unsigned long long)mine);
    }
}

template <bool ALIGNED, int NG, bool HIST>
__global__ void __launch_bounds__(kThreads, 4)
    gather_mask_cursor_kernel(GatherGeom g, const __grid_constant__ PaletteHash ph, int C, uint8_t *__restrict__ dst,
                              long long *__restrict__ px_dist) {
    __shared__ uint32_t s_tab[256];
    s_tab[threadIdx.x] = ph.tab[threadIdx.x];
    __syncthreads();
    const uint32_t mul = ph.mul;
    const uint32_t miss_e = 1u << 24;   // unmatched colours are class 1 (utils/tools.py:437)
    int first, last;
    cta_item_range(g.items, first, last);
    if (first >= last) return;
    const int row = threadIdx.x / g.gpr, grp = threadIdx.x - row * g.gpr;
    const bool on = row < g.rows_item;
    const size_t TT = (size_t)g.T * g.T;
    const size_t sstep = (size_t)g.rows_item * g.pitch;
    const int dstep = g.rows_item * g.T;

    using CC = typename CounterSel<NG>::type;
    CC cc;
    cc.reset();
    int since_flush = 0, since_widen = 0;
    Item it = decode_item(g, first);
    BlockCursor cur;
    cursor_set<3>(g, it, row, grp, dst, TT, cur);
    SrcUnit<3> q, nq;
    if (on) load_ptr<3>(cur.src, ALIGNED, q);
    for (int item = first; item < last; ++item) {
        Item it_n = it;
        const uint8_t *nsrc = cur.src;
        bool chg = false;
        if (item + 1 < last) {
            it_n = next_item(g, it);
            chg = it_n.blk != it.blk;
            nsrc = chg ? unit_src<3>(g, it_n, row, grp) : cur.src + sstep;
            if (on) load_ptr<3>(nsrc, ALIGNED, nq);
        }
        if (on) {
            const uint32_t w[12] = {q.q[0].x, q.q[0].y, q.q[0].z, q.q[0].w, q.q[1].x, q.q[1].y,
                                    q.q[1].z, q.q[1].w, q.q[2].x, q.q[2].y, q.q[2].z, q.q[2].w};
            uint32_t ow[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t a = w[3 * k], b = w[3 * k + 1], c = w[3 * k + 2];
                const uint32_t e0 = lookup_entry(a, s_tab, mul, miss_e);
                const uint32_t e1 = lookup_entry(__funnelshift_r(a, b, 24), s_tab, mul, miss_e);
                const uint32_t e2 = lookup_entry(__funnelshift_r(b, c, 16), s_tab, mul, miss_e);
                const uint32_t e3 = lookup_entry(c >> 8, s_tab, mul, miss_e);
                ow[k] = pack_top_bytes(e0, e1, e2, e3);
            }
            if (HIST) {
                cc.add16(ow[0], ow[1], ow[2], ow[3]);
                if (++since_widen == CC::kWidenUnits) {
                    cc.widen();
                    since_widen = 0;
                }
            }
            const uint4 o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            st_stream16(cur.dst[0], o);
            if (cur.n > 1) st_stream16(cur.dst[1], o);
            if (cur.n > 2) st_stream16(cur.dst[2], o);
            if (cur.n > 3) st_stream16(cur.dst[3], o);
        }
        // histograms are per destination tile: flush when the block changes, or before the packed
        // per-thread fields could overflow (CC::kFlushUnits units)
        if (HIST && (chg || item + 1 == last || ++since_flush >= CC::kFlushUnits)) {
            flush_warp_hist(g, it.blk, cc, C, px_dist);
            since_flush = since_widen = 0;
        }
        if (chg) {
            cursor_set<3>(g, it_n, row, grp, dst, TT, cur);
        } else {
            cur.src = nsrc;
#pragma unroll
            for (int i = 0; i < 4; ++i) cur.dst[i] += dstep;
        }
        it = it_n;
        q = nq;
    }
}

// ------------------------------------------------------------------------------------------------
// normalising gather: u8 source -> network-ready f32 tiles (models/model.py:416-445, 376-377)
// ------------------------------------------------------------------------------------------------
// Output is 4 bytes per pixel and plane, so here a unit is FOUR pixels: a lane reads 4 B (gray) or
// 12 B (RGB) and writes one 16-byte vector per destination plane -- a warp-level store is again a
// contiguous 512-byte run, which is what the (write-dominated: 0.75 B in, 12 B out per tile pixel)
// kernel needs.  Items are whole block rows: S/4 units, walked with a CTA-stride loop.
struct NormParams {
    float mean[3], std[3];
    float post_div;  // 255 (models/model.py:435,445) or 1 (grayscale `default` branch, 431-432)
    int out_ch;
};

template <int CH, bool ALIGNED>
__global__ void __launch_bounds__(kThreads) gather_norm_kernel(GatherGeom g, NormParams np, float *__restrict__ dst) {
    // 256-entry tables of IEEE sub/div/div in the reference's order, so the f32 tiles are bit-equal
    // to ((x - mean) / std) / 255 evaluated by torch on the CPU.
    __shared__ float s_lut[CH][256];
#pragma unroll
    for (int k = 0; k < CH; ++k)
        s_lut[k][threadIdx.x] = __fdiv_rn(__fdiv_rn(__fsub_rn((float)threadIdx.x, np.mean[k]), np.std[k]), np.post_div);
    __syncthreads();
    const size_t TT = (size_t)g.T * g.T;
    // Items are `rows_item` consecutive rows of one S-block (sized by the host so that one item is about
    // one unit per thread); CTA k takes a contiguous item range.  The item decode is CTA-uniform and the
    // thread's (row, unit) position inside an item advances by constants, so the loop has no divisions.
    const int upr = g.S / 4;                                   // units per block row
    const int per_item = g.rows_item * upr;
    const int d_row = kThreads / upr, d_col = kThreads - d_row * upr;
    int first, last;
    cta_item_range(g.items, first, last);
    for (int item = first; item < last; ++item) {
        const int blk = item / g.slabs, slab = item - blk * g.slabs;
        const int bx = blk % g.nbx, by = blk / g.nbx;
        const TileSpan ts = tile_span(g, by, bx);
        int row = threadIdx.x / upr, col = threadIdx.x - row * upr;
        for (int idx = threadIdx.x; idx < per_item; idx += kThreads, row += d_row, col += d_col) {
            if (col >= upr) {
                col -= upr;
                ++row;
            }
            const int ly = slab * g.rows_item + row, lx = col * 4;
            const uint8_t *p = g.src + (size_t)(by * g.S + ly) * g.pitch + (size_t)(bx * g.S + lx) * CH;
            uint32_t px[3];
            if (CH == 1) {
                px[0] = ALIGNED ? __ldg(reinterpret_cast<const uint32_t *>(p))
                                : (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) |
                                      ((uint32_t)__ldg(p + 3) << 24);
            } else {
                uint32_t w[3];
                if (ALIGNED) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) w[k] = __ldg(reinterpret_cast<const uint32_t *>(p) + k);
                } else {
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        w[k] = (uint32_t)__ldg(p + 4 * k) | ((uint32_t)__ldg(p + 4 * k + 1) << 8) |
                               ((uint32_t)__ldg(p + 4 * k + 2) << 16) | ((uint32_t)__ldg(p + 4 * k + 3) << 24);
                }
                deinterleave4(w[0], w[1], w[2], px[0], px[CH == 3 ? 1 : 0], px[CH == 3 ? 2 : 0]);
            }
            float4 f[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                f[k].x = s_lut[k][px[k] & 0xFF];
                f[k].y = s_lut[k][(px[k] >> 8) & 0xFF];
                f[k].z = s_lut[k][(px[k] >> 16) & 0xFF];
                f[k].w = s_lut[k][px[k] >> 24];
            }
            for (int r = ts.r_lo; r <= ts.r_hi; ++r) {
                const int ty = (by - r) * g.S + ly;
                for (int c = ts.c_lo; c <= ts.c_hi; ++c) {
                    const int tx = (bx - c) * g.S + lx;
                    float *d = dst + ((size_t)(r * g.nW + c) * np.out_ch) * TT + (size_t)ty * g.T + tx;
                    if (CH == 1) {
                        for (int k = 0; k < np.out_ch; ++k) st_stream_f4(d + k * TT, f[0]);
                    } else {
#pragma unroll
                        for (int k = 0; k < CH; ++k) st_stream_f4(d + k * TT, f[k]);
                    }
                }
            }
        }
    }
}

// Staged form for 16-byte aligned sources: one small CTA per kNormRows rows of an S-block.  The strip's
// source rows reach shared memory through coalesced 16-byte cp.async while the tables are built; a
// thread then takes four pixels (12 bytes at a 12-byte lane stride: conflict-free) from there, and the
// destination tile bases are computed once per CTA -- the per-unit work is table look-ups and stores.
constexpr int kNormRows = 16;

template <int CH>
__global__ void __launch_bounds__(kThreads) gather_norm_staged_kernel(GatherGeom g, NormParams np, float *__restrict__ dst, int gpb) {
    __shared__ float s_lut[CH][256];
    extern __shared__ __align__(16) uint8_t s_src[];          // kNormRows rows of S * CH bytes
    const int blk = blockIdx.x / gpb, y0 = (blockIdx.x - blk * gpb) * kNormRows;
    const int nrow = min(kNormRows, g.S - y0);
    const int bx = blk % g.nbx, by = blk / g.nbx;
    const int row_bytes = g.S * CH, cpr = row_bytes / 16;
    {
        const uint8_t *src0 = g.src + (size_t)(by * g.S + y0) * g.pitch + (size_t)bx * row_bytes;
        const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(s_src);
        int r = threadIdx.x / cpr, ck = threadIdx.x - r * cpr;
        const int dr = kThreads / cpr, dck = kThreads - dr * cpr;
        for (; r < nrow; r += dr, ck += dck) {
            if (ck >= cpr) {
                ck -= cpr;
                if (++r >= nrow) break;
            }
            cp_async16(s0 + (uint32_t)(r * row_bytes + ck * 16), src0 + (size_t)r * g.pitch + ck * 16);
        }
        cp_async_commit();
    }
#pragma unroll
    for (int k = 0; k < CH; ++k)
        s_lut[k][threadIdx.x] = __fdiv_rn(__fdiv_rn(__fsub_rn((float)threadIdx.x, np.mean[k]), np.std[k]), np.post_div);

    const size_t TT = (size_t)g.T * g.T;
    float *tb[4];
    int nt = 0;
    {
        const TileSpan ts = tile_span(g, by, bx);
#pragma unroll
        for (int i = 0; i < 4; ++i) tb[i] = dst;
        for (int r = ts.r_lo; r <= ts.r_hi; ++r)
            for (int c = ts.c_lo; c <= ts.c_hi; ++c) {
                float *b = dst + ((size_t)(r * g.nW + c) * np.out_ch) * TT + (size_t)((by - r) * g.S + y0) * g.T + (bx - c) * g.S;
                if (nt == 0) tb[0] = b;
                else if (nt == 1) tb[1] = b;
                else if (nt == 2) tb[2] = b;
                else if (nt == 3) tb[3] = b;
                ++nt;
            }
    }
    cp_async_wait<0>();
    __syncthreads();

    const int upr = g.S / 4;                                   // 4-pixel units per row
    int row = threadIdx.x / upr, col = threadIdx.x - row * upr;
    const int d_row = kThreads / upr, d_col = kThreads - d_row * upr;
    for (; row < nrow; row += d_row, col += d_col) {
        if (col >= upr) {
            col -= upr;
            if (++row >= nrow) break;
        }
        const uint32_t *p = reinterpret_cast<const uint32_t *>(s_src + row * row_bytes + col * 4 * CH);
        uint32_t px[3];
        if (CH == 1) px[0] = p[0];
        else deinterleave4(p[0], p[CH == 3 ? 1 : 0], p[CH == 3 ? 2 : 0], px[0], px[CH == 3 ? 1 : 0], px[CH == 3 ? 2 : 0]);
        float4 f[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            f[k].x = s_lut[k][px[k] & 0xFF];
            f[k].y = s_lut[k][(px[k] >> 8) & 0xFF];
            f[k].z = s_lut[k][(px[k] >> 16) & 0xFF];
            f[k].w = s_lut[k][px[k] >> 24];
        }
        const int off = row * g.T + col * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < nt) {
                float *d = tb[i] + off;
                if (CH == 1) {
                    for (int k = 0; k < np.out_ch; ++k) st_stream_f4(d + k * TT, f[0]);
                } else {
#pragma unroll
                    for (int k = 0; k < CH; ++k) st_stream_f4(d + k * TT, f[k]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// normalising gather, space-to-depth layout for the network stem
// ------------------------------------------------------------------------------------------------
// The ResNet stem is a 7x7 stride-2 convolution over 3 channels (models/backbone/resnet.py), the one
// layer the library runs far from its roofline (K = 3 input channels).  It equals a 4x4 stride-1
// convolution over the 2x2 space-to-depth image (12 channels, padded to 16) with the 7x7 kernel zero-
// extended to 8x8 -- models/fused.py rearranges the weights -- so this kernel writes the tiles directly in
// that layout, channels-last, border included:
//     dst[n, Y, X, (py*2+px)*3 + c] = norm(tile_n[c, 2(Y-2)+py, 2(X-2)+px]),  Y, X in [0, T/2+3)
// (two zero rows/columns before the image, one after; channels 12..15 zero), f32, same IEEE
// normalisation tables as gather_norm_kernel.
//
// Source-driven like the u8 gathers: a 2x2 source quad becomes ONE 64-byte record, built once and
// stored into every tile that contains it (up to four at S = T/2), so the table look-ups are paid per
// source pixel, not per tile pixel.  A lane owns one 16-byte quarter of a record (`part` = lane & 3,
// constant per thread because the CTA strides by a multiple of four): quarters 0..2 take four source
// bytes from two 16-bit loads, quarter 3 is the zero padding.  Consecutive lanes therefore write
// consecutive 16-byte vectors -- a warp-level store is a contiguous 512-byte run of full sectors in
// each destination tile.  A second, small loop zero-fills the border records (6*Hs - 9 per tile).
template <int CH, bool AL2>
__global__ void __launch_bounds__(kThreads) gather_norm_s2d_kernel(GatherGeom g, NormParams np, float *__restrict__ dst) {
    __shared__ float s_lut[3][256];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        s_lut[k][threadIdx.x] = __fdiv_rn(__fdiv_rn(__fsub_rn((float)threadIdx.x, np.mean[k]), np.std[k]), np.post_div);
    __syncthreads();
    const int Hs = g.T / 2 + 3;
    const int hq = g.S / 2;            // quads per block side
    const int upr = hq * 4;            // 16-byte vectors per quad row of a block
    const int part = threadIdx.x & 3;

    // thread-constant byte sources inside the quad's two 2*CH-byte rows, and the table row per byte
    size_t offA, offB;
    uint32_t sel;
    if (CH == 3) {   // record bytes 4*part .. 4*part+3 of [row0: 6 bytes][row1: 6 bytes]
        offA = part == 0 ? 0 : (part == 1 ? 4 : g.pitch + 2);
        offB = part == 0 ? 2 : (part == 1 ? g.pitch : g.pitch + 4);
        sel = 0x3210u;
    } else {         // gray replicated over c: channel k of the record is quad pixel k / 3
        offA = part == 2 ? g.pitch : 0;
        offB = g.pitch;
        sel = part == 0 ? 0x1000u : (part == 1 ? 0x2211u : 0x1110u);
    }
    const float *lut0 = s_lut[(part * 4) % 3], *lut1 = s_lut[(part * 4 + 1) % 3], *lut2 = s_lut[(part * 4 + 2) % 3],
                *lut3 = s_lut[(part * 4 + 3) % 3];
    auto load16 = [](const uint8_t *p) -> uint32_t {
        if (AL2) return (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(p));
        return (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8);
    };

    const int items = g.nbx * g.nby * hq;     // one quad row of one S-block
    int first, last;
    cta_item_range(items, first, last);
    for (int item = first; item < last; ++item) {
        const int blk = item / hq, qy = item - blk * hq;
        const int bx = blk % g.nbx, by = blk / g.nbx;
        const TileSpan ts = tile_span(g, by, bx);
        const uint8_t *rowp = g.src + (size_t)(by * g.S + 2 * qy) * g.pitch + (size_t)bx * g.S * CH;
        for (int f = threadIdx.x; f < upr; f += kThreads) {
            const int qx = f >> 2;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (part != 3) {
                const uint8_t *p = rowp + (size_t)qx * 2 * CH;
                const uint32_t w = __byte_perm(load16(p + offA) | (load16(p + offB) << 16), 0u, sel);
                v.x = lut0[w & 0xFF];
                v.y = lut1[(w >> 8) & 0xFF];
                v.z = lut2[(w >> 16) & 0xFF];
                v.w = lut3[w >> 24];
            }
            for (int r = ts.r_lo; r <= ts.r_hi; ++r) {
                const int Y = (by - r) * hq + qy + 2;
                for (int c = ts.c_lo; c <= ts.c_hi; ++c) {
                    const int X = (bx - c) * hq + qx + 2;
                    st_stream_f4(dst + ((((size_t)(r * g.nW + c) * Hs + Y) * Hs + X) * 4 + part) * 4, v);
                }
            }
        }
    }

    // border records: rows 0, 1, Hs-1 whole; columns 0, 1, Hs-1 of the T/2 rows in between
    const int nb = 6 * Hs - 9;
    const long long total = (long long)g.nH * g.nW * nb * 4;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const unsigned rec = (unsigned)(i >> 2);
        const unsigned n = rec / (unsigned)nb;
        int k = (int)(rec - n * (unsigned)nb), Y, X;
        if (k < 2 * Hs) {
            Y = k >= Hs;
            X = k - Y * Hs;
        } else if (k < 3 * Hs) {
            Y = Hs - 1;
            X = k - 2 * Hs;
        } else {
            k -= 3 * Hs;
            Y = 2 + k / 3;
            const int j = k - (Y - 2) * 3;
            X = j < 2 ? j : Hs - 1;
        }
        st_stream_f4(dst + ((((size_t)n * Hs + Y) * Hs + X) * 4 + (i & 3)) * 4, zero);
    }
}

// Staged form for 16-byte aligned sources (the pipeline's case): one small CTA per kS2dRows quad rows
// of an S-block.  The 2 * kS2dRows source rows of the strip go to shared memory with coalesced 16-byte
// cp.async while the normalisation tables are built, so the per-record work below starts from shared
// memory (no global-load latency inside a thread's chain, no 64-bit source addressing), and the
// destination tile bases are computed once per CTA.  Same records, same stores as the kernel above.
constexpr int kS2dRows = 8;

template <int CH>
__global__ void __launch_bounds__(kThreads) gather_norm_s2d_staged_kernel(GatherGeom g, NormParams np, float *__restrict__ dst, int gpb) {
    __shared__ float s_lut[3][256];
    extern __shared__ __align__(16) uint8_t s_src[];          // 2 * kS2dRows rows of S * CH bytes
    const int Hs = g.T / 2 + 3;
    const int hq = g.S / 2;            // quads per block side
    const int upr = hq * 4;            // 16-byte vectors per quad row of a block
    const int blk = blockIdx.x / gpb, qy0 = (blockIdx.x - blk * gpb) * kS2dRows;
    const int nq = min(kS2dRows, hq - qy0);
    const int bx = blk % g.nbx, by = blk / g.nbx;
    const int row_bytes = g.S * CH, cpr = row_bytes / 16;
    {
        const uint8_t *src0 = g.src + (size_t)(by * g.S + 2 * qy0) * g.pitch + (size_t)bx * row_bytes;
        const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(s_src);
        int r = threadIdx.x / cpr, ck = threadIdx.x - r * cpr;
        const int dr = kThreads / cpr, dck = kThreads - dr * cpr;
        for (; r < 2 * nq; r += dr, ck += dck) {
            if (ck >= cpr) {
                ck -= cpr;
                if (++r >= 2 * nq) break;
            }
            cp_async16(s0 + (uint32_t)(r * row_bytes + ck * 16), src0 + (size_t)r * g.pitch + ck * 16);
        }
        cp_async_commit();
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        s_lut[k][threadIdx.x] = __fdiv_rn(__fdiv_rn(__fsub_rn((float)threadIdx.x, np.mean[k]), np.std[k]), np.post_div);

    // destination: record (Y, X) = (qy + 2, qx + 2) of every tile that contains the block
    float4 *tb[4];
    int nt = 0;
    {
        const TileSpan ts = tile_span(g, by, bx);
        float4 *d4 = reinterpret_cast<float4 *>(dst);
#pragma unroll
        for (int i = 0; i < 4; ++i) tb[i] = d4;
        for (int r = ts.r_lo; r <= ts.r_hi; ++r)
            for (int c = ts.c_lo; c <= ts.c_hi; ++c) {
                float4 *b = d4 + (((size_t)(r * g.nW + c) * Hs + (by - r) * hq + qy0 + 2) * Hs + (bx - c) * hq + 2) * 4;
                if (nt == 0) tb[0] = b;
                else if (nt == 1) tb[1] = b;
                else if (nt == 2) tb[2] = b;
                else if (nt == 3) tb[3] = b;
                ++nt;
            }
    }
    const int part = threadIdx.x & 3;
    // thread-constant byte sources inside the quad's two 2*CH-byte rows (see the kernel above)
    int offA, offB;
    uint32_t sel;
    if (CH == 3) {
        offA = part == 0 ? 0 : (part == 1 ? 4 : row_bytes + 2);
        offB = part == 0 ? 2 : (part == 1 ? row_bytes : row_bytes + 4);
        sel = 0x3210u;
    } else {
        offA = part == 2 ? row_bytes : 0;
        offB = row_bytes;
        sel = part == 0 ? 0x1000u : (part == 1 ? 0x2211u : 0x1110u);
    }
    const float *lut0 = s_lut[(part * 4) % 3], *lut1 = s_lut[(part * 4 + 1) % 3], *lut2 = s_lut[(part * 4 + 2) % 3],
                *lut3 = s_lut[(part * 4 + 3) % 3];
    cp_async_wait<0>();
    __syncthreads();

    if (nt <= 4) {       // T/S <= 2; larger overlaps take the general kernel (the host checks)
        int ql = threadIdx.x / upr, f = threadIdx.x - ql * upr;
        const int dq = kThreads / upr, df = kThreads - dq * upr;       // df is a multiple of 4: `part` stays put
        for (; ql < nq; ql += dq, f += df) {
            if (f >= upr) {
                f -= upr;
                if (++ql >= nq) break;
            }
            const int qx = f >> 2;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (part != 3) {
                const uint8_t *p = s_src + (2 * ql) * row_bytes + qx * 2 * CH;
                const uint32_t a = *reinterpret_cast<const uint16_t *>(p + offA), b = *reinterpret_cast<const uint16_t *>(p + offB);
                const uint32_t w = __byte_perm(a | (b << 16), 0u, sel);
                v.x = lut0[w & 0xFF];
                v.y = lut1[(w >> 8) & 0xFF];
                v.z = lut2[(w >> 16) & 0xFF];
                v.w = lut3[w >> 24];
            }
            const int off = (ql * Hs + qx) * 4 + part;
            st_stream_f4(reinterpret_cast<float *>(tb[0] + off), v);
            if (nt > 1) st_stream_f4(reinterpret_cast<float *>(tb[1] + off), v);
            if (nt > 2) st_stream_f4(reinterpret_cast<float *>(tb[2] + off), v);
            if (nt > 3) st_stream_f4(reinterpret_cast<float *>(tb[3] + off), v);
        }
    }

    // border records: rows 0, 1, Hs-1 whole; columns 0, 1, Hs-1 of the T/2 rows in between
    const int nb = 6 * Hs - 9;
    const long long total = (long long)g.nH * g.nW * nb * 4;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const unsigned rec = (unsigned)(i >> 2);
        const unsigned n = rec / (unsigned)nb;
        int k = (int)(rec - n * (unsigned)nb), Y, X;
        if (k < 2 * Hs) {
            Y = k >= Hs;
            X = k - Y * Hs;
        } else if (k < 3 * Hs) {
            Y = Hs - 1;
            X = k - 2 * Hs;
        } else {
            k -= 3 * Hs;
            Y = 2 + k / 3;
            const int j = k - (Y - 2) * 3;
            X = j < 2 ? j : Hs - 1;
        }
        st_stream_f4(dst + ((((size_t)n * Hs + Y) * Hs + X) * 4 + (i & 3)) * 4, zero);
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
    // one unit per thread and item: the largest power-of-two row count with rows * gpr <= 256
    int rows = 1;
    while (rows * 2 * g->gpr <= kThreads && S % (rows * 2) == 0) rows *= 2;
    g->rows_item = rows;
    g->slabs = S / rows;
    const long long items = (long long)g->nbx * g->nby * g->slabs;
    if (items > 0x7FFFFFFF) return PYLC_ERR_GEOMETRY;
    g->items = (int)items;
    return PYLC_OK;
}

static bool aligned16(const void *p, size_t pitch) { return ((uintptr_t)p % 16 == 0) && (pitch % 16 == 0); }

// Persistent grid: SM count x CTAs that fit per SM, capped by the amount of work.
template <typename K>
static unsigned persistent_ctas(K kernel, long long work_items) {
    int dev = 0, sms = 148, per_sm = 4;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    long long want = (long long)sms * per_sm;
    if (want > work_items) want = work_items;
    return (unsigned)(want < 1 ? 1 : want);
}

}  // namespace pylc

using namespace pylc;

extern "C" int pylc_tile_grid(int H, int W, int T, int S, int *nH, int *nW) {
    if (T <= 0 || S <= 0 || !nH || !nW) return PYLC_ERR_ARG;
    *nH = H >= T ? (H - T) / S + 1 : 0;
    *nW = W >= T ? (W - T) / S + 1 : 0;
    return PYLC_OK;
}

extern "C" int pylc_tile_gather_u8_stack(const uint8_t *src, int n_img, size_t img_stride, int H, int W, int ch, size_t src_pitch,
                                         int T, int S, uint8_t *dst, uint64_t *stat, pylc_stream_t stream) {
    if (ch != 1 && ch != 3) return PYLC_ERR_ARG;
    if (n_img < 1 || (n_img > 1 && img_stride < (size_t)H * src_pitch)) return PYLC_ERR_ARG;
    GatherGeom g;
    int rc = make_geom(src, H, W, ch, src_pitch, T, S, &g);
    if (rc) return rc;
    if (g.items == 0) return PYLC_OK;  // source smaller than one tile: nothing to write
    if (!dst) return PYLC_ERR_ARG;
    if ((uintptr_t)dst % 16) return PYLC_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const bool al = aligned16(src, src_pitch) && (n_img == 1 || img_stride % 16 == 0);
    auto *sp = reinterpret_cast<unsigned long long *>(stat);
    // staged form: strips of up to 32 KB (64 rows) of source rows per CTA, the stack in the grid's y dimension
    static const size_t strip_cap = [] {       // PYLC_IMG_STRIP_KB: A/B knob for the strip size (1..32 KB)
        const char *e = getenv("PYLC_IMG_STRIP_KB");
        const long kb = e ? atol(e) : 0;
        return (size_t)(kb >= 1 && kb <= 32 ? kb : kImgStripKB) * 1024;
    }();
    int strip_rows = 1;
    while (strip_rows * 2 <= S && (size_t)strip_rows * 2 * S * ch <= strip_cap && strip_rows < 64) strip_rows *= 2;
    const size_t strip = (size_t)strip_rows * S * ch;
    const int spb = (S + strip_rows - 1) / strip_rows;
    if (al && g.m <= 2 && strip <= 32 * 1024 && g.nbx <= 65535 && g.nby <= 65535 && S * ch / 16 >= 1) {
        StagedArgs sa;
        sa.strip_rows = strip_rows;
        const uint32_t cpr = (uint32_t)(S * ch / 16), gpr = (uint32_t)(S / 16);
        sa.magic_cpr = cpr > 1 ? (uint32_t)((0x100000000ull + cpr - 1) / cpr) : 0u;
        sa.magic_gpr = gpr > 1 ? (uint32_t)((0x100000000ull + gpr - 1) / gpr) : 0u;
        sa.img_stride = img_stride;
        const int per_launch = 65535 / g.nby;                  // grid.z = image * nby + block row
        for (int i0 = 0; i0 < n_img; i0 += per_launch) {
            const int ni = n_img - i0 < per_launch ? n_img - i0 : per_launch;
            const dim3 grid((unsigned)spb, (unsigned)g.nbx, (unsigned)(g.nby * ni));
            GatherGeom gi = g;
            gi.src = src + (size_t)i0 * img_stride;
            uint8_t *d = dst + (size_t)i0 * g.nH * g.nW * ch * T * T;
            unsigned long long *s = sp ? sp + (size_t)i0 * g.nH * g.nW * ch * 2 : nullptr;
            if (ch == 1) {
                if (stat) gather_img_staged_kernel<1, true><<<grid, kThreads, strip, st>>>(gi, d, s, sa);
                else gather_img_staged_kernel<1, false><<<grid, kThreads, strip, st>>>(gi, d, s, sa);
            } else {
                if (stat) gather_img_staged_kernel<3, true><<<grid, kThreads, strip, st>>>(gi, d, s, sa);
                else gather_img_staged_kernel<3, false><<<grid, kThreads, strip, st>>>(gi, d, s, sa);
            }
            rc = finish_launch();
            if (rc) return rc;
        }
        return PYLC_OK;
    }
    // unaligned sources / T/S > 2: the persistent per-thread kernel, one launch per image of the stack
    const bool al1 = aligned16(src, src_pitch) && (n_img == 1 || img_stride % 16 == 0);
#define LAUNCH(CH, AL, ST) \
    gather_img_kernel<CH, AL, ST><<<persistent_ctas(gather_img_kernel<CH, AL, ST>, g.items), kThreads, 0, st>>>(gi, d, s)
    for (int i = 0; i < n_img; ++i) {
        GatherGeom gi = g;
        gi.src = src + (size_t)i * img_stride;
        uint8_t *d = dst + (size_t)i * g.nH * g.nW * ch * T * T;
        unsigned long long *s = sp ? sp + (size_t)i * g.nH * g.nW * ch * 2 : nullptr;
        if (ch == 1) {
            if (al1) { if (stat) LAUNCH(1, true, true); else LAUNCH(1, true, false); }
            else     { if (stat) LAUNCH(1, false, true); else LAUNCH(1, false, false); }
        } else {
            if (al1) { if (stat) LAUNCH(3, true, true); else LAUNCH(3, true, false); }
            else     { if (stat) LAUNCH(3, false, true); else LAUNCH(3, false, false); }
        }
        rc = finish_launch();
        if (rc) return rc;
    }
#undef LAUNCH
    return PYLC_OK;
}

extern "C" int pylc_tile_gather_u8(const uint8_t *src, int H, int W, int ch, size_t src_pitch, int T, int S,
                                   uint8_t *dst, uint64_t *stat, pylc_stream_t stream) {
    return pylc_tile_gather_u8_stack(src, 1, 0, H, W, ch, src_pitch, T, S, dst, stat, stream);
}

extern "C" int pylc_mask_gather_encode_hist_stack(const uint8_t *src, int n_img, size_t img_stride, int H, int W, size_t src_pitch,
                                                  int T, int S, const uint8_t *palette, int C, uint8_t *dst, int64_t *px_dist,
                                                  pylc_stream_t stream) {
    if (!palette) return PYLC_ERR_ARG;
    if (C < 1 || C > PYLC_MAX_CLASSES) return PYLC_ERR_CLASSES;
    if (n_img < 1 || (n_img > 1 && img_stride < (size_t)H * src_pitch)) return PYLC_ERR_ARG;
    GatherGeom g;
    int rc = make_geom(src, H, W, 3, src_pitch, T, S, &g);
    if (rc) return rc;
    PaletteHash ph;
    rc = build_palette_hash(palette, C, &ph);
    if (rc) return rc;
    if (g.items == 0) return PYLC_OK;
    if (!dst) return PYLC_ERR_ARG;
    if ((uintptr_t)dst % 16) return PYLC_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const bool al = aligned16(src, src_pitch) && (n_img == 1 || img_stride % 16 == 0);
    // TMA form (16-byte aligned rows, S a multiple of 256, T/S <= 2): the reference's two geometries; the whole
    // stack is ONE launch (3-D tensor map over [n_img][H][pitch / 4])
    rc = tma_disabled() ? -1
                        : launch_mask_gather_tma(src, n_img, img_stride, H, W, src_pitch, T, S, g.nH, g.nW, ph, C, dst,
                                                 reinterpret_cast<long long *>(px_dist), st);
    if (rc != -1) return rc;
#define LAUNCH(AL, NG, HS)                                                                                         \
    gather_mask_kernel<AL, NG, HS><<<persistent_ctas(gather_mask_kernel<AL, NG, HS>, g.items), kThreads, 0, st>>>( \
        gi, ph, C, d, pd)
#define LAUNCH_CUR(AL, NG, HS)                                                                          \
    gather_mask_cursor_kernel<AL, NG, HS>                                                              \
        <<<persistent_ctas(gather_mask_cursor_kernel<AL, NG, HS>, g.items), kThreads, 0, st>>>(gi, ph, C, d, pd)
#define PICK(L, AL)                                                 \
    switch (px_dist ? counter_groups(C) : -1) {                     \
        case -1: L(AL, 5, false); break;                            \
        case 5: L(AL, 5, true); break;                              \
        case 6: L(AL, 6, true); break;                              \
        case 7: L(AL, 7, true); break;                              \
        default: L(AL, 0, true); break;                             \
    }
    for (int i = 0; i < n_img; ++i) {       // per-thread kernels: one launch per image of the stack
        GatherGeom gi = g;
        gi.src = src + (size_t)i * img_stride;
        uint8_t *d = dst + (size_t)i * g.nH * g.nW * T * T;
        long long *pd = px_dist ? reinterpret_cast<long long *>(px_dist) + (size_t)i * g.nH * g.nW * C : nullptr;
        if (g.m <= 2) {
            if (al) { PICK(LAUNCH_CUR, true) } else { PICK(LAUNCH_CUR, false) }
        } else {
            if (al) { PICK(LAUNCH, true) } else { PICK(LAUNCH, false) }
        }
        rc = finish_launch();
        if (rc) return rc;
    }
#undef PICK
#undef LAUNCH_CUR
#undef LAUNCH
    return PYLC_OK;
}

extern "C" int pylc_mask_gather_encode_hist(const uint8_t *src, int H, int W, size_t src_pitch, int T, int S,
                                            const uint8_t *palette, int C, uint8_t *dst, int64_t *px_dist,
                                            pylc_stream_t stream) {
    return pylc_mask_gather_encode_hist_stack(src, 1, 0, H, W, src_pitch, T, S, palette, C, dst, px_dist, stream);
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
    cudaStream_t st = (cudaStream_t)stream;
    // staged form: 16-byte aligned rows, T/S <= 2, and a strip of source rows that fits shared memory
    const size_t strip = (size_t)kNormRows * S * ch;
    const int gpb = (S + kNormRows - 1) / kNormRows;
    if (aligned16(src, src_pitch) && g.m <= 2 && strip <= 40 * 1024 && (long long)g.nbx * g.nby * gpb < 0x7FFFFFFF) {
        const unsigned grid = (unsigned)(g.nbx * g.nby * gpb);
        if (ch == 1) gather_norm_staged_kernel<1><<<grid, kThreads, strip, st>>>(g, np, dst, gpb);
        else gather_norm_staged_kernel<3><<<grid, kThreads, strip, st>>>(g, np, dst, gpb);
        return finish_launch();
    }
    const bool al = ((uintptr_t)src % 4 == 0) && (src_pitch % 4 == 0);   // 4-byte units on this path
#define LAUNCH(CH, AL) \
    gather_norm_kernel<CH, AL><<<persistent_ctas(gather_norm_kernel<CH, AL>, g.items), kThreads, 0, st>>>(g, np, dst)
    if (ch == 1) { if (al) LAUNCH(1, true); else LAUNCH(1, false); }
    else         { if (al) LAUNCH(3, true); else LAUNCH(3, false); }
#undef LAUNCH
    return finish_launch();
}

extern "C" int pylc_tile_gather_norm_s2d_f32(const uint8_t *src, int H, int W, int ch, size_t src_pitch, int T, int S,
                                             const float *mean, const float *std, float post_div, float *dst,
                                             pylc_stream_t stream) {
    if (!mean || !std || (ch != 1 && ch != 3)) return PYLC_ERR_ARG;
    if (T % 2) return PYLC_ERR_GEOMETRY;
    GatherGeom g;
    int rc = make_geom(src, H, W, ch, src_pitch, T, S, &g);
    if (rc) return rc;
    if ((long long)g.nH * g.nW == 0) return PYLC_OK;
    if (!dst) return PYLC_ERR_ARG;
    if ((uintptr_t)dst % 16) return PYLC_ERR_ALIGN;
    NormParams np;
    for (int k = 0; k < 3; ++k) {
        np.mean[k] = mean[ch == 1 ? 0 : k];
        np.std[k] = std[ch == 1 ? 0 : k];
    }
    np.post_div = post_div;
    np.out_ch = 16;
    if ((long long)g.nH * g.nW * (6 * (T / 2 + 3) - 9) > 0x3FFFFFFFll) return PYLC_ERR_GEOMETRY;
    const long long items = (long long)g.nbx * g.nby * (S / 2);
    if (items > 0x7FFFFFFF) return PYLC_ERR_GEOMETRY;
    cudaStream_t st = (cudaStream_t)stream;
    // staged form: 16-byte aligned rows, T/S <= 2, and a strip of source rows that fits shared memory
    const size_t strip = (size_t)2 * kS2dRows * S * ch;
    const int gpb = (S / 2 + kS2dRows - 1) / kS2dRows;
    if (aligned16(src, src_pitch) && g.m <= 2 && strip <= 40 * 1024 && (long long)g.nbx * g.nby * gpb < 0x7FFFFFFF) {
        const unsigned grid = (unsigned)(g.nbx * g.nby * gpb);
        if (ch == 1) gather_norm_s2d_staged_kernel<1><<<grid, kThreads, strip, st>>>(g, np, dst, gpb);
        else gather_norm_s2d_staged_kernel<3><<<grid, kThreads, strip, st>>>(g, np, dst, gpb);
        return finish_launch();
    }
    const bool al2 = ((uintptr_t)src % 2 == 0) && (src_pitch % 2 == 0);
#define LAUNCH(CH, AL) \
    gather_norm_s2d_kernel<CH, AL><<<persistent_ctas(gather_norm_s2d_kernel<CH, AL>, items), kThreads, 0, st>>>(g, np, dst)
    if (ch == 1) { if (al2) LAUNCH(1, true); else LAUNCH(1, false); }
    else         { if (al2) LAUNCH(3, true); else LAUNCH(3, false); }
#undef LAUNCH
    return finish_launch();
}
