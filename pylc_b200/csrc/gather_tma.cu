// Mask tile gather + palette encode + per-tile class histogram, TMA form (sm_100a).
// Replaces Extractor.__split + tools.class_encode + the one-hot histogram of get_profile for 16-byte
// aligned sources (reference utils/extract.py:296-310, 195-214; utils/tools.py:435-449;
// utils/profile.py:109-111).
//
// Data movement is done by the copy engine, threads only look colours up:
//   * the source mask is described by a 2-D tensor map over 32-bit elements ([H][pitch/4]); a box is
//     32 rows x 256 pixels (768 B per row, 24 KB) and lands densely in shared memory, signalled on an
//     mbarrier (cp.async.bulk.tensor.2d ... mbarrier::complete_tx::bytes) -- SASS UTMALDG;
//   * the 256 threads of a CTA each own two 16-pixel units of the box, 16 rows apart (48 B in, 16 B out
//     each, lane-contiguous and bank-conflict free), and a warp encodes its 64 units with
//     warp_encode_units (palette_warp.cuh): one look-up per uniform 4-pixel group, mixed groups compacted
//     and re-encoded pixel by pixel;
//   * the encoded 32 x 256 byte box goes back through shared memory and ONE thread stores it to every
//     destination tile that contains it (1 tile for S = T, up to 4 for S = T/2) with
//     cp.async.bulk.tensor.3d stores over the [n][T][T/4] tile tensor -- SASS UTMASTG.
// CTAs are persistent over a contiguous range of boxes with a kStages-deep ring of input boxes, so the
// load of box k+1 is in flight while box k is encoded; two output boxes alternate, and thread 0 drains the
// stores of box k-1 (wait_group.read 0) before the barrier that lets box k+1 be written.
// Per-tile histograms: register counters per lane (GroupCounter: anchors count four pixels at once),
// flushed per warp with 64-bit global atomics whenever the S x S block changes -- no CTA barrier besides
// the one per box.
#include "palette_warp.cuh"

#include <stdlib.h>
#include <type_traits>

namespace pylc {

#ifndef PYLC_GT_ROWS
#define PYLC_GT_ROWS 32
#endif
#ifndef PYLC_GT_CTAS
#define PYLC_GT_CTAS 3
#endif
#ifndef PYLC_GT_STAGES
#define PYLC_GT_STAGES 2
#endif
constexpr int kBoxPx = 256, kBoxRows = PYLC_GT_ROWS; // a box: 32 rows x 256 pixels; a lane owns two 16-pixel units of it
constexpr int kNU = kBoxRows / 16;
constexpr int kUnitIn = 16 * kBoxPx * 3, kUnitOut = 16 * kBoxPx;      // the units of one lane are 16 rows apart
constexpr int kBoxIn = kBoxRows * kBoxPx * 3, kBoxOut = kBoxRows * kBoxPx;
constexpr int kStages = PYLC_GT_STAGES, kOutBufs = 2;
constexpr int kMaskCtasPerSm = PYLC_GT_CTAS;

struct TmaGeom {
    int T, S, nH, nW, m;
    int nbx, nby;          // S-blocks across / down
    int xparts, rgroups;   // boxes across / down one block
    int per_block;         // xparts * rgroups
    int items_img;         // nbx * nby * per_block: boxes of one image
    int tiles_img;         // nH * nW
    int items;             // n_img * items_img (a stack of equally sized sources is one launch)
};

struct BoxPos {
    int blk, bx, by, rg, xp;   // blk numbers the S-blocks of the whole stack (change detection only)
    int img, tile0;            // source image of the stack and its first destination tile
};
__device__ __forceinline__ BoxPos box_decode(const TmaGeom &g, int item) {
    BoxPos p;
    p.img = item / g.items_img;
    const int li = item - p.img * g.items_img;
    const int lb = li / g.per_block;
    const int r = li - lb * g.per_block;
    p.rg = r / g.xparts;
    p.xp = r - p.rg * g.xparts;
    p.by = lb / g.nbx;
    p.bx = lb - p.by * g.nbx;
    p.blk = p.img * (g.nbx * g.nby) + lb;
    p.tile0 = p.img * g.tiles_img;
    return p;
}
__device__ __forceinline__ BoxPos box_next(const TmaGeom &g, BoxPos p) {
    if (++p.xp == g.xparts) {
        p.xp = 0;
        if (++p.rg == g.rgroups) {
            p.rg = 0;
            ++p.blk;
            if (++p.bx == g.nbx) {
                p.bx = 0;
                if (++p.by == g.nby) {
                    p.by = 0;
                    ++p.img;
                    p.tile0 += g.tiles_img;
                }
            }
        }
    }
    return p;
}

template <class CC>
__device__ __forceinline__ void flush_warp_hist_tma(const TmaGeom &g, const BoxPos &p, CC &cc, int C, long long *px_dist) {
    const int lane = threadIdx.x & 31;
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
        const int r_lo = max(0, p.by - g.m + 1), r_hi = min(g.nH - 1, p.by);
        const int c_lo = max(0, p.bx - g.m + 1), c_hi = min(g.nW - 1, p.bx);
        for (int r = r_lo; r <= r_hi; ++r)
            for (int c = c_lo; c <= c_hi; ++c)
                atomicAdd((unsigned long long *)&px_dist[(size_t)(p.tile0 + r * g.nW + c) * C + lane], (unsigned long long)mine);
    }
}

// NG > 0: GroupCounter<NG> histograms (C <= 14); NG = 0: ByteCounter over the finished box (C <= 32); HIST off: none
// STACK: the source is a stack of equally sized images ([n_img][H][pitch/4] tensor map, 3-D box loads)
template <int NG, bool HIST, bool STACK>
__global__ void __launch_bounds__(kThreads, kMaskCtasPerSm)
    gather_mask_tma_kernel(const __grid_constant__ CUtensorMap tm_src, const __grid_constant__ CUtensorMap tm_dst,
                           const TmaGeom g, const __grid_constant__ PaletteHash ph, int C, long long *__restrict__ px_dist) {
    extern __shared__ __align__(128) uint8_t s_dyn[];       // kStages input boxes, then kOutBufs output boxes
    __shared__ uint32_t s_tab[256];
    __shared__ __align__(8) unsigned long long s_bar[kStages];
    __shared__ __align__(16) uint8_t s_queue[kWarps][256];

    const int tid = threadIdx.x, warp = tid >> 5;
    const int first = (int)((long long)g.items * blockIdx.x / gridDim.x);
    const int n = (int)((long long)g.items * (blockIdx.x + 1) / gridDim.x) - first;
    if (n <= 0) return;

    const uint32_t in0 = smem_u32(s_dyn), out0 = in0 + kStages * kBoxIn, bar0 = smem_u32(s_bar);
    auto issue_load = [&](const BoxPos &p, int stage) {   // source coordinates in 32-bit elements / rows
        const uint32_t bar = bar0 + 8u * stage;
        mbar_arrive_expect_tx(bar, kBoxIn);
        const int c0 = (p.bx * g.S + p.xp * kBoxPx) * 3 / 4, c1 = p.by * g.S + p.rg * kBoxRows;
        if (STACK) tma_load_3d(in0 + (uint32_t)stage * kBoxIn, &tm_src, c0, c1, p.img, bar);
        else tma_load_2d(in0 + (uint32_t)stage * kBoxIn, &tm_src, c0, c1, bar);
    };
    // thread 0 puts the first boxes in flight before anything else happens in the CTA: the table copy and
    // the barrier below overlap the loads' latency (the other threads only need the barriers initialised
    // before they WAIT on them, which the __syncthreads guarantees)
    BoxPos pl = box_decode(g, first);       // next box to load (thread 0 only)
    if (tid == 0) {
        tma_prefetch_desc(&tm_src);
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8u * s, 1);
        mbar_fence_init();
        for (int k = 0; k < kStages && k < n; ++k) {
            issue_load(pl, k);
            pl = box_next(g, pl);
        }
        tma_prefetch_desc(&tm_dst);
    }
    s_tab[tid] = ph.tab[tid];
    __syncthreads();

    const uint32_t mul = ph.mul, tab = smem_u32(s_tab);
    const uint32_t miss_e = 1u << 24;       // unmatched colours are class 1 (utils/tools.py:437)
    const uint32_t q_warp = smem_u32(&s_queue[warp][0]);
    const uint32_t in_warp0 = in0 + (uint32_t)warp * (32 * 48), out_warp0 = out0 + (uint32_t)warp * (32 * 16);
    constexpr bool GROUPS = HIST && NG > 0;
    using GC = typename std::conditional<GROUPS, GroupCounter<(NG > 0 ? NG : 1)>, NoCounter>::type;
    GC gc;
    ByteCounter bc;
    if constexpr (GROUPS) gc.reset();
    if (HIST && NG == 0) bc.reset();
    int since_flush = 0;

    BoxPos p = box_decode(g, first);
    int stage = 0, obuf = 0;
    uint32_t parity = 0;
    for (int k = 0; k < n; ++k) {
        mbar_wait(bar0 + 8u * stage, parity);
        const uint32_t in_s = (uint32_t)stage * kBoxIn, out_s = (uint32_t)obuf * kBoxOut;
        warp_encode_units<kNU, GROUPS>(in_warp0 + in_s, kUnitIn, out_warp0 + out_s, kUnitOut, q_warp, tab, mul, miss_e, gc, (1u << kNU) - 1u);
        const BoxPos pn = box_next(g, p);
        const bool chg = pn.blk != p.blk || k + 1 == n;
        if (HIST && NG == 0) {              // wide palettes: count the finished units byte by byte
            __syncwarp();
#pragma unroll
            for (int j = 0; j < kNU; ++j) {
                const uint4 r = lds128(out0 + out_s + (uint32_t)j * kUnitOut + (uint32_t)tid * 16u);
                bc.add16(r.x, r.y, r.z, r.w);
            }
            if (chg || ++since_flush >= ByteCounter::kFlushUnits / 2) {
                flush_warp_hist_tma(g, p, bc, C, px_dist);
                since_flush = 0;
            }
        }
        if constexpr (GROUPS) {
            if (chg || ++since_flush >= GC::kFlushItems) {
                flush_warp_hist_tma(g, p, gc, C, px_dist);
                since_flush = 0;
            }
        }
        if (tid == 0) tma_store_wait_read<0>();     // the stores of box k-1 have read their buffer (rewritten at box k+1)
        fence_proxy_async_smem();
        __syncthreads();                    // the output box is complete and nobody reads input box `stage` any more
        if (tid == 0) {
            const int r_lo = max(0, p.by - g.m + 1), r_hi = min(g.nH - 1, p.by);
            const int c_lo = max(0, p.bx - g.m + 1), c_hi = min(g.nW - 1, p.bx);
            for (int r = r_lo; r <= r_hi; ++r)
                for (int c = c_lo; c <= c_hi; ++c)
                    tma_store_3d(&tm_dst, ((p.bx - c) * g.S + p.xp * kBoxPx) / 4, (p.by - r) * g.S + p.rg * kBoxRows, p.tile0 + r * g.nW + c, out0 + out_s);
            tma_store_commit();
            if (k + kStages < n) {
                issue_load(pl, stage);
                pl = box_next(g, pl);
            }
        }
        p = pn;
        if (++stage == kStages) {
            stage = 0;
            parity ^= 1u;
        }
        obuf ^= 1;
    }
    if (tid == 0) tma_store_wait_read<0>();
}

// Returns PYLC_OK after launching, or -1 when this form does not apply (the caller falls back to the
// per-thread kernels): needs 16-byte aligned rows, S a multiple of 256 and T/S <= 2.  n_img > 1: a stack of
// equally sized sources `img_stride` bytes apart (a multiple of 16), tiles of image i at dst tile i * nH * nW.
int launch_mask_gather_tma(const uint8_t *src, int n_img, size_t img_stride, int H, int W, size_t pitch, int T, int S, int nH, int nW,
                           const PaletteHash &ph, int C, uint8_t *dst, long long *px_dist, cudaStream_t st) {
    if (((uintptr_t)src % 16) || (pitch % 16) || ((uintptr_t)dst % 16) || S % kBoxPx || S % kBoxRows || T % S || T / S > 2 || T % 16) return -1;
    if (n_img < 1 || (n_img > 1 && (img_stride % 16 || img_stride < (size_t)H * pitch))) return -1;
    TmaGeom g;
    g.T = T, g.S = S, g.nH = nH, g.nW = nW, g.m = T / S;
    g.nbx = nW - 1 + g.m, g.nby = nH - 1 + g.m;
    g.xparts = S / kBoxPx, g.rgroups = S / kBoxRows;
    g.per_block = g.xparts * g.rgroups;
    const long long items = (long long)g.nbx * g.nby * g.per_block * n_img;
    if (items <= 0 || items > 0x7FFFFFFF || (long long)nH * nW * n_img > 0x7FFFFFFF) return -1;
    g.items = (int)items;
    g.items_img = (int)(items / n_img);
    g.tiles_img = nH * nW;

    CUtensorMap tm_src, tm_dst;
    if (n_img == 1) {
        const uint64_t dims[2] = {(uint64_t)(pitch / 4), (uint64_t)H};
        const uint64_t strides[1] = {(uint64_t)pitch};
        const uint32_t box[2] = {kBoxPx * 3 / 4, kBoxRows};
        if (!tma_encode_u32(&tm_src, src, 2, dims, strides, box)) return -1;
    } else {
        const uint64_t dims[3] = {(uint64_t)(pitch / 4), (uint64_t)H, (uint64_t)n_img};
        const uint64_t strides[2] = {(uint64_t)pitch, (uint64_t)img_stride};
        const uint32_t box[3] = {kBoxPx * 3 / 4, kBoxRows, 1};
        if (!tma_encode_u32(&tm_src, src, 3, dims, strides, box)) return -1;
    }
    {
        const uint64_t dims[3] = {(uint64_t)(T / 4), (uint64_t)T, (uint64_t)nH * nW * n_img};
        const uint64_t strides[2] = {(uint64_t)T, (uint64_t)T * T};
        const uint32_t box[3] = {kBoxPx / 4, kBoxRows, 1};
        if (!tma_encode_u32(&tm_dst, dst, 3, dims, strides, box)) return -1;
    }
    const size_t smem = (size_t)kStages * kBoxIn + (size_t)kOutBufs * kBoxOut;
    int dev = 0, sms = 148;
    const char *cta_s = getenv("PYLC_TMA_CTAS");
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define LAUNCH(NG, HS)                                                                                         \
    do {                                                                                                       \
        if (n_img > 1) LAUNCH_K(NG, HS, true);                                                                 \
        else LAUNCH_K(NG, HS, false);                                                                          \
    } while (0)
#define LAUNCH_K(NG, HS, STK)                                                                                  \
    do {                                                                                                       \
        auto kern = gather_mask_tma_kernel<NG, HS, STK>;                                                          \
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {   \
            cudaGetLastError();                                                                                \
            return -1;                                                                                         \
        }                                                                                                      \
        int per_sm = 0;                                                                                        \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1; \
        if (per_sm > kMaskCtasPerSm) per_sm = kMaskCtasPerSm;                                                  \
        if (cta_s) per_sm = atoi(cta_s);                                                                       \
        long long ctas = (long long)sms * per_sm;                                                              \
        if (ctas > items) ctas = items;                                                                        \
        kern<<<(unsigned)ctas, kThreads, smem, st>>>(tm_src, tm_dst, g, ph, C, px_dist);                       \
    } while (0)
    switch (px_dist ? counter_groups(C) : -1) {
        case -1: LAUNCH(5, false); break;
        case 5: LAUNCH(5, true); break;
        case 6: LAUNCH(6, true); break;
        case 7: LAUNCH(7, true); break;
        default: LAUNCH(0, true); break;
    }
#undef LAUNCH
#undef LAUNCH_K
    return finish_launch();
}

}  // namespace pylc
