// Nearest-neighbour resample + ground-truth palette encode + confusion matrix, TMA form (sm_100a).
// Replaces cv2.resize(INTER_NEAREST) of the predicted mask, the two class_encode passes of Evaluator.load
// and the scikit-learn confusion matrix (reference utils/tools.py:316-317, utils/evaluate.py:87-119,150-176,
// utils/metrics.py:45-87) for the evaluation hot case: counts only (no label / RGB outputs), C <= 11,
// 16-byte aligned ground-truth rows and label rows.
//
// Data movement is the copy engine's: a dedicated producer warp keeps a kCfStages-deep ring of stages full,
// each stage = one 16-row x 256-pixel box of the RGB ground truth ([h_full][pitch/4] tensor map, 12 KB) plus
// the window of the fitted label map the box resamples from ([h][w/4] tensor map, 16 rows x 288 B), both
// cp.async.bulk.tensor.2d loads signalled on the stage's `full` mbarrier; consumer warps release a stage on
// its `empty` mbarrier.  The eight consumer warps only look things up:
//   * a lane owns one 16-pixel unit of the box (warp w: rows 2w, 2w+1; 16 lanes per row); its 16 predicted
//     labels come from the staged label window with four thread-constant PRMT selectors (the column map of
//     a CTA never changes: a CTA owns one 256-pixel column block and a contiguous range of rows);
//   * per 4-pixel group ONE palette look-up (the group's first pixel) and ONE counter update of weight 4 when
//     the 12 ground-truth bytes are one repeated pixel and the four predicted labels are equal -- label masks
//     and x4-up-sampled predictions are piecewise constant; the other ("mixed") groups are copied, 16 bytes
//     each, into a per-warp shared-memory queue at ballot-derived positions and counted pixel by pixel only
//     when 32 of them have piled up, one group per lane, so the per-pixel path always runs with a full warp.
//     Every count is exact: a group is taken from its first pixel only when all of it was compared equal;
//   * counters are lane-private 16-bit columns in shared memory, tab[warp][t*C+p][lane]: one conflict-free
//     load / add / store per update, no atomics; summed and flushed with <= C*C global atomics per CTA.
// The coverage injection of Evaluator.validate (the first n_inject flat pixels count as (i, i),
// utils/evaluate.py:172-174) is applied as a correction, one lane per injected pixel: -1 on its true pair,
// +1 on (i, i).  Lanes whose maps do not fit the staged window (down-sampling maps) gather their labels from
// global memory; results are identical.  Precondition (not checked, see include/pylc_b200.h): labels < C.
//
// Measured (B200, 6000x4000, 288 CTAs in one wave; tools/exp/cfdbg.py with -DPYLC_CF_DEBUG): set-up done 1.8 us
// after kernel entry, first box consumed at 4.8 us (every CTA requests three stages at once: 14 MB of fill),
// 21 boxes per CTA in 18.3 us = 5.2 TB/s during the loop, flush 0.7 us.  The first version's producer spun on
// the `empty` barriers without sleeping and issued 11 % of the kernel's instructions; its mixed groups were
// counted box by box with ~9 of 32 lanes active, which was half of all instructions (hence the deferred queue).
#include "palette_warp.cuh"

namespace pylc {

namespace {
constexpr int kCfWarps = 8;                           // consumer warps
constexpr int kCfThreads = (kCfWarps + 1) * 32;       // + the producer warp
constexpr int kCfBoxPx = 256, kCfBoxRows = 2 * kCfWarps;
constexpr int kCfRowIn = kCfBoxPx * 3;                // 768 B of RGB per box row
constexpr int kCfBoxIn = kCfBoxRows * kCfRowIn;       // 12288
constexpr int kCfLabW = 288, kCfLabRows = kCfBoxRows; // label window: up-sampling maps touch <= 271 B x 16 rows
constexpr int kCfLabBox = kCfLabW * kCfLabRows;       // 4608
constexpr int kCfStage = kCfBoxIn + kCfLabBox;        // 16896 (a multiple of 128)
constexpr int kCfStages = 3;
constexpr int kCfCtasPerSm = 2;
constexpr int kCfMaxBoxes = 1024;                     // per CTA: keeps every 16-bit lane counter below 1024 * (16 + 3 * 4)
constexpr int kCfDense = 48;                          // of 128 groups per warp and box: more mixed groups than this are counted in place
constexpr int kCfQueue = 96;                          // queue entries per warp: < 32 left over + <= kCfDense new ones per box
static_assert(kCfStage % 128 == 0 && kCfBoxIn % 128 == 0, "TMA destinations are 128-byte aligned");
}  // namespace

struct CfArgs {
    const uint8_t *labels;
    const int32_t *x_ofs, *y_ofs;
    const uint8_t *gt_rgb;
    size_t gt_pitch;
    int h, w, h_full, w_full, C, n_inject;
    int nbx, nby, boxes_per_slot;
    long long *conf;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// producer-side wait: sleeps between polls so the spinning thread does not eat the consumers' issue slots
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITB_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONEB_%=;\n"
        "nanosleep.u32 256;\n"
        "bra WAITB_%=;\n"
        "DONEB_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kCfWarps * 32) : "memory"); }
// producer warp and consumer warps meet once after their set-up, at one call site (compute-sanitizer's synccheck
// reports a barrier that the warps of a CTA reach from different program points)
__device__ __forceinline__ void cta_sync_all() {
    __syncwarp();                 // the producer's lane 0 ran the prologue alone: converge before the aligned barrier
    __syncthreads();
}

// lane-private 16-bit counter += w
__device__ __forceinline__ void bump16(uint32_t saddr, uint32_t w) {
    asm volatile("{ .reg .u16 t, v; cvt.u16.u32 v, %1; ld.shared.u16 t, [%0]; add.u16 t, t, v; st.shared.u16 [%0], t; }" ::"r"(saddr), "r"(w)
                 : "memory");
}

// one queued mixed group: exact per-pixel encode of its 12 ground-truth bytes, four (or nv) single counts
__device__ __forceinline__ void count_queued(uint32_t entry, uint32_t my_col, uint32_t tab, uint32_t mul, uint32_t miss_e) {
    const uint4 q = lds128(entry);
    const uint32_t idx = encode_group_s(q.x, q.y, q.z, tab, mul, miss_e) + (q.w & 0x3FFFFFFFu);
    const int nv = (int)(q.w >> 30) + 1;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
        if (jj < nv) bump16(my_col + __byte_perm(idx, 0, 0x4440u | jj) * 64u, 1u);
}

#ifdef PYLC_CF_DEBUG      // per-CTA timeline for tools/exp/cfdbg.py: make EXTRA=-DPYLC_CF_DEBUG (never in the shipped library)
__device__ unsigned long long g_cf_dbg[8 * 512];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define DBG(slot) do { if (tid == 0 && blockIdx.x < 512) { g_cf_dbg[blockIdx.x * 8 + (slot)] = gtime(); g_cf_dbg[blockIdx.x * 8 + 4 + (slot)] = clock64(); } } while (0)
#else
#define DBG(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(kCfThreads, kCfCtasPerSm)
    resample_confusion_tma_kernel(const __grid_constant__ CUtensorMap tm_gt, const __grid_constant__ CUtensorMap tm_lab, const CfArgs a,
                                  const __grid_constant__ PaletteHash ph) {
    extern __shared__ __align__(128) uint8_t s_dyn[];      // kCfStages stages | counters [warps][CC][32] u16 | queues [warps][kCfQueue] uint4
    __shared__ uint32_t s_tab[256];
    __shared__ __align__(8) unsigned long long s_full[kCfStages], s_empty[kCfStages];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    DBG(0);
    const int CC = a.C * a.C;
    const int cb = blockIdx.x % a.nbx, slot = blockIdx.x / a.nbx;
    const int by0 = slot * a.boxes_per_slot;
    const int n = min(a.boxes_per_slot, a.nby - by0);
    const int X0 = cb * kCfBoxPx;
    const uint32_t st0 = smem_u32(s_dyn), full0 = smem_u32(s_full), empty0 = smem_u32(s_empty);

    // ---- set-up: producer warp and consumer warps each do their part, then meet at ONE barrier ---------------
    int wbase_p = 0;                     // producer: first byte column of the label window (16-byte aligned)
    int X = 0, valid_x = 0, wbase = 0;   // consumers: the lane's first pixel, its valid pixels, the window origin
    int sxs[16];
    uint32_t cnt0 = 0, my_tab = 0, my_col = 0, q_warp = 0;
    if (warp == kCfWarps) {
        if (lane == 0) {
            // the map entries the label-window coordinates depend on: requested first, used after the ground-truth
            // boxes (which depend on nothing) are in flight
            const int x_first = n > 0 ? __ldg(a.x_ofs + X0) : 0;
            int sy0 = n > 0 ? __ldg(a.y_ofs + by0 * kCfBoxRows) : 0;
            int sy1 = n > 1 ? __ldg(a.y_ofs + (by0 + 1) * kCfBoxRows) : 0;
            int sy2 = n > 2 ? __ldg(a.y_ofs + (by0 + 2) * kCfBoxRows) : 0;
            tma_prefetch_desc(&tm_gt);
            tma_prefetch_desc(&tm_lab);
#pragma unroll
            for (int s = 0; s < kCfStages; ++s) {
                mbar_init(full0 + 8u * s, 1);
                mbar_init(empty0 + 8u * s, kCfWarps);
            }
            mbar_fence_init();
            static_assert(kCfStages == 3, "the prologue below fills three stages");
#pragma unroll
            for (int s = 0; s < kCfStages; ++s) {
                if (s < n) {
                    mbar_arrive_expect_tx(full0 + 8u * s, kCfStage);     // out-of-bounds parts of a box are zero-filled and still counted
                    tma_load_2d(st0 + (uint32_t)s * kCfStage, &tm_gt, X0 * 3 / 4, (by0 + s) * kCfBoxRows, full0 + 8u * s);
                }
            }
            wbase_p = x_first & ~15;
            if (0 < n) tma_load_2d(st0 + kCfBoxIn, &tm_lab, wbase_p / 4, sy0, full0);
            if (1 < n) tma_load_2d(st0 + kCfStage + kCfBoxIn, &tm_lab, wbase_p / 4, sy1, full0 + 8u);
            if (2 < n) tma_load_2d(st0 + 2 * kCfStage + kCfBoxIn, &tm_lab, wbase_p / 4, sy2, full0 + 16u);
        }
    } else {
        // column map of the lane's 16 pixels (constant for the CTA), issued before the set-up stores
        X = X0 + (lane & 15) * 16;
        valid_x = max(0, min(16, a.w_full - X));
        if (valid_x == 16 && ((uintptr_t)a.x_ofs & 15) == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int4 v = __ldg(reinterpret_cast<const int4 *>(a.x_ofs + X) + k);
                sxs[4 * k] = v.x, sxs[4 * k + 1] = v.y, sxs[4 * k + 2] = v.z, sxs[4 * k + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) sxs[j] = __ldg(a.x_ofs + min(X + j, a.w_full - 1));
        }
        wbase = __ldg(a.x_ofs + X0) & ~15;
        s_tab[tid] = (ph.tab[tid] & 0x00FFFFFFu) | (((ph.tab[tid] >> 24) * (uint32_t)a.C) << 24);   // class byte pre-multiplied by C
        cnt0 = st0 + kCfStages * kCfStage;
        my_tab = cnt0 + (uint32_t)(warp * CC) * 64u;
        my_col = my_tab + (uint32_t)lane * 2u;
        for (int i = lane; i < CC * 4; i += 32) sts128(my_tab + 16u * i, make_uint4(0u, 0u, 0u, 0u));
        q_warp = cnt0 + (uint32_t)(kCfWarps * CC) * 64u + (uint32_t)warp * (kCfQueue * 16u);
    }
    cta_sync_all();                      // every thread of the CTA, one call site: mbarriers initialised, table and counters ready

    // ---- producer warp: keeps the ring full ---------------------------------------------------------------
    if (warp == kCfWarps) {
        if (lane == 0) {
            int stage = 0;
            uint32_t round = 1;          // refill `round` of a stage waits for the consumers' release `round - 1`
            int sy0 = n > kCfStages ? __ldg(a.y_ofs + (by0 + kCfStages) * kCfBoxRows) : 0;
            for (int k = kCfStages; k < n; ++k) {
                const int Y0 = (by0 + k) * kCfBoxRows;
                const int sy0_next = k + 1 < n ? __ldg(a.y_ofs + Y0 + kCfBoxRows) : 0;
                mbar_wait_backoff(empty0 + 8u * stage, (round - 1u) & 1u);
                const uint32_t bar = full0 + 8u * stage, dst = st0 + (uint32_t)stage * kCfStage;
                mbar_arrive_expect_tx(bar, kCfStage);
                tma_load_2d(dst, &tm_gt, X0 * 3 / 4, Y0, bar);
                tma_load_2d(dst + kCfBoxIn, &tm_lab, wbase_p / 4, sy0, bar);
                sy0 = sy0_next;
                if (++stage == kCfStages) {
                    stage = 0;
                    ++round;
                }
            }
        }
        return;
    }

    // ---- consumer warps -----------------------------------------------------------------------------------
    DBG(1);
    if (n <= 0) return;
    const uint32_t mul = ph.mul, tab = smem_u32(s_tab);
    const uint32_t miss_e = (uint32_t)a.C << 24;      // unmatched colours are class 1 (utils/tools.py:437)

    // coverage injection (utils/evaluate.py:172-174): flat pixels i < n_inject count as (i, i) instead of their own
    // pair.  One lane per pixel, while the first boxes are in flight: -1 on the true pair, +1 on (i, i).
    if (blockIdx.x == 0 && warp == 0 && lane < a.n_inject) {
        const uint8_t *p = a.gt_rgb + 3 * lane;
        const uint32_t key = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
        const uint32_t t = lookup_entry_s(key, tab, mul, miss_e) >> 24;                      // class * C
        const uint32_t pr = __ldg(a.labels + (size_t)__ldg(a.y_ofs) * a.w + __ldg(a.x_ofs + lane));
        atomicAdd((unsigned long long *)&a.conf[t + pr], ~0ull);                              // -1
        atomicAdd((unsigned long long *)&a.conf[lane * a.C + lane], 1ull);
    }

    uint32_t cm_lo[4], cm_sel[4];
    bool fast = valid_x > 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int word0 = sxs[4 * k] & ~3;
        cm_sel[k] = 0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int d = sxs[4 * k + jj] - word0;
            fast = fast && d >= 0 && d <= 7;
            cm_sel[k] |= (uint32_t)(d & 7) << (4 * jj);
        }
        const int lo = word0 - wbase;
        fast = fast && lo >= 0 && lo + 8 <= kCfLabW;
        cm_lo[k] = (uint32_t)max(0, min(lo, kCfLabW - 8));
    }

    const uint32_t in_lane0 = (uint32_t)warp * (2 * kCfRowIn) + (uint32_t)lane * 48u;
    const int yl = 2 * warp + (lane >> 4);            // the lane's row inside a box

    // the row maps of the next box are fetched one box ahead
    int sy0 = __ldg(a.y_ofs + by0 * kCfBoxRows);
    int sy = __ldg(a.y_ofs + min(by0 * kCfBoxRows + yl, a.h_full - 1));
    int stage = 0, qn = 0;               // qn: entries in the warp's queue (warp-uniform)
    uint32_t parity = 0;
    for (int k = 0; k < n; ++k) {
        const int Y0 = (by0 + k) * kCfBoxRows, Y = Y0 + yl;
        int sy0_n = 0, sy_n = 0;
        if (k + 1 < n) {
            sy0_n = __ldg(a.y_ofs + Y0 + kCfBoxRows);
            sy_n = __ldg(a.y_ofs + min(Y + kCfBoxRows, a.h_full - 1));
        }
        const int valid = Y < a.h_full ? valid_x : 0;
        mbar_wait(full0 + 8u * stage, parity);
        if (k == 0) DBG(2);
        const uint32_t in_s = st0 + (uint32_t)stage * kCfStage, lab_s = in_s + kCfBoxIn;

        // ---- phase A: the lane's unit --------------------------------------------------------------------
        uint32_t pw[4];
        const int rsel = sy - sy0;
        if (fast && rsel >= 0 && rsel < kCfLabRows) {
            const uint32_t row = lab_s + (uint32_t)rsel * kCfLabW;
#pragma unroll
            for (int g = 0; g < 4; ++g) pw[g] = __byte_perm(lds32(row + cm_lo[g]), lds32(row + cm_lo[g] + 4u), cm_sel[g]);
        } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) pw[g] = 0;
            if (valid) {
                const uint8_t *lrow = a.labels + (size_t)sy * a.w;
#pragma unroll
                for (int j = 0; j < 16; ++j) pw[j >> 2] |= (uint32_t)__ldg(lrow + sxs[j]) << (8 * (j & 3));
            }
        }
        const uint32_t in_lane = in_s + in_lane0;
        const uint4 q0 = lds128(in_lane), q1 = lds128(in_lane + 16), q2 = lds128(in_lane + 32);
        const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
        uint32_t flags = 0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint32_t ga = w[3 * g], gb = w[3 * g + 1], gc = w[3 * g + 2];
            const uint32_t e = lookup_entry_s(ga, tab, mul, miss_e);
            const int nv = valid - 4 * g;                           // valid pixels of this group
            const uint32_t spread = group_spread(ga, gb, gc) | (pw[g] ^ __byte_perm(pw[g], 0, 0x0000));
            const bool whole = nv >= 4 && spread == 0;
            flags |= (nv > 0 && !whole) ? (1u << g) : 0u;
            if (whole) bump16(my_col + ((e >> 24) + (pw[g] & 0xFFu)) * 64u, 4u);
        }

        // ---- phase B: mixed groups -> queue (or counted in place when the box is noise-like) ---------------
        if (__any_sync(0xFFFFFFFFu, flags != 0)) {
            uint32_t m[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) m[g] = __ballot_sync(0xFFFFFFFFu, (flags >> g) & 1u);
            const int nq = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
            if (nq > kCfDense) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if ((flags >> g) & 1u) {
                        const uint32_t idx = encode_group_s(w[3 * g], w[3 * g + 1], w[3 * g + 2], tab, mul, miss_e) + pw[g];
                        const int nv = valid - 4 * g;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            if (jj < nv) bump16(my_col + __byte_perm(idx, 0, 0x4440u | jj) * 64u, 1u);
                    }
                }
            } else {
                const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if ((flags >> g) & 1u) {
                        const int nv = min(4, valid - 4 * g);       // 1..4; rides in the two top bits of the label word (labels < 32)
                        sts128(q_warp + (uint32_t)(qn + __popc(m[g] & lt)) * 16u,
                               make_uint4(w[3 * g], w[3 * g + 1], w[3 * g + 2], pw[g] | ((uint32_t)(nv - 1) << 30)));
                    }
                    qn += __popc(m[g]);
                }
            }
        }
        __syncwarp();                    // every lane is done with the stage; queue entries are visible to the warp
        if (lane == 0) mbar_arrive(empty0 + 8u * stage);
        while (qn >= 32) {               // 32 mixed groups piled up: one per lane, pixel by pixel
            qn -= 32;
            count_queued(q_warp + (uint32_t)(qn + lane) * 16u, my_col, tab, mul, miss_e);
            __syncwarp();                // the slots are free again before the next box writes them
        }
        sy0 = sy0_n;
        sy = sy_n;
        if (++stage == kCfStages) {
            stage = 0;
            parity ^= 1u;
        }
    }

    if (lane < qn) count_queued(q_warp + (uint32_t)lane * 16u, my_col, tab, mul, miss_e);    // the remainder (< 32)

    // ---- flush: column sums -> global matrix ----------------------------------------------------------------
    DBG(3);
    consumer_sync();
    for (int i = tid; i < CC; i += kCfWarps * 32) {
        unsigned long long t = 0;
#pragma unroll
        for (int wp = 0; wp < kCfWarps; ++wp) {
            const uint32_t col = cnt0 + (uint32_t)(wp * CC + i) * 64u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 v = lds128(col + 16u * ((q + i) & 3));        // rotated: spreads the banks
                t += (v.x & 0xFFFFu) + (v.x >> 16) + (v.y & 0xFFFFu) + (v.y >> 16) + (v.z & 0xFFFFu) + (v.z >> 16) + (v.w & 0xFFFFu) + (v.w >> 16);
            }
        }
        if (t) atomicAdd((unsigned long long *)&a.conf[i], t);
    }
}

// Returns PYLC_OK / a CUDA error after launching, or -1 when the form does not apply.
int launch_resample_confusion_tma(const uint8_t *labels, int h, int w, const int32_t *x_ofs, const int32_t *y_ofs, int h_full, int w_full,
                                  const uint8_t *gt_rgb, size_t gt_pitch, const PaletteHash &ph, int C, int n_inject, long long *conf,
                                  cudaStream_t st) {
    if (!gt_rgb || !conf || C > 11) return -1;
    if (((uintptr_t)gt_rgb % 16) || (gt_pitch % 16) || ((uintptr_t)labels % 16) || (w % 16) || w < 16) return -1;
    if (n_inject > w_full) return -1;    // the injected pixels lie in the first row
    CUtensorMap tm_gt, tm_lab;
    {
        const uint64_t dims[2] = {(uint64_t)(gt_pitch / 4), (uint64_t)h_full};
        const uint64_t strides[1] = {(uint64_t)gt_pitch};
        const uint32_t box[2] = {kCfRowIn / 4, kCfBoxRows};
        if (!tma_encode_u32(&tm_gt, gt_rgb, 2, dims, strides, box)) return -1;
    }
    {
        const uint64_t dims[2] = {(uint64_t)(w / 4), (uint64_t)h};
        const uint64_t strides[1] = {(uint64_t)w};
        const uint32_t box[2] = {kCfLabW / 4, kCfLabRows};
        if (!tma_encode_u32(&tm_lab, labels, 2, dims, strides, box)) return -1;
    }
    CfArgs a;
    a.labels = labels; a.x_ofs = x_ofs; a.y_ofs = y_ofs; a.gt_rgb = gt_rgb; a.gt_pitch = gt_pitch;
    a.h = h; a.w = w; a.h_full = h_full; a.w_full = w_full; a.C = C; a.n_inject = n_inject; a.conf = conf;
    a.nbx = (w_full + kCfBoxPx - 1) / kCfBoxPx;
    a.nby = (h_full + kCfBoxRows - 1) / kCfBoxRows;
    const size_t smem = (size_t)kCfStages * kCfStage + (size_t)kCfWarps * C * C * 64 + (size_t)kCfWarps * kCfQueue * 16;
    auto kern = resample_confusion_tma_kernel;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    int dev = 0, sms = 148, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCfThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (per_sm > kCfCtasPerSm) per_sm = kCfCtasPerSm;
    // one wave: a CTA owns a column block and a contiguous run of boxes down it
    int slots = sms * per_sm / a.nbx;
    if (slots < 1) slots = 1;
    if (slots > a.nby) slots = a.nby;
    a.boxes_per_slot = (a.nby + slots - 1) / slots;
    if (a.boxes_per_slot > kCfMaxBoxes) a.boxes_per_slot = kCfMaxBoxes;
    const int row_slots = (a.nby + a.boxes_per_slot - 1) / a.boxes_per_slot;
    kern<<<(unsigned)(a.nbx * row_slots), kCfThreads, smem, st>>>(tm_gt, tm_lab, a, ph);
    return finish_launch();
}

}  // namespace pylc
#ifdef PYLC_CF_DEBUG
extern "C" __attribute__((visibility("default"))) int pylc_debug_read(void *dst, size_t bytes) { return (int)cudaMemcpyFromSymbol(dst, pylc::g_cf_dbg, bytes); }
#endif
