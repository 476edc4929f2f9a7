// Evaluation kernels: nearest-neighbour resample of the fitted label map to full resolution,
// ground-truth palette encode, coverage injection and the confusion matrix in one pass
// (utils/tools.py:316-317, utils/evaluate.py:87-119,150-176, utils/metrics.py:45-87).
//
// Confusion counting (C <= 15, the shipped schemas have 9 and 11 classes): every (truth, prediction)
// pair is a byte code t*C + p, and every LANE owns a private column of 16-bit counters in shared
// memory -- tab[warp][code][lane] -- so a pixel costs one conflict-free load / add / store and no
// atomic at all (shared-memory atomics retire at ~2 cycles per lane on this part, which made the
// earlier run-length + ATOMS scheme LSU-bound).  The cost is independent of how coherent the label
// maps are.  Columns are summed and flushed to the i64 global matrix once per CTA.
// C > 15 keeps per-warp u32 tables with run-length compressed shared atomics.
#include "common.cuh"

namespace pylc {

struct PairRun {
    uint32_t cur, cnt;
    __device__ __forceinline__ void reset() { cur = 0; cnt = 0; }
    __device__ __forceinline__ void push(uint32_t idx, unsigned *tab) {
        if (idx != cur) {
            if (cnt) atomicAdd(&tab[cur], cnt);
            cur = idx;
            cnt = 0;
        }
        ++cnt;
    }
    __device__ __forceinline__ void flush(unsigned *tab) {
        if (cnt) atomicAdd(&tab[cur], cnt);
        cnt = 0;
    }
};

__device__ __forceinline__ void flush_tables(unsigned *s_conf, int CC, long long *conf) {
    __syncthreads();
    for (int i = threadIdx.x; i < CC; i += kThreads) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += s_conf[w * CC + i];
        if (t) atomicAdd((unsigned long long *)&conf[i], t);
    }
}

constexpr int kRsThreads = 64;  // resample CTA: 2 warps, one 1024-pixel column block of the full-res image
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kStageBytes = 544;  // staged label window per warp: 33 chunks of 16 B + pad
constexpr int kRsMaxRows = 2048;  // rows per CTA: keeps every 16-bit lane counter below 2048 * 16

struct ResampleArgs {
    const uint8_t *labels;
    const int32_t *x_ofs, *y_ofs;
    const uint8_t *gt_rgb;
    size_t gt_pitch;
    int h, w, h_full, w_full, C, n_inject;
    int col_blocks, rows_per_slot;
    long long *conf;
    uint8_t *pred_full, *pred_rgb, *gt_full;
    bool gt_aligned, out_aligned, rgb_aligned, labels_vec, labels_stage;
};

// Row-invariant nearest-neighbour column map of a thread's 16 destination pixels.  Destination
// word k (4 pixels) reads source bytes x_ofs[X+4k .. X+4k+3]; when they all lie inside the two
// aligned source words starting at lo[k] one PRMT with a precomputed selector gathers them -- true
// for every up-sampling map and for down-sampling by less than 2.
struct ColMap {
    uint32_t lo[4], hi[4];  // byte offsets (multiples of 4) of the two source words inside a label row
    uint32_t sel[4];        // PRMT selector per destination word
    bool fast;
};

// one lane-private 16-bit counter += 1 (32-bit shared-window address: one IMAD per pixel)
template <bool CTR32>
__device__ __forceinline__ void bump_counter(uint32_t saddr) {
    if (CTR32)
        asm volatile("{ .reg .u32 t; ld.shared.u32 t, [%0]; add.u32 t, t, 1; st.shared.u32 [%0], t; }" ::"r"(saddr) : "memory");
    else
        asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%0]; add.u16 t, t, 1; st.shared.u16 [%0], t; }" ::"r"(saddr) : "memory");
}
// Work decomposition: a CTA owns one column block (kRsThreads x 16 destination pixels) and a
// contiguous range of rows.  Per row a thread streams three 16-byte ground-truth loads (issued one
// row ahead) and eight 4-byte loads of the L1/L2-resident fitted label row (skipped when the row
// maps to the same source row as the previous one); the column map costs nothing per row.
//   MODE 0: C <= 15, pair codes via a pre-multiplied palette table (entry byte = class * C)
//   MODE 1: C <= 15, encoded ground truth also written out (entry byte = class)
//   MODE 2: C  > 15, u32 codes, per-warp atomic tables
// CTR32: 32-bit lane counters (one bank word per lane: conflict-free; C <= 11 keeps the table at
// <= 31 KB per CTA); otherwise 16-bit counters, two lanes per bank word (2-way conflicts, half the
// shared memory).
template <int MODE, bool CTR32>
__global__ void __launch_bounds__(kRsThreads, 12)
    resample_confusion_kernel(ResampleArgs a, const __grid_constant__ PaletteHash ph, const __grid_constant__ ColourLut lut) {
    constexpr bool PACKED = MODE != 2;
    constexpr bool PREMUL = MODE == 0;
    extern __shared__ __align__(16) unsigned s_dyn[];
    __shared__ uint32_t s_tab[256];
    __shared__ uint32_t s_lut[PYLC_MAX_CLASSES];
    const int CC = a.C * a.C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t scale = PREMUL ? (uint32_t)a.C : 1u;
#pragma unroll
    for (int i = threadIdx.x; i < 256; i += kRsThreads) {     // four independent parameter-bank reads per thread
        const uint32_t e = ph.tab[i];
        s_tab[i] = (e & 0x00FFFFFFu) | (((e >> 24) * scale) << 24);
    }
    const uint32_t miss_e = scale << 24;   // unmatched colours are class 1 (utils/tools.py:437)
    if (threadIdx.x < PYLC_MAX_CLASSES) s_lut[threadIdx.x] = lut.rgb[threadIdx.x];
    const int tab_words = PACKED ? kRsWarps * CC * (CTR32 ? 32 : 16) : kRsWarps * CC;
#pragma unroll 4
    for (int i = threadIdx.x; i < tab_words / 4; i += kRsThreads) reinterpret_cast<uint4 *>(s_dyn)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = (tab_words & ~3) + threadIdx.x; i < tab_words; i += kRsThreads) s_dyn[i] = 0;
    __syncthreads();
    const uint32_t mul = ph.mul;
    const bool do_conf = a.conf != nullptr && a.gt_rgb != nullptr;
    constexpr uint32_t kCtrBytes = CTR32 ? 4u : 2u, kCodeStride = 32u * kCtrBytes;
    const uint32_t my_col = (uint32_t)__cvta_generic_to_shared(s_dyn) + (uint32_t)(warp * CC * 32 + lane) * kCtrBytes;   // PACKED
    unsigned *my_tab = s_dyn + warp * CC;                                                                // !PACKED
    PairRun run;
    run.reset();

    const int cb = blockIdx.x % a.col_blocks;
    const int slot = blockIdx.x / a.col_blocks;
    const int X = (cb * kRsThreads + threadIdx.x) * 16;
    const bool active = X < a.w_full;
    const int valid = active ? min(16, a.w_full - X) : 0;
    const bool full = valid == 16;

    ColMap cm;
    cm.fast = a.labels_vec && active;
#pragma unroll
    for (int k = 0; k < 4; ++k) cm.lo[k] = cm.hi[k] = cm.sel[k] = 0;
    int sx_first = 0, sx_last = 0;
    if (active) {
        int sxs[16];
        if (full && ((uintptr_t)a.x_ofs & 15) == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int4 v = __ldg(reinterpret_cast<const int4 *>(a.x_ofs + X) + k);
                sxs[4 * k] = v.x, sxs[4 * k + 1] = v.y, sxs[4 * k + 2] = v.z, sxs[4 * k + 3] = v.w;
            }
        } else {
            int prev = __ldg(a.x_ofs + X);
#pragma unroll
            for (int j = 0; j < 16; ++j) prev = sxs[j] = j < valid ? __ldg(a.x_ofs + X + j) : prev;
        }
        sx_first = sxs[0];
        sx_last = sxs[15];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int word0 = sxs[4 * k] & ~3;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int d = sxs[4 * k + jj] - word0;
                cm.fast = cm.fast && d >= 0 && d <= 7;
                cm.sel[k] |= (uint32_t)(d & 7) << (4 * jj);
            }
            cm.lo[k] = (uint32_t)word0;
            cm.hi[k] = (uint32_t)min(word0 + 4, a.w - 4);   // never selected when clamped (source x < w)
        }
    }
    // Warp-staged label rows: when the 512 destination pixels of a warp read a source window of at
    // most 33 16-byte chunks (any up-sampling map), each lane copies ONE chunk of the label row into
    // the warp's shared-memory window and every lane gathers its 16 labels from there with
    // thread-constant offsets -- instead of eight dependent global loads per row.
    const int wbase = __shfl_sync(0xFFFFFFFFu, sx_first, 0) & ~15;
    const bool lane_ok = !active || (cm.fast && sx_first >= wbase && sx_last - wbase < kStageBytes - 16);
    const bool staged = a.labels_stage && __all_sync(0xFFFFFFFFu, lane_ok) && __any_sync(0xFFFFFFFFu, active);
    if (staged && active) {     // inactive lanes keep offset 0: they still execute the (discarded) gather
#pragma unroll
        for (int k = 0; k < 4; ++k) cm.lo[k] -= (uint32_t)wbase;
    }
    const size_t label_bytes = (size_t)a.h * a.w;
    const int y_lo = slot * a.rows_per_slot, y_hi = min(a.h_full, y_lo + a.rows_per_slot);
    const bool gt_vec = a.gt_rgb != nullptr && a.gt_aligned && active && (size_t)X * 3 + 48 <= a.gt_pitch;

    __shared__ __align__(16) uint8_t s_stage[kRsWarps][2][kStageBytes];
    const uint32_t stage0 = (uint32_t)__cvta_generic_to_shared(&s_stage[warp][0][0]);
    int parity = 0;

    uint4 q0 = make_uint4(0u, 0u, 0u, 0u), q1 = q0, q2 = q0;
    if (gt_vec && y_lo < y_hi) {
        const uint8_t *p = a.gt_rgb + (size_t)y_lo * a.gt_pitch + (size_t)X * 3;
        q0 = ld_stream16(p);
        q1 = ld_stream16(p + 16);
        q2 = ld_stream16(p + 32);
    }
    uint32_t pw[4] = {0u, 0u, 0u, 0u};    // predicted labels, 4 per word (kept while the source row repeats)
    int prev_sy = -1;
    // label-row chunk(s) of the next distinct source row (staged mode), fetched one row ahead
    uint4 c0 = make_uint4(0u, 0u, 0u, 0u), c1 = c0;
    auto fetch_chunks = [&](int sy) {
        const size_t off = (size_t)sy * a.w + (size_t)wbase + 16u * lane;
        c0 = off + 16 <= label_bytes ? __ldg(reinterpret_cast<const uint4 *>(a.labels + off)) : make_uint4(0u, 0u, 0u, 0u);
        if (lane == 0) {
            const size_t off1 = off + 512;
            c1 = off1 + 16 <= label_bytes ? __ldg(reinterpret_cast<const uint4 *>(a.labels + off1)) : make_uint4(0u, 0u, 0u, 0u);
        }
    };
    // source row of this row, the next one and the one after: the index map is read two rows ahead so
    // that no row waits for a dependent global load
    int sy = y_lo < y_hi ? __ldg(a.y_ofs + y_lo) : 0;
    int sy_next = y_lo + 1 < y_hi ? __ldg(a.y_ofs + y_lo + 1) : sy;
    if (staged && y_lo < y_hi) fetch_chunks(sy);

    for (int Y = y_lo; Y < y_hi; ++Y) {
        uint32_t gw[4] = {0u, 0u, 0u, 0u};    // encoded ground truth (x C when PREMUL), 4 per word
        uint4 n0 = q0, n1 = q1, n2 = q2;
        if (gt_vec && Y + 1 < y_hi) {          // next row's ground truth: the HBM stream, one row ahead
            const uint8_t *p = a.gt_rgb + (size_t)(Y + 1) * a.gt_pitch + (size_t)X * 3;
            n0 = ld_stream16(p);
            n1 = ld_stream16(p + 16);
            n2 = ld_stream16(p + 32);
        }
        const int sy_next2 = Y + 2 < y_hi ? __ldg(a.y_ofs + Y + 2) : sy_next;
        const bool reload = sy != prev_sy;
        prev_sy = sy;
        if (staged) {
            if (reload) {
                const uint32_t buf = stage0 + (uint32_t)parity * kStageBytes;
                parity ^= 1;
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(buf + 16u * lane), "r"(c0.x), "r"(c0.y), "r"(c0.z), "r"(c0.w) : "memory");
                if (lane == 0)
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(buf + 512u), "r"(c1.x), "r"(c1.y), "r"(c1.z), "r"(c1.w) : "memory");
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t lo, hi;
                    asm volatile("ld.shared.u32 %0, [%2]; ld.shared.u32 %1, [%2+4];" : "=r"(lo), "=r"(hi) : "r"(buf + cm.lo[k]) : "memory");
                    pw[k] = __byte_perm(lo, hi, cm.sel[k]);
                }
            }
            if (sy_next != sy) fetch_chunks(sy_next);      // after this row's chunks went to shared memory
        }
        if (active) {
            // prediction: nearest-neighbour gather from the fitted label map
            if (reload && !staged) {
                const uint8_t *lrow = a.labels + (size_t)sy * a.w;
                if (cm.fast) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        pw[k] = __byte_perm(__ldg(reinterpret_cast<const uint32_t *>(lrow + cm.lo[k])),
                                            __ldg(reinterpret_cast<const uint32_t *>(lrow + cm.hi[k])), cm.sel[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) pw[k] = 0;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < valid) pw[j >> 2] |= (uint32_t)__ldg(lrow + __ldg(a.x_ofs + X + j)) << (8 * (j & 3));
                }
            }
            if (a.gt_rgb) {
                uint32_t key[16];
                if (gt_vec) {
                    const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        key[4 * k] = w[3 * k];
                        key[4 * k + 1] = __funnelshift_r(w[3 * k], w[3 * k + 1], 24);
                        key[4 * k + 2] = __funnelshift_r(w[3 * k + 1], w[3 * k + 2], 16);
                        key[4 * k + 3] = w[3 * k + 2] >> 8;
                    }
                } else {
                    const uint8_t *p = a.gt_rgb + (size_t)Y * a.gt_pitch + (size_t)X * 3;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        key[j] = 0;
                        if (j < valid) key[j] = __ldg(p + 3 * j) | (__ldg(p + 3 * j + 1) << 8) | (__ldg(p + 3 * j + 2) << 16);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    gw[k] = pack_top_bytes(lookup_entry(key[4 * k], s_tab, mul, miss_e), lookup_entry(key[4 * k + 1], s_tab, mul, miss_e),
                                           lookup_entry(key[4 * k + 2], s_tab, mul, miss_e), lookup_entry(key[4 * k + 3], s_tab, mul, miss_e));
            }
        }

        if (do_conf && active) {
            if (PACKED) {
                // pair codes t*C + p, four per word (no carries: every code is < 256)
                uint32_t idxw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (PREMUL) {
                        idxw[k] = gw[k] + pw[k];
                    } else {
                        const uint32_t even = (gw[k] & 0x00FF00FFu) * (uint32_t)a.C + (pw[k] & 0x00FF00FFu);
                        const uint32_t odd = ((gw[k] >> 8) & 0x00FF00FFu) * (uint32_t)a.C + ((pw[k] >> 8) & 0x00FF00FFu);
                        idxw[k] = even | (odd << 8);
                    }
                }
                // coverage injection (utils/evaluate.py:172-174): the first n_inject flat pixels count as
                // (i, i).  Only the counts see it -- the label / RGB outputs stay the plain resample.
                if (Y == 0 && X < a.n_inject) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (X + j < a.n_inject) {
                            const uint32_t code = (uint32_t)(X + j) * (uint32_t)(a.C + 1);
                            idxw[j >> 2] = (idxw[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | (code << (8 * (j & 3)));
                        }
                }
                if (full) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) bump_counter<CTR32>(my_col + __byte_perm(idxw[j >> 2], 0, 0x4440u | (j & 3)) * kCodeStride);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < valid) bump_counter<CTR32>(my_col + __byte_perm(idxw[j >> 2], 0, 0x4440u | (j & 3)) * kCodeStride);
                }
            } else {
                const long long flat0 = (long long)Y * a.w_full + X;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < valid) {
                        uint32_t t = (gw[j >> 2] >> (8 * (j & 3))) & 0xFFu, p = (pw[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                        if (flat0 + j < a.n_inject) t = p = (uint32_t)(flat0 + j);
                        run.push(t * a.C + p, my_tab);
                    }
            }
        }

        q0 = n0;
        q1 = n1;
        q2 = n2;
        sy = sy_next;
        sy_next = sy_next2;

        if (!active) continue;
        const size_t o = (size_t)Y * a.w_full + X;
        if (a.pred_full) {
            if (full && a.out_aligned) {
                st_stream16(a.pred_full + o, make_uint4(pw[0], pw[1], pw[2], pw[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < valid) a.pred_full[o + j] = (uint8_t)(pw[j >> 2] >> (8 * (j & 3)));
            }
        }
        if (!PREMUL && a.gt_full && a.gt_rgb) {
            if (full && a.out_aligned) {
                st_stream16(a.gt_full + o, make_uint4(gw[0], gw[1], gw[2], gw[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < valid) a.gt_full[o + j] = (uint8_t)(gw[j >> 2] >> (8 * (j & 3)));
            }
        }
        if (a.pred_rgb) {
            uint8_t *d = a.pred_rgb + o * 3;
            if (full && a.rgb_aligned) {
                uint32_t ow[12];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t p0 = s_lut[pw[k] & 31], p1 = s_lut[(pw[k] >> 8) & 31], p2 = s_lut[(pw[k] >> 16) & 31],
                                   p3 = s_lut[(pw[k] >> 24) & 31];
                    ow[3 * k] = p0 | (p1 << 24);
                    ow[3 * k + 1] = (p1 >> 8) | (p2 << 16);
                    ow[3 * k + 2] = (p2 >> 16) | (p3 << 8);
                }
                st_stream16(d, make_uint4(ow[0], ow[1], ow[2], ow[3]));
                st_stream16(d + 16, make_uint4(ow[4], ow[5], ow[6], ow[7]));
                st_stream16(d + 32, make_uint4(ow[8], ow[9], ow[10], ow[11]));
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < valid) {
                        const uint32_t p = s_lut[(pw[j >> 2] >> (8 * (j & 3))) & 31];
                        d[3 * j] = (uint8_t)p;
                        d[3 * j + 1] = (uint8_t)(p >> 8);
                        d[3 * j + 2] = (uint8_t)(p >> 16);
                    }
            }
        }
    }
    if (do_conf) {
        if (!PACKED) run.flush(my_tab);
        __syncthreads();
        for (int i = threadIdx.x; i < CC; i += kRsThreads) {
            unsigned long long t = 0;
            if (PACKED) {
#pragma unroll
                for (int w = 0; w < kRsWarps; ++w) {
                    if (CTR32) {
                        const unsigned *col = s_dyn + (size_t)(w * CC + i) * 32;   // 32 lane counters of code i
#pragma unroll
                        for (int k = 0; k < 32; ++k) t += col[(k + i) & 31];       // rotated: spreads the banks
                    } else {
                        const unsigned *col = s_dyn + (size_t)(w * CC + i) * 16;   // 32 u16 lane counters of code i
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            const unsigned v = col[(k + i) & 15];
                            t += (v & 0xFFFFu) + (v >> 16);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int w = 0; w < kRsWarps; ++w) t += s_dyn[w * CC + i];
            }
            if (t) atomicAdd((unsigned long long *)&a.conf[i], t);
        }
    }
}

__global__ void __launch_bounds__(kThreads)
    confusion_u8_kernel(const uint8_t *__restrict__ y_true, const uint8_t *__restrict__ y_pred, long long n, int C,
                        int n_inject, long long *__restrict__ conf, bool aligned) {
    extern __shared__ unsigned s_dyn[];
    const int CC = C * C;
    for (int i = threadIdx.x; i < CC * kWarps; i += kThreads) s_dyn[i] = 0;
    __syncthreads();
    unsigned *my_tab = s_dyn + (threadIdx.x >> 5) * CC;
    PairRun run;
    run.reset();
    const long long units = (n + 15) / 16;
    for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < units; u += (long long)gridDim.x * kThreads) {
        const long long x = u * 16;
        const int valid = (int)min(16ll, n - x);
        uint32_t t[16], p[16];
        if (aligned && valid == 16) {
            const uint4 a = ld_stream16(y_true + x), b = ld_stream16(y_pred + x);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                t[j] = (aw[j >> 2] >> ((j & 3) * 8)) & 0xFF;
                p[j] = (bw[j >> 2] >> ((j & 3) * 8)) & 0xFF;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                t[j] = j < valid ? y_true[x + j] : 0;
                p[j] = j < valid ? y_pred[x + j] : 0;
            }
        }
        if (x < n_inject) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (x + j < n_inject) t[j] = p[j] = (uint32_t)(x + j);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (j < valid && t[j] < (uint32_t)C && p[j] < (uint32_t)C) run.push(t[j] * C + p[j], my_tab);
    }
    run.flush(my_tab);
    flush_tables(s_dyn, CC, conf);
}

}  // namespace pylc

using namespace pylc;

static unsigned persistent_grid(long long units) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (units + kThreads - 1) / kThreads;
    const long long cap = (long long)sms * 8;
    return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

extern "C" int pylc_resample_encode_confusion(const uint8_t *labels, int h, int w, const int32_t *x_ofs,
                                              const int32_t *y_ofs, int h_full, int w_full, const uint8_t *gt_rgb,
                                              size_t gt_pitch, const uint8_t *palette, const uint8_t *lut_rgb, int C,
                                              int n_inject, int64_t *conf, uint8_t *pred_full, uint8_t *pred_rgb,
                                              uint8_t *gt_full, pylc_stream_t stream) {
    if (!labels || !x_ofs || !y_ofs || h < 1 || w < 1 || h_full < 1 || w_full < 1) return PYLC_ERR_ARG;
    if (C < 1 || C > PYLC_MAX_CLASSES) return PYLC_ERR_CLASSES;
    if (gt_rgb && (!palette || gt_pitch < (size_t)w_full * 3)) return PYLC_ERR_ARG;
    if (pred_rgb && !lut_rgb) return PYLC_ERR_ARG;
    if (conf && !gt_rgb) return PYLC_ERR_ARG;
    if (n_inject < 0 || n_inject > C) return PYLC_ERR_ARG;
    PaletteHash ph;
    if (gt_rgb) {
        int rc = build_palette_hash(palette, C, &ph);
        if (rc) return rc;
    } else {
        for (int i = 0; i < 256; ++i) ph.tab[i] = 1u << 24;
        ph.mul = 1;
    }
    // counts only, aligned rows, C <= 11: the TMA form (confusion_tma.cu); PYLC_NO_TMA=1 keeps the per-thread kernel
    if (conf && gt_rgb && !pred_full && !pred_rgb && !gt_full && !tma_disabled()) {
        const int rc = launch_resample_confusion_tma(labels, h, w, x_ofs, y_ofs, h_full, w_full, gt_rgb, gt_pitch, ph, C, n_inject,
                                                     reinterpret_cast<long long *>(conf), (cudaStream_t)stream);
        if (rc >= 0) return rc;
    }
    ColourLut lut;
    if (lut_rgb) build_colour_lut(lut_rgb, C, &lut);
    else for (int i = 0; i < PYLC_MAX_CLASSES; ++i) lut.rgb[i] = 0;
    ResampleArgs a;
    a.labels = labels; a.x_ofs = x_ofs; a.y_ofs = y_ofs; a.gt_rgb = gt_rgb; a.gt_pitch = gt_pitch;
    a.h = h; a.w = w; a.h_full = h_full; a.w_full = w_full; a.C = C; a.n_inject = n_inject;
    const int groups_per_row = (w_full + 15) / 16;
    a.col_blocks = (groups_per_row + kRsThreads - 1) / kRsThreads;
    a.conf = reinterpret_cast<long long *>(conf);
    a.pred_full = pred_full; a.pred_rgb = pred_rgb; a.gt_full = gt_full;
    a.gt_aligned = gt_rgb && ((uintptr_t)gt_rgb % 16 == 0) && (gt_pitch % 16 == 0);
    a.out_aligned = (w_full % 16 == 0) && ((uintptr_t)pred_full % 16 == 0) && ((uintptr_t)gt_full % 16 == 0);
    a.rgb_aligned = (w_full % 16 == 0) && ((uintptr_t)pred_rgb % 16 == 0);
    a.labels_vec = ((uintptr_t)labels % 4 == 0) && (w % 4 == 0) && w >= 4;
    a.labels_stage = a.labels_vec && ((uintptr_t)labels % 16 == 0) && (w % 16 == 0);
    const int mode = C > 15 ? 2 : (gt_full ? 1 : 0);
    const bool ctr32 = false;   // measured: 16-bit counters (half the table set-up and flush) win at every C
    const size_t smem = mode == 2 ? (size_t)C * C * kRsWarps * sizeof(unsigned)
                                  : (size_t)C * C * kRsWarps * 32 * (ctr32 ? 4 : 2);
    // persistent-style grid: one wave of as many CTAs as fit (registers / shared memory decide),
    // each with a contiguous row range of its column block
    using Kern = void (*)(ResampleArgs, const PaletteHash, const ColourLut);
    const Kern kern = mode == 2 ? (Kern)resample_confusion_kernel<2, false>
                    : mode == 1 ? (ctr32 ? (Kern)resample_confusion_kernel<1, true> : (Kern)resample_confusion_kernel<1, false>)
                                : (ctr32 ? (Kern)resample_confusion_kernel<0, true> : (Kern)resample_confusion_kernel<0, false>);
    int dev = 0, sms = 148, per_sm = 8;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRsThreads, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 4;
    }
    if (per_sm > 16) per_sm = 16;
    int slots = sms * per_sm / a.col_blocks;
    // At least two rows per CTA amortise the column map.  More rows per CTA is slower, not faster (measured
    // on a 3000x2000 image: 4 rows 27 us, 8 rows 31 us, 16 rows 49 us, 32 rows 87 us): a warp's row is a chain
    // of dependent shared-memory updates, so throughput comes from the number of resident warps.
    if (slots > (h_full + 1) / 2) slots = (h_full + 1) / 2;
    if (slots < 1) slots = 1;
    a.rows_per_slot = (h_full + slots - 1) / slots;
    if (a.rows_per_slot > kRsMaxRows) a.rows_per_slot = kRsMaxRows;   // 16-bit lane counters
    const int row_slots = (h_full + a.rows_per_slot - 1) / a.rows_per_slot;
    const unsigned grid = (unsigned)(a.col_blocks * row_slots);
    kern<<<grid, kRsThreads, smem, (cudaStream_t)stream>>>(a, ph, lut);
    return finish_launch();
}

extern "C" int pylc_confusion_u8(const uint8_t *y_true, const uint8_t *y_pred, int64_t n, int C, int n_inject,
                                 int64_t *conf, pylc_stream_t stream) {
    if (!y_true || !y_pred || !conf || n < 0) return PYLC_ERR_ARG;
    if (C < 1 || C > PYLC_MAX_CLASSES) return PYLC_ERR_CLASSES;
    if (n_inject < 0 || n_inject > C) return PYLC_ERR_ARG;
    if (n == 0) return PYLC_OK;
    const bool aligned = ((uintptr_t)y_true % 16 == 0) && ((uintptr_t)y_pred % 16 == 0);
    const size_t smem = (size_t)C * C * kWarps * sizeof(unsigned);
    confusion_u8_kernel<<<persistent_grid((n + 15) / 16), kThreads, smem, (cudaStream_t)stream>>>(
        y_true, y_pred, n, C, n_inject, reinterpret_cast<long long *>(conf), aligned);
    return finish_launch();
}
