// Evaluation kernels: nearest-neighbour resample of the fitted label map to full resolution,
// ground-truth palette encode, coverage injection and the confusion matrix in one pass
// (utils/tools.py:316-317, utils/evaluate.py:87-119,150-176, utils/metrics.py:45-87).
//
// Confusion counting: each thread walks 16 consecutive pixels and run-length encodes the
// (truth, prediction) pair, so spatially coherent masks issue roughly one shared-memory atomic per
// run instead of one per pixel.  Every warp owns a private C x C table in shared memory; tables are
// summed and flushed to the i64 global matrix once per CTA.
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

struct ResampleArgs {
    const uint8_t *labels;
    const int32_t *x_ofs, *y_ofs;
    const uint8_t *gt_rgb;
    size_t gt_pitch;
    int h, w, h_full, w_full, C, n_inject;
    int col_blocks, row_slots, rows_per_slot, replicas;
    long long *conf;
    uint8_t *pred_full, *pred_rgb, *gt_full;
    bool gt_aligned, out_aligned, rgb_aligned, labels_vec;
};

// 16 bytes starting at byte offset `start` of an 8-byte-aligned buffer of `total` bytes
// (total % 8 == 0), as four little-endian words.  Three aligned 8-byte loads + funnel shifts.
__device__ __forceinline__ void load_window16(const uint8_t *buf, size_t start, size_t total, uint32_t (&x)[4]) {
    const size_t a8 = start & ~(size_t)7;
    const unsigned r = (unsigned)(start & 7);
    uint2 q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        q[k] = (a8 + 8 * k < total) ? __ldg(reinterpret_cast<const uint2 *>(buf + a8 + 8 * k)) : make_uint2(0u, 0u);
    const uint32_t w[6] = {q[0].x, q[0].y, q[1].x, q[1].y, q[2].x, q[2].y};
    const bool up = (r & 4) != 0;
    const unsigned sh = (r & 3) * 8;
    uint32_t v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = up ? w[i + 1] : w[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = __funnelshift_r(v[i], v[i + 1], sh);
}

// word i (0..3, dynamic) of a 4-word window; i + 1 == 4 yields 0 (never selected, see ColMap)
__device__ __forceinline__ uint32_t pick_word(const uint32_t (&x)[4], uint32_t i) {
    const uint32_t lo = (i & 1) ? x[1] : x[0];
    const uint32_t hi = (i & 1) ? x[3] : x[2];
    return (i & 4) ? 0u : ((i & 2) ? hi : lo);
}

// byte d (0..15, dynamic) of 16 bytes held in four words
__device__ __forceinline__ uint32_t pick_byte(const uint32_t (&x)[4], uint32_t d) {
    return __byte_perm(pick_word(x, d >> 2), 0, 0x4440u | (d & 3));
}

// Row-invariant nearest-neighbour column map of a thread's 16 destination pixels: source offsets
// d_j = x_ofs[X+j] - x_ofs[X].  When they rise by 0 or 1 per pixel (any up-sampling map) the four
// source bytes of destination word k lie inside the 8-byte pair starting at word d_4k / 4 of a
// 16-byte window, so one PRMT with a precomputed selector gathers them.
struct ColMap {
    int base;
    uint32_t sel[4];   // PRMT selector per destination word
    uint32_t wsel;     // 4 x 2 bits: first window word per destination word
    bool fast;
};

// Counts (truth, prediction) pairs of one thread-row.  `idxw` holds the 16 pair codes t*C + p as
// bytes.  Runs of equal codes are peeled with FFS from a boundary bitmap and added with one shared
// atomic each; the trip count is the warp maximum of the run counts, so the loop is convergent.
// Every lane adds into one of `replicas` copies of its warp's table (lane % replicas), which cuts
// same-address serialisation on the dominant class pair.
__device__ __forceinline__ void count_runs(const uint32_t (&idxw)[4], uint32_t bounds, int valid, unsigned *tab) {
    const int n = __popc(bounds);
    const int rounds = __reduce_max_sync(0xFFFFFFFFu, n);
    for (int r = 0; r < rounds; ++r) {
        if (bounds) {
            const int start = __ffs(bounds) - 1;
            bounds &= bounds - 1;
            const int end = bounds ? __ffs(bounds) - 1 : valid;
            atomicAdd(&tab[pick_byte(idxw, (uint32_t)start)], (unsigned)(end - start));
        }
    }
}

// Work decomposition: a CTA owns one column block (kRsThreads x 16 destination pixels) and a
// contiguous range of rows.  Per row a thread needs three 16-byte ground-truth loads (the HBM
// stream) and a 24-byte window of the L2-resident fitted label row; the column map costs nothing
// per row.  PACKED: C <= 15, pair codes fit a byte (the shipped schemas have 9 and 11 classes).
template <bool PACKED>
__global__ void __launch_bounds__(kRsThreads, 16)
    resample_confusion_kernel(ResampleArgs a, const __grid_constant__ PaletteHash ph, const __grid_constant__ ColourLut lut) {
    extern __shared__ unsigned s_dyn[];
    __shared__ uint32_t s_tab[256];
    __shared__ uint32_t s_lut[PYLC_MAX_CLASSES];
    const int CC = a.C * a.C;
    const int n_tabs = kRsWarps * a.replicas;
    for (int i = threadIdx.x; i < 256; i += kRsThreads) s_tab[i] = ph.tab[i];
    if (threadIdx.x < PYLC_MAX_CLASSES) s_lut[threadIdx.x] = lut.rgb[threadIdx.x];
    for (int i = threadIdx.x; i < CC * n_tabs; i += kRsThreads) s_dyn[i] = 0;
    __syncthreads();
    const uint32_t mul = ph.mul;
    unsigned *my_tab = s_dyn + ((threadIdx.x >> 5) * a.replicas + (threadIdx.x & 31) % a.replicas) * CC;
    const bool do_conf = a.conf != nullptr && a.gt_rgb != nullptr;
    PairRun run;   // !PACKED path only
    run.reset();

    const int cb = blockIdx.x % a.col_blocks;
    const int slot = blockIdx.x / a.col_blocks;
    const int X = (cb * kRsThreads + threadIdx.x) * 16;
    const bool active = X < a.w_full;
    const int valid = active ? min(16, a.w_full - X) : 0;
    const bool full = valid == 16;

    ColMap cm;
    cm.base = 0;
    cm.wsel = 0;
    cm.fast = a.labels_vec && active;
#pragma unroll
    for (int k = 0; k < 4; ++k) cm.sel[k] = 0;
    if (active) {
        cm.base = __ldg(a.x_ofs + X);
        int prev = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int first = 0;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = 4 * k + jj;
                const int d = j < valid ? __ldg(a.x_ofs + X + j) - cm.base : prev;
                cm.fast = cm.fast && d >= prev && d <= prev + 1;
                prev = d;
                if (jj == 0) first = d >> 2;
                cm.sel[k] |= (uint32_t)((d - 4 * first) & 7) << (4 * jj);
            }
            cm.wsel |= (uint32_t)(first & 3) << (2 * k);
        }
    }
    const size_t label_bytes = (size_t)a.h * a.w;
    const int y_lo = slot * a.rows_per_slot, y_hi = min(a.h_full, y_lo + a.rows_per_slot);

    for (int Y = y_lo; Y < y_hi; ++Y) {
        uint32_t pw[4] = {0u, 0u, 0u, 0u};    // predicted labels, 4 per word
        uint32_t gw[4] = {0u, 0u, 0u, 0u};    // encoded ground truth, 4 per word
        if (active) {
            // ground truth first: these are the HBM-streaming loads
            uint32_t key[16];
            if (a.gt_rgb) {
                const uint8_t *p = a.gt_rgb + (size_t)Y * a.gt_pitch + (size_t)X * 3;
                if (a.gt_aligned && (size_t)X * 3 + 48 <= a.gt_pitch) {
                    const uint4 q0 = __ldg((const uint4 *)p), q1 = __ldg((const uint4 *)(p + 16)),
                                q2 = __ldg((const uint4 *)(p + 32));
                    const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        key[4 * k] = w[3 * k];
                        key[4 * k + 1] = __funnelshift_r(w[3 * k], w[3 * k + 1], 24);
                        key[4 * k + 2] = __funnelshift_r(w[3 * k + 1], w[3 * k + 2], 16);
                        key[4 * k + 3] = w[3 * k + 2] >> 8;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        key[j] = 0;
                        if (j < valid) key[j] = __ldg(p + 3 * j) | (__ldg(p + 3 * j + 1) << 8) | (__ldg(p + 3 * j + 2) << 16);
                    }
                }
            }
            // prediction: nearest-neighbour gather from the fitted label map
            const size_t row_off = (size_t)__ldg(a.y_ofs + Y) * a.w;
            if (cm.fast) {
                uint32_t x[4];
                load_window16(a.labels, row_off + cm.base, label_bytes, x);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t i = (cm.wsel >> (2 * k)) & 3u;
                    pw[k] = __byte_perm(pick_word(x, i), pick_word(x, i + 1), cm.sel[k]);
                }
            } else {
                const uint8_t *lrow = a.labels + row_off;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < valid) pw[j >> 2] |= (uint32_t)__ldg(lrow + __ldg(a.x_ofs + X + j)) << (8 * (j & 3));
            }
            if (a.gt_rgb) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    gw[k] = encode_key(key[4 * k], s_tab, mul) | (encode_key(key[4 * k + 1], s_tab, mul) << 8) |
                            (encode_key(key[4 * k + 2], s_tab, mul) << 16) | (encode_key(key[4 * k + 3], s_tab, mul) << 24);
            }
        }

        if (do_conf) {
            if (PACKED) {
                // pair codes t*C + p for four pixels at a time (two 16-bit lanes per multiply)
                uint32_t idxw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t even = (gw[k] & 0x00FF00FFu) * (uint32_t)a.C + (pw[k] & 0x00FF00FFu);
                    const uint32_t odd = ((gw[k] >> 8) & 0x00FF00FFu) * (uint32_t)a.C + ((pw[k] >> 8) & 0x00FF00FFu);
                    idxw[k] = even | (odd << 8);
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
                // run boundaries: byte j differs from byte j-1
                uint32_t bounds = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t prevw = k == 0 ? (idxw[0] << 8) | (~idxw[0] & 0xFFu)   // byte 0 always opens a run
                                                  : __funnelshift_l(idxw[k - 1], idxw[k], 8);
                    const uint32_t ne = __vcmpne4(idxw[k], prevw);                       // 0xFF per differing byte
                    bounds |= ((ne & 1u) | ((ne >> 7) & 2u) | ((ne >> 14) & 4u) | ((ne >> 21) & 8u)) << (4 * k);
                }
                bounds &= valid >= 16 ? 0xFFFFu : ((1u << valid) - 1u);
                count_runs(idxw, bounds, valid, my_tab);
            } else if (active) {
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
        if (a.gt_full && a.gt_rgb) {
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
            for (int w = 0; w < n_tabs; ++w) t += s_dyn[w * CC + i];
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
    ColourLut lut;
    if (lut_rgb) build_colour_lut(lut_rgb, C, &lut);
    else for (int i = 0; i < PYLC_MAX_CLASSES; ++i) lut.rgb[i] = 0;
    ResampleArgs a;
    a.labels = labels; a.x_ofs = x_ofs; a.y_ofs = y_ofs; a.gt_rgb = gt_rgb; a.gt_pitch = gt_pitch;
    a.h = h; a.w = w; a.h_full = h_full; a.w_full = w_full; a.C = C; a.n_inject = n_inject;
    const int groups_per_row = (w_full + 15) / 16;
    a.col_blocks = (groups_per_row + kRsThreads - 1) / kRsThreads;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int slots = sms * 16 / a.col_blocks;                       // ~16 resident CTAs per SM
    if (slots > (h_full + 1) / 2) slots = (h_full + 1) / 2;    // at least two rows per CTA amortise the column map
    if (slots < 1) slots = 1;
    a.rows_per_slot = (h_full + slots - 1) / slots;
    a.row_slots = (h_full + a.rows_per_slot - 1) / a.rows_per_slot;
    a.replicas = C <= 11 ? 8 : (C <= 15 ? 4 : 1);
    a.conf = reinterpret_cast<long long *>(conf);
    a.pred_full = pred_full; a.pred_rgb = pred_rgb; a.gt_full = gt_full;
    a.gt_aligned = gt_rgb && ((uintptr_t)gt_rgb % 16 == 0) && (gt_pitch % 16 == 0);
    a.out_aligned = (w_full % 16 == 0) && ((uintptr_t)pred_full % 16 == 0) && ((uintptr_t)gt_full % 16 == 0);
    a.rgb_aligned = (w_full % 16 == 0) && ((uintptr_t)pred_rgb % 16 == 0);
    a.labels_vec = ((uintptr_t)labels % 8 == 0) && (((size_t)h * w) % 8 == 0);
    const size_t smem = (size_t)C * C * kRsWarps * a.replicas * sizeof(unsigned);
    const unsigned grid = (unsigned)(a.col_blocks * a.row_slots);
    if (C <= 15) resample_confusion_kernel<true><<<grid, kRsThreads, smem, (cudaStream_t)stream>>>(a, ph, lut);
    else resample_confusion_kernel<false><<<grid, kRsThreads, smem, (cudaStream_t)stream>>>(a, ph, lut);
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
