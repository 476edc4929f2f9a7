// Stand-alone palette encode, tile profiling (histogram + moments) and colourise kernels, plus the
// host-side palette hash builder shared by every kernel that encodes RGB masks.
#include "common.cuh"

namespace pylc {

std::atomic<int64_t> g_launches{0};

int build_palette_hash(const uint8_t *palette, int C, PaletteHash *out) {
    uint32_t keys[PYLC_MAX_CLASSES];
    uint32_t cls[PYLC_MAX_CLASSES];
    int n = 0;
    for (int i = 0; i < C; ++i) {
        const uint32_t k = palette[3 * i] | (palette[3 * i + 1] << 8) | (palette[3 * i + 2] << 16);
        int j = 0;
        for (; j < n; ++j)
            if (keys[j] == k) break;
        keys[j] = k;
        cls[j] = (uint32_t)i;  // a later duplicate overwrites the earlier index
        if (j == n) ++n;
    }
    uint32_t mul = 0x9E3779B1u;
    for (int attempt = 0; attempt < 1 << 16; ++attempt, mul = mul * 0x01000193u + 0x632BE5ABu) {
        mul |= 1u;
        bool used[256] = {false};
        bool ok = true;
        for (int j = 0; j < n && ok; ++j) {
            const uint32_t slot = ((keys[j] * mul) >> 16) & 0xFFu;
            ok = !used[slot];
            used[slot] = true;
        }
        if (!ok) continue;
        for (int s = 0; s < 256; ++s) out->tab[s] = (1u << 24) | 0x00FFFFFFu;  // empty: class 1
        for (int j = 0; j < n; ++j) out->tab[((keys[j] * mul) >> 16) & 0xFFu] = keys[j] | (cls[j] << 24);
        out->mul = mul;
        return PYLC_OK;
    }
    return PYLC_ERR_PALETTE;
}

void build_colour_lut(const uint8_t *lut_rgb, int C, ColourLut *out) {
    for (int i = 0; i < PYLC_MAX_CLASSES; ++i)
        out->rgb[i] = i < C ? (lut_rgb[3 * i] | (lut_rgb[3 * i + 1] << 8) | (lut_rgb[3 * i + 2] << 16)) : 0u;
}

constexpr int kUnitsPerThread = 8;
constexpr int kPxPerCta = kThreads * kUnitsPerThread * 16;

// ------------------------------------------------------------------------------------------------
// class_encode, interleaved [n_img, rows, cols, 3] (layout 0) or planar [n_img, 3, rows*cols] (layout 1)
// ------------------------------------------------------------------------------------------------
struct EncodeArgs {
    const uint8_t *rgb;
    uint8_t *out;
    long long *hist;
    size_t pitch;          // layout 0: source row pitch
    long long rows, cols;  // layout 1: rows = 1, cols = HW
    long long groups_per_row;
    long long total_units;  // n_img * rows * groups_per_row
    int C;
    bool src_aligned, out_aligned;
};

template <int LAYOUT, int NG, bool HIST>
__global__ void __launch_bounds__(kThreads) class_encode_kernel(EncodeArgs a, const __grid_constant__ PaletteHash ph) {
    using CC = typename CounterSel<NG>::type;
    __shared__ uint32_t s_tab[256];
    __shared__ unsigned s_hist[PYLC_MAX_CLASSES];
    s_tab[threadIdx.x] = ph.tab[threadIdx.x];
    if (threadIdx.x < PYLC_MAX_CLASSES) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t mul = ph.mul;
    CC cc;
    cc.reset();
    int since_widen = 0;
    // a thread's units are kThreads apart: (row, group) advance by constants with a carry -- one 64-bit
    // division per thread instead of one per unit
    const long long first = (long long)blockIdx.x * kThreads * kUnitsPerThread + threadIdx.x;
    long long rowi = first / a.groups_per_row;        // global row index (img * rows + row)
    long long grp = first - rowi * a.groups_per_row;
    const long long d_row = kThreads / a.groups_per_row, d_grp = kThreads - d_row * a.groups_per_row;
#pragma unroll 1
    for (int it = 0; it < kUnitsPerThread; ++it, rowi += d_row, grp += d_grp) {
        if (grp >= a.groups_per_row) {
            grp -= a.groups_per_row;
            ++rowi;
        }
        if (first + (long long)it * kThreads >= a.total_units) break;
        const long long x = grp * 16;
        const int valid = (int)min(16ll, a.cols - x);
        uint32_t key[16];
        if (LAYOUT == 0) {
            const uint8_t *p = a.rgb + (size_t)rowi * a.pitch + (size_t)x * 3;
            if (a.src_aligned && (size_t)(x * 3 + 48) <= a.pitch) {
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
        } else {
            const long long img = rowi;  // rows == 1
            const uint8_t *p = a.rgb + (size_t)img * 3 * a.cols + x;
            if (a.src_aligned && valid == 16) {
                const uint4 r = __ldg((const uint4 *)p), g = __ldg((const uint4 *)(p + a.cols)),
                            b = __ldg((const uint4 *)(p + 2 * a.cols));
                const uint32_t rw[4] = {r.x, r.y, r.z, r.w}, gw[4] = {g.x, g.y, g.z, g.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t t0 = __byte_perm(rw[k], gw[k], 0x5140), t1 = __byte_perm(rw[k], gw[k], 0x7362);
                    key[4 * k] = __byte_perm(t0, bw[k], 0x0410);
                    key[4 * k + 1] = __byte_perm(t0, bw[k], 0x0532);
                    key[4 * k + 2] = __byte_perm(t1, bw[k], 0x0610);
                    key[4 * k + 3] = __byte_perm(t1, bw[k], 0x0732);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    key[j] = 0;
                    if (j < valid) key[j] = __ldg(p + j) | (__ldg(p + a.cols + j) << 8) | (__ldg(p + 2 * a.cols + j) << 16);
                }
            }
        }
        uint32_t ow[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t c = encode_key(key[j], s_tab, mul);
            ow[j >> 2] |= c << ((j & 3) * 8);
        }
        if (HIST) {
            if (valid == 16) cc.add16(ow[0], ow[1], ow[2], ow[3]);
            else cc.add16(void_tail<CC>(ow[0], 0, valid), void_tail<CC>(ow[1], 1, valid), void_tail<CC>(ow[2], 2, valid),
                          void_tail<CC>(ow[3], 3, valid));
            if (++since_widen == CC::kWidenUnits) {
                cc.widen();
                since_widen = 0;
            }
        }
        uint8_t *o = a.out + (size_t)rowi * a.cols + x;
        if (a.out_aligned && valid == 16) {
            st_stream16(o, make_uint4(ow[0], ow[1], ow[2], ow[3]));
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < valid) o[j] = (uint8_t)(ow[j >> 2] >> ((j & 3) * 8));
        }
    }
    if (HIST) {
        cc.widen();
        flush_counter(cc, a.C, s_hist);
        __syncthreads();
        if (threadIdx.x < a.C && s_hist[threadIdx.x])
            atomicAdd((unsigned long long *)&a.hist[threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
    }
}

// ------------------------------------------------------------------------------------------------
// profiling over extracted tiles: per-tile class histogram of u8 masks
// ------------------------------------------------------------------------------------------------
template <int NG>
__global__ void __launch_bounds__(kThreads)
    tile_hist_kernel(const uint8_t *__restrict__ masks, long long tile_px, int chunks, int C,
                     long long *__restrict__ px_dist) {
    __shared__ unsigned s_hist[PYLC_MAX_CLASSES];
    if (threadIdx.x < PYLC_MAX_CLASSES) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const long long tile = blockIdx.x / chunks;
    const int chunk = blockIdx.x % chunks;
    const uint8_t *base = masks + (size_t)tile * tile_px;
    const long long units = tile_px / 16;
    using CC = typename CounterSel<NG>::type;
    static_assert(kUnitsPerThread <= CC::kFlushUnits, "one flush per thread");
    CC cc;
    cc.reset();
    uint4 v[kUnitsPerThread];
#pragma unroll
    for (int it = 0; it < kUnitsPerThread; ++it) {
        const long long u = ((long long)chunk * kUnitsPerThread + it) * kThreads + threadIdx.x;
        v[it] = u < units ? ld_stream16(base + u * 16) : make_uint4(~0u, ~0u, ~0u, ~0u);  // 0xFF is never a class
    }
#pragma unroll
    for (int it = 0; it < kUnitsPerThread; ++it) {
        // masks come from a database: any byte value may occur; ids the counter cannot hold are dropped
        uint32_t w0 = v[it].x, w1 = v[it].y, w2 = v[it].z, w3 = v[it].w;
        if (CC::kNeedsSanitize && ((w0 | w1 | w2 | w3) & 0xF0F0F0F0u)) {   // rare: some id >= 16 in this unit
            w0 = CC::sanitize(w0);
            w1 = CC::sanitize(w1);
            w2 = CC::sanitize(w2);
            w3 = CC::sanitize(w3);
        }
        cc.add16(w0, w1, w2, w3);
        if ((it + 1) % CC::kWidenUnits == 0) cc.widen();
    }
    cc.widen();
    flush_counter(cc, C, s_hist);
    __syncthreads();
    if (threadIdx.x < C && s_hist[threadIdx.x])
        atomicAdd((unsigned long long *)&px_dist[tile * C + threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
}

// per-plane sum(x), sum(x*x) of u8 tiles: planes = n * ch
__global__ void __launch_bounds__(kThreads)
    tile_moments_kernel(const uint8_t *__restrict__ imgs, long long tile_px, int chunks,
                        unsigned long long *__restrict__ stat) {
    __shared__ unsigned long long s_sum[2];
    if (threadIdx.x < 2) s_sum[threadIdx.x] = 0;
    __syncthreads();
    const long long plane = blockIdx.x / chunks;
    const int chunk = blockIdx.x % chunks;
    const uint8_t *base = imgs + (size_t)plane * tile_px;
    const long long units = tile_px / 16;
    uint4 v[kUnitsPerThread];
#pragma unroll
    for (int it = 0; it < kUnitsPerThread; ++it) {
        const long long u = ((long long)chunk * kUnitsPerThread + it) * kThreads + threadIdx.x;
        v[it] = u < units ? ld_stream16(base + u * 16) : make_uint4(0, 0, 0, 0);
    }
    uint32_t s1 = 0, s2 = 0;
#pragma unroll
    for (int it = 0; it < kUnitsPerThread; ++it) {
        const uint32_t w[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s1 = __dp4a(w[j], 0x01010101u, s1);
            s2 = __dp4a(w[j], w[j], s2);
        }
    }
    const uint32_t a = __reduce_add_sync(0xFFFFFFFFu, s1);  // <= 32*128*255 fits
    const uint32_t b_lo = __reduce_add_sync(0xFFFFFFFFu, s2 & 0xFFFFu);
    const uint32_t b_hi = __reduce_add_sync(0xFFFFFFFFu, s2 >> 16);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_sum[0], (unsigned long long)a);
        atomicAdd(&s_sum[1], (unsigned long long)b_lo + ((unsigned long long)b_hi << 16));
    }
    __syncthreads();
    if (threadIdx.x < 2) atomicAdd(&stat[plane * 2 + threadIdx.x], s_sum[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// colourise: u8 labels -> interleaved RGB u8
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    colourise_kernel(const uint8_t *__restrict__ labels, long long n_px, const __grid_constant__ ColourLut lut,
                     uint8_t *__restrict__ rgb, bool aligned) {
    __shared__ uint32_t s_lut[256];
    s_lut[threadIdx.x] = threadIdx.x < PYLC_MAX_CLASSES ? lut.rgb[threadIdx.x] : 0u;
    __syncthreads();
    const long long units = (n_px + 15) / 16;
    for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < units; u += (long long)gridDim.x * kThreads) {
        const long long x = u * 16;
        if (aligned && x + 16 <= n_px) {
            const uint4 v = ld_stream16(labels + x);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            uint32_t o[12];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t p0 = s_lut[w[k] & 0xFF], p1 = s_lut[(w[k] >> 8) & 0xFF],
                               p2 = s_lut[(w[k] >> 16) & 0xFF], p3 = s_lut[w[k] >> 24];
                o[3 * k] = p0 | (p1 << 24);
                o[3 * k + 1] = (p1 >> 8) | (p2 << 16);
                o[3 * k + 2] = (p2 >> 16) | (p3 << 8);
            }
            uint8_t *d = rgb + x * 3;
            st_stream16(d, make_uint4(o[0], o[1], o[2], o[3]));
            st_stream16(d + 16, make_uint4(o[4], o[5], o[6], o[7]));
            st_stream16(d + 32, make_uint4(o[8], o[9], o[10], o[11]));
        } else {
            for (long long i = x; i < min(x + 16, n_px); ++i) {
                const uint32_t p = s_lut[labels[i]];
                rgb[3 * i] = (uint8_t)p;
                rgb[3 * i + 1] = (uint8_t)(p >> 8);
                rgb[3 * i + 2] = (uint8_t)(p >> 16);
            }
        }
    }
}

}  // namespace pylc

using namespace pylc;

extern "C" int pylc_class_encode(const uint8_t *rgb, int n_img, int rows, int cols, size_t pitch, int layout,
                                 const uint8_t *palette, int C, uint8_t *out, int64_t *hist, pylc_stream_t stream) {
    if (!rgb || !out || !palette || n_img < 0 || rows < 0 || cols < 0 || (layout != 0 && layout != 1)) return PYLC_ERR_ARG;
    if (C < 1 || C > PYLC_MAX_CLASSES) return PYLC_ERR_CLASSES;
    if (layout == 0 && pitch < (size_t)cols * 3) return PYLC_ERR_ARG;
    PaletteHash ph;
    int rc = build_palette_hash(palette, C, &ph);
    if (rc) return rc;
    EncodeArgs a;
    a.rgb = rgb;
    a.out = out;
    a.hist = reinterpret_cast<long long *>(hist);
    a.C = C;
    if (layout == 0) {
        a.pitch = pitch;
        a.rows = (long long)n_img * rows;
        a.cols = cols;
        a.src_aligned = ((uintptr_t)rgb % 16 == 0) && (pitch % 16 == 0);
    } else {
        a.pitch = 0;
        a.rows = n_img;
        a.cols = (long long)rows * cols;
        a.src_aligned = ((uintptr_t)rgb % 16 == 0) && (a.cols % 16 == 0);
    }
    a.out_aligned = ((uintptr_t)out % 16 == 0) && (a.cols % 16 == 0);
    a.groups_per_row = (a.cols + 15) / 16;
    a.total_units = a.rows * a.groups_per_row;
    if (a.total_units == 0) return PYLC_OK;
    const long long per_cta = (long long)kThreads * kUnitsPerThread;
    const unsigned grid = (unsigned)((a.total_units + per_cta - 1) / per_cta);
    cudaStream_t st = (cudaStream_t)stream;
    if (layout == 0 && !tma_disabled()) {   // TMA form: 16-byte aligned rows, whole 16-pixel units per output row
        rc = launch_class_encode_tma(rgb, a.rows, a.cols, pitch, ph, C, out, a.hist, st);
        if (rc != -1) return rc;
    }
#define LAUNCH(L, NG, HS) class_encode_kernel<L, NG, HS><<<grid, kThreads, 0, st>>>(a, ph)
#define LAUNCH_L(L)                                                                    \
    switch (hist ? counter_groups(C) : -1) {                                           \
        case -1: LAUNCH(L, 5, false); break;                                           \
        case 5: LAUNCH(L, 5, true); break;                                             \
        case 6: LAUNCH(L, 6, true); break;                                             \
        case 7: LAUNCH(L, 7, true); break;                                             \
        default: LAUNCH(L, 0, true); break;                                            \
    }
    if (layout == 0) { LAUNCH_L(0) } else { LAUNCH_L(1) }
#undef LAUNCH_L
#undef LAUNCH
    return finish_launch();
}

extern "C" int pylc_profile_tiles(const uint8_t *imgs, int ch, const uint8_t *masks, int n, int64_t tile_px, int C,
                                  uint64_t *stat, int64_t *px_dist, pylc_stream_t stream) {
    if (n < 0 || tile_px <= 0) return PYLC_ERR_ARG;
    if (tile_px % 16) return PYLC_ERR_GEOMETRY;
    if ((imgs && (!stat || (ch != 1 && ch != 3))) || (masks && !px_dist)) return PYLC_ERR_ARG;
    if (masks && (C < 1 || C > PYLC_MAX_CLASSES)) return PYLC_ERR_CLASSES;
    if (((uintptr_t)imgs % 16) || ((uintptr_t)masks % 16)) return PYLC_ERR_ALIGN;
    if (n == 0) return PYLC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = (int)((tile_px + kPxPerCta - 1) / kPxPerCta);
    int rc = PYLC_OK;
    if (masks) {
        const unsigned grid = (unsigned)((long long)n * chunks);
        switch (counter_groups(C)) {
            case 5: tile_hist_kernel<5><<<grid, kThreads, 0, st>>>(masks, tile_px, chunks, C, (long long *)px_dist); break;
            case 6: tile_hist_kernel<6><<<grid, kThreads, 0, st>>>(masks, tile_px, chunks, C, (long long *)px_dist); break;
            case 7: tile_hist_kernel<7><<<grid, kThreads, 0, st>>>(masks, tile_px, chunks, C, (long long *)px_dist); break;
            default: tile_hist_kernel<0><<<grid, kThreads, 0, st>>>(masks, tile_px, chunks, C, (long long *)px_dist); break;
        }
        rc = finish_launch();
        if (rc) return rc;
    }
    if (imgs) {
        const unsigned grid = (unsigned)((long long)n * ch * chunks);
        tile_moments_kernel<<<grid, kThreads, 0, st>>>(imgs, tile_px, chunks, (unsigned long long *)stat);
        rc = finish_launch();
    }
    return rc;
}

extern "C" int pylc_colourise_u8(const uint8_t *labels, int64_t n_px, const uint8_t *lut_rgb, int C, uint8_t *rgb,
                                 pylc_stream_t stream) {
    if (!labels || !rgb || !lut_rgb || n_px < 0) return PYLC_ERR_ARG;
    if (C < 1 || C > PYLC_MAX_CLASSES) return PYLC_ERR_CLASSES;
    if (n_px == 0) return PYLC_OK;
    ColourLut lut;
    build_colour_lut(lut_rgb, C, &lut);
    const bool aligned = ((uintptr_t)labels % 16 == 0) && ((uintptr_t)rgb % 16 == 0);
    const long long units = (n_px + 15) / 16;
    const long long want = (units + kThreads - 1) / kThreads;
    const unsigned grid = (unsigned)(want < 148 * 16 ? want : 148 * 16);
    colourise_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(labels, n_px, lut, rgb, aligned);
    return finish_launch();
}

// ------------------------------------------------------------------------------------------------
// Augmentor.optimize grid search (utils/augment.py:92-187): for every (rate_coef, threshold) pair the
// per-tile over-sampling rates and the class histogram of the over-sampled dataset.
// ------------------------------------------------------------------------------------------------
// One CTA per grid point; threads stride over the tiles and keep the C class sums in registers.
// All arithmetic the reference does in float64 / int64 is done in the same types (one IEEE multiply,
// truncation, clip, integer multiply-add), so every output integer equals NumPy's.  The inputs are
// a few hundred KB and stay in L2: the kernel is latency-, not bandwidth-bound.
namespace pylc {

template <int CMAX>
__global__ void __launch_bounds__(kThreads)
    sample_rate_grid_kernel(const double *__restrict__ scores, const long long *__restrict__ px_dist, int N, int C,
                            const double *__restrict__ rate_coefs, const double *__restrict__ thresholds, int n_thr,
                            int rate_lo, int rate_hi, long long *__restrict__ sum_rates, long long *__restrict__ full_px_dist) {
    const int g = blockIdx.x;
    const double rc = __ldg(rate_coefs + g / n_thr), th = __ldg(thresholds + g % n_thr);
    long long acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[c] = 0;
    long long rsum = 0;
    for (int n = threadIdx.x; n < N; n += kThreads) {
        const double s = __ldg(scores + n);
        // np.multiply(scores > threshold, rate_coef * scores).astype(int), then np.clip (augment.py:142-152)
        long long rate = s > th ? (long long)__dmul_rn(rc, s) : 0ll;
        rate = rate < rate_lo ? rate_lo : (rate > rate_hi ? rate_hi : rate);
        rsum += rate;
        const long long mult = rate + 1;          // px_dist + rates * px_dist (augment.py:156-157)
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) acc[c] += mult * __ldg(px_dist + (size_t)n * C + c);
    }
    __shared__ unsigned long long s_acc[PYLC_MAX_CLASSES + 1];
    if (threadIdx.x <= PYLC_MAX_CLASSES) s_acc[threadIdx.x] = 0;
    __syncthreads();
    auto warp_sum = [](long long v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        return v;
    };
    const int lane = threadIdx.x & 31;
    rsum = warp_sum(rsum);
    if (lane == 0) atomicAdd(&s_acc[PYLC_MAX_CLASSES], (unsigned long long)rsum);
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
        if (c < C) {
            const long long v = warp_sum(acc[c]);
            if (lane == 0) atomicAdd(&s_acc[c], (unsigned long long)v);
        }
    __syncthreads();
    if (threadIdx.x < C) full_px_dist[(size_t)g * C + threadIdx.x] = (long long)s_acc[threadIdx.x];
    if (threadIdx.x == 0) sum_rates[g] = (long long)s_acc[PYLC_MAX_CLASSES];
}

}  // namespace pylc

extern "C" int pylc_sample_rate_grid(const double *scores, const int64_t *px_dist, int N, int C, const double *rate_coefs,
                                     int n_coefs, const double *thresholds, int n_thresholds, int rate_lo, int rate_hi,
                                     int64_t *sum_rates, int64_t *full_px_dist, pylc_stream_t stream) {
    if (!scores || !px_dist || !rate_coefs || !thresholds || !sum_rates || !full_px_dist) return PYLC_ERR_ARG;
    if (N < 1 || n_coefs < 1 || n_thresholds < 1 || rate_lo > rate_hi) return PYLC_ERR_ARG;
    if (C < 1 || C > PYLC_MAX_CLASSES) return PYLC_ERR_CLASSES;
    const unsigned grid = (unsigned)(n_coefs * n_thresholds);
    auto *pd = reinterpret_cast<const long long *>(px_dist);
    auto *sr = reinterpret_cast<long long *>(sum_rates);
    auto *fp = reinterpret_cast<long long *>(full_px_dist);
    if (C <= 12)
        pylc::sample_rate_grid_kernel<12><<<grid, pylc::kThreads, 0, (cudaStream_t)stream>>>(
            scores, pd, N, C, rate_coefs, thresholds, n_thresholds, rate_lo, rate_hi, sr, fp);
    else
        pylc::sample_rate_grid_kernel<PYLC_MAX_CLASSES><<<grid, pylc::kThreads, 0, (cudaStream_t)stream>>>(
            scores, pd, N, C, rate_coefs, thresholds, n_thresholds, rate_lo, rate_hi, sr, fp);
    return pylc::finish_launch();
}
