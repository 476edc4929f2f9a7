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

struct ResampleArgs {
    const uint8_t *labels;
    const int32_t *x_ofs, *y_ofs;
    const uint8_t *gt_rgb;
    size_t gt_pitch;
    int h, w, h_full, w_full, C, n_inject;
    long long groups_per_row, total_units;
    long long *conf;
    uint8_t *pred_full, *pred_rgb, *gt_full;
    bool gt_aligned, out_aligned, rgb_aligned;
};

__global__ void __launch_bounds__(kThreads)
    resample_confusion_kernel(ResampleArgs a, const __grid_constant__ PaletteHash ph, const __grid_constant__ ColourLut lut) {
    extern __shared__ unsigned s_dyn[];
    __shared__ uint32_t s_tab[256];
    __shared__ uint32_t s_lut[PYLC_MAX_CLASSES];
    const int CC = a.C * a.C;
    unsigned *s_conf = s_dyn;
    s_tab[threadIdx.x] = ph.tab[threadIdx.x];
    if (threadIdx.x < PYLC_MAX_CLASSES) s_lut[threadIdx.x] = lut.rgb[threadIdx.x];
    for (int i = threadIdx.x; i < CC * kWarps; i += kThreads) s_conf[i] = 0;
    __syncthreads();
    const uint32_t mul = ph.mul;
    unsigned *my_tab = s_conf + (threadIdx.x >> 5) * CC;
    const bool do_conf = a.conf != nullptr && a.gt_rgb != nullptr;
    PairRun run;
    run.reset();

    for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < a.total_units;
         u += (long long)gridDim.x * kThreads) {
        const long long Y = u / a.groups_per_row;
        const int X = (int)(u - Y * a.groups_per_row) * 16;
        const int valid = min(16, a.w_full - X);
        const bool full = valid == 16;

        // prediction: nearest-neighbour gather from the fitted label map
        const uint8_t *lrow = a.labels + (size_t)__ldg(a.y_ofs + Y) * a.w;
        uint32_t pred[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pred[j] = j < valid ? (uint32_t)__ldg(lrow + __ldg(a.x_ofs + X + j)) : 0u;

        // ground truth: palette encode
        uint32_t gt[16];
        if (a.gt_rgb) {
            const uint8_t *p = a.gt_rgb + (size_t)Y * a.gt_pitch + (size_t)X * 3;
            uint32_t key[16];
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
#pragma unroll
            for (int j = 0; j < 16; ++j) gt[j] = encode_key(key[j], s_tab, mul);
        }

        // coverage injection (utils/evaluate.py:172-174): the first n_inject flat pixels become (i, i)
        const long long flat0 = Y * a.w_full + X;
        if (flat0 < a.n_inject) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (flat0 + j < a.n_inject) {
                    pred[j] = (uint32_t)(flat0 + j);
                    gt[j] = (uint32_t)(flat0 + j);
                }
        }

        if (do_conf) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < valid) run.push(gt[j] * a.C + pred[j], my_tab);
        }

        const size_t o = (size_t)Y * a.w_full + X;
        if (a.pred_full) {
            if (full && a.out_aligned) {
                uint32_t ow[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) ow[k] = pred[4 * k] | (pred[4 * k + 1] << 8) | (pred[4 * k + 2] << 16) | (pred[4 * k + 3] << 24);
                st_stream16(a.pred_full + o, make_uint4(ow[0], ow[1], ow[2], ow[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < valid) a.pred_full[o + j] = (uint8_t)pred[j];
            }
        }
        if (a.gt_full && a.gt_rgb) {
            if (full && a.out_aligned) {
                uint32_t ow[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) ow[k] = gt[4 * k] | (gt[4 * k + 1] << 8) | (gt[4 * k + 2] << 16) | (gt[4 * k + 3] << 24);
                st_stream16(a.gt_full + o, make_uint4(ow[0], ow[1], ow[2], ow[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < valid) a.gt_full[o + j] = (uint8_t)gt[j];
            }
        }
        if (a.pred_rgb) {
            uint8_t *d = a.pred_rgb + o * 3;
            if (full && a.rgb_aligned) {
                uint32_t ow[12];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t p0 = s_lut[pred[4 * k] & 31], p1 = s_lut[pred[4 * k + 1] & 31],
                                   p2 = s_lut[pred[4 * k + 2] & 31], p3 = s_lut[pred[4 * k + 3] & 31];
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
                        const uint32_t p = s_lut[pred[j] & 31];
                        d[3 * j] = (uint8_t)p;
                        d[3 * j + 1] = (uint8_t)(p >> 8);
                        d[3 * j + 2] = (uint8_t)(p >> 16);
                    }
            }
        }
    }
    if (do_conf) {
        run.flush(my_tab);
        flush_tables(s_conf, CC, a.conf);
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
    a.groups_per_row = (w_full + 15) / 16;
    a.total_units = a.groups_per_row * h_full;
    a.conf = reinterpret_cast<long long *>(conf);
    a.pred_full = pred_full; a.pred_rgb = pred_rgb; a.gt_full = gt_full;
    a.gt_aligned = gt_rgb && ((uintptr_t)gt_rgb % 16 == 0) && (gt_pitch % 16 == 0);
    a.out_aligned = (w_full % 16 == 0) && ((uintptr_t)pred_full % 16 == 0) && ((uintptr_t)gt_full % 16 == 0);
    a.rgb_aligned = (w_full % 16 == 0) && ((uintptr_t)pred_rgb % 16 == 0);
    const size_t smem = (size_t)C * C * kWarps * sizeof(unsigned);
    resample_confusion_kernel<<<persistent_grid(a.total_units), kThreads, smem, (cudaStream_t)stream>>>(a, ph, lut);
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
