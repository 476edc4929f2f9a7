// MultiLoss = ce_w*CE + dice_w*Dice + focal_w*Focal (models/modules/loss.py:71-194), forward and
// backward, as two streaming passes over the logits:
//
//   reduce : softmax once per pixel, accumulate I_c, K_c, the CE numerator/denominator and the
//            focal sum (2C+3 numbers) -- Dice needs these batch-global sums before any gradient
//            exists, and under data parallelism they are all-reduced between the two passes.
//   grad   : softmax again, closed-form gradient written with 128-bit stores.
//
// Logits are [B, C, HW]; a thread handles PX consecutive pixels and reads one 16-byte vector per
// class plane, so every warp-level load is a contiguous 512-byte segment.
#include <cooperative_groups.h>

#include "common.cuh"

namespace pylc {

struct LossArgs {
    const float *logits;
    const void *target;
    const float *class_w;  // device, nullable
    int target_is_i64;
    int B, C;
    long long HW, units_per_img, total_units;
    pylc_loss_cfg cfg;
    uint8_t *t8_out;       // nullable: the reduce pass leaves a u8 copy of the targets here (1 B/px instead of 8 in the second pass)
    int t8_coherent;       // u8 targets were written earlier in THIS launch by other CTAs: read them through L2 (ld.global.cg)
};

__device__ __forceinline__ uint32_t ld_cg_u32(const void *p) {
    uint32_t r;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_cg_u16(const void *p) {
    unsigned short r;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_cg_u8(const void *p) {
    uint32_t r;
    asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

template <int PX>
struct PxVec;
template <>
struct PxVec<4> {
    static __device__ __forceinline__ void load(const float *p, float (&o)[4]) {
        const float4 t = ld_stream_f4(p);
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[4]) {
        st_stream_f4(p, make_float4(v[0], v[1], v[2], v[3]));
    }
    static __device__ __forceinline__ void load_target(const void *t, int is_i64, int cg, long long idx, int (&o)[4]) {
        if (is_i64) {
            const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(static_cast<const long long *>(t) + idx));
            const longlong2 b = __ldg(reinterpret_cast<const longlong2 *>(static_cast<const long long *>(t) + idx + 2));
            o[0] = (int)a.x; o[1] = (int)a.y; o[2] = (int)b.x; o[3] = (int)b.y;
        } else {
            const uint8_t *p = static_cast<const uint8_t *>(t) + idx;
            const uint32_t w = cg ? ld_cg_u32(p) : __ldg(reinterpret_cast<const uint32_t *>(p));
            o[0] = w & 0xFF; o[1] = (w >> 8) & 0xFF; o[2] = (w >> 16) & 0xFF; o[3] = w >> 24;
        }
    }
    static __device__ __forceinline__ void store_t8(uint8_t *p, const int (&t)[4]) {
        *reinterpret_cast<uint32_t *>(p) = (uint32_t)(t[0] & 0xFF) | ((uint32_t)(t[1] & 0xFF) << 8) | ((uint32_t)(t[2] & 0xFF) << 16) | ((uint32_t)t[3] << 24);
    }
};
template <>
struct PxVec<2> {
    static __device__ __forceinline__ void load(const float *p, float (&o)[2]) {
        const float2 t = ld_stream_f2(p);
        o[0] = t.x; o[1] = t.y;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[2]) {
        asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
    }
    static __device__ __forceinline__ void load_target(const void *t, int is_i64, int cg, long long idx, int (&o)[2]) {
        if (is_i64) {
            const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(static_cast<const long long *>(t) + idx));
            o[0] = (int)a.x; o[1] = (int)a.y;
        } else {
            const uint8_t *p = static_cast<const uint8_t *>(t) + idx;
            const uint32_t w = cg ? ld_cg_u16(p) : (uint32_t)__ldg(reinterpret_cast<const unsigned short *>(p));
            o[0] = w & 0xFF; o[1] = w >> 8;
        }
    }
    static __device__ __forceinline__ void store_t8(uint8_t *p, const int (&t)[2]) {
        *reinterpret_cast<unsigned short *>(p) = (unsigned short)((t[0] & 0xFF) | ((t[1] & 0xFF) << 8));
    }
};
template <>
struct PxVec<1> {
    static __device__ __forceinline__ void load(const float *p, float (&o)[1]) { o[0] = __ldg(p); }
    static __device__ __forceinline__ void store(float *p, const float (&v)[1]) { *p = v[0]; }
    static __device__ __forceinline__ void load_target(const void *t, int is_i64, int cg, long long idx, int (&o)[1]) {
        o[0] = is_i64 ? (int)__ldg(static_cast<const long long *>(t) + idx)
                      : (cg ? (int)ld_cg_u8(static_cast<const uint8_t *>(t) + idx) : (int)__ldg(static_cast<const uint8_t *>(t) + idx));
    }
    static __device__ __forceinline__ void store_t8(uint8_t *p, const int (&t)[1]) { *p = (uint8_t)t[0]; }
};

__device__ __forceinline__ float pow_gamma(float base, float gamma) {
    if (gamma == 2.f) return base * base;
    if (gamma == 1.f) return base;
    if (gamma == 0.f) return 1.f;
    return powf(base, gamma);
}

// softmax of one pixel column; returns max and sum, leaves exp(z - max) in e[].  One FMA + one
// MUFU.EX2 per class (see common.cuh): ~1e-6 relative, two orders inside the 1e-4 loss tolerance.
template <int CMAX>
__device__ __forceinline__ void softmax_px(const float (&z)[CMAX], int C, float (&e)[CMAX], float &m, float &s) {
    m = z[0];
#pragma unroll
    for (int c = 1; c < CMAX; ++c)
        if (c < C) m = fmaxf(m, z[c]);
    const float ms = -m * kLog2e;
    s = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
        if (c < C) {
            e[c] = ex2_approx(fmaf(z[c], kLog2e, ms));
            s += e[c];
        }
}

__device__ __forceinline__ float ln_fast(float x) { return lg2_approx(x) * kLn2; }

// (image, unit-in-image) position of a thread's grid-stride walk over the units, forwards or backwards:
// one 64-bit division when the walk starts, constant steps with a carry afterwards.
template <bool REVERSE>
struct UnitWalk {
    int b, r, left;                  // image, unit inside the image, units this thread still has to do
    int step_b, step_r, upi;         // (the host rejects images of 2^31 units or more)
    __device__ __forceinline__ UnitWalk(const LossArgs &a) {
        const long long stride = (long long)gridDim.x * kThreads, first = (long long)blockIdx.x * kThreads + threadIdx.x;
        upi = (int)a.units_per_img;
        left = first < a.total_units ? (int)((a.total_units - first + stride - 1) / stride) : 0;
        const long long u = REVERSE ? a.total_units - 1 - first : first;
        b = left ? (int)(u / upi) : 0;
        r = left ? (int)(u - (long long)b * upi) : 0;
        step_b = (int)(stride / upi);
        step_r = (int)(stride - (long long)step_b * upi);
    }
    __device__ __forceinline__ void next() {
        --left;
        if (REVERSE) {
            b -= step_b;
            r -= step_r;
            if (r < 0) {
                r += upi;
                --b;
            }
        } else {
            b += step_b;
            r += step_r;
            if (r >= upi) {
                r -= upi;
                ++b;
            }
        }
    }
};

template <int C_T, int CMAX, int PX>
__device__ __forceinline__ void loss_reduce_pass(const LossArgs &a, double *partials) {
    const int C = C_T > 0 ? C_T : a.C;
    __shared__ float s_w[PYLC_MAX_CLASSES];
    __shared__ double s_part[2 * PYLC_MAX_CLASSES + 3];
    if (threadIdx.x < PYLC_MAX_CLASSES) s_w[threadIdx.x] = (a.class_w && threadIdx.x < C) ? a.class_w[threadIdx.x] : 1.f;
    if (threadIdx.x < 2 * PYLC_MAX_CLASSES + 3) s_part[threadIdx.x] = 0.0;
    __syncthreads();

    float accI[CMAX], accP[CMAX];
    uint32_t accN[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        accI[c] = 0.f;
        accP[c] = 0.f;
        accN[c] = 0;
    }
    float ce_num = 0.f, ce_den = 0.f, focal = 0.f;
    bool bad = false;      // a target outside [0, C): the reference's CrossEntropyLoss / one_hot raise (loss.py:66-69,137)
    const float eps = a.cfg.eps, gamma = a.cfg.fl_gamma, alpha = a.cfg.fl_alpha;

    for (UnitWalk<false> w(a); w.left > 0; w.next()) {
        const long long b = w.b, off = (long long)w.r * PX;
        const float *p = a.logits + ((size_t)b * C) * a.HW + off;
        float z[CMAX][PX];
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) PxVec<PX>::load(p + (size_t)c * a.HW, z[c]);
        int t[PX];
        PxVec<PX>::load_target(a.target, a.target_is_i64, 0, b * a.HW + off, t);
        if (a.t8_out) PxVec<PX>::store_t8(a.t8_out + b * a.HW + off, t);
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            bad |= (unsigned)t[j] >= (unsigned)C;
            float zc[CMAX], e[CMAX], m, s;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) zc[c] = z[c][j];
            softmax_px<CMAX>(zc, C, e, m, s);
            const float r = rcp_approx(s);
            // target-class picks: FMA with a 0/1 mask for register arrays, a shared-memory read for
            // the class weight.  (A `hit ? arr[c] : x` select chain gets turned into a dynamically
            // indexed local-memory table by the compiler, which costs an L1 round trip per pixel.)
            float zt = 0.f, pt = 0.f;
            const float wt = s_w[t[j] & (PYLC_MAX_CLASSES - 1)];
#pragma unroll
            for (int c = 0; c < CMAX; ++c)
                if (c < C) {
                    const float pc = e[c] * r;
                    const bool hit = t[j] == c;
                    const float hf = hit ? 1.f : 0.f;
                    accP[c] += pc;
                    accI[c] = fmaf(hf, pc, accI[c]);
                    accN[c] += hit ? 1u : 0u;
                    zt = fmaf(hf, zc[c], zt);
                    pt = fmaf(hf, pc, pt);
                }
            ce_num += wt * (ln_fast(s) + (m - zt));
            ce_den += wt;
            const float q = pt + eps;
            focal += -alpha * pow_gamma(1.f - q, gamma) * ln_fast(q);
        }
    }

    // an out-of-range target poisons the CE numerator: every loss value and every gradient of this batch
    // becomes NaN instead of silently using another class's weight (the host mirror turns it into an error)
    if (bad) ce_num = ce_den = __int_as_float(0x7FC00000);      // ce_den reaches every gradient through 1 / sum(w_t)
    // block reduction: warp shuffles in f32, cross-warp in f64 shared atomics
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            float vi = accI[c], vp = accP[c];
            uint32_t vn = __reduce_add_sync(0xFFFFFFFFu, accN[c]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                vi += __shfl_xor_sync(0xFFFFFFFFu, vi, o);
                vp += __shfl_xor_sync(0xFFFFFFFFu, vp, o);
            }
            if (lane == 0) {
                atomicAdd(&s_part[c], (double)vi);
                atomicAdd(&s_part[C + c], (double)vp + (double)vn);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ce_num += __shfl_xor_sync(0xFFFFFFFFu, ce_num, o);
        ce_den += __shfl_xor_sync(0xFFFFFFFFu, ce_den, o);
        focal += __shfl_xor_sync(0xFFFFFFFFu, focal, o);
    }
    if (lane == 0) {
        atomicAdd(&s_part[2 * C], (double)ce_num);
        atomicAdd(&s_part[2 * C + 1], (double)ce_den);
        atomicAdd(&s_part[2 * C + 2], (double)focal);
    }
    __syncthreads();
    if (threadIdx.x < 2 * C + 3) atomicAdd(&partials[threadIdx.x], s_part[threadIdx.x]);
}

template <int C_T, int CMAX, int PX>
__global__ void __launch_bounds__(kThreads, (PX == 4 || CMAX > 9) ? 2 : 3) loss_reduce_kernel(LossArgs a, double *__restrict__ partials) {
    loss_reduce_pass<C_T, CMAX, PX>(a, partials);
}

// REVERSE walks the units back to front: in the fused kernel the reduce pass has just streamed the
// logits front to back, so the tail it left in the 126 MB L2 is what this pass reads first.
template <int C_T, int CMAX, int PX, bool REVERSE>
__device__ __forceinline__ void loss_grad_pass(const LossArgs &a, const double *partials, long long n_px_total, float grad_scale,
                                               const float *grad_scale_dev, float *grad) {
    const int C = C_T > 0 ? C_T : a.C;
    __shared__ float s_w[PYLC_MAX_CLASSES], s_a[PYLC_MAX_CLASSES], s_b[PYLC_MAX_CLASSES];
    __shared__ float s_inv_den;
    if (threadIdx.x < PYLC_MAX_CLASSES) {
        const int c = threadIdx.x;
        s_w[c] = (a.class_w && c < C) ? a.class_w[c] : 1.f;
        float av = 0.f, bv = 0.f;
        if (c < C) {
            const double I = __ldcg(partials + c), K = __ldcg(partials + C + c), sm = (double)a.cfg.dice_smooth;
            av = (float)(-2.0 / ((K + sm) * C));
            bv = (float)((2.0 * I + sm) / ((K + sm) * (K + sm) * C));
        }
        s_a[c] = av;
        s_b[c] = bv;
    }
    if (threadIdx.x == 0) s_inv_den = (float)(1.0 / __ldcg(partials + 2 * C + 1));
    __syncthreads();
    if (grad_scale_dev) grad_scale *= __ldg(grad_scale_dev);   // upstream dL/d(loss) without a host sync
    const float eps = a.cfg.eps, gamma = a.cfg.fl_gamma;
    const float lce = a.cfg.ce_weight * s_inv_den, ld = a.cfg.dice_weight;
    const float lf = a.cfg.focal_weight * a.cfg.fl_alpha / (float)n_px_total;
    float bb[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) bb[c] = c < C ? s_b[c] : 0.f;

    for (UnitWalk<REVERSE> w(a); w.left > 0; w.next()) {
        const long long b = w.b, off = (long long)w.r * PX;
        const size_t base = ((size_t)b * C) * a.HW + off;
        float z[CMAX][PX];
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) PxVec<PX>::load(a.logits + base + (size_t)c * a.HW, z[c]);
        int t[PX];
        PxVec<PX>::load_target(a.target, a.target_is_i64, a.t8_coherent, b * a.HW + off, t);
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            float zc[CMAX], e[CMAX], m, s;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) zc[c] = z[c][j];
            softmax_px<CMAX>(zc, C, e, m, s);
            const float r = rcp_approx(s);
            float pt = 0.f, sb = 0.f;
            const int tj = t[j] & (PYLC_MAX_CLASSES - 1);
            const float wt = s_w[tj], at = s_a[tj];   // shared-memory picks (see loss_reduce_kernel)
#pragma unroll
            for (int c = 0; c < CMAX; ++c)
                if (c < C) {
                    e[c] *= r;  // p_c
                    pt = fmaf((t[j] == c) ? 1.f : 0.f, e[c], pt);
                    sb = fmaf(bb[c], e[c], sb);
                }
            const float q = pt + eps, omq = 1.f - q;
            const float dq = gamma * pow_gamma(omq, gamma - 1.f) * ln_fast(q) - pow_gamma(omq, gamma) * rcp_approx(q);
            const float f = lf * dq * pt;
            const float ca = lce * wt;
            const float common = ca - f - ld * (at * pt + sb);
            const float hit_term = fmaf(ld * pt, at, f - ca);   // extra gradient of the target class
#pragma unroll
            for (int c = 0; c < CMAX; ++c)
                if (c < C) {
                    float g = e[c] * fmaf(ld, bb[c], common);
                    g = fmaf((t[j] == c) ? 1.f : 0.f, hit_term, g);
                    z[c][j] = g * grad_scale;
                }
        }
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) PxVec<PX>::store(grad + base + (size_t)c * a.HW, z[c]);
    }
}

template <int C_T, int CMAX, int PX>
__global__ void __launch_bounds__(kThreads, (PX == 4 || CMAX > 9) ? 2 : 3)
    loss_grad_kernel(LossArgs a, const double *partials, long long n_px_total, float grad_scale,
                     const float *__restrict__ grad_scale_dev, float *__restrict__ grad) {
    loss_grad_pass<C_T, CMAX, PX, false>(a, partials, n_px_total, grad_scale, grad_scale_dev, grad);
}

__device__ __forceinline__ void loss_finalize(const double *partials, int C, long long n_px_total, const pylc_loss_cfg &cfg,
                                              float *out4) {
    const double ce = __ldcg(partials + 2 * C) / __ldcg(partials + 2 * C + 1);
    double dice = 0.0;
    for (int c = 0; c < C; ++c)
        dice += 1.0 - (2.0 * __ldcg(partials + c) + cfg.dice_smooth) / (__ldcg(partials + C + c) + cfg.dice_smooth);
    dice /= C;
    const double focal = __ldcg(partials + 2 * C + 2) / (double)n_px_total;
    out4[0] = (float)(cfg.ce_weight * ce + cfg.dice_weight * dice + cfg.focal_weight * focal);
    out4[1] = (float)ce;
    out4[2] = (float)dice;
    out4[3] = (float)focal;
}

__global__ void loss_finalize_kernel(const double *partials, int C, long long n_px_total, pylc_loss_cfg cfg, float *__restrict__ out4) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    loss_finalize(partials, C, n_px_total, cfg, out4);
}

// Forward and backward in ONE cooperative launch (single-GPU training step): reduce pass, grid-wide
// barrier, loss values written by one thread, gradient pass in reverse order.  Co-residency of the
// whole grid is what cudaLaunchCooperativeKernel guarantees; the grid is the same one-wave
// persistent grid the two-launch path uses.
// ---- in-kernel all-reduce of the partials over peer memory (data-parallel training step) -------------------
// Every rank owns one exchange workspace in memory its peers can address (symmetric memory: the same virtual
// allocation mapped into every process over NVLink / NVSwitch); `peer_ws[r]` is rank r's, seen from here.  Two
// slots alternate by call parity; a slot holds one arrival flag per rank and the rank's 2C+3 sums:
//     struct { uint64 flag[PYLC_DP_MAX_RANKS]; double part[PYLC_DP_MAX_PARTIALS]; } slot[2];
// Protocol of call number `epoch` (the same on every rank, starting at 1), run by CTA 0 after the grid has
// finished the local reduction:
//   1. publish the local sums in the own slot, fence at system scope;
//   2. thread p stores `epoch` into flag[rank] of peer p's slot (st.release.sys -- a store over the fabric);
//   3. thread p spins on the own slot's flag[p] (ld.acquire.sys) until it reads `epoch`: rank p's sums are visible;
//   4. thread t adds the ranks' part[t] in rank order (the same order everywhere: bit-identical results on every
//      rank) and writes the total over the local partials.
// A rank can run at most one call ahead of its slowest peer (it needs that peer's flag of the current call), so
// two slots are enough.  The spin is bounded: a peer that never arrives (a crashed rank, mismatched epochs) turns
// the partials into NaN after ~1 s instead of hanging the device.
struct DpArgs {
    double *const *peer_ws;     // DEVICE array [world] of workspace pointers (every rank's, in rank order)
    int rank, world;
    unsigned long long epoch;
};
constexpr int kDpSlotDoubles = PYLC_DP_MAX_RANKS + PYLC_DP_MAX_PARTIALS;
static_assert(2 * kDpSlotDoubles * 8 <= PYLC_DP_WS_BYTES, "workspace size in the header");

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// called by every thread of ONE CTA; n = 2C+3
__device__ __forceinline__ void dp_all_reduce_partials(const DpArgs &dp, double *partials, int n) {
    const int t = threadIdx.x;
    const size_t slot = (size_t)(dp.epoch & 1ull) * kDpSlotDoubles;
    double *mine = dp.peer_ws[dp.rank] + slot;
    if (t < n) mine[PYLC_DP_MAX_RANKS + t] = __ldcg(partials + t);
    __threadfence_system();
    __syncthreads();
    __shared__ int s_ok;
    if (t == 0) s_ok = 1;
    __syncthreads();
    if (t < dp.world) {
        st_release_sys_u64(reinterpret_cast<unsigned long long *>(dp.peer_ws[t] + slot) + dp.rank, dp.epoch);
        const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(mine) + t;
        long long spins = 0;
        while (ld_acquire_sys_u64(flag) != dp.epoch) {
            if (++spins > (1ll << 24)) {        // ~1 s with the sleep below
                s_ok = 0;
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (t < n) {
        double sum = 0.0;
        for (int r = 0; r < dp.world; ++r) sum += ld_relaxed_sys_f64(dp.peer_ws[r] + slot + PYLC_DP_MAX_RANKS + t);
        partials[t] = s_ok ? sum : __longlong_as_double(0x7FF8000000000000ll);
    }
    __threadfence();
}

template <int C_T, int CMAX, int PX>
__global__ void __launch_bounds__(kThreads, (PX == 4 || CMAX > 9) ? 2 : 3)
    loss_fused_kernel(LossArgs a, double *partials, long long n_px_total, float grad_scale, const float *grad_scale_dev,
                      float *grad, float *out4, DpArgs dp) {
    loss_reduce_pass<C_T, CMAX, PX>(a, partials);
    __threadfence();
    cooperative_groups::this_grid().sync();
    if (dp.world > 1) {          // data parallel: the sums of the single large batch, exchanged inside the launch
        if (blockIdx.x == 0) dp_all_reduce_partials(dp, partials, 2 * (C_T > 0 ? C_T : a.C) + 3);
        cooperative_groups::this_grid().sync();
    }
    if (out4 && blockIdx.x == 0 && threadIdx.x == 0) loss_finalize(partials, C_T > 0 ? C_T : a.C, n_px_total, a.cfg, out4);
    if (a.t8_out) {            // the gradient pass reads the 1-byte targets the reduce pass left behind
        a.target = a.t8_out;
        a.target_is_i64 = 0;
        a.t8_coherent = 1;
        a.t8_out = nullptr;
    }
    loss_grad_pass<C_T, CMAX, PX, true>(a, partials, n_px_total, grad_scale, grad_scale_dev, grad);
}

// grad *= *scale unless *scale == 1 (the usual loss.backward()): lets the fused kernel's gradient be
// handed to autograd without a host sync and, in the usual case, without touching it again.
__global__ void __launch_bounds__(kThreads) scale_unless_one_kernel(float *__restrict__ g, long long n4, long long n,
                                                                    const float *__restrict__ scale) {
    const float s = __ldg(scale);
    if (s == 1.f) return;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kThreads) {
        float4 v = reinterpret_cast<float4 *>(g)[i];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        reinterpret_cast<float4 *>(g)[i] = v;
    }
    if (blockIdx.x == 0)
        for (long long i = n4 * 4 + threadIdx.x; i < n; i += kThreads) g[i] *= s;
}

static int fill_args(const float *logits, const void *target, int target_is_i64, int B, int C, int64_t HW,
                     const float *class_w, const pylc_loss_cfg *cfg, LossArgs *a, int *px, uint8_t *t8_out = nullptr) {
    if (!logits || !target || !cfg || B < 1 || HW < 1) return PYLC_ERR_ARG;
    if (C < 1 || C > PYLC_MAX_CLASSES) return PYLC_ERR_CLASSES;
    if (HW > 0x7FFFFFFFll) return PYLC_ERR_GEOMETRY;   // 32-bit position inside an image (UnitWalk)
    // pixels per lane: 4 (one 128-bit load per class plane) while the per-thread state fits the
    // register file (C <= 6), else 2 (64-bit loads, 3 CTAs per SM) -- the warp-level access stays one
    // contiguous run per plane either way, and nothing is demoted to local memory
    const int want = C <= 6 ? 4 : 2;
    const uintptr_t talign = (uintptr_t)(target_is_i64 ? 8 : 1) * want;
    const bool vec = (HW % want == 0) && ((uintptr_t)logits % (4 * want) == 0) && ((uintptr_t)target % talign == 0) && C <= 16 &&
                     ((uintptr_t)t8_out % want == 0);
    *px = vec ? want : 1;
    a->t8_out = target_is_i64 ? t8_out : nullptr;      // u8 targets need no copy
    a->t8_coherent = 0;
    a->logits = logits;
    a->target = target;
    a->class_w = class_w;
    a->target_is_i64 = target_is_i64;
    a->B = B;
    a->C = C;
    a->HW = HW;
    a->units_per_img = HW / *px;
    a->total_units = a->units_per_img * B;
    a->cfg = *cfg;
    return PYLC_OK;
}

// persistent grid: exactly one wave of resident CTAs, capped by the amount of work
template <typename K>
static unsigned loss_grid(K kernel, long long units) {
    int dev = 0, sms = 148, per_sm = 2;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    const long long want = (units + kThreads - 1) / kThreads;
    const long long cap = (long long)sms * per_sm;
    return (unsigned)(want < cap ? want : cap);
}

}  // namespace pylc

using namespace pylc;

#define LAUNCH_LOSS(K, ...) K<<<loss_grid(K, a.total_units), kThreads, 0, st>>>(__VA_ARGS__)
#define DISPATCH_LOSS(KERNEL, ...)                                                                      \
    do {                                                                                                \
        if (px == 4) {                                                                                  \
            if (C <= 4) LAUNCH_LOSS((KERNEL<0, 4, 4>), __VA_ARGS__);                        \
            else LAUNCH_LOSS((KERNEL<0, 6, 4>), __VA_ARGS__);                               \
        } else if (px == 2) {                                                                           \
            if (C == 9) LAUNCH_LOSS((KERNEL<9, 9, 2>), __VA_ARGS__);                        \
            else if (C == 11) LAUNCH_LOSS((KERNEL<11, 11, 2>), __VA_ARGS__);                \
            else if (C <= 8) LAUNCH_LOSS((KERNEL<0, 8, 2>), __VA_ARGS__);                   \
            else if (C <= 12) LAUNCH_LOSS((KERNEL<0, 12, 2>), __VA_ARGS__);                 \
            else LAUNCH_LOSS((KERNEL<0, 16, 2>), __VA_ARGS__);                              \
        } else {                                                                                        \
            if (C <= 12) LAUNCH_LOSS((KERNEL<0, 12, 1>), __VA_ARGS__);                      \
            else LAUNCH_LOSS((KERNEL<0, 32, 1>), __VA_ARGS__);                              \
        }                                                                                               \
    } while (0)

extern "C" int pylc_multiloss_reduce(const float *logits, const void *target, int target_is_i64, int B, int C,
                                     int64_t HW, const float *class_w, const pylc_loss_cfg *cfg, double *partials,
                                     uint8_t *target_u8_out, pylc_stream_t stream) {
    if (!partials) return PYLC_ERR_ARG;
    LossArgs a;
    int px;
    int rc = fill_args(logits, target, target_is_i64, B, C, HW, class_w, cfg, &a, &px, target_u8_out);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_LOSS(loss_reduce_kernel, a, partials);
    return finish_launch();
}

extern "C" int pylc_multiloss_grad(const float *logits, const void *target, int target_is_i64, int B, int C,
                                   int64_t HW, const float *class_w, const pylc_loss_cfg *cfg, const double *partials,
                                   int64_t n_px_total, float grad_scale, const float *grad_scale_dev, float *grad,
                                   pylc_stream_t stream) {
    if (!partials || !grad || n_px_total < 1) return PYLC_ERR_ARG;
    if ((uintptr_t)grad % 16) return PYLC_ERR_ALIGN;
    LossArgs a;
    int px;
    int rc = fill_args(logits, target, target_is_i64, B, C, HW, class_w, cfg, &a, &px);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_LOSS(loss_grad_kernel, a, partials, (long long)n_px_total, grad_scale, grad_scale_dev, grad);
    return finish_launch();
}

extern "C" int pylc_multiloss_finalize(const double *partials, int C, int64_t n_px_total, const pylc_loss_cfg *cfg,
                                       float *out4, pylc_stream_t stream) {
    if (!partials || !cfg || !out4 || n_px_total < 1) return PYLC_ERR_ARG;
    if (C < 1 || C > PYLC_MAX_CLASSES) return PYLC_ERR_CLASSES;
    loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(partials, C, (long long)n_px_total, *cfg, out4);
    return finish_launch();
}

static int launch_loss_fused(const float *logits, const void *target, int target_is_i64, int B, int C, int64_t HW,
                             const float *class_w, const pylc_loss_cfg *cfg, double *partials, float grad_scale,
                             const float *grad_scale_dev, float *grad, float *out4, uint8_t *target_u8_ws, DpArgs dp,
                             pylc_stream_t stream) {
    if (!partials || !grad) return PYLC_ERR_ARG;
    if ((uintptr_t)grad % 16) return PYLC_ERR_ALIGN;
    LossArgs a;
    int px;
    int rc = fill_args(logits, target, target_is_i64, B, C, HW, class_w, cfg, &a, &px, target_u8_ws);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    long long n_px_total = (long long)B * HW * dp.world;
    void *args[] = {&a, &partials, &n_px_total, &grad_scale, &grad_scale_dev, &grad, &out4, &dp};
    cudaError_t e = cudaSuccess;
#define LAUNCH_FUSED(K) e = cudaLaunchCooperativeKernel((void *)K, dim3(loss_grid(K, a.total_units)), dim3(kThreads), args, 0, st)
    if (px == 4) {
        if (C <= 4) LAUNCH_FUSED((loss_fused_kernel<0, 4, 4>));
        else LAUNCH_FUSED((loss_fused_kernel<0, 6, 4>));
    } else if (px == 2) {
        if (C == 9) LAUNCH_FUSED((loss_fused_kernel<9, 9, 2>));
        else if (C == 11) LAUNCH_FUSED((loss_fused_kernel<11, 11, 2>));
        else if (C <= 8) LAUNCH_FUSED((loss_fused_kernel<0, 8, 2>));
        else if (C <= 12) LAUNCH_FUSED((loss_fused_kernel<0, 12, 2>));
        else LAUNCH_FUSED((loss_fused_kernel<0, 16, 2>));
    } else {
        if (C <= 12) LAUNCH_FUSED((loss_fused_kernel<0, 12, 1>));
        else LAUNCH_FUSED((loss_fused_kernel<0, 32, 1>));
    }
#undef LAUNCH_FUSED
    if (e != cudaSuccess) return (int)e;
    return finish_launch();
}

extern "C" int pylc_multiloss_fwd_bwd(const float *logits, const void *target, int target_is_i64, int B, int C, int64_t HW,
                                      const float *class_w, const pylc_loss_cfg *cfg, double *partials, float grad_scale,
                                      const float *grad_scale_dev, float *grad, float *out4, uint8_t *target_u8_ws,
                                      pylc_stream_t stream) {
    DpArgs dp;
    dp.peer_ws = nullptr; dp.rank = 0; dp.world = 1; dp.epoch = 0;
    return launch_loss_fused(logits, target, target_is_i64, B, C, HW, class_w, cfg, partials, grad_scale, grad_scale_dev, grad, out4,
                             target_u8_ws, dp, stream);
}

extern "C" int pylc_multiloss_fwd_bwd_dp(const float *logits, const void *target, int target_is_i64, int B, int C, int64_t HW,
                                         const float *class_w, const pylc_loss_cfg *cfg, double *partials, float grad_scale,
                                         const float *grad_scale_dev, float *grad, float *out4, uint8_t *target_u8_ws,
                                         double *const *peer_ws, int rank, int world, uint64_t epoch, pylc_stream_t stream) {
    if (!peer_ws || world < 1 || world > PYLC_DP_MAX_RANKS || rank < 0 || rank >= world || epoch == 0) return PYLC_ERR_ARG;
    if (2 * C + 3 > PYLC_DP_MAX_PARTIALS) return PYLC_ERR_CLASSES;
    DpArgs dp;
    dp.peer_ws = peer_ws; dp.rank = rank; dp.world = world; dp.epoch = epoch;
    return launch_loss_fused(logits, target, target_is_i64, B, C, HW, class_w, cfg, partials, grad_scale, grad_scale_dev, grad, out4,
                             target_u8_ws, dp, stream);
}

extern "C" int pylc_scale_unless_one_f32(float *data, int64_t n, const float *scale_dev, pylc_stream_t stream) {
    if (!data || !scale_dev || n < 0) return PYLC_ERR_ARG;
    if ((uintptr_t)data % 16) return PYLC_ERR_ALIGN;
    if (n == 0) return PYLC_OK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long n4 = n / 4;
    long long want = (n4 + kThreads - 1) / kThreads;
    if (want > (long long)sms * 8) want = (long long)sms * 8;
    if (want < 1) want = 1;
    scale_unless_one_kernel<<<(unsigned)want, kThreads, 0, (cudaStream_t)stream>>>(data, n4, (long long)n, scale_dev);
    return finish_launch();
}
