// Shared helpers for the pylc_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>

#include "../../include/pylc_b200.h"

namespace pylc {

extern std::atomic<int64_t> g_launches;

// Launch errors are reported by the call itself.  A fault INSIDE a kernel is asynchronous and would surface in
// a later CUDA call of the process; PYLC_SYNC_CHECK=1 in the environment makes every entry point wait for its
// kernel and return that error from the call that caused it (debugging aid: serialises the stream).
inline bool sync_check_enabled() {
    static const bool on = [] {
        const char *e = getenv("PYLC_SYNC_CHECK");
        return e && e[0] == '1';
    }();
    return on;
}
inline int finish_launch() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && sync_check_enabled()) e = cudaDeviceSynchronize();
    return e == cudaSuccess ? PYLC_OK : (int)e;
}

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ---- palette hash ---------------------------------------------------------------------------
// Exact RGB -> class lookup.  key = R | G<<8 | B<<16; slot = ((key * mul) >> 16) & 0xFF depends
// only on the low 24 bits of key, so callers may leave garbage in the top byte.  Each slot holds
// key | cls<<24; empty slots hold class 1, which is also what a miss yields (tools.py:437).
struct PaletteHash {
    uint32_t tab[256];
    uint32_t mul;
};

// Host: builds the table; later duplicates of a colour win (tools.py:441-444).
int build_palette_hash(const uint8_t *palette, int C, PaletteHash *out);

struct ColourLut {
    uint32_t rgb[PYLC_MAX_CLASSES];  // R | G<<8 | B<<16
};
void build_colour_lut(const uint8_t *lut_rgb, int C, ColourLut *out);

// TMA forms (gather_tma.cu).  Return PYLC_OK / a CUDA error after launching, or -1 when the form does not
// apply to the arguments (alignment, geometry, driver without tensor maps): the caller then launches the
// per-thread kernel.
int launch_mask_gather_tma(const uint8_t *src, int n_img, size_t img_stride, int H, int W, size_t pitch, int T, int S, int nH, int nW,
                           const PaletteHash &ph, int C, uint8_t *dst, long long *px_dist, cudaStream_t st);
int launch_class_encode_tma(const uint8_t *rgb, long long rows, long long cols, size_t pitch, const PaletteHash &ph, int C,
                            uint8_t *out, long long *hist, cudaStream_t st);

// PYLC_NO_TMA=1 in the environment keeps every entry point on its per-thread kernel (A/B measurements, tests)
inline bool tma_disabled() {
    const char *e = getenv("PYLC_NO_TMA");
    return e && e[0] == '1';
}
int launch_resample_confusion_tma(const uint8_t *labels, int h, int w, const int32_t *x_ofs, const int32_t *y_ofs, int h_full, int w_full,
                                  const uint8_t *gt_rgb, size_t gt_pitch, const PaletteHash &ph, int C, int n_inject, long long *conf,
                                  cudaStream_t st);

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t encode_key(uint32_t key, const uint32_t *tab, uint32_t mul) {
    uint32_t prod = key * mul;
    uint32_t e = tab[__byte_perm(prod, 0, 0x4442)];
    return ((e ^ key) & 0x00FFFFFFu) ? 1u : (e >> 24);
}

// Table entry of a colour with the class byte replaced by `miss_e`'s when the colour is not in the
// palette; four entries' class bytes pack into one word with three PRMTs.
__device__ __forceinline__ uint32_t lookup_entry(uint32_t key, const uint32_t *tab, uint32_t mul, uint32_t miss_e) {
    const uint32_t e = tab[__byte_perm(key * mul, 0, 0x4442)];
    return ((e ^ key) & 0x00FFFFFFu) ? miss_e : e;
}
__device__ __forceinline__ uint32_t pack_top_bytes(uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3) {
    return __byte_perm(__byte_perm(e0, e1, 0x0073), __byte_perm(e2, e3, 0x0073), 0x5410);
}

// Streaming 16-byte load (read once: bypass L1 allocation).
__device__ __forceinline__ uint4 ld_stream16(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ld_stream_f2(const float *p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream16(void *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream_f4(float *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

// Asynchronous 16-byte global -> shared copy (LDGSTS, L2 only).  A thread that later reads only the
// bytes it copied itself needs no barrier, just cp_async_wait<N>() on its own groups.
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t smem_addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_addr) : "memory");
    return r;
}

// 16 bytes from an arbitrarily aligned address (slow path for unpitched sources).
__device__ __forceinline__ uint4 ld_bytes16(const uint8_t *p) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        w[i] = (uint32_t)__ldg(p + 4 * i) | ((uint32_t)__ldg(p + 4 * i + 1) << 8) |
               ((uint32_t)__ldg(p + 4 * i + 2) << 16) | ((uint32_t)__ldg(p + 4 * i + 3) << 24);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <bool ALIGNED>
__device__ __forceinline__ uint4 ld16(const uint8_t *p) {
    if (ALIGNED) return __ldg(reinterpret_cast<const uint4 *>(p));
    return ld_bytes16(p);
}

// ---- fast transcendental pieces ---------------------------------------------------------------
// One MUFU instruction each.  ex2/lg2.approx carry <= 2 ulp relative error, rcp.approx <= 1 ulp;
// the softmax built from them stays ~1e-6 relative of an exactly rounded one, well inside the
// 1e-5 (stitched map) and 1e-4 (loss) parity tolerances, at a fraction of the issue slots of
// expf / logf / IEEE division -- the stitch and loss kernels are issue-bound otherwise.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// PTX shl clamps shift amounts >= 32 to zero (C's << is undefined there).
__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t s) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
    return r;
}

// ---- per-thread class counters ----------------------------------------------------------------
// Register-resident histograms of u8 class ids, fed SIXTEEN packed ids at a time (four words of four
// class bytes, the form every encode kernel already holds them in); no shared-memory traffic in the
// pixel loop.  Protocol: add16() per 16-pixel unit, widen() after at most kWidenUnits units, read
// count() / reset() after at most kFlushUnits units.
//
// NibbleCounter<NG> (C <= 2*NG <= 14; ids >= 2*NG are ignored, ids must be < 16):
//   PRMT is an 8-entry byte table look-up for four 4-bit indices at once.  Two words of class bytes
//   fold into one word of eight nibbles (w0 + 16*w1), whose halves are the selectors.  The table of
//   group g holds 0x01 for class 2g and 0x10 for class 2g+1 (zero elsewhere), so one PRMT turns four
//   pixels into four byte lanes of two 4-bit one-hot fields, and the lanes add up with IADD3.  An
//   index with bit 3 set makes PRMT replicate the sign bit of the table byte -- zero for these
//   tables -- so classes 8..15 vanish from groups 0..3 and, with the selector XOR 0x8888, classes
//   0..7 vanish from groups 4..7.  Cost: (4 PRMT + 2 IADD3) per group and 16 pixels, about 2.5
//   instructions per pixel at C = 9 against ~6.5 for a shift-and-add of one-hot fields per pixel.
//   A field grows by at most 4 per unit: widen() every 3 units moves the nibble fields into byte
//   fields (<= 12 each time), which hold 21 widenings = 63 units.
template <int NG>
struct NibbleCounter {
    static_assert(NG >= 1 && NG <= 7, "ids 14 and 15 must stay free: 15 is the 'no class' id");
    static constexpr int kWidenUnits = 3, kFlushUnits = 63, kMaxClasses = 2 * NG;
    static constexpr uint32_t kVoid = 0x0F0F0F0Fu;     // four ids that are never counted
    uint32_t nib[NG], wide[NG][2];
    // Arbitrary mask bytes (tiles read back from a database): ids >= 16 become 15.
    static constexpr bool kNeedsSanitize = true;
    static __device__ __forceinline__ uint32_t sanitize(uint32_t w) {
        uint32_t t = w & 0xF0F0F0F0u;
        t |= t >> 1;
        t |= t >> 2;
        const uint32_t bad = ((t >> 4) & 0x01010101u) * 0xFFu;   // 0xFF in every byte that was >= 16
        return (w & ~bad) | (bad & 0x0F0F0F0Fu);
    }
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int g = 0; g < NG; ++g) nib[g] = wide[g][0] = wide[g][1] = 0u;
    }
    static __device__ __forceinline__ uint32_t look(int g, uint32_t sel) {
        const uint32_t a = (g & 3) == 0 ? 0x00001001u : ((g & 3) == 1 ? 0x10010000u : 0u);
        const uint32_t b = (g & 3) == 2 ? 0x00001001u : ((g & 3) == 3 ? 0x10010000u : 0u);
        uint32_t d;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
        return d;
    }
    __device__ __forceinline__ void add16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
        const uint32_t p0 = w0 + (w1 << 4), p1 = w2 + (w3 << 4);
        const uint32_t q0 = p0 >> 16, q1 = p1 >> 16;
#pragma unroll
        for (int g = 0; g < (NG < 4 ? NG : 4); ++g) {
            nib[g] += look(g, p0) + look(g, q0);
            nib[g] += look(g, p1) + look(g, q1);
        }
        if (NG > 4) {
            const uint32_t x0 = p0 ^ 0x88888888u, x1 = p1 ^ 0x88888888u;
            const uint32_t y0 = x0 >> 16, y1 = x1 >> 16;
#pragma unroll
            for (int g = 4; g < NG; ++g) {
                nib[g] += look(g, x0) + look(g, y0);
                nib[g] += look(g, x1) + look(g, y1);
            }
        }
    }
    __device__ __forceinline__ void widen() {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            wide[g][0] += nib[g] & 0x0F0F0F0Fu;
            wide[g][1] += (nib[g] >> 4) & 0x0F0F0F0Fu;
            nib[g] = 0u;
        }
    }
    // c must be a compile-time constant after unrolling (register arrays)
    __device__ __forceinline__ uint32_t count(int c) const { return __dp4a(wide[c >> 1][c & 1], 0x01010101u, 0u); }
};

// C <= 32: four u64 of 8-bit fields, one pixel at a time (<= 15 units between flushes)
struct ByteCounter {
    static constexpr int kWidenUnits = 1 << 30, kFlushUnits = 15, kMaxClasses = PYLC_MAX_CLASSES;
    static constexpr uint32_t kVoid = 0xFFFFFFFFu;
    unsigned long long acc[4];
    static constexpr bool kNeedsSanitize = false;
    static __device__ __forceinline__ uint32_t sanitize(uint32_t w) { return w; }
    __device__ __forceinline__ void reset() { acc[0] = acc[1] = acc[2] = acc[3] = 0; }
    __device__ __forceinline__ void add(uint32_t cls) {   // ids >= 32 are ignored
        unsigned long long one = 1ull << ((cls & 7u) * 8u);
        uint32_t w = cls >> 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += (w == (uint32_t)i) ? one : 0ull;
    }
    __device__ __forceinline__ void add16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
        const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            add(w[k] & 0xFFu);
            add((w[k] >> 8) & 0xFFu);
            add((w[k] >> 16) & 0xFFu);
            add(w[k] >> 24);
        }
    }
    __device__ __forceinline__ void widen() {}
    __device__ __forceinline__ uint32_t count(int c) const {
        unsigned long long src = acc[0];
        if ((c >> 3) == 1) src = acc[1];
        if ((c >> 3) == 2) src = acc[2];
        if ((c >> 3) == 3) src = acc[3];
        return (uint32_t)(src >> ((c & 7) * 8)) & 0xFFu;
    }
};

// NG = 0 selects the byte counter (15 <= C <= 32), otherwise NibbleCounter<NG>
template <int NG>
struct CounterSel {
    using type = NibbleCounter<NG>;
};
template <>
struct CounterSel<0> {
    using type = ByteCounter;
};
// host: counter variant for C classes
inline int counter_groups(int C) { return C <= 10 ? 5 : (C <= 12 ? 6 : (C <= 14 ? 7 : 0)); }

// Turns the bytes j >= valid of word k (of a 16-pixel unit) into the id that counter CC ignores.
template <class CC>
__device__ __forceinline__ uint32_t void_tail(uint32_t w, int k, int valid) {
    const int n = valid - 4 * k;                        // valid bytes in this word
    return n >= 4 ? w : (n <= 0 ? CC::kVoid : ((w & ~(0xFFFFFFFFu << (8 * n))) | (CC::kVoid << (8 * n))));
}

// Warp-aggregated flush of a counter into a shared-memory histogram (one atomic per class per
// warp).  Must be called by all 32 lanes.
template <class CC>
__device__ __forceinline__ void flush_counter(const CC &cc, int C, unsigned *s_hist) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < CC::kMaxClasses; ++c) {
        if (c < C) {
            unsigned v = __reduce_add_sync(0xFFFFFFFFu, cc.count(c));
            if (lane == 0 && v) atomicAdd(&s_hist[c], v);
        }
    }
}

#endif  // __CUDACC__

}  // namespace pylc
