// Shared helpers for the pylc_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/pylc_b200.h"

namespace pylc {

extern std::atomic<int64_t> g_launches;

inline int finish_launch() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
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
// Register-resident histogram of u8 class ids without shared-memory atomics in the pixel loop.
//   C <= 12 : one u64 of 5-bit fields per 16-pixel unit, widened into two u64 of 10-bit fields
//             (<= 63 units between flushes)
//   C <= 32 : four u64 of 8-bit fields (<= 15 units between flushes)
template <bool WIDE>
struct ClassCounter;

template <>
struct ClassCounter<false> {
    static constexpr unsigned long long kEven = 0x01F07C1F07C1Full | (0x1Full << 50);  // fields 0,2,..,10
    unsigned long long unit, even, odd;
    __device__ __forceinline__ void reset() { unit = even = odd = 0; }
    // PTX shl.b64 clamps amounts >= 64 to zero, so out-of-range class ids (>= 13) vanish
    // instead of being undefined; id 12 lands in bits 60..63, which count() never reads.
    __device__ __forceinline__ void add(uint32_t cls) {
        unsigned long long one;
        asm("shl.b64 %0, 1, %1;" : "=l"(one) : "r"(cls * 5u));
        unit += one;
    }
    // call after every <= 31 pixels
    __device__ __forceinline__ void end_unit() {
        even += unit & kEven;
        odd += (unit >> 5) & kEven;
        unit = 0;
    }
    __device__ __forceinline__ uint32_t count(int c) const {
        unsigned long long src = (c & 1) ? odd : even;
        return (uint32_t)(src >> ((c >> 1) * 10)) & 0x3FFu;
    }
};

template <>
struct ClassCounter<true> {
    unsigned long long acc[4];
    __device__ __forceinline__ void reset() { acc[0] = acc[1] = acc[2] = acc[3] = 0; }
    __device__ __forceinline__ void add(uint32_t cls) {
        unsigned long long one = 1ull << ((cls & 7u) * 8u);
        uint32_t w = cls >> 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += (w == (uint32_t)i) ? one : 0ull;
    }
    __device__ __forceinline__ void end_unit() {}
    __device__ __forceinline__ uint32_t count(int c) const {
        unsigned long long src = acc[0];
        if ((c >> 3) == 1) src = acc[1];
        if ((c >> 3) == 2) src = acc[2];
        if ((c >> 3) == 3) src = acc[3];
        return (uint32_t)(src >> ((c & 7) * 8)) & 0xFFu;
    }
};

// Warp-aggregated flush of a ClassCounter into a shared-memory histogram (one atomic per class
// per warp).  Must be called by all 32 lanes.
template <bool WIDE>
__device__ __forceinline__ void flush_counter(const ClassCounter<WIDE> &cc, int C, unsigned *s_hist) {
    const int lane = threadIdx.x & 31;
    for (int c = 0; c < C; ++c) {
        unsigned v = __reduce_add_sync(0xFFFFFFFFu, cc.count(c));
        if (lane == 0 && v) atomicAdd(&s_hist[c], v);
    }
}

#endif  // __CUDACC__

}  // namespace pylc
