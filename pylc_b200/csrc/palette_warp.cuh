// Warp-cooperative exact palette encode of RGB staged in shared memory (utils/tools.py:435-449).
//
// The exact per-pixel look-up (key extract, hash multiply, slot, LDS, compare, select, pack) costs ~9
// instructions per pixel, which is what bound the round-1 palette kernels (ncu: 59 % issue, 29 % DRAM).
// Label masks are piecewise constant, so this form looks up ONE pixel per pair of 4-pixel groups and proves the
// other seven equal to it with byte permutes and compares on the pair's 24 bytes (three per group, plus one
// compare of the groups' first words):
//
//   phase A  every lane: NU units of 16 pixels = 4 NU groups; per group the anchor pixel's table entry is
//            broadcast to four class bytes and a flag says whether the 12 bytes are NOT one repeated
//            pixel; the 16 bytes go to the unit's slot of the output box in shared memory
//   phase B  the flagged groups of the whole warp are compacted into a small shared-memory queue (one
//            byte per group) and re-encoded pixel by pixel, one group per lane, straight into the output
//            box; when more than kDenseGroups groups are flagged (noise-like masks) the lanes re-encode
//            their own flagged units instead, so the worst case is the per-pixel kernel plus the flag test
//
// Both phases are exact: a group is taken from its anchor only when all 12 bytes were compared equal.
//
// Class counting rides on the same structure (GroupCounter): an unflagged group counts its anchor class
// with weight 4 -- one table look-up for four pixels -- and the lane that re-encodes a flagged group
// counts its four true classes with weight 1.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace pylc {

// 12 bytes = 4 interleaved RGB pixels.  Zero iff all four pixels are equal.
__device__ __forceinline__ uint32_t group_spread(uint32_t a, uint32_t b, uint32_t c) {
    // with P(x) = bytes (x1, x2, x0, x1): one repeated pixel <=> b == P(a), c == P(b), a == P(c)
    return (b ^ __byte_perm(a, 0, 0x1021)) | (c ^ __byte_perm(b, 0, 0x1021)) | (a ^ __byte_perm(c, 0, 0x1021));
}

// table entry of a colour from a table at a 32-bit shared-memory address (no generic pointer in the loop)
__device__ __forceinline__ uint32_t lookup_entry_s(uint32_t key, uint32_t tab, uint32_t mul, uint32_t miss_e) {
    const uint32_t e = lds32(tab + __byte_perm(key * mul, 0, 0x4442) * 4u);
    return ((e ^ key) & 0x00FFFFFFu) ? miss_e : e;
}
// per-pixel encode of one group: four class bytes packed into a word
__device__ __forceinline__ uint32_t encode_group_s(uint32_t a, uint32_t b, uint32_t c, uint32_t tab, uint32_t mul, uint32_t miss_e) {
    const uint32_t e0 = lookup_entry_s(a, tab, mul, miss_e);
    const uint32_t e1 = lookup_entry_s(__funnelshift_r(a, b, 24), tab, mul, miss_e);
    const uint32_t e2 = lookup_entry_s(__funnelshift_r(b, c, 16), tab, mul, miss_e);
    const uint32_t e3 = lookup_entry_s(c >> 8, tab, mul, miss_e);
    return pack_top_bytes(e0, e1, e2, e3);
}

// ---- weighted class counters ------------------------------------------------------------------
// NibbleCounter's PRMT-as-table trick (common.cuh) with two weights sharing the accumulators: anchors of
// unflagged groups add 4 per pixel-of-the-table (0x04 / 0x40 table bytes), re-encoded pixels add 1.
// Ids must be < 16; id 15 (kVoid) is never counted; C <= 2 NG <= 14.
// A nibble field grows by at most 8 (two anchor words) + 3 (re-encoded words that land on one byte lane)
// per item, so widen() runs after every item and the byte fields hold kFlushItems = 16 items.
template <int NG>
struct GroupCounter {
    static_assert(NG >= 1 && NG <= 7, "ids 14 and 15 must stay free");
    static constexpr int kFlushItems = 16, kMaxClasses = 2 * NG;
    uint32_t nib[NG], wide[NG][2];
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int g = 0; g < NG; ++g) nib[g] = wide[g][0] = wide[g][1] = 0u;
    }
    template <int W>
    static __device__ __forceinline__ uint32_t look(int g, uint32_t sel) {
        constexpr uint32_t lo = W == 4 ? 0x00004004u : 0x00001001u, hi = W == 4 ? 0x40040000u : 0x10010000u;
        const uint32_t a = (g & 3) == 0 ? lo : ((g & 3) == 1 ? hi : 0u);
        const uint32_t b = (g & 3) == 2 ? lo : ((g & 3) == 3 ? hi : 0u);
        uint32_t d;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
        return d;
    }
    // two words of four anchor class bytes each (one per unflagged group; kVoid for flagged ones): weight 4
    __device__ __forceinline__ void add_anchors(uint32_t aw0, uint32_t aw1) {
        const uint32_t p = aw0 + (aw1 << 4), q = p >> 16;
#pragma unroll
        for (int g = 0; g < (NG < 4 ? NG : 4); ++g) nib[g] += look<4>(g, p) + look<4>(g, q);
        if (NG > 4) {
            const uint32_t x = p ^ 0x88888888u, y = x >> 16;
#pragma unroll
            for (int g = 4; g < NG; ++g) nib[g] += look<4>(g, x) + look<4>(g, y);
        }
    }
    // one word of four class bytes: weight 1
    __device__ __forceinline__ void add_pixels(uint32_t w) {
        const uint32_t t = w | (w >> 4);
        const uint32_t p = __byte_perm(t, 0, 0x4420);       // nibbles id0, id1, id2, id3
#pragma unroll
        for (int g = 0; g < (NG < 4 ? NG : 4); ++g) nib[g] += look<1>(g, p);
        if (NG > 4) {
            const uint32_t x = p ^ 0x8888u;
#pragma unroll
            for (int g = 4; g < NG; ++g) nib[g] += look<1>(g, x);
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
    __device__ __forceinline__ uint32_t count(int c) const { return __dp4a(wide[c >> 1][c & 1], 0x01010101u, 0u); }
};

constexpr uint32_t kVoidId = 0x0Fu;

// ---- the warp encode ----------------------------------------------------------------------------
// NU units per lane; unit j of lane l lives at in_warp + j * in_stride + 48 l (48 bytes) and its 16 output
// bytes go to out_warp + j * out_stride + 16 l.  Group ids are j * 128 + 4 l + k (k = group in the unit).
// `tab` holds key | byte << 24 per slot; the byte is what gets emitted (a class id, or class * C for the
// confusion kernel).  COUNT: feed `gc` (GroupCounter) -- anchors with weight 4, re-encoded pixels with 1;
// requires the emitted bytes to be class ids < 15.
// `unit_ok`: bit j set = unit j of this lane lies inside the image (edge boxes of a TMA load are zero- or
// padding-filled outside it): the bytes of an outside unit are still written (a clipped store drops them)
// but never counted and never sent to the fix-up.
// Must be called by all 32 lanes.  On return the output box holds the exact encode of all 32 NU units
// once the warp's shared-memory writes are made visible (__syncwarp / __syncthreads by the caller).
template <int NU, bool COUNT, class GC>
__device__ __forceinline__ void warp_encode_units(uint32_t in_warp, uint32_t in_stride, uint32_t out_warp, uint32_t out_stride,
                                                  uint32_t q_warp, uint32_t tab, uint32_t mul, uint32_t miss_e, GC &gc,
                                                  uint32_t unit_ok = 3u) {
    static_assert(NU == 1 || NU == 2, "group ids are one byte");
    constexpr int kDenseGroups = NU * 48;            // of NU * 128 groups per warp
    const uint32_t lane = threadIdx.x & 31;
    uint32_t flags = 0, aw[2] = {0x0F0F0F0Fu, 0x0F0F0F0Fu};
#pragma unroll
    for (int j = 0; j < NU; ++j) {
        const uint32_t in_lane = in_warp + (uint32_t)j * in_stride + lane * 48u;
        const uint4 q0 = lds128(in_lane), q1 = lds128(in_lane + 16), q2 = lds128(in_lane + 32);
        const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
        uint32_t ow[4], e[4];
        const bool ok = (unit_ok >> j) & 1u;
        // two 4-pixel groups share ONE look-up: both are uniform and their first words agree <=> the 8 pixels are
        // one colour; otherwise BOTH groups go to the fix-up (the exact per-pixel path decides each of them)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t a = w[6 * h], b = w[6 * h + 1], c = w[6 * h + 2], d = w[6 * h + 3], f = w[6 * h + 4], g = w[6 * h + 5];
            const uint32_t ent = lookup_entry_s(a, tab, mul, miss_e);
            ow[2 * h] = ow[2 * h + 1] = __byte_perm(ent, 0, 0x3333);
            const bool mixed = ok && (group_spread(a, b, c) | group_spread(d, f, g) | (a ^ d)) != 0;
            flags |= mixed ? (3u << (4 * j + 2 * h)) : 0u;
            if (COUNT) e[2 * h] = e[2 * h + 1] = (mixed || !ok) ? (kVoidId << 24) : ent;
        }
        sts128(out_warp + (uint32_t)j * out_stride + lane * 16u, make_uint4(ow[0], ow[1], ow[2], ow[3]));
        if (COUNT) aw[j] = pack_top_bytes(e[0], e[1], e[2], e[3]);
    }
    if (__any_sync(0xFFFFFFFFu, flags != 0)) {
        // exclusive prefix of the per-lane flagged-group counts
        const int mine = __popc(flags);
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((int)lane >= d) incl += v;
        }
        const int n = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (n > kDenseGroups) {
#pragma unroll
            for (int j = 0; j < NU; ++j) {
                if ((flags >> (4 * j)) & 15u) {
                    const uint32_t in_lane = in_warp + (uint32_t)j * in_stride + lane * 48u;
                    const uint4 q0 = lds128(in_lane), q1 = lds128(in_lane + 16), q2 = lds128(in_lane + 32);
                    const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
                    uint32_t ow[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        ow[k] = encode_group_s(w[3 * k], w[3 * k + 1], w[3 * k + 2], tab, mul, miss_e);
                        if (COUNT) gc.add_pixels(ow[k]);
                    }
                    sts128(out_warp + (uint32_t)j * out_stride + lane * 16u, make_uint4(ow[0], ow[1], ow[2], ow[3]));
                    if (COUNT) aw[j] = 0x0F0F0F0Fu;      // the whole unit was counted pixel by pixel
                }
            }
        } else {
            uint32_t f = flags, pos = q_warp + (uint32_t)(incl - mine);
            while (f) {
                const uint32_t bit = (uint32_t)__ffs(f) - 1u;
                f &= f - 1u;
                asm volatile("st.shared.u8 [%0], %1;" ::"r"(pos), "r"(lane * 4u + (bit & 3u) + ((bit & 4u) << 5)) : "memory");
                ++pos;
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                uint32_t gid;
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(gid) : "r"(q_warp + i) : "memory");
                const uint32_t u = gid >> 7, g = gid & 127u;
                const uint32_t ia = in_warp + u * in_stride + g * 12u;
                const uint32_t a = lds32(ia), b = lds32(ia + 4), c = lds32(ia + 8);
                const uint32_t word = encode_group_s(a, b, c, tab, mul, miss_e);
                sts32(out_warp + u * out_stride + g * 4u, word);
                if (COUNT) gc.add_pixels(word);
            }
        }
    }
    if (COUNT) {
        gc.add_anchors(aw[0], aw[1]);
        gc.widen();
    }
}

// placeholder counter for COUNT = false
struct NoCounter {
    __device__ __forceinline__ void add_pixels(uint32_t) {}
    __device__ __forceinline__ void add_anchors(uint32_t, uint32_t) {}
    __device__ __forceinline__ void widen() {}
};

}  // namespace pylc
