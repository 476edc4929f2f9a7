// class_encode on interleaved RGB rows, TMA form (sm_100a): tools.class_encode (reference
// utils/tools.py:412-449; called on whole masks by Evaluator.load, utils/evaluate.py:103-108).
//
// Same machinery as the TMA mask gather (gather_tma.cu): 32-row x 256-pixel boxes of the [rows][pitch/4]
// source tensor land in shared memory through cp.async.bulk.tensor loads signalled on mbarriers, a warp
// encodes its 64 units with warp_encode_units (one look-up per uniform 4-pixel group, mixed groups compacted
// and re-encoded), and the encoded box leaves through one cp.async.bulk.tensor store into the [rows][cols/4]
// label tensor.  Boxes on the right / bottom edge are zero-filled by the load and clipped by the store; their
// outside units are neither fixed up nor counted.  The optional class histogram is one [C] vector for the
// whole call: GroupCounter registers per lane, one shared-memory reduction and C global atomics per CTA.
#include "palette_warp.cuh"

#include <type_traits>

namespace pylc {

namespace {
constexpr int kBoxPx = 256, kBoxRows = 32;
constexpr int kUnitIn = 16 * kBoxPx * 3, kUnitOut = 16 * kBoxPx;
constexpr int kBoxIn = kBoxRows * kBoxPx * 3, kBoxOut = kBoxRows * kBoxPx;
constexpr int kStages = 2, kOutBufs = 2;
constexpr int kCtasPerSm = 3;
}  // namespace

struct EncGeom {
    int rows, cols;      // pixels
    int nbx, nby, items;
};

template <int NG, bool HIST>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
    class_encode_tma_kernel(const __grid_constant__ CUtensorMap tm_src, const __grid_constant__ CUtensorMap tm_dst, const EncGeom g,
                            const __grid_constant__ PaletteHash ph, int C, long long *__restrict__ hist) {
    extern __shared__ __align__(128) uint8_t s_dyn[];
    __shared__ uint32_t s_tab[256];
    __shared__ unsigned s_hist[PYLC_MAX_CLASSES];
    __shared__ __align__(8) unsigned long long s_bar[kStages];
    __shared__ __align__(16) uint8_t s_queue[kWarps][256];

    const int tid = threadIdx.x, warp = tid >> 5;
    const int first = (int)((long long)g.items * blockIdx.x / gridDim.x);
    const int n = (int)((long long)g.items * (blockIdx.x + 1) / gridDim.x) - first;
    if (n <= 0) return;

    const uint32_t in0 = smem_u32(s_dyn), out0 = in0 + kStages * kBoxIn, bar0 = smem_u32(s_bar);
    // boxes are walked row-major; box `item` covers pixels [bx*256, +256) x rows [by*32, +32)
    auto issue_load = [&](int item, int stage) {
        const int by = item / g.nbx, bx = item - by * g.nbx;
        const uint32_t bar = bar0 + 8u * stage;
        mbar_arrive_expect_tx(bar, kBoxIn);       // out-of-bounds parts of an edge box are zero-filled and still counted
        tma_load_2d(in0 + (uint32_t)stage * kBoxIn, &tm_src, bx * (kBoxPx * 3 / 4), by * kBoxRows, bar);
    };
    // thread 0 puts the first boxes in flight before the table copy and the barrier (see gather_tma.cu)
    if (tid == 0) {
        tma_prefetch_desc(&tm_src);
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8u * s, 1);
        mbar_fence_init();
        for (int k = 0; k < kStages && k < n; ++k) issue_load(first + k, k);
        tma_prefetch_desc(&tm_dst);
    }
    s_tab[tid] = ph.tab[tid];
    if (tid < PYLC_MAX_CLASSES) s_hist[tid] = 0;
    __syncthreads();

    const uint32_t mul = ph.mul, tab = smem_u32(s_tab);
    const uint32_t miss_e = 1u << 24;             // unmatched colours are class 1 (utils/tools.py:437)
    const uint32_t q_warp = smem_u32(&s_queue[warp][0]);
    const uint32_t in_warp0 = in0 + (uint32_t)warp * (32 * 48), out_warp0 = out0 + (uint32_t)warp * (32 * 16);
    constexpr bool GROUPS = HIST && NG > 0;
    using GC = typename std::conditional<GROUPS, GroupCounter<(NG > 0 ? NG : 1)>, NoCounter>::type;
    GC gc;
    ByteCounter bc;
    if constexpr (GROUPS) gc.reset();
    if (HIST && NG == 0) bc.reset();
    int since_flush = 0;
    const int ux = (tid & 15) * 16, uy = tid >> 4;       // the lane's unit inside a 16-row half box

    int by = first / g.nbx, bx = first - by * g.nbx;
    int stage = 0, obuf = 0;
    uint32_t parity = 0;
    for (int k = 0; k < n; ++k) {
        mbar_wait(bar0 + 8u * stage, parity);
        const uint32_t in_s = (uint32_t)stage * kBoxIn, out_s = (uint32_t)obuf * kBoxOut;
        const bool x_ok = bx * kBoxPx + ux < g.cols;
        const uint32_t ok = (x_ok && by * kBoxRows + uy < g.rows ? 1u : 0u) | (x_ok && by * kBoxRows + 16 + uy < g.rows ? 2u : 0u);
        warp_encode_units<2, GROUPS>(in_warp0 + in_s, kUnitIn, out_warp0 + out_s, kUnitOut, q_warp, tab, mul, miss_e, gc, ok);
        const bool last = k + 1 == n;
        if (HIST && NG == 0) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint4 r = lds128(out0 + out_s + (uint32_t)j * kUnitOut + (uint32_t)tid * 16u);
                if ((ok >> j) & 1u) bc.add16(r.x, r.y, r.z, r.w);
            }
            if (last || ++since_flush >= ByteCounter::kFlushUnits / 2) {
                flush_counter(bc, C, s_hist);
                bc.reset();
                since_flush = 0;
            }
        }
        if constexpr (GROUPS) {
            if (last || ++since_flush >= GC::kFlushItems) {
                flush_counter(gc, C, s_hist);
                gc.reset();
                since_flush = 0;
            }
        }
        if (tid == 0) tma_store_wait_read<0>();
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tma_store_2d(&tm_dst, bx * (kBoxPx / 4), by * kBoxRows, out0 + out_s);
            tma_store_commit();
            if (k + kStages < n) issue_load(first + k + kStages, stage);
        }
        if (++bx == g.nbx) {
            bx = 0;
            ++by;
        }
        if (++stage == kStages) {
            stage = 0;
            parity ^= 1u;
        }
        obuf ^= 1;
    }
    if (tid == 0) tma_store_wait_read<0>();
    if (HIST) {
        __syncthreads();
        if (tid < C && s_hist[tid]) atomicAdd((unsigned long long *)&hist[tid], (unsigned long long)s_hist[tid]);
    }
}

// Returns PYLC_OK / a CUDA error after launching, or -1 when the form does not apply: needs 16-byte aligned
// source rows and an output whose rows are whole 16-pixel units (cols % 16 == 0, 16-byte aligned).
int launch_class_encode_tma(const uint8_t *rgb, long long rows, long long cols, size_t pitch, const PaletteHash &ph, int C,
                            uint8_t *out, long long *hist, cudaStream_t st) {
    if (((uintptr_t)rgb % 16) || (pitch % 16) || ((uintptr_t)out % 16) || cols % 16 || cols < 16 || rows < 1) return -1;
    if (rows > 0x7FFFFFF || cols > 0x7FFFFFF) return -1;
    EncGeom g;
    g.rows = (int)rows, g.cols = (int)cols;
    g.nbx = (int)((cols + kBoxPx - 1) / kBoxPx), g.nby = (int)((rows + kBoxRows - 1) / kBoxRows);
    const long long items = (long long)g.nbx * g.nby;
    if (items > 0x7FFFFFFF) return -1;
    g.items = (int)items;
    CUtensorMap tm_src, tm_dst;
    {
        const uint64_t dims[2] = {(uint64_t)(pitch / 4), (uint64_t)rows};
        const uint64_t strides[1] = {(uint64_t)pitch};
        const uint32_t box[2] = {kBoxPx * 3 / 4, kBoxRows};
        if (!tma_encode_u32(&tm_src, rgb, 2, dims, strides, box)) return -1;
    }
    {
        const uint64_t dims[2] = {(uint64_t)(cols / 4), (uint64_t)rows};
        const uint64_t strides[1] = {(uint64_t)cols};
        const uint32_t box[2] = {kBoxPx / 4, kBoxRows};
        if (!tma_encode_u32(&tm_dst, out, 2, dims, strides, box)) return -1;
    }
    const size_t smem = (size_t)kStages * kBoxIn + (size_t)kOutBufs * kBoxOut;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define LAUNCH(NG, HS)                                                                                         \
    do {                                                                                                       \
        auto kern = class_encode_tma_kernel<NG, HS>;                                                           \
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { \
            cudaGetLastError();                                                                                \
            return -1;                                                                                         \
        }                                                                                                      \
        int per_sm = 0;                                                                                        \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1; \
        if (per_sm > kCtasPerSm) per_sm = kCtasPerSm;                                                          \
        long long ctas = (long long)sms * per_sm;                                                              \
        if (ctas > items) ctas = items;                                                                        \
        kern<<<(unsigned)ctas, kThreads, smem, st>>>(tm_src, tm_dst, g, ph, C, hist);                          \
    } while (0)
    switch (hist ? counter_groups(C) : -1) {
        case -1: LAUNCH(5, false); break;
        case 5: LAUNCH(5, true); break;
        case 6: LAUNCH(6, true); break;
        case 7: LAUNCH(7, true); break;
        default: LAUNCH(0, true); break;
    }
#undef LAUNCH
    return finish_launch();
}

}  // namespace pylc
