// Library-level entry points of the pylc_b200 C ABI.
#include "common.cuh"
#include "tma.cuh"

#include <mutex>

using namespace pylc;

namespace pylc {

// cuTensorMapEncodeTiled through the runtime's driver entry-point query: the library keeps linking only
// the (static) runtime, and a driver without the symbol makes the TMA kernels fall back, not fail.
typedef CUresult(CUDAAPI *EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                         const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult status = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &status) == cudaSuccess &&
            status == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    });
    return fn;
}

bool tma_encode_u32(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                    const uint32_t *box) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || rank < 1 || rank > 5) return false;
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bdim, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace pylc

extern "C" int pylc_abi_version(void) { return PYLC_ABI_VERSION; }

extern "C" const char *pylc_error_string(int code) {
    switch (code) {
        case PYLC_OK: return "ok";
        case PYLC_ERR_ARG: return "pylc: invalid argument (null pointer or non-positive size)";
        case PYLC_ERR_CLASSES: return "pylc: n_classes outside [1, 32]";
        case PYLC_ERR_GEOMETRY: return "pylc: unsupported tile/stride geometry";
        case PYLC_ERR_ALIGN: return "pylc: pointer not sufficiently aligned";
        case PYLC_ERR_PALETTE: return "pylc: could not build a collision-free palette hash";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "pylc: unknown error";
}

extern "C" int pylc_device_info(int *sm_count, int *compute_capability) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    int sms = 0, major = 0, minor = 0;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev)) != cudaSuccess) return (int)e;
    if (sm_count) *sm_count = sms;
    if (compute_capability) *compute_capability = major * 10 + minor;
    return PYLC_OK;
}

extern "C" int64_t pylc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
