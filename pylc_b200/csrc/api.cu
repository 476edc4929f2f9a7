// Library-level entry points of the pylc_b200 C ABI.
#include "common.cuh"

using namespace pylc;

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
