// Augmentation copies of tiles on the device: tools.augment_transform (reference utils/tools.py:452-594) --
// perspective jitter (cv2.warpPerspective, bilinear image / nearest mask, reflect-101 border), 30-pixel crop,
// resize back to the tile size (cv2.resize INTER_AREA / INTER_NEAREST) and the brightness shift -- for every
// over-sampled copy Augmentor.oversample appends (utils/augment.py:184-239), bit-identical to the OpenCV call
// chain.  The arithmetic is augment_math.cuh (one function, also compiled for the host and held against OpenCV
// and the reference's golden vectors by tests that need no GPU); this file only maps pixels to threads.
//
// One thread = one output pixel of one copy, all channels and the mask: the four warped samples the enlarging
// resize blends share their (double precision) source coordinates across channels; source tiles are read through
// L1 / L2 (every source pixel is used by ~1.3 output pixels; a 512 x 512 gray tile is 256 KB).  The random draws
// (control-point jitter -> cv2.getPerspectiveTransform -> cv2.invert, brightness shift) stay on the host: 17 numbers
// per copy, uploaded as the job table.
#include "augment_math.cuh"
#include "common.cuh"

namespace pylc {

constexpr int kAugThreads = 128;

__global__ void __launch_bounds__(kAugThreads)
    augment_tiles_kernel(const uint8_t *__restrict__ src_img, const uint8_t *__restrict__ src_mask, int n_src, int ch, int T,
                         const int32_t *__restrict__ job_src, const double *__restrict__ job_minv, const int32_t *__restrict__ job_shift,
                         int job0, uint8_t *__restrict__ dst_img, uint8_t *__restrict__ dst_mask) {
    const int dx = blockIdx.x * kAugThreads + threadIdx.x, dy = blockIdx.y, j = job0 + blockIdx.z;
    if (dx >= T) return;
    const int s = __ldg(job_src + j);
    if (s < 0 || s >= n_src) return;                 // rejected on the host as well; never dereference a bad index
    double M[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) M[k] = __ldg(job_minv + (size_t)j * 9 + k);
    const size_t TT = (size_t)T * T;
    pylc_aug::augment_pixel(src_img + (size_t)s * ch * TT, src_mask + (size_t)s * TT, ch, T, M, __ldg(job_shift + j), dx, dy,
                            dst_img + (size_t)j * ch * TT, dst_mask + (size_t)j * TT);
}

}  // namespace pylc

using namespace pylc;

extern "C" int pylc_augment_tiles_u8(const uint8_t *src_img, const uint8_t *src_mask, int n_src, int ch, int T, const int32_t *job_src,
                                     const double *job_minv, const int32_t *job_shift, int n_jobs, uint8_t *dst_img, uint8_t *dst_mask,
                                     pylc_stream_t stream) {
    if (ch != 1 && ch != 3) return PYLC_ERR_ARG;
    if (T <= 2 * pylc_aug::kCrop + 1 || T > 32768 || n_src < 1 || n_jobs < 0) return PYLC_ERR_GEOMETRY;
    if (n_jobs == 0) return PYLC_OK;
    if (!src_img || !src_mask || !job_src || !job_minv || !job_shift || !dst_img || !dst_mask) return PYLC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
        const int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
        const dim3 grid((unsigned)((T + kAugThreads - 1) / kAugThreads), (unsigned)T, (unsigned)nj);
        augment_tiles_kernel<<<grid, kAugThreads, 0, st>>>(src_img, src_mask, n_src, ch, T, job_src, job_minv, job_shift, j0, dst_img,
                                                          dst_mask);
        const int rc = finish_launch();
        if (rc) return rc;
    }
    return PYLC_OK;
}
