// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product.
// Host twin of the augmentation kernel: the SAME per-pixel function the sm_100a kernel calls
// (pylc_b200/csrc/augment_math.cuh, compiled here by the host C++ compiler) looped over a tile on the CPU, so
// that tests without a GPU can hold the kernel's arithmetic against OpenCV and the reference's golden vectors.
#include "../pylc_b200/csrc/augment_math.cuh"

extern "C" int pylc_oracle_augment_tiles_host(const uint8_t *src_img, const uint8_t *src_mask, int n_src, int ch, int T,
                                              const int32_t *job_src, const double *job_minv, const int32_t *job_shift, int n_jobs,
                                              uint8_t *dst_img, uint8_t *dst_mask) {
    if (T <= 2 * pylc_aug::kCrop + 1 || (ch != 1 && ch != 3)) return -1;
    const size_t TT = (size_t)T * T;
    for (int j = 0; j < n_jobs; ++j) {
        const int s = job_src[j];
        if (s < 0 || s >= n_src) return -2;
        for (int dy = 0; dy < T; ++dy)
            for (int dx = 0; dx < T; ++dx)
                pylc_aug::augment_pixel(src_img + (size_t)s * ch * TT, src_mask + (size_t)s * TT, ch, T, job_minv + (size_t)j * 9,
                                        job_shift[j], dx, dy, dst_img + (size_t)j * ch * TT, dst_mask + (size_t)j * TT);
    }
    return 0;
}
