// Non-convolution glue of the DeepLabv3+ inference plan (models/architectures/deeplab.py:38,
// models/decoder.py:46-48, models/backbone/resnet.py max-pool of the stem): the three data-movement
// steps between the library convolutions that the stock PyTorch kernels run at 10-30 % of HBM speed on
// channels-last tensors.  All three are pure HBM streams (every output byte written once, every input
// byte fetched from DRAM once) and keep PyTorch's arithmetic:
//     bilinear, align_corners=True:  src = dst * (in-1)/(out-1);  i0 = (int)src;  l1 = src - i0;  l0 = 1 - l1
//     out = h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11)              (upsample_bilinear2d, UpSample.cuh)
// so results agree with the eager ops to float rounding (tests: 1e-6 relative; max-pool exact).
#include "common.cuh"

namespace pylc {

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 lerp4(float a0, const float4 &u, float a1, const float4 &v) {
    return make_float4(a0 * u.x + a1 * v.x, a0 * u.y + a1 * v.y, a0 * u.z + a1 * v.z, a0 * u.w + a1 * v.w);
}
__device__ __forceinline__ float4 max4(const float4 &a, const float4 &b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// ------------------------------------------------------------------------------------------------
// decoder: bilinear up-sample of the ASPP output to the low-level feature size + channel concat
//   x [B, h, w, Cx] NHWC, low [B, H, W, Cl] NHWC  ->  out [B, H, W, Cx + Cl] NHWC   (decoder.py:46-48)
// ------------------------------------------------------------------------------------------------
// General form (any geometry; the pipeline's shapes take the staged kernel below).
// A warp takes 16-pixel segments of output rows (persistent grid, segments dealt round-robin so the
// load balances to a fraction of a segment).  Per segment it first issues all loads of the low-level
// channels it has to copy (independent, so they overlap), then walks X left to right: a lane owns VPL
// float4 channel groups and keeps the horizontally adjacent source taps of both source rows in
// registers, re-loading a tap only when the source column advances (every ~(W-1)/(w-1) output
// pixels).  Each step the warp writes one contiguous Cx*4-byte run; the low-level channels follow.
constexpr int kUpSeg = 16;      // output pixels per work item
constexpr int kUpLowMax = 8;    // float4 per lane of low-level data per segment: Cl <= 64

template <int VPL>
__global__ void __launch_bounds__(kThreads) upsample_concat_kernel(const float *__restrict__ x, const float *__restrict__ low,
                                                                  float *__restrict__ out, int B, int h, int w, int Cx, int H,
                                                                  int W, int Cl) {
    const int lane = threadIdx.x & 31;
    const int segs = (W + kUpSeg - 1) / kUpSeg;
    const long long items = (long long)B * H * segs;
    const long long nwarps = (long long)gridDim.x * kWarps;
    const float rh = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, rw = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const int Co = Cx + Cl, lq = Cl / 4;                    // float4 per pixel of low-level channels
    for (long long item = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5); item < items; item += nwarps) {
        const int seg = (int)(item % segs);
        const long long row = item / segs;
        const int b = (int)(row / H), Y = (int)(row - (long long)b * H);
        const int X0 = seg * kUpSeg, X1 = min(W, X0 + kUpSeg);
        const float h1r = rh * Y;
        const int y1 = (int)h1r, y1p = y1 < h - 1 ? 1 : 0;
        const float hl1 = h1r - y1, hl0 = 1.f - hl1;
        const float *r0 = x + ((size_t)b * h + y1) * w * Cx, *r1 = r0 + (size_t)y1p * w * Cx;
        const float *lrow = low + (((size_t)b * H + Y) * W + X0) * Cl;
        float *orow = out + ((size_t)b * H + Y) * W * Co;

        // low-level channels of the segment: (X1 - X0) * lq float4, lane-strided, all loads in flight at once
        const int nlow = (X1 - X0) * lq;
        float4 lv[kUpLowMax];
#pragma unroll
        for (int k = 0; k < kUpLowMax; ++k)
            if (k * 32 + lane < nlow) lv[k] = ld_stream_f4(lrow + (size_t)(k * 32 + lane) * 4);

        float4 t00[VPL], t01[VPL], t10[VPL], t11[VPL];   // taps (row 0/1, column x1 / x1+1) of this lane's channels
        int x1_cur = -2;
        for (int X = X0; X < X1; ++X) {
            const float w1r = rw * X;
            const int x1 = (int)w1r, x1p = x1 < w - 1 ? 1 : 0;
            const float wl1 = w1r - x1, wl0 = 1.f - wl1;
            if (x1 != x1_cur) {
                const bool shift = x1 == x1_cur + 1;          // the old right tap becomes the left tap
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    const int c = (v * 32 + lane) * 4;
                    if (c < Cx) {
                        if (shift) {
                            t00[v] = t01[v];
                            t10[v] = t11[v];
                        } else {
                            t00[v] = ldg4(r0 + (size_t)x1 * Cx + c);
                            t10[v] = ldg4(r1 + (size_t)x1 * Cx + c);
                        }
                        t01[v] = ldg4(r0 + (size_t)(x1 + x1p) * Cx + c);
                        t11[v] = ldg4(r1 + (size_t)(x1 + x1p) * Cx + c);
                    }
                }
                x1_cur = x1;
            }
            float *o = orow + (size_t)X * Co;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c = (v * 32 + lane) * 4;
                if (c < Cx) {
                    const float4 top = lerp4(wl0, t00[v], wl1, t01[v]), bot = lerp4(wl0, t10[v], wl1, t11[v]);
                    st_stream_f4(o + c, lerp4(hl0, top, hl1, bot));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kUpLowMax; ++k) {
            const int idx = k * 32 + lane;
            if (idx < nlow) {
                const int px = idx / lq, q = idx - px * lq;
                st_stream_f4(orow + (size_t)(X0 + px) * Co + Cx + q * 4, lv[k]);
            }
        }
    }
}

// Staged form: one CTA per output row.  Every pixel of an output row interpolates between the same two
// source rows, so the CTA blends those two rows ONCE into shared memory (w x Cx floats: 32 KB for the
// decoder's 32 x 256) and an output pixel is one horizontal blend of two shared-memory columns: 2 LDS.128
// + 8 FP instructions + 1 STG.128 per float4, against ~30 in the warp-per-segment kernel above, which is
// issue-bound at 64 % of the roofline with 118 registers.  Many small CTAs (six resident per SM), lanes
// walk the channel groups of a pixel, so loads, shared-memory accesses and stores are all contiguous.
// (Rows are blended before columns: results differ from h0*(w0*a + w1*b) + h1*(...) by an ulp or two.)
__global__ void __launch_bounds__(kThreads) upsample_concat_staged_kernel(const float *__restrict__ x, const float *__restrict__ low,
                                                                         float *__restrict__ out, int B, int h, int w, int Cx, int H,
                                                                         int W, int Cl) {
    extern __shared__ __align__(16) float4 s_v[];            // [w][Cx/4] row-blended source, then x1[W], wl1[W]
    const int cq = Cx / 4, lq = Cl / 4, Co4 = (Cx + Cl) / 4;
    int *s_x1 = reinterpret_cast<int *>(s_v + (size_t)w * cq);
    float *s_w1 = reinterpret_cast<float *>(s_x1 + W);
    const int b = blockIdx.x / H, Y = blockIdx.x - b * H;
    const float rh = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, rw = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const float h1r = rh * Y;
    const int y1 = (int)h1r, y1p = y1 < h - 1 ? 1 : 0;
    const float hl1 = h1r - y1, hl0 = 1.f - hl1;
    const float4 *r0 = reinterpret_cast<const float4 *>(x + ((size_t)b * h + y1) * w * Cx);
    const float4 *r1 = r0 + (size_t)y1p * w * cq;
    for (int i = threadIdx.x; i < w * cq; i += kThreads) s_v[i] = lerp4(hl0, __ldg(r0 + i), hl1, __ldg(r1 + i));
    for (int X = threadIdx.x; X < W; X += kThreads) {
        const float w1r = rw * X;
        const int x1 = (int)w1r;
        s_x1[X] = x1 * cq + ((x1 < w - 1 ? cq : 0) << 16);   // left column offset | right-column step << 16 (w * cq < 65536)
        s_w1[X] = w1r - x1;
    }
    __syncthreads();

    float4 *orow = reinterpret_cast<float4 *>(out) + ((size_t)b * H + Y) * W * Co4;
    {   // up-sampled channels: thread walks (X, g) with constant steps, no divisions in the loop
        int X = threadIdx.x / cq, g = threadIdx.x - X * cq;
        const int dX = kThreads / cq, dg = kThreads - dX * cq;
        for (; X < W; X += dX, g += dg) {
            if (g >= cq) {
                g -= cq;
                if (++X >= W) break;
            }
            const int pk = s_x1[X];
            const float wl1 = s_w1[X], wl0 = 1.f - wl1;
            const float4 *c0 = s_v + (pk & 0xFFFF) + g;
            st_stream_f4(reinterpret_cast<float *>(orow + (size_t)X * Co4 + g), lerp4(wl0, c0[0], wl1, c0[pk >> 16]));
        }
    }
    if (lq > 0) {   // low-level channels: a contiguous row of W * lq float4 copied behind each pixel's up-sampled channels
        const float *lrow = low + ((size_t)b * H + Y) * W * Cl;
        int X = threadIdx.x / lq, q = threadIdx.x - X * lq;
        const int dX = kThreads / lq, dq = kThreads - dX * lq;
        for (int idx = threadIdx.x; idx < W * lq; idx += kThreads, X += dX, q += dq) {
            if (q >= lq) {
                q -= lq;
                ++X;
            }
            st_stream_f4(reinterpret_cast<float *>(orow + (size_t)X * Co4 + cq + q), ld_stream_f4(lrow + (size_t)idx * 4));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// stem: max_pool2d(kernel 3, stride 2, padding 1) on NHWC                      (backbone/resnet.py)
//   in [B, H, W, C] -> out [B, Ho, Wo, C],  Ho = (H + 2 - 3)/2 + 1
// ------------------------------------------------------------------------------------------------
// One thread = one float4 channel group of TWO horizontally adjacent output pixels (they share an input
// column): 15 loads for 2 outputs.  Consecutive threads walk the channel groups of a pixel pair, so a
// warp reads and writes contiguous runs.  Rows shared by consecutive output rows come from L2.
// One CTA per output row; a thread's (pixel pair, channel group) position advances by constants, so the
// loop carries no divisions (the earlier grid-stride form spent most of its instructions on 64-bit index
// arithmetic).  Column maxima over the three input rows are formed first, then combined per output.
__global__ void __launch_bounds__(kThreads) maxpool3x3s2_kernel(const float *__restrict__ in, float *__restrict__ out, int B, int H,
                                                               int W, int C, int Ho, int Wo) {
    const int cg = C / 4, pairs = (Wo + 1) / 2;
    const int b = blockIdx.x / Ho, Y = blockIdx.x - b * Ho;
    const float ninf = -INFINITY;
    const float4 lowest = make_float4(ninf, ninf, ninf, ninf);
    const float *rows[3];
    bool have[3];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int y = 2 * Y - 1 + dy;
        have[dy] = y >= 0 && y < H;
        rows[dy] = in + (((size_t)b * H + (have[dy] ? y : 0)) * W) * C;
    }
    float *orow = out + (((size_t)b * Ho + Y) * Wo) * C;
    int px = threadIdx.x / cg, g = threadIdx.x - px * cg;
    const int dpx = kThreads / cg, dg = kThreads - dpx * cg;
    for (; px < pairs; px += dpx, g += dg) {
        if (g >= cg) {
            g -= cg;
            if (++px >= pairs) break;
        }
        const int X0 = px * 2;
        float4 col[5];
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
            const int xx = 2 * X0 - 1 + dx;
            col[dx] = lowest;
            if (xx >= 0 && xx < W) {
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
                    if (have[dy]) col[dx] = max4(col[dx], ldg4(rows[dy] + (size_t)xx * C + g * 4));
            }
        }
        float *o = orow + (size_t)X0 * C + g * 4;
        st_stream_f4(o, max4(max4(col[0], col[1]), col[2]));
        if (X0 + 1 < Wo) st_stream_f4(o + C, max4(max4(col[2], col[3]), col[4]));
    }
}

// ------------------------------------------------------------------------------------------------
// head: final bilinear up-sample of the decoder output to tile size, NHWC in -> NCHW logits out
//   in [B, h, w, C] NHWC  ->  out [B, C, H, W] planar (what the stitch kernel reads)   (deeplab.py:38)
// ------------------------------------------------------------------------------------------------
// General form (any geometry; up-sampling shapes take the staged kernel below).
// One thread = 4 consecutive output pixels of one row, all classes: one 16-byte store per class plane
// (a warp writes 512 contiguous bytes per plane).  The <= 26 MB input stays in L1/L2.
__global__ void __launch_bounds__(kThreads) upsample_to_nchw_kernel(const float *__restrict__ in, float *__restrict__ out, int B, int h,
                                                                   int w, int C, int H, int W) {
    const int gpr = W / 4;
    const long long total = (long long)B * H * gpr;
    const float rh = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, rw = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    // up-sampling by more than 3x: the four pixels of a thread touch at most three source columns, so the
    // taps are loaded once per class and selected per pixel
    const bool few_cols = rw * 3.f < 1.f;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const int gx = (int)(i % gpr);
        const long long r = i / gpr;
        const int Y = (int)(r % H), b = (int)(r / H);
        const float h1r = rh * Y;
        const int y1 = (int)h1r, y1p = y1 < h - 1 ? 1 : 0;
        const float hl1 = h1r - y1, hl0 = 1.f - hl1;
        const float *r0 = in + ((size_t)b * h + y1) * w * C, *r1 = r0 + (size_t)y1p * w * C;
        int x1[4];
        float wl0[4], wl1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float w1r = rw * (gx * 4 + j);
            x1[j] = (int)w1r;
            wl1[j] = w1r - x1[j];
            wl0[j] = 1.f - wl1[j];
        }
        float *o = out + (((size_t)b * C) * H + Y) * W + gx * 4;
        if (few_cols) {
            const int xa = x1[0], xb = min(xa + 1, w - 1), xc = min(xa + 2, w - 1);
            for (int c = 0; c < C; ++c) {
                const float a0 = __ldg(r0 + (size_t)xa * C + c), a1 = __ldg(r0 + (size_t)xb * C + c), a2 = __ldg(r0 + (size_t)xc * C + c);
                const float b0 = __ldg(r1 + (size_t)xa * C + c), b1 = __ldg(r1 + (size_t)xb * C + c), b2 = __ldg(r1 + (size_t)xc * C + c);
                float res[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool first = x1[j] == xa;
                    const float tl = first ? a0 : a1, tr = first ? a1 : a2, bl = first ? b0 : b1, br = first ? b1 : b2;
                    res[j] = hl0 * (wl0[j] * tl + wl1[j] * tr) + hl1 * (wl0[j] * bl + wl1[j] * br);
                }
                st_stream_f4(o + (size_t)c * H * W, make_float4(res[0], res[1], res[2], res[3]));
            }
        } else {
            for (int c = 0; c < C; ++c) {
                float res[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int xr = min(x1[j] + 1, w - 1);
                    res[j] = hl0 * (wl0[j] * __ldg(r0 + (size_t)x1[j] * C + c) + wl1[j] * __ldg(r0 + (size_t)xr * C + c)) +
                             hl1 * (wl0[j] * __ldg(r1 + (size_t)x1[j] * C + c) + wl1[j] * __ldg(r1 + (size_t)xr * C + c));
                }
                st_stream_f4(o + (size_t)c * H * W, make_float4(res[0], res[1], res[2], res[3]));
            }
        }
    }
}

// Staged form (the pipeline's case: x4 up-sampling of [128,128,C] decoder outputs).  The kernel above
// reads its taps with scalar loads whose lanes are C floats apart: 54 load instructions per thread, each
// touching nine 128-byte lines, which makes it L1-tag bound at 69 % of the write roofline.  Here a CTA
// owns kUpRows consecutive output rows of one image: the two or three source rows they interpolate are
// one contiguous run of the NHWC input, copied to shared memory with coalesced 16-byte cp.async, and the
// taps come from there (lane stride C words: conflict-free for odd C).  Same arithmetic, same order.
constexpr int kUpRows = 4;
constexpr int kUpStageBytes = 40 * 1024;

__global__ void __launch_bounds__(kThreads, 4) upsample_to_nchw_staged_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                                          int B, int h, int w, int C, int H, int W, int cap_f) {
    extern __shared__ __align__(16) float s_src[];
    const int groups = (H + kUpRows - 1) / kUpRows;
    const int b = blockIdx.x / groups, Y0 = (blockIdx.x - b * groups) * kUpRows;
    const int Y1 = min(H, Y0 + kUpRows);
    const float rh = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, rw = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const int y_first = (int)(rh * Y0), y_last = min(h - 1, (int)(rh * (Y1 - 1)) + 1);
    const int row_f = w * C;                                   // floats per source row
    const int n_f = (y_last - y_first + 1) * row_f;
    const float *src = in + ((size_t)b * h + y_first) * row_f;
    if (n_f > cap_f) __trap();                                 // the host sizes the buffer with a row to spare; never taken
    if ((((uintptr_t)src) & 15) == 0 && (n_f & 3) == 0) {
        const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(s_src);
        for (int i = threadIdx.x; i < n_f / 4; i += kThreads) cp_async16(s0 + 16u * i, src + 4 * i);
        cp_async_commit();
        cp_async_wait<0>();
    } else {
        for (int i = threadIdx.x; i < n_f; i += kThreads) s_src[i] = __ldg(src + i);
    }
    __syncthreads();

    const int gpr = W / 4;
    const bool few_cols = rw * 3.f < 1.f;
    for (int u = threadIdx.x; u < (Y1 - Y0) * gpr; u += kThreads) {
        const int row = u / gpr, gx = u - row * gpr;
        const int Y = Y0 + row;
        const float h1r = rh * Y;
        const int y1 = (int)h1r, y1p = y1 < h - 1 ? 1 : 0;
        const float hl1 = h1r - y1, hl0 = 1.f - hl1;
        const float *r0 = s_src + (y1 - y_first) * row_f, *r1 = r0 + y1p * row_f;
        int x1[4], xr[4];
        float wl0[4], wl1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float w1r = rw * (gx * 4 + j);
            x1[j] = (int)w1r;
            xr[j] = min(x1[j] + 1, w - 1) * C;
            wl1[j] = w1r - x1[j];
            wl0[j] = 1.f - wl1[j];
            x1[j] *= C;
        }
        float *o = out + (((size_t)b * C) * H + Y) * W + gx * 4;
        if (few_cols) {
            // The four pixels touch at most three source columns.  Each pixel's bilinear weights are spread
            // over the three columns once per thread (zero where a column is not used), so a class costs six
            // taps from shared memory and no per-class selects; the kernel is issue-bound, not store-bound.
            // (The rows are blended before the columns: results differ from h0*(w0*a + w1*b) + h1*(...)
            // by an ulp or two.)
            const int xa = x1[0], xb = min(xa + C, (w - 1) * C), xc = min(xa + 2 * C, (w - 1) * C);
            float cw[4][3];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool first = x1[j] == xa;
                const float l = wl0[j], r = wl1[j];
                const bool right_same = xr[j] == x1[j];     // clamped at the last column: both taps coincide
                float c0 = 0.f, c1 = 0.f, c2 = 0.f;
                if (first) { c0 = l; if (right_same) c0 += r; else c1 = r; }
                else { c1 = l; if (right_same) c1 += r; else c2 = r; }
                cw[j][0] = c0; cw[j][1] = c1; cw[j][2] = c2;
            }
            for (int c = 0; c < C; ++c) {
                // vertical blend of the three columns, then each pixel's horizontal blend: 6 + 12 FP ops per class
                const float v0 = fmaf(hl0, r0[xa + c], hl1 * r1[xa + c]);
                const float v1 = fmaf(hl0, r0[xb + c], hl1 * r1[xb + c]);
                const float v2 = fmaf(hl0, r0[xc + c], hl1 * r1[xc + c]);
                float res[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) res[j] = fmaf(cw[j][0], v0, fmaf(cw[j][1], v1, cw[j][2] * v2));
                st_stream_f4(o + (size_t)c * H * W, make_float4(res[0], res[1], res[2], res[3]));
            }
        } else {
            for (int c = 0; c < C; ++c) {
                float res[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    res[j] = hl0 * (wl0[j] * r0[x1[j] + c] + wl1[j] * r0[xr[j] + c]) + hl1 * (wl0[j] * r1[x1[j] + c] + wl1[j] * r1[xr[j] + c]);
                st_stream_f4(o + (size_t)c * H * W, make_float4(res[0], res[1], res[2], res[3]));
            }
        }
    }
}

static unsigned glue_grid(long long threads_wanted, int per_sm) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = (threads_wanted + kThreads - 1) / kThreads;
    const long long cap = (long long)sms * per_sm;
    if (want > cap) want = cap;
    return (unsigned)(want < 1 ? 1 : want);
}

}  // namespace pylc

using namespace pylc;

extern "C" int pylc_upsample_concat_nhwc_f32(const float *x, int B, int h, int w, int Cx, const float *low, int H, int W, int Cl,
                                             float *out, pylc_stream_t stream) {
    if (!x || !low || !out || B < 1 || h < 1 || w < 1 || H < 1 || W < 1) return PYLC_ERR_ARG;
    if (Cx < 4 || Cx % 4 || Cl < 0 || Cl % 4 || Cx > 512 || Cl * kUpSeg > kUpLowMax * 128) return PYLC_ERR_GEOMETRY;
    if (((uintptr_t)x | (uintptr_t)low | (uintptr_t)out) % 16) return PYLC_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    // staged form when one blended source row fits the default shared-memory window
    const size_t smem = (size_t)w * Cx * 4 + (size_t)W * 8;
    if (smem <= 48 * 1024 && (long long)w * (Cx / 4) < 65536 && (long long)B * H < 0x7FFFFFFF) {
        upsample_concat_staged_kernel<<<(unsigned)(B * H), kThreads, smem, st>>>(x, low, out, B, h, w, Cx, H, W, Cl);
        return finish_launch();
    }
    const long long items = (long long)B * H * ((W + kUpSeg - 1) / kUpSeg);
    const unsigned grid = glue_grid(items * 32, 3);            // persistent: 3 CTAs of 8 warps per SM
    if (Cx <= 128) upsample_concat_kernel<1><<<grid, kThreads, 0, st>>>(x, low, out, B, h, w, Cx, H, W, Cl);
    else if (Cx <= 256) upsample_concat_kernel<2><<<grid, kThreads, 0, st>>>(x, low, out, B, h, w, Cx, H, W, Cl);
    else upsample_concat_kernel<4><<<grid, kThreads, 0, st>>>(x, low, out, B, h, w, Cx, H, W, Cl);
    return finish_launch();
}

// ---- tap combine: the reduction step of the plan's tap-split convolution -------------------------------
// out[b, y, x, :] = relu(out[b, y, x, :] + sum over taps t whose source pixel (y + dy_t, x + dx_t) lies inside the
// map of z_t[b, p - p0_t, :]), p = y*W + x.  z_t holds the product of tap t for the flat pixel range
// [p0_t, p0_t + m_t) of every image (a GEMM over a contiguous run of the channels-last activation shifted by
// dy_t*W + dx_t pixels; rows whose source wrapped into the neighbouring image row are the ones masked out here).
// One streaming pass: every float4 of `out` is read and written once, every valid float4 of the z_t once.
constexpr int kMaxTaps = 8;
struct TapArgs {
    float *out;
    const float *z[kMaxTaps];
    int p0[kMaxTaps], m[kMaxTaps], dy[kMaxTaps], dx[kMaxTaps];
    int ntaps, B, H, W, O4;          // O4 = channels / 4
};

__global__ void __launch_bounds__(kThreads) tap_combine_relu_kernel(const __grid_constant__ TapArgs a) {
    const int HW = a.H * a.W;
    const long long total = (long long)a.B * HW * a.O4;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const int o4 = (int)(i % a.O4);
        const long long bp = i / a.O4;
        const int p = (int)(bp % HW), b = (int)(bp / HW);
        const int y = p / a.W, x = p - y * a.W;
        float4 v = reinterpret_cast<const float4 *>(a.out)[i];
#pragma unroll
        for (int t = 0; t < kMaxTaps; ++t) {
            if (t < a.ntaps && (unsigned)(y + a.dy[t]) < (unsigned)a.H && (unsigned)(x + a.dx[t]) < (unsigned)a.W) {
                const float4 z = ld_stream_f4(a.z[t] + (((long long)b * a.m[t] + (p - a.p0[t])) * a.O4 + o4) * 4);
                v.x += z.x; v.y += z.y; v.z += z.z; v.w += z.w;
            }
        }
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        reinterpret_cast<float4 *>(a.out)[i] = v;
    }
}

extern "C" int pylc_tap_combine_relu_f32(float *out, int B, int H, int W, int O, const float *const *z, const int32_t *p0,
                                         const int32_t *m, const int32_t *dy, const int32_t *dx, int ntaps, pylc_stream_t stream) {
    if (!out || B < 1 || H < 1 || W < 1 || ntaps < 0 || ntaps > kMaxTaps) return PYLC_ERR_ARG;
    if (ntaps && (!z || !p0 || !m || !dy || !dx)) return PYLC_ERR_ARG;
    if (O < 4 || O % 4) return PYLC_ERR_GEOMETRY;
    if ((uintptr_t)out % 16) return PYLC_ERR_ALIGN;
    TapArgs a;
    a.out = out; a.ntaps = ntaps; a.B = B; a.H = H; a.W = W; a.O4 = O / 4;
    for (int t = 0; t < kMaxTaps; ++t) {
        const bool on = t < ntaps;
        a.z[t] = on ? z[t] : nullptr;
        a.p0[t] = on ? p0[t] : 0; a.m[t] = on ? m[t] : 0; a.dy[t] = on ? dy[t] : 0; a.dx[t] = on ? dx[t] : 0;
        if (on) {
            if (!z[t] || (uintptr_t)z[t] % 16) return PYLC_ERR_ALIGN;
            // every pixel the mask lets through must lie inside [p0, p0 + m): source inside the map <=> p + dy*W + dx in [0, HW)
            const long long s = (long long)dy[t] * W + dx[t];
            const long long lo = s < 0 ? -s : 0, hi = (long long)H * W - (s > 0 ? s : 0);
            if (p0[t] > lo || (long long)p0[t] + m[t] < hi) return PYLC_ERR_GEOMETRY;
        }
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long total = (long long)B * H * W * a.O4;
    long long want = (total + kThreads - 1) / kThreads;
    if (want > (long long)sms * 8) want = (long long)sms * 8;
    tap_combine_relu_kernel<<<(unsigned)want, kThreads, 0, (cudaStream_t)stream>>>(a);
    return finish_launch();
}

extern "C" int pylc_maxpool3x3s2_nhwc_f32(const float *in, int B, int H, int W, int C, float *out, pylc_stream_t stream) {
    if (!in || !out || B < 1 || H < 1 || W < 1) return PYLC_ERR_ARG;
    if (C < 4 || C % 4) return PYLC_ERR_GEOMETRY;
    if (((uintptr_t)in | (uintptr_t)out) % 16) return PYLC_ERR_ALIGN;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    if ((long long)B * Ho > 0x7FFFFFFF) return PYLC_ERR_GEOMETRY;
    maxpool3x3s2_kernel<<<(unsigned)(B * Ho), kThreads, 0, (cudaStream_t)stream>>>(in, out, B, H, W, C, Ho, Wo);
    return finish_launch();
}

extern "C" int pylc_upsample_nhwc_to_nchw_f32(const float *in, int B, int h, int w, int C, float *out, int H, int W,
                                              pylc_stream_t stream) {
    if (!in || !out || B < 1 || h < 1 || w < 1 || H < 1 || W < 1) return PYLC_ERR_ARG;
    if (C < 1 || C > PYLC_MAX_CLASSES) return PYLC_ERR_CLASSES;
    if (W % 4) return PYLC_ERR_GEOMETRY;
    if ((uintptr_t)out % 16) return PYLC_ERR_ALIGN;
    const long long total = (long long)B * H * (W / 4);
    cudaStream_t st = (cudaStream_t)stream;
    // staged form when the source rows of kUpRows output rows fit the staging buffer (any up-sampling)
    const double rh = H > 1 ? (double)(h - 1) / (H - 1) : 0.0;
    long long stage_rows = (long long)(rh * (kUpRows - 1)) + 4;   // floor bound + one row for float rounding
    if (stage_rows > h) stage_rows = h;
    const long long groups = (long long)B * ((H + kUpRows - 1) / kUpRows);
    if (stage_rows * w * C * 4 <= kUpStageBytes && groups < 0x7FFFFFFF) {
        const size_t smem = (size_t)stage_rows * w * C * 4;
        upsample_to_nchw_staged_kernel<<<(unsigned)groups, kThreads, smem, st>>>(in, out, B, h, w, C, H, W, (int)(smem / 4));
        return finish_launch();
    }
    upsample_to_nchw_kernel<<<glue_grid(total, 8), kThreads, 0, st>>>(in, out, B, h, w, C, H, W);
    return finish_launch();
}
