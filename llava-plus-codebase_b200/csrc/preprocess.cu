// Image preprocessing of the LLaVA path on the GPU: uint8 RGB (HWC, any size) -> CLIP pixel_values [3, out, out] bf16.
//
// Replaces, for throughput runs (BASELINE configs[2]: >= 2 k images/s/GPU would otherwise be bound by PIL on the host
// cores), the reference's llava/mm_utils.py:16-44 — `expand2square` (pad to a square with the processor's mean colour) and
// CLIPImageProcessor.preprocess as the pinned transformers 4.31 runs it: PIL bicubic resize of the shortest edge to `out`,
// centre crop out x out, rescale by 1/255, normalise by mean / std.
//
// The resize is PIL's ImagingResample, restated bit for bit: separable, horizontal pass first, 8-bit intermediate, filter
// support scaled by the downscale factor (antialiasing), coefficients in 22-bit fixed point, rounding by adding 2^21 before
// the shift, clip to [0, 255]. The coefficient tables (bounds + fixed-point weights per output pixel) are computed on the
// host in double precision exactly as PIL's precompute_coeffs / normalize_coeffs_8bpc do (llava/_b2/preprocess.py) and passed
// in; the kernels below only do the integer arithmetic. Padding is virtual: a read outside the pasted image returns the
// background colour, so the padded square is never materialised.
#include <stdint.h>

#include "common.cuh"
#include "kernels.h"

namespace b2 {
namespace {

constexpr int PP_PRECISION_BITS = 22;  // PIL: 32 - 8 - 2

__device__ __forceinline__ int clip8(int v) {
    v >>= PP_PRECISION_BITS;  // arithmetic shift, like PIL's lookup on `in >> PRECISION_BITS`
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// pixel of the virtual (padded) source image
__device__ __forceinline__ uchar3 src_pixel(const uint8_t* __restrict__ img, int H, int W, int pad_top, int pad_left, int y, int x,
                                            uchar3 bg) {
    const int yy = y - pad_top, xx = x - pad_left;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) return bg;
    const uint8_t* p = img + ((size_t)yy * W + xx) * 3;
    return make_uchar3(p[0], p[1], p[2]);
}

// horizontal pass: tmp[y, xo, c] for the source rows [y0, y0 + rows) the vertical pass will read; xo in [x_lo, x_lo + cols)
__global__ void pp_horizontal_kernel(const uint8_t* __restrict__ img, int H, int W, int pad_top, int pad_left, uchar3 bg,
                                     const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize, int identity,
                                     int y0, int rows, int x_lo, int cols, uint8_t* __restrict__ tmp) {
    const int xo = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (xo >= cols || r >= rows) return;
    const int y = y0 + r, xs = x_lo + xo;
    uint8_t* o = tmp + ((size_t)r * cols + xo) * 3;
    if (identity) {  // in_size == out_size: PIL skips the pass
        const uchar3 p = src_pixel(img, H, W, pad_top, pad_left, y, xs, bg);
        o[0] = p.x; o[1] = p.y; o[2] = p.z;
        return;
    }
    const int xmin = bounds[2 * xs], n = bounds[2 * xs + 1];
    const int32_t* k = kk + (size_t)xs * ksize;
    int s0 = 1 << (PP_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int i = 0; i < n; ++i) {
        const uchar3 p = src_pixel(img, H, W, pad_top, pad_left, y, xmin + i, bg);
        const int w = k[i];
        s0 += (int)p.x * w; s1 += (int)p.y * w; s2 += (int)p.z * w;
    }
    o[0] = (uint8_t)clip8(s0); o[1] = (uint8_t)clip8(s1); o[2] = (uint8_t)clip8(s2);
}

// vertical pass over tmp + centre crop + rescale + normalise + bf16, CHW output. One thread per output pixel (3 channels).
__global__ void pp_vertical_norm_kernel(const uint8_t* __restrict__ tmp, int y0, int rows, int cols, const int32_t* __restrict__ bounds,
                                        const int32_t* __restrict__ kk, int ksize, int identity, int y_lo, int out, float3 mean,
                                        float3 stdv, float rescale, __nv_bfloat16* __restrict__ pixels, uint8_t* __restrict__ u8_out) {
    const int xo = blockIdx.x * blockDim.x + threadIdx.x;
    const int yo = blockIdx.y;
    if (xo >= out || yo >= out) return;
    const int ys = y_lo + yo;  // row of the resized (uncropped) image
    int c0, c1, c2;
    if (identity) {
        const uint8_t* p = tmp + ((size_t)(ys - y0) * cols + xo) * 3;
        c0 = p[0]; c1 = p[1]; c2 = p[2];
    } else {
        const int ymin = bounds[2 * ys], n = bounds[2 * ys + 1];
        const int32_t* k = kk + (size_t)ys * ksize;
        int s0 = 1 << (PP_PRECISION_BITS - 1), s1 = s0, s2 = s0;
        for (int i = 0; i < n; ++i) {
            const uint8_t* p = tmp + ((size_t)(ymin + i - y0) * cols + xo) * 3;
            const int w = k[i];
            s0 += (int)p[0] * w; s1 += (int)p[1] * w; s2 += (int)p[2] * w;
        }
        c0 = clip8(s0); c1 = clip8(s1); c2 = clip8(s2);
    }
    const size_t plane = (size_t)out * out, at = (size_t)yo * out + xo;
    if (u8_out != nullptr) {  // the resized + cropped 8-bit image itself (parity tests compare it with PIL bit for bit)
        u8_out[at * 3 + 0] = (uint8_t)c0; u8_out[at * 3 + 1] = (uint8_t)c1; u8_out[at * 3 + 2] = (uint8_t)c2;
    }
    if (pixels != nullptr) {
        // HF: image * rescale_factor, then (image - mean) / std, all in fp32
        // (separately rounded mul / sub / div, as numpy evaluates it: no FMA contraction)
        pixels[at] = __float2bfloat16_rn(__fdiv_rn(__fsub_rn(__fmul_rn((float)c0, rescale), mean.x), stdv.x));
        pixels[plane + at] = __float2bfloat16_rn(__fdiv_rn(__fsub_rn(__fmul_rn((float)c1, rescale), mean.y), stdv.y));
        pixels[2 * plane + at] = __float2bfloat16_rn(__fdiv_rn(__fsub_rn(__fmul_rn((float)c2, rescale), mean.z), stdv.z));
    }
}

}  // namespace

int preprocess_clip_image(const PreprocessArgs& a, cudaStream_t stream) {
    B2_CHECK_ARG(a.img && a.tmp && (a.pixels || a.u8_out) && a.H > 0 && a.W > 0 && a.out > 0, "preprocess: bad argument");
    B2_CHECK_ARG(a.rows > 0 && a.cols == a.out && a.y0 >= 0 && a.x_lo >= 0 && a.y_lo >= 0, "preprocess: bad plan");
    B2_CHECK_ARG(a.h_identity || (a.h_bounds && a.h_kk && a.h_ksize > 0), "preprocess: horizontal coefficient table missing");
    B2_CHECK_ARG(a.v_identity || (a.v_bounds && a.v_kk && a.v_ksize > 0), "preprocess: vertical coefficient table missing");
    const uchar3 bg = make_uchar3(a.bg[0], a.bg[1], a.bg[2]);
    {
        dim3 grid((a.cols + 127) / 128, a.rows);
        pp_horizontal_kernel<<<grid, 128, 0, stream>>>(a.img, a.H, a.W, a.pad_top, a.pad_left, bg, a.h_bounds, a.h_kk, a.h_ksize,
                                                       a.h_identity, a.y0, a.rows, a.x_lo, a.cols, a.tmp);
        B2_LAUNCH_CHECK();
    }
    {
        dim3 grid((a.out + 127) / 128, a.out);
        pp_vertical_norm_kernel<<<grid, 128, 0, stream>>>(a.tmp, a.y0, a.rows, a.cols, a.v_bounds, a.v_kk, a.v_ksize, a.v_identity,
                                                          a.y_lo, a.out, make_float3(a.mean[0], a.mean[1], a.mean[2]),
                                                          make_float3(a.stdv[0], a.stdv[1], a.stdv[2]), a.rescale,
                                                          reinterpret_cast<__nv_bfloat16*>(a.pixels), a.u8_out);
        B2_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace b2
