// CLIP ViT glue kernels around the tcgen05 GEMMs (all HBM-bound, one read + one write per element):
//   vit_im2col   : Conv2d(3->D, k=s=14, no bias) of transformers modeling_clip.py:202-217 restated as an
//                  im2col gather so patch embedding is a [B*576, 588(+pad)] x [588, D] GEMM on the tensor cores.
//   vit_embed_ln : cat(CLS, patches) + position_embedding, then pre_layrnorm (modeling_clip.py:213-217, :677).
//   vit_drop_cls : feature_select 'patch' (llava/model/multimodal_encoder/clip_encoder.py:29-37).
#include "common.cuh"
#include "kernels.h"

namespace b2 {
namespace {

// one CTA per patch; threads sweep the kpad output elements (k = c*ps*ps + i*ps + j)
__global__ void vit_im2col_kernel(const __nv_bfloat16* __restrict__ pix, __nv_bfloat16* __restrict__ out, int img,
                                  int ps, int kpad) {
    pdl_trigger();
    pdl_wait();  // inputs are outputs of the upstream kernel (programmatic dependent launch)
    const int grid_w = img / ps;
    const int P = grid_w * grid_w;
    const int bp = blockIdx.x;
    const int b = bp / P, pidx = bp % P;
    const int py = pidx / grid_w, px = pidx % grid_w;
    const int kk = 3 * ps * ps;
    const __nv_bfloat16* base = pix + (size_t)b * 3 * img * img;
    __nv_bfloat16* o = out + (size_t)bp * kpad;
    for (int k = threadIdx.x; k < kpad; k += blockDim.x) {
        __nv_bfloat16 v = __float2bfloat16_rn(0.f);
        if (k < kk) {
            const int c = k / (ps * ps), r = k % (ps * ps);
            const int i = r / ps, j = r % ps;
            v = base[((size_t)c * img + (py * ps + i)) * img + (px * ps + j)];
        }
        o[k] = v;
    }
}

// one warp per token row; D % 256 == 0, D <= 2048
__global__ void __launch_bounds__(128)
vit_embed_ln_kernel(const __nv_bfloat16* __restrict__ patch_out, const __nv_bfloat16* __restrict__ cls,
                    const __nv_bfloat16* __restrict__ pos, const __nv_bfloat16* __restrict__ gamma,
                    const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ hidden, int B, int P,
                    int D, float eps) {
    pdl_trigger();
    pdl_wait();  // inputs are outputs of the upstream kernel (programmatic dependent launch)
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int T = P + 1;
    if (row >= B * T) return;
    const int b = row / T, t = row % T;
    const __nv_bfloat16* src = (t == 0) ? cls : patch_out + ((size_t)b * P + (t - 1)) * D;
    const __nv_bfloat16* pr = pos + (size_t)t * D;
    const int nvec = D / 256;
    float v[8][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < nvec) {
            const int c = i * 256 + lane * 8;
            const uint4 a = *reinterpret_cast<const uint4*>(src + c);
            const uint4 q = *reinterpret_cast<const uint4*>(pr + c);
            // HF adds in the model dtype: embeddings (bf16) + position (bf16) -> bf16
            v[i][0] = round_bf16(bf16_lo(a.x) + bf16_lo(q.x)); v[i][1] = round_bf16(bf16_hi(a.x) + bf16_hi(q.x));
            v[i][2] = round_bf16(bf16_lo(a.y) + bf16_lo(q.y)); v[i][3] = round_bf16(bf16_hi(a.y) + bf16_hi(q.y));
            v[i][4] = round_bf16(bf16_lo(a.z) + bf16_lo(q.z)); v[i][5] = round_bf16(bf16_hi(a.z) + bf16_hi(q.z));
            v[i][6] = round_bf16(bf16_lo(a.w) + bf16_lo(q.w)); v[i][7] = round_bf16(bf16_hi(a.w) + bf16_hi(q.w));
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += v[i][e];
        }
    }
    const float mean = warp_sum(sum) / D;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < nvec) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float d = v[i][e] - mean;
                var += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(var) / D + eps);
    __nv_bfloat16* yr = hidden + (size_t)row * D;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < nvec) {
            const int c = i * 256 + lane * 8;
            const uint4 g = *reinterpret_cast<const uint4*>(gamma + c);
            const uint4 bq = *reinterpret_cast<const uint4*>(beta + c);
            const float gg[8] = {bf16_lo(g.x), bf16_hi(g.x), bf16_lo(g.y), bf16_hi(g.y),
                                 bf16_lo(g.z), bf16_hi(g.z), bf16_lo(g.w), bf16_hi(g.w)};
            const float bb[8] = {bf16_lo(bq.x), bf16_hi(bq.x), bf16_lo(bq.y), bf16_hi(bq.y),
                                 bf16_lo(bq.z), bf16_hi(bq.z), bf16_lo(bq.w), bf16_hi(bq.w)};
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mean) * rstd * gg[e] + bb[e];
            *reinterpret_cast<uint4*>(yr + c) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                                                           pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        }
    }
}

__global__ void vit_drop_cls_kernel(const uint4* __restrict__ hidden, uint4* __restrict__ out, int P, int vec_per_row,
                                    int64_t total_vec) {
    pdl_trigger();
    pdl_wait();  // inputs are outputs of the upstream kernel (programmatic dependent launch)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_vec) return;
    const int64_t row = i / vec_per_row;  // b*P + p
    const int c = (int)(i % vec_per_row);
    const int64_t b = row / P, pp = row % P;
    out[i] = hidden[(b * (P + 1) + 1 + pp) * vec_per_row + c];
}

}  // namespace

int vit_im2col(const void* pixels, void* out, int B, int img, int patch, int kpad, cudaStream_t stream) {
    B2_CHECK_ARG(img % patch == 0 && kpad >= 3 * patch * patch && kpad % 8 == 0,
                 "vit_im2col: bad geometry img=%d patch=%d kpad=%d", img, patch, kpad);
    const int P = (img / patch) * (img / patch);
    B2_CUDA_CHECK(launch_pdl(vit_im2col_kernel, dim3(B * P), dim3(128), 0, stream, reinterpret_cast<const __nv_bfloat16*>(pixels),
                             reinterpret_cast<__nv_bfloat16*>(out), img, patch, kpad));
    B2_LAUNCH_CHECK();
    return 0;
}

int vit_embed_ln(const void* patch_out, const void* cls, const void* pos, const void* gamma, const void* beta,
                 void* hidden, int B, int P, int D, float eps, cudaStream_t stream) {
    B2_CHECK_ARG(D % 256 == 0 && D <= 2048, "vit_embed_ln: D must be a multiple of 256 and <= 2048 (D=%d)", D);
    const int rows = B * (P + 1);
    B2_CUDA_CHECK(launch_pdl(vit_embed_ln_kernel, dim3((rows + 3) / 4), dim3(128), 0, stream,
        reinterpret_cast<const __nv_bfloat16*>(patch_out), reinterpret_cast<const __nv_bfloat16*>(cls),
        reinterpret_cast<const __nv_bfloat16*>(pos), reinterpret_cast<const __nv_bfloat16*>(gamma),
        reinterpret_cast<const __nv_bfloat16*>(beta), reinterpret_cast<__nv_bfloat16*>(hidden), B, P, D, eps));
    B2_LAUNCH_CHECK();
    return 0;
}

int vit_drop_cls(const void* hidden, void* out, int B, int P, int D, cudaStream_t stream) {
    B2_CHECK_ARG(D % 8 == 0, "vit_drop_cls: D %% 8 != 0");
    const int vec_per_row = D / 8;
    const int64_t total = (int64_t)B * P * vec_per_row;
    const int grid = (int)((total + 255) / 256);
    B2_CUDA_CHECK(launch_pdl(vit_drop_cls_kernel, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const uint4*>(hidden),
                             reinterpret_cast<uint4*>(out), P, vec_per_row, total));
    B2_LAUNCH_CHECK();
    return 0;
}

}  // namespace b2
