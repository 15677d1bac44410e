// Helpers of the detection-head path (the student's RetinaNet head applied to the teacher
// pyramid, distillator.py:107-112 / customized_detectors/retinanet.py:36-45): the loss hands back d(logits) /
// d(deltas) as per-level (N, H*W*A, K) fp32 tensors, which in the pixel-major pyramid order are (B*P, A*K) matrices
// with 720 / 36 columns. One pass measures their l2 norm (power-of-two scale of the fp16 operands), a second one
// writes the scaled fp16 operand pyramids of the 256-column dgrad / wgrad launches (zero padded) and the bias gradient.
#include <cuda_fp16.h>

#include "common.cuh"

namespace lgd {

constexpr int HG_ROWS = 32;   // pixel rows per block

struct HeadGradSrc {
  const float* ptr[LGD_MAX_LEVELS];
  long long bstride[LGD_MAX_LEVELS];   // elements between consecutive images of a level
};

// pixel p of the pyramid order (level-major, then image, then y*w+x) -> source row pointer
__device__ __forceinline__ const float* head_row(const HeadGradSrc& src, const Pyr& p, long long pix, int ncols) {
  int l = 0;
  while (l + 1 < p.num_levels && pix >= (long long)p.batch * p.pix_start[l + 1]) ++l;
  const long long r = pix - (long long)p.batch * p.pix_start[l];
  const int hw = p.h[l] * p.w[l];
  const int b = (int)(r / hw);
  const int q = (int)(r - (long long)b * hw);
  return src.ptr[l] + b * src.bstride[l] + (long long)q * ncols;
}

__global__ void __launch_bounds__(256)
head_grad_sumsq_kernel(HeadGradSrc src, Pyr p, int ncols, long long npix, float* __restrict__ partial) {
  __shared__ float red[32];
  const long long p0 = (long long)blockIdx.x * HG_ROWS;
  float s = 0.f;
  for (int r = 0; r < HG_ROWS && p0 + r < npix; ++r) {
    const float* row = head_row(src, p, p0 + r, ncols);
    for (int c = threadIdx.x; c < ncols; c += 256) {
      const float v = row[c];
      s = fmaf(v, v, s);
    }
  }
  const float t = block_sum<float>(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// out_half: nchunks pyramids of (npix, 256) halves back to back; colsum_partial: [block][ncols]
__global__ void __launch_bounds__(256)
head_grad_split_kernel(HeadGradSrc src, Pyr p, int ncols, int nchunks, long long npix, const float* __restrict__ scale3,
                       __half* __restrict__ out_half, float* __restrict__ colsum_partial) {
  const long long p0 = (long long)blockIdx.x * HG_ROWS;
  const float s = __ldg(scale3);
  float cs[4] = {0.f, 0.f, 0.f, 0.f};   // up to 4 chunks of 256 columns
  for (int r = 0; r < HG_ROWS && p0 + r < npix; ++r) {
    const float* row = head_row(src, p, p0 + r, ncols);
    for (int j = 0; j < nchunks; ++j) {
      const int c = j * 256 + threadIdx.x;
      const float v = c < ncols ? row[c] : 0.f;
      cs[j] += v;
      const float sv = fminf(fmaxf(v * s, -65504.f), 65504.f);
      out_half[((long long)j * npix + p0 + r) * C + threadIdx.x] = __float2half_rn(sv);
    }
  }
  for (int j = 0; j < nchunks; ++j) {
    const int c = j * 256 + threadIdx.x;
    if (c < ncols) colsum_partial[(long long)blockIdx.x * ncols + c] = cs[j];
  }
}

__global__ void head_colsum_finalize_kernel(const float* __restrict__ partial, int nblocks, int ncols,
                                            float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  double s0 = 0.0, s1 = 0.0;
  int i = 0;
  for (; i + 2 <= nblocks; i += 2) {
    s0 += (double)partial[(long long)i * ncols + c];
    s1 += (double)partial[(long long)(i + 1) * ncols + c];
  }
  if (i < nblocks) s0 += (double)partial[(long long)i * ncols + c];
  out[c] = (float)(s0 + s1);
}

}  // namespace lgd

using namespace lgd;

extern "C" size_t lgd_head_grad_workspace(const lgd_pyramid_t* pyr, int ncols) {
  Pyr p;
  if (make_pyr(pyr, &p) != LGD_OK) return 0;
  const long long npix = (long long)p.batch * p.pix_start[p.num_levels];
  const long long nblk = (npix + HG_ROWS - 1) / HG_ROWS;
  return (size_t)nblk * (ncols + 1) * sizeof(float);
}

extern "C" int lgd_head_grad_prepare(const lgd_pyramid_t* pyr, const float* const* grad_levels_host,
                                     const int64_t* batch_strides_host, int ncols, void* out_half, float* scale3,
                                     float* gbias, void* workspace, size_t workspace_bytes, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(grad_levels_host && batch_strides_host && out_half && scale3 && gbias && workspace,
                "lgd_head_grad_prepare: null pointer");
  LGD_CHECK_ARG(ncols > 0 && ncols <= 4 * C, "lgd_head_grad_prepare: between 1 and 1024 columns");
  LGD_CHECK_ARG(workspace_bytes >= lgd_head_grad_workspace(pyr, ncols), "lgd_head_grad_prepare: workspace too small");
  HeadGradSrc src;
  for (int l = 0; l < LGD_MAX_LEVELS; ++l) {
    src.ptr[l] = l < p.num_levels ? grad_levels_host[l] : nullptr;
    src.bstride[l] = l < p.num_levels ? batch_strides_host[l] : 0;
    LGD_CHECK_ARG(l >= p.num_levels || src.ptr[l] != nullptr, "lgd_head_grad_prepare: null level pointer");
  }
  const long long npix = (long long)p.batch * p.pix_start[p.num_levels];
  const int nblk = (int)((npix + HG_ROWS - 1) / HG_ROWS);
  const int nchunks = (ncols + C - 1) / C;
  float* partial = static_cast<float*>(workspace);
  float* colsum_partial = partial + nblk;
  cudaStream_t s = (cudaStream_t)stream;
  head_grad_sumsq_kernel<<<nblk, 256, 0, s>>>(src, p, ncols, npix, partial);
  LGD_LAUNCH_CHECK();
  rc = lgd_grad_scale(partial, nblk, 1, nullptr, nullptr, 1.0f, scale3, stream);
  if (rc != LGD_OK) return rc;
  head_grad_split_kernel<<<nblk, 256, 0, s>>>(src, p, ncols, nchunks, npix, scale3, static_cast<__half*>(out_half),
                                              colsum_partial);
  LGD_LAUNCH_CHECK();
  head_colsum_finalize_kernel<<<(ncols + 255) / 256, 256, 0, s>>>(colsum_partial, nblk, ncols, gbias);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}
