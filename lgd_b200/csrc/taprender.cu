// local_inst_proj_2D without a convolution (a7, dynamic_teacher.py:137-146).
//
// The convolution's input is the rendered map sum_t mask_t (x) e_t: piecewise constant over box rectangles. Hence
//     conv3x3(rendered)[y, x] = sum_t sum_tap [(y+dy, x+dx) in box_t] * V[t][tap],      V[t][tap] = W_tap e_t,
// one token-sized GEMM (rows x 256 x 2304) followed by a paint pass: interior pixels of a box take the precomputed
// sum_tap V[t][tap], only the one-pixel ring inside and outside a box border sums individual taps. Backward:
//     S[t][tap] = sum_{q in box_t} g[q - tap]      (ONE box-sum pass with nine accumulators per channel)
//     d e_t = sum_tap W_tap^T S[t][tap],           d W_tap = sum_t S[t][tap] (x) e_t          (two token-sized GEMMs).
// Exact fp32 arithmetic (the convolution path rounds the rendered map and the weights to fp16). The identity is checked on
// the CPU in tests/test_oracle.py::test_local_inst_conv_over_the_rendered_map_equals_per_box_tap_sums; the kernels against
// F.conv2d + autograd in tests/test_gpu_kernels.py::test_tap_render_matches_convolution.
// Rasterised polygon masks (LOAD_LABELMAP) are not rectangles: that recipe keeps the convolution.
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace lgd {

constexpr int TAP_PIX = 32;        // pixels per strip
constexpr int TAP_SPB = 4;         // strips per block
constexpr int TAP_STAGE = 16;      // rows whose tap sums are staged in shared memory
constexpr int TAP_MAX_ROWS = LGD_TAP_MAX_ROWS;
constexpr int NT = 9 * C;          // columns of a row's tap vectors: [tap][co]
constexpr int TS_ITEM_PX = 128;    // pixels per box-sum item (same cut as region.cu)

__host__ __device__ inline int tap_blocks_of_level(int hw) {
  const int strips = (hw + TAP_PIX - 1) / TAP_PIX;
  return (strips + TAP_SPB - 1) / TAP_SPB;
}
__device__ __forceinline__ unsigned bit_run(int n) { return n >= 32 ? 0xffffffffu : (1u << n) - 1u; }
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// wp[(tap*C + co)*C + ci] = w[co][ci][tap]   (nn.Conv2d weight (co, ci, 3, 3), tap = ky*3 + kx)
__global__ void tap_weights_kernel(const float* __restrict__ w, float* __restrict__ wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * C * C) return;
  const int ci = i % C, co = (i / C) % C, tap = i / (C * C);
  wp[i] = w[((long long)co * C + ci) * 9 + tap];
}
// gw[co][ci][tap] = gwp[(tap*C + co)*C + ci]
__global__ void tap_weights_grad_kernel(const float* __restrict__ gwp, float* __restrict__ gw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * C * C) return;
  const int tap = i % 9, ci = (i / 9) % C, co = i / (9 * C);
  gw[i] = gwp[((long long)tap * C + co) * C + ci];
}
// vsum[row][c] = sum_tap V[row][tap][c]   (fixed order)
__global__ void tap_vsum_kernel(const float* __restrict__ V, float* __restrict__ vsum) {
  const long long row = blockIdx.x;
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) s += V[row * NT + tap * C + c];
  vsum[row * C + c] = s;
}

// out[l, b, pixel, :] = relu(bias + sum over the rendered rows t of image b, over the taps (dy, dx) with pixel + (dy, dx)
// inside box t, of V[l*T + t][tap][:]).
// A block paints TAP_SPB consecutive 32-pixel strips of one (level, image). Per strip, thread k turns row k's interval
// into two coverage masks -- `any` (the box dilated by one pixel: at least one tap inside) and `all` (eroded by one
// pixel: all nine inside, the row contributes its tap sum) -- and marks the pixels where the set of in-box taps can differ
// from the left neighbour's (first pixel of an image row; x in {x0-1, x0, x0+1, x1-1, x1, x1+1}). A thread owns eight
// consecutive pixels of one channel quad: it evaluates the sum for its first pixel and for marked pixels and stores the
// previous value otherwise.
__global__ void __launch_bounds__(256, 4)
tap_paint_kernel(Pyr p, const float* __restrict__ V, const float* __restrict__ vsum, const int* __restrict__ ranges,
                 const int* __restrict__ img_start, const int* __restrict__ n_rows, int T,
                 const float* __restrict__ bias, int bias_stride_level, int bias_stride_img,
                 __half* __restrict__ out_half, float* __restrict__ out32) {
  __shared__ float4 se[TAP_STAGE][64];
  __shared__ int4 sr[TAP_MAX_ROWS];
  __shared__ unsigned sany[TAP_MAX_ROWS];
  __shared__ unsigned sall[TAP_MAX_ROWS];
  __shared__ unsigned sdiff[2];
  const int b = blockIdx.y;
  int l = 0, grp = blockIdx.x;
  while (l + 1 < p.num_levels) {
    const int nbk = tap_blocks_of_level(p.h[l] * p.w[l]);
    if (grp < nbk) break;
    grp -= nbk;
    ++l;
  }
  const int H = p.h[l], W = p.w[l], HW = H * W;
  const int q = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const int t0 = img_start[b];
  const int nb = min((n_rows != nullptr) ? n_rows[b] : (img_start[b + 1] - t0), TAP_MAX_ROWS);   // host checks the limit
  const long long row0 = (long long)l * T + t0;
  const int k = threadIdx.x;
  if (k < nb) sr[k] = *reinterpret_cast<const int4*>(ranges + (row0 + k) * 4);
#pragma unroll
  for (int jj = 0; jj < TAP_STAGE / 4; ++jj) {
    const int j = sub + 4 * jj;
    if (j < nb) se[j][q] = ld4(vsum + (row0 + j) * C + q * 4);
  }
  if (threadIdx.x < 2) sdiff[threadIdx.x] = 0u;
  const float4 b4 = ld4(bias + (long long)l * bias_stride_level + (long long)b * bias_stride_img + q * 4);
  const long long obase = p.off[l] + (long long)b * HW * C + q * 4;
  for (int s = 0; s < TAP_SPB; ++s) {
    const int pix0 = (grp * TAP_SPB + s) * TAP_PIX;
    if (pix0 >= HW) break;   // block-uniform
    const int ya = pix0 / W, xa = pix0 - ya * W;
    const int npx = min(TAP_PIX, HW - pix0);
    __syncthreads();   // staging done / the previous strip's masks have been consumed
    {
      unsigned m_any = 0u, m_all = 0u, mark = 0u;
      int4 r = make_int4(0, 0, 0, 0);
      if (k < nb) r = sr[k];
      const bool box = k < nb && r.y > r.x && r.w > r.z;
      int x = xa, y = ya;
      for (int j0 = 0; j0 < npx; ++y) {
        const int run = min(W - x, npx - j0);   // pixels j0 .. j0+run-1 of the strip = image row y, columns x .. x+run-1
        mark |= 1u << j0;
        if (box && y >= r.z - 1 && y < r.w + 1) {
          const int lo = max(r.x - 1, x), hi = min(r.y + 1, x + run);
          if (hi > lo) {
            m_any |= bit_run(hi - lo) << (j0 + lo - x);
            const int pos[6] = {r.x - 1, r.x, r.x + 1, r.y - 1, r.y, r.y + 1};
#pragma unroll
            for (int u = 0; u < 6; ++u)
              if (pos[u] >= x && pos[u] < x + run) mark |= 1u << (j0 + pos[u] - x);
          }
          if (y >= r.z + 1 && y < r.w - 1) {
            const int lo2 = max(r.x + 1, x), hi2 = min(r.y - 1, x + run);
            if (hi2 > lo2) m_all |= bit_run(hi2 - lo2) << (j0 + lo2 - x);
          }
        }
        j0 += run;
        x = 0;
      }
      if (k < nb) {
        sany[k] = m_any;
        sall[k] = m_all;
      }
      mark = __reduce_or_sync(0xffffffffu, mark);
      if ((threadIdx.x & 31) == 0) atomicOr(&sdiff[s & 1], mark);   // integer OR: order-free
    }
    __syncthreads();
    const unsigned differs = sdiff[s & 1];
    if (threadIdx.x == 0) sdiff[(s + 1) & 1] = 0u;   // its readers finished before this strip's first barrier
    const int j_first = sub * (TAP_PIX / 4);
    __half* o16 = out_half ? out_half + obase + (long long)(pix0 + j_first) * C : nullptr;
    float* o32 = out32 ? out32 + obase + (long long)(pix0 + j_first) * C : nullptr;
    float4 res = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < TAP_PIX / 4; ++i) {
      const int j = j_first + i;
      if (j >= npx) break;   // warp-uniform
      if (i == 0 || (differs >> j & 1u)) {
        const int pix = pix0 + j;
        const int y = pix / W, x = pix - y * W;
        float4 cur = b4;
        for (int kk = 0; kk < nb; ++kk) {
          if (sany[kk] >> j & 1u) {
            if (sall[kk] >> j & 1u) {
              add4(cur, kk < TAP_STAGE ? se[kk][q] : ld4(vsum + (row0 + kk) * C + q * 4));
            } else {
              const int4 r = sr[kk];
              const float* vrow = V + (row0 + kk) * NT + q * 4;
#pragma unroll
              for (int dy = -1; dy <= 1; ++dy) {
                if (y + dy < r.z || y + dy >= r.w) continue;
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx)
                  if (x + dx >= r.x && x + dx < r.y) add4(cur, ld4(vrow + ((dy + 1) * 3 + (dx + 1)) * C));
              }
            }
          }
        }
        res.x = relu_keep_nan(cur.x); res.y = relu_keep_nan(cur.y);
        res.z = relu_keep_nan(cur.z); res.w = relu_keep_nan(cur.w);
      }
      if (o16 != nullptr) {
        const __half2 h0 = __floats2half2_rn(res.x, res.y), h1 = __floats2half2_rn(res.z, res.w);
        uint2 hv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h0);
        hv.y = *reinterpret_cast<const uint32_t*>(&h1);
        // the copy doubles as the ReLU mask of the backward: a positive value never becomes 0 (as in the conv epilogue)
        if (res.x > 0.f && (hv.x & 0xffffu) == 0) hv.x |= 1u;
        if (res.y > 0.f && (hv.x >> 16) == 0) hv.x |= 0x10000u;
        if (res.z > 0.f && (hv.y & 0xffffu) == 0) hv.y |= 1u;
        if (res.w > 0.f && (hv.y >> 16) == 0) hv.y |= 0x10000u;
        *reinterpret_cast<uint2*>(o16 + i * C) = hv;
      }
      if (o32 != nullptr) *reinterpret_cast<float4*>(o32 + i * C) = res;
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward: tap sums
__device__ __forceinline__ int ts_rows_per_item(int bw) { return bw >= TS_ITEM_PX ? 1 : TS_ITEM_PX / max(bw, 1); }

int boxsum_plan(const int* ranges, int n_entries, int* item_start, cudaStream_t stream);   // region.cu (same item cut)

// partial[(item*9 + tap)*C + c] = sum over the item's box pixels q of g[q - tap, c]   (0 outside the image)
__global__ void __launch_bounds__(256, 2)
tapsum_kernel(Pyr p, const float* __restrict__ g, const int* __restrict__ ranges, const int* __restrict__ img_of,
              const int* __restrict__ img_start, const int* __restrict__ n_rows, int T,
              const int* __restrict__ item_start, float* __restrict__ partial) {
  __shared__ float4 sh[4][9][64];
  __shared__ int s_entry;
  const int n_entries = p.num_levels * T;
  const int total = item_start[n_entries];
  const int q = threadIdx.x & 63, sub = threadIdx.x >> 6;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    if (threadIdx.x == 0) {  // last entry whose first item is <= item
      int lo = 0, hi = n_entries - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (item_start[mid] <= item) lo = mid; else hi = mid - 1;
      }
      s_entry = lo;
    }
    __syncthreads();
    const int e = s_entry;
    const int l = e / T, t = e - l * T;
    const int b = img_of[t];
    const bool active = n_rows == nullptr || (t - img_start[b]) < n_rows[b];
    if (active) {   // block-uniform
      const int4 r = *reinterpret_cast<const int4*>(ranges + (long long)e * 4);
      const int bw = r.y - r.x;
      const int rp = ts_rows_per_item(bw);
      const int y_begin = r.z + (item - item_start[e]) * rp;
      const int y_end = min(r.w, y_begin + rp);
      const int H = p.h[l], W = p.w[l];
      const float* gl = g + p.off[l] + (long long)b * H * W * C + q * 4;
      float4 acc[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) acc[tap] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int n = (y_end - y_begin) * bw;
      for (int i = sub; i < n; i += 4) {
        const int ry = i / bw;
        const int yy = y_begin + ry, xx = r.x + (i - ry * bw);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int qy = yy - (tap / 3 - 1), qx = xx - (tap % 3 - 1);
          if (qy >= 0 && qy < H && qx >= 0 && qx < W) add4(acc[tap], ld4(gl + ((long long)qy * W + qx) * C));
        }
      }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) sh[sub][tap][q] = acc[tap];
      __syncthreads();
      for (int tap = sub; tap < 9; tap += 4) {
        float4 a = sh[0][tap][q];
        add4(a, sh[1][tap][q]);
        add4(a, sh[2][tap][q]);
        add4(a, sh[3][tap][q]);
        *reinterpret_cast<float4*>(partial + ((long long)item * 9 + tap) * C + q * 4) = a;
      }
    }
    __syncthreads();  // sh / s_entry are reused by the next item
  }
}

// S[(l*T + t)*NT + tap*C + c] = sum over the entry's items (fixed order); rows outside the rendered subset -> 0
__global__ void tapsum_finalize_kernel(const float* __restrict__ partial, const int* __restrict__ item_start,
                                       const int* __restrict__ img_of, const int* __restrict__ img_start,
                                       const int* __restrict__ n_rows, int T, float* __restrict__ S) {
  const int t = blockIdx.x, l = blockIdx.y, tap = blockIdx.z, c = threadIdx.x;
  const long long row = (long long)l * T + t;
  bool active = true;
  if (n_rows != nullptr) {
    const int b = img_of[t];
    active = (t - img_start[b]) < n_rows[b];
  }
  float s = 0.f;
  if (active) {
    const int i0 = item_start[row], i1 = item_start[row + 1];
    for (int i = i0; i < i1; ++i) s += partial[((long long)i * 9 + tap) * C + c];
  }
  S[row * NT + tap * C + c] = s;
}

static size_t align256(size_t n) { return (n + 255) & ~size_t(255); }
static size_t tap_items_bound(const lgd_pyramid_t* pyr, int T) {
  size_t rows = 0;   // an item is at least one row of its box -> at most h[l] items per (box, level)
  for (int l = 0; l < pyr->num_levels; ++l) rows += (size_t)pyr->h[l];
  return (size_t)T * rows;
}

}  // namespace lgd

using namespace lgd;

extern "C" size_t lgd_tap_render_workspace(const lgd_pyramid_t* pyr, int T, int backward) {
  if (pyr == nullptr || T <= 0) return 0;
  const size_t rows = (size_t)pyr->num_levels * T;
  size_t n = align256((size_t)9 * C * C * 4);                 // tap-major weights
  n += align256(lgd_linear_workspace((int)rows, NT, C));      // split-K scratch of the token-sized GEMMs
  if (!backward) {
    n += align256(rows * NT * 4) + align256(rows * C * 4);    // V, tap sums
  } else {
    n += align256(rows * NT * 4) + align256((size_t)9 * C * C * 4);           // S, tap-major weight gradient
    n += align256((rows + 1) * 4) + align256(tap_items_bound(pyr, T) * 9 * C * 4);   // item plan, item partials
  }
  return n + 256;
}

extern "C" int lgd_tap_render_fwd(const lgd_pyramid_t* pyr, const float* emb, const float* weight,
                                  const int32_t* ranges, const int32_t* img_start, const int32_t* n_render, int T,
                                  int max_rows, const float* bias, int bias_stride_level, int bias_stride_img,
                                  void* out_half, float* out32, void* workspace, size_t workspace_bytes, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(emb && weight && ranges && img_start && bias && (out_half || out32) && workspace && T > 0,
                "lgd_tap_render_fwd: bad arguments");
  LGD_CHECK_ARG(max_rows > 0 && max_rows <= TAP_MAX_ROWS, "lgd_tap_render_fwd: at most %d rows per image (got %d)",
                TAP_MAX_ROWS, max_rows);
  LGD_CHECK_ARG(workspace_bytes >= lgd_tap_render_workspace(pyr, T, 0), "lgd_tap_render_fwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int rows = p.num_levels * T;
  char* w8 = static_cast<char*>(workspace);
  float* wp = reinterpret_cast<float*>(w8);
  w8 += align256((size_t)9 * C * C * 4);
  void* lin_ws = w8;
  const size_t lin_bytes = lgd_linear_workspace(rows, NT, C);
  w8 += align256(lin_bytes);
  float* V = reinterpret_cast<float*>(w8);
  w8 += align256((size_t)rows * NT * 4);
  float* vsum = reinterpret_cast<float*>(w8);
  tap_weights_kernel<<<(9 * C * C + 255) / 256, 256, 0, s>>>(weight, wp);
  LGD_LAUNCH_CHECK();
  rc = lgd_linear_fwd(emb, C, wp, C, nullptr, V, NT, rows, NT, C, lin_ws, lin_bytes, stream);   // V = emb Wp^T
  if (rc != LGD_OK) return rc;
  tap_vsum_kernel<<<rows, C, 0, s>>>(V, vsum);
  LGD_LAUNCH_CHECK();
  int groups = 0;
  for (int l = 0; l < p.num_levels; ++l) groups += tap_blocks_of_level(p.h[l] * p.w[l]);
  tap_paint_kernel<<<dim3(groups, p.batch), 256, 0, s>>>(p, V, vsum, ranges, img_start, n_render, T, bias,
                                                          bias_stride_level, bias_stride_img,
                                                          static_cast<__half*>(out_half), out32);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_tap_render_bwd(const lgd_pyramid_t* pyr, const float* gout, const float* emb, const float* weight,
                                  const int32_t* ranges, const int32_t* img_of, const int32_t* img_start,
                                  const int32_t* n_render, int T, float* gemb, float* gweight, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(gout && emb && weight && ranges && img_of && img_start && gemb && gweight && workspace && T > 0,
                "lgd_tap_render_bwd: bad arguments");
  LGD_CHECK_ARG(workspace_bytes >= lgd_tap_render_workspace(pyr, T, 1), "lgd_tap_render_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int rows = p.num_levels * T;
  char* w8 = static_cast<char*>(workspace);
  float* wp = reinterpret_cast<float*>(w8);
  w8 += align256((size_t)9 * C * C * 4);
  void* lin_ws = w8;
  const size_t lin_bytes = lgd_linear_workspace(rows, NT, C);
  w8 += align256(lin_bytes);
  float* S = reinterpret_cast<float*>(w8);
  w8 += align256((size_t)rows * NT * 4);
  float* gwp = reinterpret_cast<float*>(w8);
  w8 += align256((size_t)9 * C * C * 4);
  int* item_start = reinterpret_cast<int*>(w8);
  w8 += align256(((size_t)rows + 1) * 4);
  float* partial = reinterpret_cast<float*>(w8);
  int dev = 0, sms = 0;
  LGD_CUDA(cudaGetDevice(&dev));
  LGD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  tap_weights_kernel<<<(9 * C * C + 255) / 256, 256, 0, s>>>(weight, wp);
  LGD_LAUNCH_CHECK();
  rc = boxsum_plan(ranges, rows, item_start, s);
  if (rc != LGD_OK) return rc;
  tapsum_kernel<<<sms * 2, 256, 0, s>>>(p, gout, ranges, img_of, img_start, n_render, T, item_start, partial);
  LGD_LAUNCH_CHECK();
  tapsum_finalize_kernel<<<dim3(T, p.num_levels, 9), C, 0, s>>>(partial, item_start, img_of, img_start, n_render, T, S);
  LGD_LAUNCH_CHECK();
  // d emb[rows, C] = S[rows, NT] Wp[NT, C];   d Wp[NT, C] = S^T emb
  rc = lgd_linear_bwd_input(S, NT, wp, C, gemb, C, rows, NT, C, 0, lin_ws, lin_bytes, stream);
  if (rc != LGD_OK) return rc;
  rc = lgd_linear_bwd_weight(S, NT, emb, C, gwp, C, nullptr, rows, NT, C, 0, lin_ws, lin_bytes, stream);
  if (rc != LGD_OK) return rc;
  tap_weights_grad_kernel<<<(9 * C * C + 255) / 256, 256, 0, s>>>(gwp, gweight);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}
