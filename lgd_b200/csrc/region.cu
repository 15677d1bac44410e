// Label-guided region kernels: the exact box->pixel membership (a4, utils.py:53-89), mask average pooling
// (a5, dynamic_teacher.py:81-103), rendering (a7, dynamic_teacher.py:106-206) and their transposes.
// The reference materialises a float mask per (image, level) and runs a dense GEMM against it; here the
// mask never exists: membership is reduced ONCE per step to integer pixel intervals (exactly, by evaluating
// the reference's fp32 predicate on every coordinate) and every consumer works from those intervals.
#include <cuda_fp16.h>

#include "common.cuh"

namespace lgd {

constexpr int PAINT_PIX = 32;    // pixels per block in the paint kernels

struct LevelScale {
  float rh[LGD_MAX_LEVELS];
  float rw[LGD_MAX_LEVELS];
};

// The reference's test, same fp32 operations in the same order (utils.py:67-88):
//   s1 = a*r ; s2 = b*r ; c = (s1+s2)*0.5 ; size = s2-s1 ; |c - p| / size <= 0.5
__device__ __forceinline__ bool inside_1d(float a, float b, float r, int p) {
  const float s1 = __fmul_rn(a, r), s2 = __fmul_rn(b, r);
  const float c = __fmul_rn(__fadd_rn(s1, s2), 0.5f);
  const float size = __fsub_rn(s2, s1);
  const float d = __fdiv_rn(fabsf(__fsub_rn(c, (float)p)), size);
  return d <= 0.5f;  // false for NaN / inf (zero-size boxes) exactly like torch
}

// one block per (box, level): exact solution interval of the predicate along x and along y
__global__ void box_ranges_kernel(const float* __restrict__ boxes, int T, Pyr p, LevelScale sc, int* __restrict__ ranges) {
  __shared__ int lo[2], hi[2];
  const int t = blockIdx.x, l = blockIdx.y;
  const int H = p.h[l], W = p.w[l];
  if (threadIdx.x < 2) {
    lo[threadIdx.x] = 1 << 30;
    hi[threadIdx.x] = -1;
  }
  __syncthreads();
  const float x1 = boxes[4 * t + 0], y1 = boxes[4 * t + 1], x2 = boxes[4 * t + 2], y2 = boxes[4 * t + 3];
  for (int x = threadIdx.x; x < W; x += blockDim.x)
    if (inside_1d(x1, x2, sc.rw[l], x)) {
      atomicMin(&lo[0], x);
      atomicMax(&hi[0], x);
    }
  for (int y = threadIdx.x; y < H; y += blockDim.x)
    if (inside_1d(y1, y2, sc.rh[l], y)) {
      atomicMin(&lo[1], y);
      atomicMax(&hi[1], y);
    }
  __syncthreads();
  if (threadIdx.x == 0) {
    int* r = ranges + ((long long)l * T + t) * 4;
    const bool ex = hi[0] >= 0, ey = hi[1] >= 0;
    // the predicate is monotone in |c-p| so its solution set is an interval; empty -> [0,0)
    r[0] = (ex && ey) ? lo[0] : 0;
    r[1] = (ex && ey) ? hi[0] + 1 : 0;
    r[2] = (ex && ey) ? lo[1] : 0;
    r[3] = (ex && ey) ? hi[1] + 1 : 0;
  }
}

__global__ void masks_from_ranges_kernel(const int* __restrict__ ranges, int T, Pyr p, float* __restrict__ masks) {
  const int t = blockIdx.x, l = blockIdx.y;
  const int H = p.h[l], W = p.w[l];
  const int4 r = *reinterpret_cast<const int4*>(ranges + ((long long)l * T + t) * 4);
  float* m = masks + (long long)T * p.pix_start[l] + (long long)t * H * W;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    m[i] = (x >= r.x && x < r.y && y >= r.z && y < r.w) ? 1.f : 0.f;
  }
}

// ------------------------------------------------------------------------------------ box sums
// Work decomposition: boxes differ in area by four orders of magnitude (a one-pixel box at p7, the context box =
// the whole p3 level), so the sum over a box is cut into ITEMS of about ITEM_PX pixels (whole rows of the box) and a
// persistent grid walks the item list: every block does the same amount of HBM work whatever the box mix.
//   plan:     items(e) for every entry e = l*T + t, exclusive scan -> item_start[e], item_start[L*T] = total
//   boxsum:   partial[item][c] = sum over the item's pixels of f(x[pixel, c])
//   finalize: out[e][c] = sum over the entry's items, in item order (deterministic), optional division by the area
// f = identity, or relu((x-mean)*rstd) when gn_stats is given (student_proj_2D's GroupNorm+ReLU applied on the fly,
// layers.py:22-32, so the normalised map is never written).
constexpr int ITEM_PX = 128;
constexpr int PLAN_THREADS = 1024;

__device__ __forceinline__ int rows_per_item(int bw) { return bw >= ITEM_PX ? 1 : ITEM_PX / max(bw, 1); }

__global__ void __launch_bounds__(PLAN_THREADS) boxsum_plan_kernel(const int* __restrict__ ranges, int n_entries,
                                                                   int* __restrict__ item_start) {
  __shared__ int warp_tot[PLAN_THREADS / 32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_entries; base += PLAN_THREADS) {
    const int e = base + tid;
    int n = 0;
    if (e < n_entries) {
      const int4 r = *reinterpret_cast<const int4*>(ranges + (long long)e * 4);
      const int bw = r.y - r.x, bh = r.w - r.z;
      if (bw > 0 && bh > 0) {
        const int rp = rows_per_item(bw);
        n = (bh + rp - 1) / rp;
      }
    }
    int incl = n;  // inclusive scan inside the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, w, off);
        if (lane >= off) w += v;
      }
      warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int before = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + incl - n;
    if (e < n_entries) item_start[e] = before;
    __syncthreads();
    if (tid == PLAN_THREADS - 1) carry = before + n;
    __syncthreads();
  }
  if (tid == 0) item_start[n_entries] = carry;
}

__global__ void __launch_bounds__(256, 4) boxsum_kernel(Pyr p, const float* __restrict__ x,
                                                     const float* __restrict__ gn_stats,
                                                     const int* __restrict__ ranges, const int* __restrict__ img_of,
                                                     int T, const int* __restrict__ item_start,
                                                     float* __restrict__ partial) {
  __shared__ float4 sh[4][64];
  __shared__ int s_entry;
  const int n_entries = p.num_levels * T;
  const int total = item_start[n_entries];
  const int q = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const bool norm = gn_stats != nullptr;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    if (threadIdx.x == 0) {  // last entry whose first item is <= item (entries without items share their successor's start)
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
    const int4 r = *reinterpret_cast<const int4*>(ranges + (long long)e * 4);
    const int bw = r.y - r.x;
    const int rp = rows_per_item(bw);
    const int y_begin = r.z + (item - item_start[e]) * rp;
    const int y_end = min(r.w, y_begin + rp);
    const int b = img_of[t];
    const int W = p.w[l];
    const float* base = x + p.off[l] + ((long long)b * p.h[l] * W + (long long)y_begin * W + r.x) * C + q * 4;
    float mean = 0.f, rstd = 1.f;
    if (norm) {
      mean = gn_stats[2 * (l * p.batch + b)];
      rstd = gn_stats[2 * (l * p.batch + b) + 1];
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n = (y_end - y_begin) * bw;
    const bool full_rows = bw == W;  // context boxes: the item is one contiguous run of pixels
    for (int i = sub; i < n; i += 32) {  // eight pixels in flight per thread
      float4 v[8];
      bool ok[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ii = i + 4 * j;
        ok[j] = ii < n;
        int off = ii;
        if (!full_rows) {
          const int yy = ii / bw;
          off = yy * W + (ii - yy * bw);
        }
        v[j] = ok[j] ? ldg4(base + (long long)off * C) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!ok[j]) continue;
        float4 w = v[j];
        if (norm) {
          w.x = relu_keep_nan((w.x - mean) * rstd); w.y = relu_keep_nan((w.y - mean) * rstd);
          w.z = relu_keep_nan((w.z - mean) * rstd); w.w = relu_keep_nan((w.w - mean) * rstd);
        }
        acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
      }
    }
    sh[sub][q] = acc;
    __syncthreads();
    if (sub == 0) {
      float4 a = sh[0][q];
#pragma unroll
      for (int j = 1; j < 4; ++j) { const float4 v = sh[j][q]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
      stg4(partial + (long long)item * C + q * 4, a);
    }
    __syncthreads();  // sh / s_entry are reused by the next item
  }
}

// out[l,t,c] = sum over the entry's items of partial * (divide ? 1/max(count,1) : 1); rows outside the rendered
// subset -> 0
// Four item lanes per channel (a context box on the finest level has ~100 items: one serial chain of loads per channel
// made this the long pole of the launch, 15 us); lanes and their four partial sums are combined in a fixed order.
constexpr int FIN_LANES = 4;
__global__ void __launch_bounds__(C * FIN_LANES)
boxsum_finalize_kernel(const float* __restrict__ partial, const int* __restrict__ item_start,
                       const int* __restrict__ ranges, const int* __restrict__ img_of,
                       const int* __restrict__ img_start, const int* __restrict__ n_rows, int T,
                       int divide, float* __restrict__ out) {
  __shared__ float sh[FIN_LANES][C];
  const int t = blockIdx.x, l = blockIdx.y, c = threadIdx.x % C, lane = threadIdx.x / C;
  const long long row = (long long)l * T + t;
  bool active = true;
  if (n_rows != nullptr) {
    const int b = img_of[t];
    active = (t - img_start[b]) < n_rows[b];
  }
  float s = 0.f;
  if (active) {
    const int i0 = item_start[row], i1 = item_start[row + 1];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int i = i0 + lane;
    for (; i + 3 * FIN_LANES < i1; i += 4 * FIN_LANES) {
      s0 += partial[(long long)i * C + c];
      s1 += partial[(long long)(i + FIN_LANES) * C + c];
      s2 += partial[(long long)(i + 2 * FIN_LANES) * C + c];
      s3 += partial[(long long)(i + 3 * FIN_LANES) * C + c];
    }
    for (; i < i1; i += FIN_LANES) s0 += partial[(long long)i * C + c];
    s = (s0 + s1) + (s2 + s3);
  }
  sh[lane][c] = s;
  __syncthreads();
  if (lane == 0) {
    s = (sh[0][c] + sh[1][c]) + (sh[2][c] + sh[3][c]);
    if (active && divide) {
      const int4 r = *reinterpret_cast<const int4*>(ranges + row * 4);
      const float cnt = (float)((r.y - r.x) * (r.w - r.z));
      s = s / fmaxf(cnt, 1.f);  // normalizer = max(mask.sum(), 1), dynamic_teacher.py:97-98
    }
    out[row * C + c] = s;
  }
}

// the item plan of a (level, row) interval table, also used by taprender.cu (same item cut: whole box rows of ~ITEM_PX pixels)
int boxsum_plan(const int* ranges, int n_entries, int* item_start, cudaStream_t stream) {
  boxsum_plan_kernel<<<1, PLAN_THREADS, 0, stream>>>(ranges, n_entries, item_start);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

// shared host side of the two box-sum users
static size_t boxsum_plan_bytes(int n_entries) { return (((size_t)n_entries + 1) * sizeof(int) + 255) & ~size_t(255); }

static size_t boxsum_workspace_bytes(const lgd_pyramid_t* pyr, int T) {
  // an item is at least one row of its box -> at most h[l] items per (box, level)
  size_t rows = 0;
  for (int l = 0; l < pyr->num_levels; ++l) rows += (size_t)pyr->h[l];
  return boxsum_plan_bytes(pyr->num_levels * T) + (size_t)T * rows * C * sizeof(float);
}

static int boxsum(const Pyr& p, const float* x, const float* gn_stats, const int32_t* ranges, const int32_t* img_of,
                  const int32_t* img_start, const int32_t* n_rows, int T, int divide, float* out, void* workspace,
                  cudaStream_t stream) {
  const int n_entries = p.num_levels * T;
  int* item_start = static_cast<int*>(workspace);
  float* partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + boxsum_plan_bytes(n_entries));
  int dev = 0, sms = 0;
  LGD_CUDA(cudaGetDevice(&dev));
  LGD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  boxsum_plan_kernel<<<1, PLAN_THREADS, 0, stream>>>(ranges, n_entries, item_start);
  LGD_LAUNCH_CHECK();
  boxsum_kernel<<<sms * 4, 256, 0, stream>>>(p, x, gn_stats, ranges, img_of, T, item_start, partial);
  LGD_LAUNCH_CHECK();
  boxsum_finalize_kernel<<<dim3(T, p.num_levels), C * FIN_LANES, 0, stream>>>(partial, item_start, ranges, img_of, img_start, n_rows,
                                                                 T, divide, out);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

// ------------------------------------------------------------------------------------ paint
// out[l,b,pixel,:] = sum over rows t of image b (first n_rows[b] rows if given) whose box covers the pixel of
//                    src[l,t,:] * (divide ? 1/max(count_t,1) : 1)
// Measured bounds of the first version (every thread tested every row against every one of its pixels, one strip per
// block): instruction issue (ncu: sm throughput 69 %, DRAM 20 %) and, per block, a chain of three dependent L2 round
// trips (image -> intervals -> embedding rows) with three blocks per SM to hide it. Now:
//  * a block paints PAINT_SPB consecutive strips of one (level, image) and loads that image's intervals and embedding
//    rows ONCE into shared memory (at most PAINT_STAGE rows at a time; images with more rows restage per strip);
//  * rows are rectangles, so neighbouring pixels of a strip are almost always covered by the SAME rows: per strip one
//    32-bit coverage mask per row is built from the intervals (exact, one bit range per image row the strip touches)
//    and the masks' edge bits are ORed into a "differs from its left neighbour" word; every thread owns eight
//    CONSECUTIVE pixels of one channel quad, sums the rows (in row order, the same fused multiply-adds as before: results
//    are bit-identical) only for its first pixel and for pixels whose coverage changed, and copies its left neighbour's
//    accumulator otherwise.
constexpr int PAINT_STAGE = 16;   // rows staged in shared memory at a time
constexpr int PAINT_SPB = 4;      // strips per block
constexpr int PAINT_MAX_ROWS = 1024;   // rows per image paint_kernel keeps coverage masks for (more: paint_general_kernel)
__host__ __device__ inline int paint_blocks_of_level(int hw) {
  const int strips = (hw + PAINT_PIX - 1) / PAINT_PIX;
  return (strips + PAINT_SPB - 1) / PAINT_SPB;
}
__global__ void __launch_bounds__(256, 3)
paint_general_kernel(Pyr p, const float* __restrict__ src, const int* __restrict__ ranges,
                     const int* __restrict__ img_start, const int* __restrict__ n_rows, int T, int divide,
                     float* __restrict__ out, int do_round, __half* __restrict__ out_half) {
  __shared__ float4 se[PAINT_STAGE][64];
  __shared__ int4 sr[PAINT_STAGE];
  __shared__ float ssc[PAINT_STAGE];
  __shared__ unsigned smask[PAINT_STAGE];
  __shared__ unsigned sdiff;
  // blockIdx.x walks groups of PAINT_SPB strips of all levels of image blockIdx.y (no empty blocks)
  const int b = blockIdx.y;
  int l = 0, grp = blockIdx.x;
  while (l + 1 < p.num_levels) {
    const int nbk = paint_blocks_of_level(p.h[l] * p.w[l]);
    if (grp < nbk) break;
    grp -= nbk;
    ++l;
  }
  const int H = p.h[l], W = p.w[l], HW = H * W;
  const int q = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const int t0 = img_start[b];
  const int nb = (n_rows != nullptr) ? n_rows[b] : (img_start[b + 1] - t0);
  if (nb <= PAINT_MAX_ROWS) return;   // painted by paint_kernel
  const bool stage_once = nb <= PAINT_STAGE;
  for (int s = 0; s < PAINT_SPB; ++s) {
    const int pix0 = (grp * PAINT_SPB + s) * PAINT_PIX;
    if (pix0 >= HW) break;   // block-uniform
    const int ya = pix0 / W, xa = pix0 - ya * W;
    const int npx = min(PAINT_PIX, HW - pix0);
    float4 acc[PAINT_PIX / 4];
#pragma unroll
    for (int i = 0; i < PAINT_PIX / 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned differs = 0u;   // bit j: some row so far covers pixel j and pixel j-1 differently
    for (int g0 = 0; g0 < nb || g0 == 0; g0 += PAINT_STAGE) {
      const int ng = max(0, min(PAINT_STAGE, nb - g0));
      __syncthreads();   // the previous strip / group has been consumed
      if (!stage_once || s == 0) {
#pragma unroll
        for (int jj = 0; jj < PAINT_STAGE / 4; ++jj) {
          const int j = sub + 4 * jj;
          if (j < ng) {
            se[j][q] = ldg4(src + ((long long)l * T + t0 + g0 + j) * C + q * 4);
            if (q == 0) {
              const int4 r = *reinterpret_cast<const int4*>(ranges + ((long long)l * T + t0 + g0 + j) * 4);
              sr[j] = r;
              ssc[j] = divide ? 1.f / fmaxf((float)((r.y - r.x) * (r.w - r.z)), 1.f) : 1.f;
            }
          }
        }
        __syncthreads();
      }
      if (threadIdx.x < 32) {
        unsigned m = 0u;
        if (threadIdx.x < ng) {
          const int4 r = sr[threadIdx.x];
          if (r.y > r.x && r.w > ya && r.z <= (pix0 + npx - 1) / W) {
            // the strip is a few runs of consecutive pixels of consecutive image rows: one bit range per run
            int x = xa, y = ya;
            for (int j0 = 0; j0 < npx; ++y) {
              const int run = min(W - x, npx - j0);
              const int lo = max(r.x, x), hi = min(r.y, x + run);
              if (y >= r.z && y < r.w && hi > lo)
                m |= (hi - lo >= 32 ? 0xffffffffu : (1u << (hi - lo)) - 1u) << (j0 + lo - x);
              j0 += run;
              x = 0;
            }
          }
          smask[threadIdx.x] = m;
        }
        const unsigned d = __reduce_or_sync(0xffffffffu, (m ^ (m << 1)) & ~1u);
        if (threadIdx.x == 0) sdiff = d;
      }
      __syncthreads();
      differs |= sdiff;
#pragma unroll
      for (int i = 0; i < PAINT_PIX / 4; ++i) {
        const int j = sub * (PAINT_PIX / 4) + i;
        if (i == 0 || (differs >> j & 1u)) {
          for (int kk = 0; kk < ng; ++kk) {
            if (smask[kk] >> j & 1u) {
              const float4 e = se[kk][q];
              const float sc = ssc[kk];
              acc[i].x += e.x * sc; acc[i].y += e.y * sc; acc[i].z += e.z * sc; acc[i].w += e.w * sc;
            }
          }
        } else {
          acc[i] = acc[i - 1];
        }
      }
    }
    float* o = out ? out + p.off[l] + (long long)b * HW * C + q * 4 : nullptr;
#pragma unroll
    for (int i = 0; i < PAINT_PIX / 4; ++i) {
      const int pix = pix0 + sub * (PAINT_PIX / 4) + i;
      if (pix < HW) {
        float4 v = acc[i];
        if (out_half != nullptr) {
          const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
          uint2 hv;
          hv.x = *reinterpret_cast<const uint32_t*>(&h0);
          hv.y = *reinterpret_cast<const uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(out_half + p.off[l] + ((long long)b * HW + pix) * C + q * 4) = hv;
        }
        if (do_round) { v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w); }
        if (out != nullptr) stg4(o + (long long)pix * C, v);
      }
    }
  }
}

// The kernel every step runs (images with at most PAINT_MAX_ROWS rows): same method, but the coverage masks of ALL rows of
// the image are kept in shared memory, so a pixel's sum is formed in one go in ONE accumulator and stored at once -- no
// per-pixel accumulator array (40 instead of 80 registers: twice the resident blocks) and no copies between accumulators.
// The first PAINT_STAGE embedding rows are read from shared memory, later ones through L1.
__global__ void __launch_bounds__(256, 5)
paint_kernel(Pyr p, const float* __restrict__ src, const int* __restrict__ ranges,
             const int* __restrict__ img_start, const int* __restrict__ n_rows, int T, int divide,
             float* __restrict__ out, int do_round, __half* __restrict__ out_half) {
  __shared__ float4 se[PAINT_STAGE][64];
  __shared__ float ssc[PAINT_MAX_ROWS];
  __shared__ unsigned smask[PAINT_MAX_ROWS];
  __shared__ unsigned sdiff[2];
  const int b = blockIdx.y;
  int l = 0, grp = blockIdx.x;
  while (l + 1 < p.num_levels) {
    const int nbk = paint_blocks_of_level(p.h[l] * p.w[l]);
    if (grp < nbk) break;
    grp -= nbk;
    ++l;
  }
  const int H = p.h[l], W = p.w[l], HW = H * W;
  const int q = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const int t0 = img_start[b];
  const int nb = (n_rows != nullptr) ? n_rows[b] : (img_start[b + 1] - t0);
  if (nb > PAINT_MAX_ROWS) return;   // painted by paint_general_kernel
  const int* rg = ranges + ((long long)l * T + t0) * 4;
  const float* sb = src + ((long long)l * T + t0) * C + q * 4;
  // once per block: the image's scales, its first embedding rows, this thread's first interval
  int4 r0 = make_int4(0, 0, 0, 0);
  for (int k = threadIdx.x; k < nb; k += 256) {
    const int4 r = *reinterpret_cast<const int4*>(rg + 4 * k);
    if (k == threadIdx.x) r0 = r;
    ssc[k] = divide ? 1.f / fmaxf((float)((r.y - r.x) * (r.w - r.z)), 1.f) : 1.f;
  }
#pragma unroll
  for (int jj = 0; jj < PAINT_STAGE / 4; ++jj) {
    const int j = sub + 4 * jj;
    if (j < nb) se[j][q] = ldg4(sb + (long long)j * C);
  }
  if (threadIdx.x < 2) sdiff[threadIdx.x] = 0u;
  const long long obase = p.off[l] + (long long)b * HW * C + q * 4;
  for (int s = 0; s < PAINT_SPB; ++s) {
    const int pix0 = (grp * PAINT_SPB + s) * PAINT_PIX;
    if (pix0 >= HW) break;   // block-uniform
    const int ya = pix0 / W, xa = pix0 - ya * W;
    const int npx = min(PAINT_PIX, HW - pix0);
    const int yb = (pix0 + npx - 1) / W;
    __syncthreads();   // staging done / the previous strip's masks have been consumed
    {
      unsigned edges = 0u;
      for (int k = threadIdx.x; k < nb; k += 256) {
        const int4 r = k == threadIdx.x ? r0 : *reinterpret_cast<const int4*>(rg + 4 * k);
        unsigned m = 0u;
        if (r.y > r.x && r.w > ya && r.z <= yb) {
          // the strip is a few runs of consecutive pixels of consecutive image rows: one bit range per run
          int x = xa, y = ya;
          for (int j0 = 0; j0 < npx; ++y) {
            const int run = min(W - x, npx - j0);
            const int lo = max(r.x, x), hi = min(r.y, x + run);
            if (y >= r.z && y < r.w && hi > lo)
              m |= (hi - lo >= 32 ? 0xffffffffu : (1u << (hi - lo)) - 1u) << (j0 + lo - x);
            j0 += run;
            x = 0;
          }
        }
        smask[k] = m;
        edges |= (m ^ (m << 1)) & ~1u;
      }
      edges = __reduce_or_sync(0xffffffffu, edges);
      if ((threadIdx.x & 31) == 0 && edges != 0u) atomicOr(&sdiff[s & 1], edges);   // integer OR: order-free
    }
    __syncthreads();
    const unsigned differs = sdiff[s & 1];   // bit j: some row covers pixel j and pixel j-1 differently
    if (threadIdx.x == 0) sdiff[(s + 1) & 1] = 0u;   // its readers finished before this strip's first barrier
    const int j_first = sub * (PAINT_PIX / 4);
    float* o32 = out ? out + obase + (long long)(pix0 + j_first) * C : nullptr;
    __half* o16 = out_half ? out_half + obase + (long long)(pix0 + j_first) * C : nullptr;
    float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < PAINT_PIX / 4; ++i) {
      const int j = j_first + i;
      if (j >= npx) break;   // warp-uniform
      if (i == 0 || (differs >> j & 1u)) {
        cur = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kk = 0; kk < nb; ++kk) {
          if (smask[kk] >> j & 1u) {
            const float4 e = kk < PAINT_STAGE ? se[kk][q] : ldg4(sb + (long long)kk * C);
            const float sc = ssc[kk];
            cur.x += e.x * sc; cur.y += e.y * sc; cur.z += e.z * sc; cur.w += e.w * sc;
          }
        }
      }
      if (o16 != nullptr) {
        const __half2 h0 = __floats2half2_rn(cur.x, cur.y), h1 = __floats2half2_rn(cur.z, cur.w);
        uint2 hv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h0);
        hv.y = *reinterpret_cast<const uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(o16 + i * C) = hv;
      }
      if (o32 != nullptr) {
        float4 v = cur;
        if (do_round) { v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w); }
        stg4(o32 + i * C, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ rasterised masks (LOAD_LABELMAP)
// The Mask R-CNN recipe pools and renders with rasterised polygon masks (dynamic_teacher/utils.py:92-132) instead of
// box masks: arbitrary bitmaps, so the interval machinery above does not apply. The masks are the reference's own
// float 0/1 tensors (level l at T*pix_start[l], row t = h_l*w_l floats -- the layout of lgd_masks_from_ranges).
constexpr int DM_CHUNK = 512;   // pixels per block of the gather

// partial[((l*T+t)*nch + chunk)*C + c] = sum over the chunk's pixels with mask != 0 of f(x[pixel, c]);
// cnt_partial[(l*T+t)*nch + chunk] = number of such pixels. f = identity or relu((x-mean)*rstd).
__global__ void __launch_bounds__(256)
mask_gather_kernel(Pyr p, const float* __restrict__ x, const float* __restrict__ gn_stats,
                   const float* __restrict__ masks, const int* __restrict__ img_of, const int* __restrict__ img_start,
                   const int* __restrict__ n_rows, int T, int nch, float* __restrict__ partial,
                   float* __restrict__ cnt_partial) {
  const int chunk = blockIdx.x, t = blockIdx.y, l = blockIdx.z;
  const int HW = p.h[l] * p.w[l];
  const int c = threadIdx.x;
  const int b = img_of[t];
  float acc = 0.f, cnt = 0.f;
  const bool active = n_rows == nullptr || (t - img_start[b]) < n_rows[b];
  const int p0 = chunk * DM_CHUNK;
  if (active && p0 < HW) {
    const float* m = masks + (long long)T * p.pix_start[l] + (long long)t * HW;
    const float* xb = x + p.off[l] + (long long)b * HW * C + c;
    float mean = 0.f, rstd = 1.f;
    if (gn_stats != nullptr) {
      mean = gn_stats[2 * (l * p.batch + b)];
      rstd = gn_stats[2 * (l * p.batch + b) + 1];
    }
    const int p1 = min(p0 + DM_CHUNK, HW);
    for (int q = p0; q < p1; ++q) {
      const float mv = __ldg(m + q);   // warp-uniform
      if (mv != 0.f) {
        float v = __ldg(xb + (long long)q * C);
        if (gn_stats != nullptr) v = relu_keep_nan((v - mean) * rstd);
        acc = fmaf(mv, v, acc);
        cnt += mv;
      }
    }
  }
  partial[((long long)(l * T + t) * nch + chunk) * C + c] = acc;
  if (c == 0) cnt_partial[(long long)(l * T + t) * nch + chunk] = cnt;
}

// out[(l*T+t)*C + c] = sum of the chunk partials (fixed order) [/ max(count, 1)]; count[l*T+t] (optional)
__global__ void mask_gather_finalize_kernel(Pyr p, const float* __restrict__ partial, const float* __restrict__ cnt_partial,
                                            int T, int nch, int divide, float* __restrict__ out,
                                            float* __restrict__ count) {
  const int t = blockIdx.x, l = blockIdx.y, c = threadIdx.x;
  const int HW = p.h[l] * p.w[l];
  const int n = (HW + DM_CHUNK - 1) / DM_CHUNK;
  float s = 0.f, cnt = 0.f;
  for (int i = 0; i < n; ++i) {
    s += partial[((long long)(l * T + t) * nch + i) * C + c];
    cnt += cnt_partial[(long long)(l * T + t) * nch + i];
  }
  if (divide) s = s / fmaxf(cnt, 1.f);
  out[(long long)(l * T + t) * C + c] = s;
  if (count != nullptr && c == 0) count[l * T + t] = cnt;
}

// out[l,b,pixel,:] = sum over rows t of image b (first n_rows[b] rows if given) of mask[l,t,pixel] * src[l,t,:]
//                    * (count != NULL ? 1/max(count[l,t],1) : 1)
__global__ void __launch_bounds__(256)
mask_paint_kernel(Pyr p, const float* __restrict__ src, const float* __restrict__ masks,
                  const int* __restrict__ img_start, const int* __restrict__ n_rows, const float* __restrict__ count,
                  int T, float* __restrict__ out, __half* __restrict__ out_half) {
  const int b = blockIdx.y;
  int l = 0, strip = blockIdx.x;
  while (l + 1 < p.num_levels) {
    const int ns = (p.h[l] * p.w[l] + PAINT_PIX - 1) / PAINT_PIX;
    if (strip < ns) break;
    strip -= ns;
    ++l;
  }
  const int HW = p.h[l] * p.w[l];
  const int pix0 = strip * PAINT_PIX;
  if (pix0 >= HW) return;
  const int c = threadIdx.x;
  const int t0 = img_start[b];
  const int nb = (n_rows != nullptr) ? n_rows[b] : (img_start[b + 1] - t0);
  float acc[PAINT_PIX];
#pragma unroll
  for (int i = 0; i < PAINT_PIX; ++i) acc[i] = 0.f;
  for (int k = 0; k < nb; ++k) {
    const int t = t0 + k;
    const float* m = masks + (long long)T * p.pix_start[l] + (long long)t * HW + pix0;
    float e = __ldg(src + (long long)(l * T + t) * C + c);
    if (count != nullptr) e = e / fmaxf(count[l * T + t], 1.f);
#pragma unroll
    for (int i = 0; i < PAINT_PIX; ++i) {
      if (pix0 + i < HW) {
        const float mv = __ldg(m + i);   // warp-uniform
        if (mv != 0.f) acc[i] = fmaf(mv, e, acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < PAINT_PIX; ++i) {
    if (pix0 + i < HW) {
      const long long o = p.off[l] + ((long long)b * HW + pix0 + i) * C + c;
      if (out != nullptr) out[o] = acc[i];
      if (out_half != nullptr) out_half[o] = __float2half_rn(acc[i]);
    }
  }
}

__global__ void bytes_to_float_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] ? 1.f : 0.f;
}

// ------------------------------------------------------------------------------------ descriptors
__global__ void encode_desc_kernel(const float* __restrict__ boxes, const int* __restrict__ labels, int T, float fw,
                                   float fh, float* __restrict__ desc) {
  const int t = blockIdx.x, j = threadIdx.x;
  if (j >= LGD_DESC_DIM) return;
  float v;
  if (j < 4) {
    v = __fdiv_rn(boxes[4 * t + j], (j & 1) ? fh : fw);  // bboxes[:, [0,2]] /= img_w ; [1,3] /= img_h
  } else {
    v = (labels[t] == j - 4) ? 1.f : 0.f;
  }
  // range_scaling (utils.py:16-24): (b-a)/(Max-Min) * (x-Min) + a with a=-1,b=1,Min=0,Max=1
  desc[(long long)t * LGD_DESC_DIM + j] = __fadd_rn(__fmul_rn(2.0f, __fsub_rn(v, 0.0f)), -1.0f);
}

}  // namespace lgd

using namespace lgd;

// descriptors with the 49 mask dimensions of LOAD_LABELMAP (label_encoder.py:31-32,101-103): (T, 133)
__global__ void encode_desc_mask_kernel(const float* __restrict__ boxes, const int* __restrict__ labels,
                                        const float* __restrict__ mask49, int T, float fw, float fh,
                                        float* __restrict__ desc) {
  const int t = blockIdx.x, j = threadIdx.x;
  constexpr int D = LGD_DESC_DIM + 49;
  if (j >= D) return;
  float v;
  if (j < 4) {
    v = __fdiv_rn(boxes[4 * t + j], (j & 1) ? fh : fw);
  } else if (j < LGD_DESC_DIM) {
    v = (labels[t] == j - 4) ? 1.f : 0.f;
  } else {
    v = mask49[(long long)t * 49 + (j - LGD_DESC_DIM)];
  }
  desc[(long long)t * D + j] = __fadd_rn(__fmul_rn(2.0f, __fsub_rn(v, 0.0f)), -1.0f);
}

extern "C" int lgd_encode_descriptors_masks(const float* boxes, const int32_t* labels, const float* mask49, int T,
                                            int img_h, int img_w, float* desc, void* stream) {
  LGD_CHECK_ARG(boxes && labels && mask49 && desc && T > 0 && img_h > 0 && img_w > 0,
                "lgd_encode_descriptors_masks: bad arguments");
  encode_desc_mask_kernel<<<T, 160, 0, (cudaStream_t)stream>>>(boxes, labels, mask49, T, (float)img_w, (float)img_h, desc);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

// CATEGORY_FORMAT norm_classes (label_encoder.py:24-25,91-93): (T, 5 | 54) = boxes/W,H | class index / num_classes
// [| mask49], scaled to [-1,1]. labels[t] < 0 (the dummy row of an image without GT) encodes as class value 0.
__global__ void encode_desc_norm_kernel(const float* __restrict__ boxes, const int* __restrict__ labels,
                                        const float* __restrict__ mask49, int T, float fw, float fh, float ncls,
                                        float* __restrict__ desc) {
  const int t = blockIdx.x, j = threadIdx.x;
  const int D = 5 + (mask49 != nullptr ? 49 : 0);
  if (j >= D) return;
  float v;
  if (j < 4) {
    v = __fdiv_rn(boxes[4 * t + j], (j & 1) ? fh : fw);
  } else if (j == 4) {
    v = __fdiv_rn((float)max(labels[t], 0), ncls);   // labels / num_classes: int64 / int -> fp32 division in torch
  } else {
    v = mask49[(long long)t * 49 + (j - 5)];
  }
  desc[(long long)t * D + j] = __fadd_rn(__fmul_rn(2.0f, __fsub_rn(v, 0.0f)), -1.0f);
}

extern "C" int lgd_encode_descriptors_norm(const float* boxes, const int32_t* labels, const float* mask49, int T,
                                           int img_h, int img_w, int num_classes, float* desc, void* stream) {
  LGD_CHECK_ARG(boxes && labels && desc && T > 0 && img_h > 0 && img_w > 0 && num_classes > 0,
                "lgd_encode_descriptors_norm: bad arguments");
  encode_desc_norm_kernel<<<T, 64, 0, (cudaStream_t)stream>>>(boxes, labels, mask49, T, (float)img_w, (float)img_h,
                                                               (float)num_classes, desc);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_masks_from_bytes(const uint8_t* bytes, int64_t n, float* masks, void* stream) {
  LGD_CHECK_ARG(bytes && masks && n > 0, "lgd_masks_from_bytes: bad arguments");
  bytes_to_float_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bytes, masks, (long long)n);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

static int dm_chunks(const Pyr& p) {
  int n = 1;
  for (int l = 0; l < p.num_levels; ++l) n = max(n, (p.h[l] * p.w[l] + DM_CHUNK - 1) / DM_CHUNK);
  return n;
}

extern "C" size_t lgd_dense_mask_workspace(const lgd_pyramid_t* pyr, int T) {
  Pyr p;
  if (make_pyr(pyr, &p) != LGD_OK || T <= 0) return 0;
  return (size_t)p.num_levels * T * dm_chunks(p) * (C + 1) * sizeof(float);
}

extern "C" int lgd_mask_gather(const lgd_pyramid_t* pyr, const float* x, const float* gn_stats, const float* masks,
                               const int32_t* img_of, const int32_t* img_start, const int32_t* n_rows, int T, int divide,
                               float* out, float* count, void* workspace, size_t workspace_bytes, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(x && masks && img_of && img_start && out && workspace && T > 0, "lgd_mask_gather: bad arguments");
  LGD_CHECK_ARG(workspace_bytes >= lgd_dense_mask_workspace(pyr, T), "lgd_mask_gather: workspace too small");
  const int nch = dm_chunks(p);
  float* partial = static_cast<float*>(workspace);
  float* cnt_partial = partial + (size_t)p.num_levels * T * nch * C;
  mask_gather_kernel<<<dim3(nch, T, p.num_levels), C, 0, (cudaStream_t)stream>>>(p, x, gn_stats, masks, img_of, img_start,
                                                                               n_rows, T, nch, partial, cnt_partial);
  LGD_LAUNCH_CHECK();
  mask_gather_finalize_kernel<<<dim3(T, p.num_levels), C, 0, (cudaStream_t)stream>>>(p, partial, cnt_partial, T, nch,
                                                                                    divide, out, count);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_mask_paint(const lgd_pyramid_t* pyr, const float* src, const float* masks, const int32_t* img_start,
                              const int32_t* n_rows, const float* count, int T, float* out, void* out_half,
                              void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(src && masks && img_start && (out || out_half) && T > 0, "lgd_mask_paint: bad arguments");
  int strips = 0;
  for (int l = 0; l < p.num_levels; ++l) strips += (p.h[l] * p.w[l] + PAINT_PIX - 1) / PAINT_PIX;
  mask_paint_kernel<<<dim3(strips, p.batch), C, 0, (cudaStream_t)stream>>>(p, src, masks, img_start, n_rows, count, T, out,
                                                                         static_cast<__half*>(out_half));
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_encode_descriptors(const float* boxes, const int32_t* labels, int T, int img_h, int img_w,
                                      float* desc, void* stream) {
  LGD_CHECK_ARG(boxes && labels && desc && T > 0 && img_h > 0 && img_w > 0, "lgd_encode_descriptors: bad arguments");
  encode_desc_kernel<<<T, 96, 0, (cudaStream_t)stream>>>(boxes, labels, T, (float)img_w, (float)img_h, desc);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_box_ranges(const float* boxes, int T, int img_h, int img_w, const lgd_pyramid_t* pyr,
                              int32_t* ranges, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(boxes && ranges && T > 0 && img_h > 0 && img_w > 0, "lgd_box_ranges: bad arguments");
  LevelScale sc;
  for (int l = 0; l < LGD_MAX_LEVELS; ++l) {
    // r_h, r_w = dst.h / src.h, dst.w / src.w in double, used as an fp32 scalar (utils.py:67-72)
    sc.rh[l] = l < p.num_levels ? (float)((double)p.h[l] / (double)img_h) : 0.f;
    sc.rw[l] = l < p.num_levels ? (float)((double)p.w[l] / (double)img_w) : 0.f;
  }
  box_ranges_kernel<<<dim3(T, p.num_levels), 128, 0, (cudaStream_t)stream>>>(boxes, T, p, sc, ranges);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_masks_from_ranges(const int32_t* ranges, int T, const lgd_pyramid_t* pyr, float* masks,
                                     void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(ranges && masks && T > 0, "lgd_masks_from_ranges: bad arguments");
  masks_from_ranges_kernel<<<dim3(T, p.num_levels), 256, 0, (cudaStream_t)stream>>>(ranges, T, p, masks);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" size_t lgd_maskpool_workspace(const lgd_pyramid_t* pyr, int T) {
  return boxsum_workspace_bytes(pyr, T);
}

extern "C" int lgd_maskpool_fwd(const lgd_pyramid_t* pyr, const float* x, const float* gn_stats, const int32_t* ranges,
                                const int32_t* img_of, int T, float* pooled, void* workspace, size_t workspace_bytes,
                                void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(x && ranges && img_of && pooled && workspace && T > 0, "lgd_maskpool_fwd: bad arguments");
  LGD_CHECK_ARG(workspace_bytes >= lgd_maskpool_workspace(pyr, T), "lgd_maskpool_fwd: workspace too small");
  return boxsum(p, x, gn_stats, ranges, img_of, nullptr, nullptr, T, 1, pooled, workspace, (cudaStream_t)stream);
}

static int paint(const Pyr& p, const float* src, const int32_t* ranges, const int32_t* img_start, const int32_t* n_rows,
                 int T, int divide, float* out, int round_out, void* out_half, void* stream) {
  int groups = 0;
  for (int l = 0; l < p.num_levels; ++l) groups += paint_blocks_of_level(p.h[l] * p.w[l]);
  dim3 grid(groups, p.batch);
  paint_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, src, ranges, img_start, n_rows, T, divide, out, round_out,
                                                       static_cast<__half*>(out_half));
  LGD_LAUNCH_CHECK();
  if (T > PAINT_MAX_ROWS) {   // only then can an image have more rows than paint_kernel keeps masks for
    paint_general_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, src, ranges, img_start, n_rows, T, divide, out,
                                                                 round_out, static_cast<__half*>(out_half));
    LGD_LAUNCH_CHECK();
  }
  return LGD_OK;
}

extern "C" int lgd_maskpool_bwd(const lgd_pyramid_t* pyr, const float* gpooled, const int32_t* ranges,
                                const int32_t* img_start, int T, float* gy, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(gpooled && ranges && img_start && gy && T > 0, "lgd_maskpool_bwd: bad arguments");
  return paint(p, gpooled, ranges, img_start, nullptr, T, 1, gy, 0, nullptr, stream);
}

extern "C" int lgd_render_fwd(const lgd_pyramid_t* pyr, const float* emb, const int32_t* ranges,
                              const int32_t* img_start, const int32_t* n_render, int T, float* out, int round_out,
                              void* out_half, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(emb && ranges && img_start && n_render && (out || out_half) && T > 0, "lgd_render_fwd: bad arguments");
  return paint(p, emb, ranges, img_start, n_render, T, 0, out, round_out, out_half, stream);
}

extern "C" int lgd_render_bwd(const lgd_pyramid_t* pyr, const float* gout, const int32_t* ranges, const int32_t* img_of,
                              const int32_t* img_start, const int32_t* n_render, int T, float* gemb, void* workspace,
                              size_t workspace_bytes, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(gout && ranges && img_of && img_start && n_render && gemb && workspace && T > 0,
                "lgd_render_bwd: bad arguments");
  LGD_CHECK_ARG(workspace_bytes >= lgd_maskpool_workspace(pyr, T), "lgd_render_bwd: workspace too small");
  return boxsum(p, gout, nullptr, ranges, img_of, img_start, n_render, T, 0, gemb, workspace, (cudaStream_t)stream);
}
