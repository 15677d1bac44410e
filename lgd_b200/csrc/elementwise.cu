// HBM-bound kernels of the hot path: layout movers, GroupNorm(1 group) apply / backward (K2),
// InstanceNorm statistics + MSE forward / backward (K8), channel sums. All reductions are two-stage and
// deterministic (no floating-point atomics): stage 1 writes per-block partials, stage 2 sums them in a
// fixed order in double precision.
// Reference: layers.py:6-7 (GroupNorm 1 group, no affine), base_distillator.py:16-17,59-64.
#include <cuda_fp16.h>

#include <atomic>

#include "common.cuh"

namespace lgd {

// ------------------------------------------------------------------------------------ api basics
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------ layout movers
// One block = 32 pixels x all 256 channels of one image: 256 coalesced 128-byte row reads in flight per block, then
// 32 contiguous 1 KiB pixel rows written (or the reverse). tile pitch 33 keeps both phases bank-conflict free.
// src: (256, HW) of one image (NCHW plane), dst: (HW, 256).
// strips of 32 pixels of every level, numbered level after level: strip s -> level, first pixel
struct MoverLevels {
  const float* src[LGD_MAX_LEVELS];   // NCHW side of each level
  float* dst[LGD_MAX_LEVELS];
  int strip_start[LGD_MAX_LEVELS + 1];
};
__device__ __forceinline__ int mover_level(const MoverLevels& m, const Pyr& p, int strip, int& p0) {
  int l = 0;
  while (l + 1 < p.num_levels && strip >= m.strip_start[l + 1]) ++l;
  p0 = (strip - m.strip_start[l]) * 32;
  return l;
}

// ONE launch for the whole pyramid: grid = (strips of all levels, batch)
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(MoverLevels m, Pyr py, float* __restrict__ dst_pyr, int do_round, __half* __restrict__ dst_half_pyr) {
  __shared__ float tile[C][33];
  int p0;
  const int l = mover_level(m, py, blockIdx.x, p0);
  const int HW = py.h[l] * py.w[l];
  const int b = blockIdx.y;
  const float* s = m.src[l] + (long long)b * C * HW;
  float* d = dst_pyr ? dst_pyr + py.off[l] + (long long)b * C * HW : nullptr;
  __half* dh = dst_half_pyr ? dst_half_pyr + py.off[l] + (long long)b * C * HW : nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = p0 + lane;
#pragma unroll 8
  for (int i = 0; i < 32; ++i) {
    const int c = warp + 8 * i;
    tile[c][lane] = (p < HW) ? __ldg(s + (long long)c * HW + p) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pp = warp + 8 * j;
    if (p0 + pp < HW) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = tile[k * 32 + lane][pp];
        if (dh) dh[(long long)(p0 + pp) * C + k * 32 + lane] = __float2half_rn(v);
        if (do_round) v = tf32_rna(v);
        if (d) d[(long long)(p0 + pp) * C + k * 32 + lane] = v;
      }
    }
  }
}

// src: (HW, 256) -> dst: (256, HW), optional accumulate; one launch for the whole pyramid
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const float* __restrict__ src_pyr, MoverLevels m, Pyr py, int accumulate) {
  __shared__ float tile[C][33];
  int p0;
  const int l = mover_level(m, py, blockIdx.x, p0);
  const int HW = py.h[l] * py.w[l];
  const int b = blockIdx.y;
  const float* s = src_pyr + py.off[l] + (long long)b * C * HW;
  float* d = m.dst[l] + (long long)b * C * HW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pp = warp + 8 * j;
    if (p0 + pp < HW) {
      const float* in = s + (long long)(p0 + pp) * C;
#pragma unroll
      for (int k = 0; k < 8; ++k) tile[k * 32 + lane][pp] = __ldg(in + k * 32 + lane);
    }
  }
  __syncthreads();
  const int p = p0 + lane;
  if (p < HW) {
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const int c = warp + 8 * i;
      float* q = d + (long long)c * HW + p;
      const float v = tile[c][lane];
      *q = accumulate ? *q + v : v;
    }
  }
}

// ------------------------------------------------------------------------------------ segment helpers
// A "segment" is one (level, image) block of a pyramid buffer: h*w pixels x 256 channels, contiguous.
__device__ __forceinline__ void segment_of(const Pyr& p, int seg, int& l, int& b, long long& base, int& npix) {
  l = seg / p.batch;
  b = seg - l * p.batch;
  npix = p.h[l] * p.w[l];
  base = p.off[l] + (long long)b * npix * C;
}

// ------------------------------------------------------------------------------------ GroupNorm
// stats[(l*B+b)*2] = {mean, rstd} from the conv epilogue's per-tile (sum, sumsq)
__global__ void gn_finalize_kernel(Pyr p, const float* __restrict__ tile_stats, float* __restrict__ stats) {
  const int seg = blockIdx.x;
  int l = seg / p.batch, b = seg - l * p.batch;
  int tile_start = 0;
  for (int j = 0; j < l; ++j) tile_start += tiles_per_level(p.h[j], p.w[j], p.batch);
  const int per_img = tiles_per_image(p.h[l], p.w[l]);
  const float* ts = tile_stats + 2ll * (tile_start + b * per_img);
  double s = 0.0, ss = 0.0;
  for (int i = threadIdx.x; i < per_img; i += 32) {
    s += (double)ts[2 * i];
    ss += (double)ts[2 * i + 1];
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  if (threadIdx.x == 0) {
    const double n = (double)p.h[l] * p.w[l] * C;
    const double mean = s / n;
    double var = ss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[2 * seg + 0] = (float)mean;
    stats[2 * seg + 1] = (float)(1.0 / sqrt(var + (double)EPS));
  }
}

// in_partial (optional): [seg][block][2][256] per-block sums of (d, d^2), d = y - y[first pixel of the segment], of the
// STORED values per channel -- the InstanceNorm statistics of y (base_distillator.py:60) come out of the pass that
// writes y. Every thread owns one channel quad (see gn_bwd_apply_kernel).
template <bool STATS>
__global__ void gn_apply_kernel(Pyr p, const float* __restrict__ x, const float* __restrict__ stats,
                                float* __restrict__ y, int relu, int do_round, float* __restrict__ in_partial,
                                __half* __restrict__ y_half) {
  __shared__ float4 shc[STATS ? 2 : 1][4][STATS ? 64 : 1];
  const int seg = blockIdx.y;
  int l, b, npix;
  long long base;
  segment_of(p, seg, l, b, base, npix);
  const float mean = stats[2 * seg], rstd = stats[2 * seg + 1];
  const long long n4 = (long long)npix * C / 4;
  const float4* xs = reinterpret_cast<const float4*>(x + base);
  float4* ys = y ? reinterpret_cast<float4*>(y + base) : nullptr;
  float4 sh = make_float4(0.f, 0.f, 0.f, 0.f), a = sh, c = sh;
  if (STATS) {  // the shift is exactly the value stored for the first pixel of this thread's channels
    sh = __ldg(xs + (threadIdx.x & 63));
    sh.x = (sh.x - mean) * rstd; sh.y = (sh.y - mean) * rstd; sh.z = (sh.z - mean) * rstd; sh.w = (sh.w - mean) * rstd;
    if (relu) { sh.x = relu_keep_nan(sh.x); sh.y = relu_keep_nan(sh.y); sh.z = relu_keep_nan(sh.z); sh.w = relu_keep_nan(sh.w); }
    if (do_round) { sh.x = tf32_rna(sh.x); sh.y = tf32_rna(sh.y); sh.z = tf32_rna(sh.z); sh.w = tf32_rna(sh.w); }
  }
  auto one = [&](long long i, float4 v) {
    v.x = (v.x - mean) * rstd; v.y = (v.y - mean) * rstd; v.z = (v.z - mean) * rstd; v.w = (v.w - mean) * rstd;
    if (relu) { v.x = relu_keep_nan(v.x); v.y = relu_keep_nan(v.y); v.z = relu_keep_nan(v.z); v.w = relu_keep_nan(v.w); }
    if (y_half != nullptr) {  // fp16 copy of the un-rounded value: operand of the next forward convolution
      const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
      uint2 hv;
      hv.x = *reinterpret_cast<const uint32_t*>(&h0);
      hv.y = *reinterpret_cast<const uint32_t*>(&h1);
      if (relu) {  // the copy doubles as the activation pattern of the backward: a positive value never becomes 0
        if (v.x > 0.f && (hv.x & 0xffffu) == 0) hv.x |= 1u;
        if (v.y > 0.f && (hv.x >> 16) == 0) hv.x |= 0x10000u;
        if (v.z > 0.f && (hv.y & 0xffffu) == 0) hv.y |= 1u;
        if (v.w > 0.f && (hv.y >> 16) == 0) hv.y |= 0x10000u;
      }
      reinterpret_cast<uint2*>(y_half + base)[i] = hv;
    }
    if (do_round) { v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w); }
    if (ys) ys[i] = v;
    if (STATS) {
      const float dx = v.x - sh.x, dy = v.y - sh.y, dz = v.z - sh.z, dw = v.w - sh.w;
      a.x += dx; a.y += dy; a.z += dz; a.w += dw;
      c.x += dx * dx; c.y += dy * dy; c.z += dz * dz; c.w += dw * dw;
    }
  };
  // four loads in flight per thread (HBM-bound: memory-level parallelism), on 16 KiB of consecutive addresses per block
  // and iteration; every offset is a multiple of 64 float4, so a thread keeps its channel quad
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  long long i = (long long)blockIdx.x * blockDim.x * 4 + threadIdx.x;
  for (; i + 3 * 256 < n4; i += stride) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __ldg(xs + i + j * 256);
#pragma unroll
    for (int j = 0; j < 4; ++j) one(i + j * 256, v[j]);
  }
  for (int j = 0; j < 4; ++j)   // the last, partial chunk of the segment
    if (i + j * 256 < n4) one(i + j * 256, __ldg(xs + i + j * 256));
  if (STATS) {
    const int q = threadIdx.x & 63, sub = threadIdx.x >> 6;
    shc[0][sub][q] = a;
    shc[1][sub][q] = c;
    __syncthreads();
    if (sub == 0) {
      float4 r0 = shc[0][0][q], r1 = shc[1][0][q];
#pragma unroll
      for (int j = 1; j < 4; ++j) {
        const float4 t0 = shc[0][j][q], t1 = shc[1][j][q];
        r0.x += t0.x; r0.y += t0.y; r0.z += t0.z; r0.w += t0.w;
        r1.x += t1.x; r1.y += t1.y; r1.z += t1.z; r1.w += t1.w;
      }
      float* o = in_partial + ((long long)seg * gridDim.x + blockIdx.x) * 2 * C;
      stg4(o + q * 4, r0);
      stg4(o + C + q * 4, r1);
    }
  }
}

constexpr int GN_BWD_PARTS = 3;  // per-block partials of the GroupNorm backward: sum g, sum g*xhat, sum g^2
// stage 1 of GN backward: per block partial sums of g and g*xhat (g = gy masked by relu)
__global__ void gn_bwd_sums_kernel(Pyr p, const float* __restrict__ gy, const float* __restrict__ x,
                                   const float* __restrict__ stats, int relu, double* __restrict__ partial) {
  __shared__ double red[32];
  const int seg = blockIdx.y;
  int l, b, npix;
  long long base;
  segment_of(p, seg, l, b, base, npix);
  const float mean = stats[2 * seg], rstd = stats[2 * seg + 1];
  const long long n4 = (long long)npix * C / 4;
  const float4* xs = reinterpret_cast<const float4*>(x + base);
  const float4* gs = reinterpret_cast<const float4*>(gy + base);
  float s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 xv = __ldg(xs + i);
    float4 g = __ldg(gs + i);
    const float h0 = (xv.x - mean) * rstd, h1 = (xv.y - mean) * rstd, h2 = (xv.z - mean) * rstd, h3 = (xv.w - mean) * rstd;
    if (relu) {
      g.x = h0 > 0.f ? g.x : 0.f; g.y = h1 > 0.f ? g.y : 0.f; g.z = h2 > 0.f ? g.z : 0.f; g.w = h3 > 0.f ? g.w : 0.f;
    }
    s1 += (g.x + g.y) + (g.z + g.w);
    s2 += (g.x * h0 + g.y * h1) + (g.z * h2 + g.w * h3);
    s3 += (g.x * g.x + g.y * g.y) + (g.z * g.z + g.w * g.w);
  }
  const double t1 = block_sum<double>((double)s1, red);
  const double t2 = block_sum<double>((double)s2, red);
  const double t3 = block_sum<double>((double)s3, red);
  if (threadIdx.x == 0) {
    double* o = partial + GN_BWD_PARTS * ((long long)seg * gridDim.x + blockIdx.x);
    o[0] = t1;
    o[1] = t2;
    o[2] = t3;   // sum g^2: ||gx_seg||_2 <= rstd_seg * ||g_seg||_2 (the backward is rstd times a projection of g)
  }
}

// The same partials from the per-tile sums a dgrad epilogue emitted (lgd_conv3x3_dgrad_f16_gnsums): one block per
// segment, tiles summed in a fixed order in double; written as ONE part per segment.
__global__ void gn_bwd_tile_sums_kernel(Pyr p, const float* __restrict__ tile_gn, double* __restrict__ partial) {
  __shared__ double red[32];
  const int seg = blockIdx.x;
  const int l = seg / p.batch, b = seg - l * p.batch;
  int tile_start = 0;
  for (int j = 0; j < l; ++j) tile_start += tiles_per_level(p.h[j], p.w[j], p.batch);
  const int per_img = tiles_per_image(p.h[l], p.w[l]);
  const float* ts = tile_gn + 4ll * (tile_start + b * per_img);
  double s[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < per_img; i += blockDim.x) {
    s[0] += (double)ts[4 * i];
    s[1] += (double)ts[4 * i + 1];
    s[2] += (double)ts[4 * i + 2];
  }
  for (int j = 0; j < 3; ++j) {
    const double t = block_sum<double>(s[j], red);
    if (threadIdx.x == 0) partial[GN_BWD_PARTS * (long long)seg + j] = t;
    __syncthreads();
  }
}

// Power-of-two scale of an fp16 gradient copy from an upper bound U of the tensor's l2 norm: max|g| <= ||g||_2 <= U, so
// with U * s <= 2^14 nothing can overflow, and the RMS lands at >= 2^14 / sqrt(n) (2^0.8 for 9e7 elements): more
// than 14 binades of full fp16 precision below the RMS. out3 = {s, 1/s, U}.
__device__ __forceinline__ void write_scale(double U, float* __restrict__ out3) {
  float s = 1.f;
  const float u = (float)U;
  if (u > 0.f && isfinite(u)) {
    int e;
    frexpf(u, &e);                       // u = m * 2^e, m in [0.5, 1)  ->  u <= 2^e
    e = max(-100, min(100, 14 - e));
    s = ldexpf(1.f, e);
  }
  out3[0] = s;
  out3[1] = 1.f / s;
  out3[2] = u;
}

// GroupNorm backward: U^2 = sum_seg rstd_seg^2 * sum g_seg^2
__global__ void gn_bwd_scale_kernel(const double* __restrict__ partial, int nparts, const float* __restrict__ stats,
                                    int nseg, float* __restrict__ out3) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nseg * nparts; i += blockDim.x) {
    const double r = (double)stats[2 * (i / nparts) + 1];
    acc += r * r * partial[GN_BWD_PARTS * (long long)i + 2];
  }
  const double t = block_sum<double>(acc, red);
  if (threadIdx.x == 0) write_scale(sqrt(t), out3);
}

// generic: U = m3 * |m1[0]| * |m2[0]| * sqrt(sum_i terms[i*stride])   (absent factors = 1)
__global__ void grad_scale_kernel(const float* __restrict__ terms, int n, int stride, const float* __restrict__ m1,
                                  const float* __restrict__ m2, float m3, float* __restrict__ out3) {
  __shared__ double red[32];
  double acc = 0.0;
  if (terms != nullptr)
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)terms[(long long)i * stride];
  const double t = block_sum<double>(acc, red);
  if (threadIdx.x == 0) {
    double U = terms != nullptr ? sqrt(t > 0.0 ? t : 0.0) : 1.0;
    U *= fabs((double)m3);
    if (m1 != nullptr) U *= fabs((double)m1[0]);
    if (m2 != nullptr) U *= fabs((double)m2[0]);
    write_scale(U, out3);
  }
}

__device__ __forceinline__ uint2 half4_scaled_sat(const float4& v, float s) {
  const float a = fminf(fmaxf(v.x * s, -65504.f), 65504.f), b = fminf(fmaxf(v.y * s, -65504.f), 65504.f);
  const float c = fminf(fmaxf(v.z * s, -65504.f), 65504.f), d = fminf(fmaxf(v.w * s, -65504.f), 65504.f);
  const __half2 h0 = __floats2half2_rn(a, b), h1 = __floats2half2_rn(c, d);
  uint2 r;
  r.x = *reinterpret_cast<const uint32_t*>(&h0);
  r.y = *reinterpret_cast<const uint32_t*>(&h1);
  return r;
}

// csum_partial (optional): [seg][block][256] per-block channel sums of the UN-rounded gx (bias gradient of the
// convolution in front of this GroupNorm). Every thread owns one fixed channel quad: the flat float4 index advances by
// gridDim.x*256, a multiple of 64, so (i & 63) == (threadIdx.x & 63) for all its elements.
__global__ void gn_bwd_apply_kernel(Pyr p, const float* __restrict__ gy, const float* __restrict__ x,
                                    const float* __restrict__ stats, int relu, const double* __restrict__ partial,
                                    int nparts, float* __restrict__ gx, int do_round, float* __restrict__ csum_partial,
                                    __half* __restrict__ gx_half, const float* __restrict__ scale3) {
  __shared__ float sh[2];
  __shared__ float4 shc[4][64];
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  const int seg = blockIdx.y;
  int l, b, npix;
  long long base;
  segment_of(p, seg, l, b, base, npix);
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int i = 0; i < nparts; ++i) {
      a += partial[GN_BWD_PARTS * ((long long)seg * nparts + i)];
      c += partial[GN_BWD_PARTS * ((long long)seg * nparts + i) + 1];
    }
    const double n = (double)npix * C;
    sh[0] = (float)(a / n);
    sh[1] = (float)(c / n);
  }
  __syncthreads();
  const float mg = sh[0], mgh = sh[1];
  const float hs = gx_half != nullptr ? __ldg(scale3) : 1.f;
  const float mean = stats[2 * seg], rstd = stats[2 * seg + 1];
  const long long n4 = (long long)npix * C / 4;
  const float4* xs = reinterpret_cast<const float4*>(x + base);
  const float4* gs = reinterpret_cast<const float4*>(gy + base);
  float4* os = gx ? reinterpret_cast<float4*>(gx + base) : nullptr;
  auto one = [&](long long i, const float4& xv, float4 g) {
    const float h0 = (xv.x - mean) * rstd, h1 = (xv.y - mean) * rstd, h2 = (xv.z - mean) * rstd, h3 = (xv.w - mean) * rstd;
    if (relu) {
      g.x = h0 > 0.f ? g.x : 0.f; g.y = h1 > 0.f ? g.y : 0.f; g.z = h2 > 0.f ? g.z : 0.f; g.w = h3 > 0.f ? g.w : 0.f;
    }
    float4 o;
    o.x = rstd * (g.x - mg - h0 * mgh); o.y = rstd * (g.y - mg - h1 * mgh);
    o.z = rstd * (g.z - mg - h2 * mgh); o.w = rstd * (g.w - mg - h3 * mgh);
    cs.x += o.x; cs.y += o.y; cs.z += o.z; cs.w += o.w;
    if (gx_half != nullptr) reinterpret_cast<uint2*>(gx_half + base)[i] = half4_scaled_sat(o, hs);
    if (do_round) { o.x = tf32_rna(o.x); o.y = tf32_rna(o.y); o.z = tf32_rna(o.z); o.w = tf32_rna(o.w); }
    if (os) os[i] = o;
  };
  // two (x, g) pairs in flight per thread on 8 KiB of consecutive addresses per block, stream and iteration (four pairs
  // cost a third of the resident blocks and were slower); every offset is a multiple of 64 float4 (channel quad kept)
  const long long stride = (long long)gridDim.x * blockDim.x * 2;
  long long i = (long long)blockIdx.x * blockDim.x * 2 + threadIdx.x;
  for (; i + 256 < n4; i += stride) {
    float4 xv[2], gv[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      xv[j] = __ldg(xs + i + j * 256);
      gv[j] = __ldg(gs + i + j * 256);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) one(i + j * 256, xv[j], gv[j]);
  }
  if (i < n4) one(i, __ldg(xs + i), __ldg(gs + i));   // the last, partial chunk of the segment
  if (csum_partial != nullptr) {
    const int q = threadIdx.x & 63, sub = threadIdx.x >> 6;
    shc[sub][q] = cs;
    __syncthreads();
    if (sub == 0) {
      float4 r = shc[0][q];
#pragma unroll
      for (int j = 1; j < 4; ++j) { const float4 t = shc[j][q]; r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w; }
      stg4(csum_partial + ((long long)seg * gridDim.x + blockIdx.x) * C + q * 4, r);
    }
  }
}

// Two-stage, fixed-order finalize of per-block channel partials:
//   stage A (one block per segment): seg_out[seg][c] = sum_i partial[(seg*nparts + i)*stride + c]
//   stage B (one block):             total[c]        = sum_seg seg_out[seg][c]
__global__ void chan_partial_seg_kernel(int nparts, int stride, const float* __restrict__ partial,
                                        float* __restrict__ seg_out) {
  const int seg = blockIdx.x, c = threadIdx.x;
  const float* pp = partial + (long long)seg * nparts * stride + c;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int i = 0;
  for (; i + 4 <= nparts; i += 4) {
    s0 += (double)pp[(long long)(i + 0) * stride];
    s1 += (double)pp[(long long)(i + 1) * stride];
    s2 += (double)pp[(long long)(i + 2) * stride];
    s3 += (double)pp[(long long)(i + 3) * stride];
  }
  for (; i < nparts; ++i) s0 += (double)pp[(long long)i * stride];
  seg_out[(long long)seg * C + c] = (float)((s0 + s1) + (s2 + s3));
}
__global__ void chan_total_kernel(int nseg, const float* __restrict__ seg_out, float* __restrict__ total) {
  const int c = threadIdx.x;
  double s0 = 0.0, s1 = 0.0;
  int i = 0;
  for (; i + 2 <= nseg; i += 2) {
    s0 += (double)seg_out[(long long)i * C + c];
    s1 += (double)seg_out[(long long)(i + 1) * C + c];
  }
  if (i < nseg) s0 += (double)seg_out[(long long)i * C + c];
  total[c] = (float)(s0 + s1);
}

// ------------------------------------------------------------------------------------ per-channel sums
// Generic stage 1 of the per-channel reductions: block (split, seg), 256 threads = 64 channel quads x 4 pixel lanes.
// A functor F provides  Inv prepare(seg, base, c)  (loop invariants of this thread's channel quad),
// In load(idx)  (the global loads of one pixel) and  eval(inv, in, u, v)  -> two float4 to be summed over the pixels.
// F::UNROLL pixels are loaded before any is consumed (memory-level parallelism for an HBM-bound loop).
template <typename F>
__global__ void __launch_bounds__(256)
chan_sums_kernel(Pyr p, F f, float* __restrict__ partial /* [seg][NSPLIT][2][256] */) {
  __shared__ float4 sh[2][4][64];
  const int seg = blockIdx.y, split = blockIdx.x;
  int l, b, npix;
  long long base;
  segment_of(p, seg, l, b, base, npix);
  const int q = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const int p_begin = (int)((long long)npix * split / NSPLIT), p_end = (int)((long long)npix * (split + 1) / NSPLIT);
  const typename F::Inv inv = f.prepare(seg, base, q * 4);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
  const long long cbase = base + q * 4;
  int px = p_begin + sub;
  constexpr int U = F::UNROLL;  // pixels in flight per thread
  for (; px + 4 * (U - 1) < p_end; px += 4 * U) {
    typename F::In in[U];
#pragma unroll
    for (int j = 0; j < U; ++j) in[j] = f.load(cbase + (long long)(px + 4 * j) * C);
#pragma unroll
    for (int j = 0; j < U; ++j) {
      float4 u, v;
      f.eval(inv, in[j], u, v);
      a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
      c.x += v.x; c.y += v.y; c.z += v.z; c.w += v.w;
    }
  }
  for (; px < p_end; px += 4) {
    float4 u, v;
    f.eval(inv, f.load(cbase + (long long)px * C), u, v);
    a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
    c.x += v.x; c.y += v.y; c.z += v.z; c.w += v.w;
  }
  sh[0][sub][q] = a;
  sh[1][sub][q] = c;
  __syncthreads();
  if (sub == 0) {
    float4 r0 = sh[0][0][q], r1 = sh[1][0][q];
#pragma unroll
    for (int j = 1; j < 4; ++j) {
      const float4 t0 = sh[0][j][q], t1 = sh[1][j][q];
      r0.x += t0.x; r0.y += t0.y; r0.z += t0.z; r0.w += t0.w;
      r1.x += t1.x; r1.y += t1.y; r1.z += t1.z; r1.w += t1.w;
    }
    float* o = partial + ((long long)seg * NSPLIT + split) * 2 * C;
    stg4(o + q * 4, r0);
    stg4(o + C + q * 4, r1);
  }
}

// (d, d^2) with d = x - x[first pixel of the segment] -> InstanceNorm statistics. The per-channel shift keeps
// E[d^2] - E[d]^2 well conditioned when the values of a channel are close to each other (tiny levels, flat maps).
struct SumSqF {
  const float* x;
  static constexpr int UNROLL = 4;
  typedef float4 Inv;
  typedef float4 In;
  __device__ Inv prepare(int, long long base, int c) const { return ldg4(x + base + c); }
  __device__ In load(long long idx) const { return ldg4(x + idx); }
  __device__ void eval(const Inv& s, const In& xv, float4& u, float4& v) const {
    u = make_float4(xv.x - s.x, xv.y - s.y, xv.z - s.z, xv.w - s.w);
    v = make_float4(u.x * u.x, u.y * u.y, u.z * u.z, u.w * u.w);
  }
};
struct SumF {  // (g, 0) -> bias gradients
  const float* g;
  static constexpr int UNROLL = 4;
  typedef int Inv;
  typedef float4 In;
  __device__ Inv prepare(int, long long, int) const { return 0; }
  __device__ In load(long long idx) const { return ldg4(g + idx); }
  __device__ void eval(const Inv&, const In& gv, float4& u, float4& v) const {
    u = gv;
    v = make_float4(0.f, 0.f, 0.f, 0.f);
  }
};
// IN-MSE: u_s = IN(s), u_t = IN(t); d = u_s - u_t.  sums of (d, d*u_s) for the backward,
// and (d^2, 0) for the loss value.
struct MseDiffF {
  const float *s, *t, *st_s, *st_t;
  int mode;  // 0: (d^2, 0)   1: (d, d*u_s)
  static constexpr int UNROLL = 4;
  struct Inv { float4 a0, a1, b0, b1; };   // {mean, rstd} x 4 channels of s and of t
  struct In { float4 sv, tv; };
  __device__ Inv prepare(int seg, long long, int c) const {
    const float* ps = st_s + ((long long)seg * C + c) * 2;
    const float* pt = st_t + ((long long)seg * C + c) * 2;
    Inv r;
    r.a0 = ldg4(ps); r.a1 = ldg4(ps + 4); r.b0 = ldg4(pt); r.b1 = ldg4(pt + 4);
    return r;
  }
  __device__ In load(long long idx) const {
    In r;
    r.sv = ldg4(s + idx);
    r.tv = ldg4(t + idx);
    return r;
  }
  __device__ void eval(const Inv& k, const In& in, float4& u, float4& v) const {
    const float4 sv = in.sv, tv = in.tv;
    const float us0 = (sv.x - k.a0.x) * k.a0.y, us1 = (sv.y - k.a0.z) * k.a0.w, us2 = (sv.z - k.a1.x) * k.a1.y, us3 = (sv.w - k.a1.z) * k.a1.w;
    const float ut0 = (tv.x - k.b0.x) * k.b0.y, ut1 = (tv.y - k.b0.z) * k.b0.w, ut2 = (tv.z - k.b1.x) * k.b1.y, ut3 = (tv.w - k.b1.z) * k.b1.w;
    const float d0 = us0 - ut0, d1 = us1 - ut1, d2 = us2 - ut2, d3 = us3 - ut3;
    if (mode == 0) {
      u = make_float4(d0 * d0, d1 * d1, d2 * d2, d3 * d3);
      v = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      u = make_float4(d0, d1, d2, d3);
      v = make_float4(d0 * us0, d1 * us1, d2 * us2, d3 * us3);
    }
  }
};

// stage 2 for InstanceNorm statistics: (seg, c) -> {mean, rstd}
__global__ void in_stats_finalize_kernel(Pyr p, const float* __restrict__ x, const float* __restrict__ partial,
                                         int nparts, float* __restrict__ stats) {
  const int seg = blockIdx.x, c = threadIdx.x;
  int l, b, npix;
  long long base;
  segment_of(p, seg, l, b, base, npix);
  const double n = (double)npix;
  const double shift = (double)x[base + c];
  double s = 0.0, ss = 0.0;
  for (int i = 0; i < nparts; ++i) {
    const float* o = partial + ((long long)seg * nparts + i) * 2 * C;
    s += (double)o[c];
    ss += (double)o[C + c];
  }
  const double md = s / n;
  double var = ss / n - md * md;
  if (var < 0.0) var = 0.0;
  const double mean = shift + md;
  stats[((long long)seg * C + c) * 2 + 0] = (float)mean;
  stats[((long long)seg * C + c) * 2 + 1] = (float)(1.0 / sqrt(var + (double)EPS));
}

// stage 2 for channel sums: out[seg][c], optional total[c]
__global__ void chan_sums_finalize_kernel(int nseg, const float* __restrict__ partial, float* __restrict__ out,
                                          float* __restrict__ total) {
  const int c = threadIdx.x;
  double tot = 0.0;
  for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
    double s = 0.0;
    for (int i = 0; i < NSPLIT; ++i) s += (double)partial[((long long)seg * NSPLIT + i) * 2 * C + c];
    if (out) out[(long long)seg * C + c] = (float)s;
    tot += s;
  }
  if (total) total[c] = (float)tot;  // only valid with gridDim.x == 1
}

// stage 2 for the loss value, in a fixed order: one block per segment sums its NSPLIT x 256 partials in double ...
__global__ void mse_seg_kernel(const float* __restrict__ partial, double* __restrict__ seg_sum) {
  __shared__ double red[32];
  const int seg = blockIdx.x, c = threadIdx.x;
  const float* pp = partial + (long long)seg * NSPLIT * 2 * C + c;
  double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
  for (int i = 0; i < NSPLIT; i += 2) {
    s0 += (double)pp[(long long)i * 2 * C];
    s1 += (double)pp[(long long)(i + 1) * 2 * C];
  }
  const double s = block_sum<double>(s0 + s1, red);
  if (threadIdx.x == 0) seg_sum[seg] = s;
}
// ... and one warp sums the segments
__global__ void mse_total_kernel(int nseg, const double* __restrict__ seg_sum, double scale, float* __restrict__ loss) {
  double s = 0.0;
  for (int i = threadIdx.x; i < nseg; i += 32) s += seg_sum[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) loss[0] = (float)(s * scale);
}

// ------------------------------------------------------------------------------------ IN-MSE from moments
// One pass over (s, t) yields everything the loss needs. With a = s - s[first pixel], b = t - t[first pixel] (per
// channel shifts, for conditioning) the five sums  A1 = sum a, A2 = sum a^2, B1, B2, AB = sum a*b  give
//   mean_s, var_s, mean_t, var_t, cov  ->  the InstanceNorm statistics of both sides,
//   sum (IN(s) - IN(t))^2 = n (var_s rs^2 + var_t rt^2 - 2 cov rs rt)                       (the loss), and
//   sum d = 0,  sum d*IN(s) = n (var_s rs^2 - cov rs rt)   with d = IN(s) - IN(t)           (the backward's means),
// so neither a separate statistics pass nor the backward's reduction pass is needed: 2 F1 read instead of 5 F1.
// partial: [seg][NSPLIT][5][256]
__global__ void __launch_bounds__(256, 4)
in_moments_kernel(Pyr p, const float* __restrict__ s, const float* __restrict__ t, float* __restrict__ partial) {
  __shared__ float4 sh[5][4][64];
  const int seg = blockIdx.y, split = blockIdx.x;
  int l, b, npix;
  long long base;
  segment_of(p, seg, l, b, base, npix);
  const int q = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const int p_begin = (int)((long long)npix * split / NSPLIT), p_end = (int)((long long)npix * (split + 1) / NSPLIT);
  const long long cbase = base + q * 4;
  const float4 ps = ldg4(s + cbase), pt = ldg4(t + cbase);
  float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1, b1 = a1, b2 = a1, ab = a1;
  auto acc = [&](const float4& sv, const float4& tv) {
    const float x0 = sv.x - ps.x, x1 = sv.y - ps.y, x2 = sv.z - ps.z, x3 = sv.w - ps.w;
    const float y0 = tv.x - pt.x, y1 = tv.y - pt.y, y2 = tv.z - pt.z, y3 = tv.w - pt.w;
    a1.x += x0; a1.y += x1; a1.z += x2; a1.w += x3;
    b1.x += y0; b1.y += y1; b1.z += y2; b1.w += y3;
    a2.x = fmaf(x0, x0, a2.x); a2.y = fmaf(x1, x1, a2.y); a2.z = fmaf(x2, x2, a2.z); a2.w = fmaf(x3, x3, a2.w);
    b2.x = fmaf(y0, y0, b2.x); b2.y = fmaf(y1, y1, b2.y); b2.z = fmaf(y2, y2, b2.z); b2.w = fmaf(y3, y3, b2.w);
    ab.x = fmaf(x0, y0, ab.x); ab.y = fmaf(x1, y1, ab.y); ab.z = fmaf(x2, y2, ab.z); ab.w = fmaf(x3, y3, ab.w);
  };
  int px = p_begin + sub;
  constexpr int U = 4;  // pixels in flight per thread (two tensors -> eight 16-byte loads)
  for (; px + 4 * (U - 1) < p_end; px += 4 * U) {
    float4 sv[U], tv[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      sv[j] = ldg4(s + cbase + (long long)(px + 4 * j) * C);
      tv[j] = ldg4(t + cbase + (long long)(px + 4 * j) * C);
    }
#pragma unroll
    for (int j = 0; j < U; ++j) acc(sv[j], tv[j]);
  }
  for (; px < p_end; px += 4) acc(ldg4(s + cbase + (long long)px * C), ldg4(t + cbase + (long long)px * C));
  sh[0][sub][q] = a1; sh[1][sub][q] = a2; sh[2][sub][q] = b1; sh[3][sub][q] = b2; sh[4][sub][q] = ab;
  __syncthreads();
  for (int k = sub; k < 5; k += 4) {
    float4 r = sh[k][0][q];
#pragma unroll
    for (int j = 1; j < 4; ++j) { const float4 v = sh[k][j][q]; r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w; }
    stg4(partial + (((long long)seg * NSPLIT + split) * 5 + k) * C + q * 4, r);
  }
}

// (seg, c): statistics of both sides, the backward's totals, and the segment's share of sum (IN(s)-IN(t))^2
__global__ void in_moments_finalize_kernel(Pyr p, const float* __restrict__ s, const float* __restrict__ t,
                                           const float* __restrict__ partial, float* __restrict__ st_s,
                                           float* __restrict__ st_t, float* __restrict__ bwd_sums,
                                           double* __restrict__ seg_sum, float* __restrict__ gs_terms) {
  __shared__ double red[32];
  const int seg = blockIdx.x, c = threadIdx.x;
  int l, b, npix;
  long long base;
  segment_of(p, seg, l, b, base, npix);
  double m[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  static_assert(NSPLIT % 8 == 0, "finalize loads eight splits at a time");
  for (int i = 0; i < NSPLIT; i += 8) {  // 40 independent loads in flight, then the (ordered) additions
    float v[8][5];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* o = partial + ((long long)seg * NSPLIT + i + j) * 5 * C + c;
#pragma unroll
      for (int k = 0; k < 5; ++k) v[j][k] = __ldg(o + k * C);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int k = 0; k < 5; ++k) m[k] += (double)v[j][k];
  }
  const double n = (double)npix;
  const double ma = m[0] / n, mb = m[2] / n;
  double va = m[1] / n - ma * ma, vb = m[3] / n - mb * mb;
  if (va < 0.0) va = 0.0;
  if (vb < 0.0) vb = 0.0;
  const double cov = m[4] / n - ma * mb;
  // the apply kernels normalise with the stored fp32 statistics; use exactly those here
  const float mean_s = (float)((double)s[base + c] + ma), mean_t = (float)((double)t[base + c] + mb);
  const float rs_f = (float)(1.0 / sqrt(va + (double)EPS)), rt_f = (float)(1.0 / sqrt(vb + (double)EPS));
  st_s[((long long)seg * C + c) * 2 + 0] = mean_s;
  st_s[((long long)seg * C + c) * 2 + 1] = rs_f;
  st_t[((long long)seg * C + c) * 2 + 0] = mean_t;
  st_t[((long long)seg * C + c) * 2 + 1] = rt_f;
  const double rs = (double)rs_f, rt = (double)rt_f;
  const double uss = va * rs * rs, utt = vb * rt * rt, ust = cov * rs * rt;  // means of u_s^2, u_t^2, u_s u_t
  bwd_sums[((long long)seg * 2 + 0) * C + c] = 0.f;                          // sum d
  bwd_sums[((long long)seg * 2 + 1) * C + c] = (float)(n * (uss - ust));     // sum d * u_s
  double lc = n * (uss + utt - 2.0 * ust);   // sum over the channel's pixels of d^2
  if (lc < 0.0) lc = 0.0;
  const double tot = block_sum<double>(lc, red);
  if (threadIdx.x == 0) seg_sum[seg] = tot;
  if (gs_terms != nullptr) {
    // the backward is k * rs_c * (a projection of d): ||gs||_2^2 <= k^2 * sum_(seg,c) rs_c^2 * sum d^2
    __syncthreads();
    const double b = block_sum<double>(rs * rs * lc, red);
    if (threadIdx.x == 0) gs_terms[seg] = (float)b;
  }
}

// IN-MSE backward stage 2+3: per (seg,c) means of (d, d*u_s), then gs = k*rs*(d - mean_d - u_s*mean_dus)
__global__ void in_mse_bwd_apply_kernel(Pyr p, const float* __restrict__ s, const float* __restrict__ t,
                                        const float* __restrict__ st_s, const float* __restrict__ st_t,
                                        const float* __restrict__ partial, int nparts, float two_k,
                                        const float* __restrict__ gloss, float* __restrict__ gs, int do_round,
                                        float* __restrict__ csum_partial, __half* __restrict__ gs_half,
                                        const float* __restrict__ scale3) {
  __shared__ float4 m1[64], m2[64];
  __shared__ float4 shc[4][64];
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  const int seg = blockIdx.y, split = blockIdx.x;
  int l, b, npix;
  long long base;
  segment_of(p, seg, l, b, base, npix);
  const int q = threadIdx.x & 63, sub = threadIdx.x >> 6;
  if (sub == 0) {
    double a[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0};
    for (int i = 0; i < nparts; ++i) {  // nparts = 1: the totals the moments forward already derived
      const float* o = partial + ((long long)seg * nparts + i) * 2 * C;
      const float4 u = ldg4(o + q * 4), v = ldg4(o + C + q * 4);
      a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w;
      c[0] += v.x; c[1] += v.y; c[2] += v.z; c[3] += v.w;
    }
    const double n = (double)npix;
    m1[q] = make_float4((float)(a[0] / n), (float)(a[1] / n), (float)(a[2] / n), (float)(a[3] / n));
    m2[q] = make_float4((float)(c[0] / n), (float)(c[1] / n), (float)(c[2] / n), (float)(c[3] / n));
  }
  __syncthreads();
  const float4 md = m1[q], mdu = m2[q];
  const float* ps = st_s + ((long long)seg * C + q * 4) * 2;
  const float* pt = st_t + ((long long)seg * C + q * 4) * 2;
  const float4 a0 = ldg4(ps), a1 = ldg4(ps + 4), b0 = ldg4(pt), b1 = ldg4(pt + 4);
  const float k = two_k * __ldg(gloss);
  const float hs = gs_half != nullptr ? __ldg(scale3) : 1.f;
  const int p_begin = (int)((long long)npix * split / NSPLIT), p_end = (int)((long long)npix * (split + 1) / NSPLIT);
  auto emit = [&](long long idx, const float4& sv, const float4& tv) {
    const float us0 = (sv.x - a0.x) * a0.y, us1 = (sv.y - a0.z) * a0.w, us2 = (sv.z - a1.x) * a1.y, us3 = (sv.w - a1.z) * a1.w;
    const float ut0 = (tv.x - b0.x) * b0.y, ut1 = (tv.y - b0.z) * b0.w, ut2 = (tv.z - b1.x) * b1.y, ut3 = (tv.w - b1.z) * b1.w;
    float4 o;
    o.x = k * a0.y * ((us0 - ut0) - md.x - us0 * mdu.x);
    o.y = k * a0.w * ((us1 - ut1) - md.y - us1 * mdu.y);
    o.z = k * a1.y * ((us2 - ut2) - md.z - us2 * mdu.z);
    o.w = k * a1.w * ((us3 - ut3) - md.w - us3 * mdu.w);
    cs.x += o.x; cs.y += o.y; cs.z += o.z; cs.w += o.w;
    if (gs_half != nullptr) *reinterpret_cast<uint2*>(gs_half + idx) = half4_scaled_sat(o, hs);
    if (do_round) { o.x = tf32_rna(o.x); o.y = tf32_rna(o.y); o.z = tf32_rna(o.z); o.w = tf32_rna(o.w); }
    if (gs != nullptr) stg4(gs + idx, o);
  };
  int px = p_begin + sub;
  constexpr int U = 4;  // pixels in flight per thread (two tensors -> eight 16-byte loads)
  for (; px + 4 * (U - 1) < p_end; px += 4 * U) {
    float4 sv[U], tv[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const long long idx = base + (long long)(px + 4 * j) * C + q * 4;
      sv[j] = ldg4(s + idx);
      tv[j] = ldg4(t + idx);
    }
#pragma unroll
    for (int j = 0; j < U; ++j) emit(base + (long long)(px + 4 * j) * C + q * 4, sv[j], tv[j]);
  }
  for (; px < p_end; px += 4) {
    const long long idx = base + (long long)px * C + q * 4;
    emit(idx, ldg4(s + idx), ldg4(t + idx));
  }
  if (csum_partial != nullptr) {  // [seg][split][256] channel sums of the un-rounded gradient (bias gradient)
    shc[sub][q] = cs;
    __syncthreads();
    if (sub == 0) {
      float4 r = shc[0][q];
#pragma unroll
      for (int j = 1; j < 4; ++j) { const float4 t = shc[j][q]; r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w; }
      stg4(csum_partial + ((long long)seg * NSPLIT + split) * C + q * 4, r);
    }
  }
}

__global__ void relu_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y, float* __restrict__ gx,
                                long long n4, int do_round) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 g = __ldg(reinterpret_cast<const float4*>(gy) + i);
    const float4 v = __ldg(reinterpret_cast<const float4*>(y) + i);
    g.x = v.x > 0.f ? g.x : 0.f; g.y = v.y > 0.f ? g.y : 0.f; g.z = v.z > 0.f ? g.z : 0.f; g.w = v.w > 0.f ? g.w : 0.f;
    if (do_round) { g.x = tf32_rna(g.x); g.y = tf32_rna(g.y); g.z = tf32_rna(g.z); g.w = tf32_rna(g.w); }
    reinterpret_cast<float4*>(gx)[i] = g;
  }
}

__global__ void round_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = tf32_rna(__ldg(x + i));
}

__global__ void ctx_bias_table_kernel(const float* __restrict__ ctx, const int* __restrict__ ctx_row,
                                      const float* __restrict__ conv_bias, int B, int T, float* __restrict__ out) {
  const int l = blockIdx.y, b = blockIdx.x, c = threadIdx.x;
  const int r = ctx_row[b];
  float v = conv_bias ? conv_bias[c] : 0.f;
  if (r >= 0) v += ctx[((long long)l * T + r) * C + c];
  out[((long long)l * B + b) * C + c] = v;
}

__global__ void ctx_bias_table_bwd_kernel(const float* __restrict__ gtable, const int* __restrict__ ctx_row,
                                          const int* __restrict__ img_of, int B, int T, float* __restrict__ gctx) {
  const int l = blockIdx.y, t = blockIdx.x, c = threadIdx.x;
  const int b = img_of[t];
  gctx[((long long)l * T + t) * C + c] = (ctx_row[b] == t) ? gtable[((long long)l * B + b) * C + c] : 0.f;
}

// seg_scratch: nseg*256 floats used when the caller does not want the per-segment sums themselves
static int finalize_chan_partials(int nseg, int nparts, int stride, const float* partial, float* chan_sums,
                                  float* chan_total, float* seg_scratch, cudaStream_t st) {
  float* seg_out = chan_sums ? chan_sums : seg_scratch;
  chan_partial_seg_kernel<<<nseg, C, 0, st>>>(nparts, stride, partial, seg_out);
  LGD_LAUNCH_CHECK();
  if (chan_total) {
    chan_total_kernel<<<1, C, 0, st>>>(nseg, seg_out, chan_total);
    LGD_LAUNCH_CHECK();
  }
  return LGD_OK;
}

static int grid_for(long long n_items, int max_blocks) {
  long long g = (n_items + 255) / 256;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

// ------------------------------------------------------------------------------------ GroupNorm(32, 256), affine
// The towers of the FCOS-family detection heads (thirdparty_heads/fcos.py:455-476, poto.py:545-566): Conv2d ->
// GroupNorm(32, 256) (affine gamma / beta, eps 1e-5, biased variance over the 8 channels x H x W of a group) -> ReLU.
// A thread owns one channel quad (4 of the 8 channels of group q >> 1) like everywhere in this file.
// stats32: (F*B, 32, 2) = {mean, rstd};  chsum: (F*B, 256) = per-channel sum of x (the backward's sum of xhat).
constexpr int GN32_G = 32;

// the normalised, affine value: ONE expression shared by the forward apply and the backward's ReLU decision, so both
// take the same decision on every element
__device__ __forceinline__ float gn32_affine(float x, float mean, float rstd, float gamma, float beta) {
  return __fmaf_rn(__fmul_rn(__fsub_rn(x, mean), rstd), gamma, beta);
}

// stage 2 of the statistics: per-channel shifted sums (chan_sums_kernel<SumSqF>) -> group {mean, rstd}, channel sums
__global__ void gn32_finalize_kernel(Pyr p, const float* __restrict__ x, const float* __restrict__ partial,
                                     float* __restrict__ stats32, float* __restrict__ chsum) {
  __shared__ double s1[C], s2[C];
  const int seg = blockIdx.x, c = threadIdx.x;
  int l, b, npix;
  long long base;
  segment_of(p, seg, l, b, base, npix);
  const double n = (double)npix, shift = (double)x[base + c];
  double s = 0.0, ss = 0.0;
  for (int i = 0; i < NSPLIT; ++i) {
    const float* o = partial + ((long long)seg * NSPLIT + i) * 2 * C;
    s += (double)o[c];
    ss += (double)o[C + c];
  }
  const double sx = n * shift + s;                              // sum of x
  s1[c] = sx;
  s2[c] = ss + 2.0 * shift * s + n * shift * shift;            // sum of x^2
  chsum[(long long)seg * C + c] = (float)sx;
  __syncthreads();
  if (c < GN32_G) {
    double a = 0.0, q = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a += s1[c * 8 + j];
      q += s2[c * 8 + j];
    }
    const double mean = a / (8.0 * n);
    double var = q / (8.0 * n) - mean * mean;
    if (var < 0.0) var = 0.0;
    stats32[((long long)seg * GN32_G + c) * 2 + 0] = (float)mean;
    stats32[((long long)seg * GN32_G + c) * 2 + 1] = (float)(1.0 / sqrt(var + (double)EPS));
  }
}

__global__ void gn32_apply_kernel(Pyr p, const float* __restrict__ x, const float* __restrict__ stats32,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                                  __half* __restrict__ y_half, float* __restrict__ y) {
  const int seg = blockIdx.y;
  int l, b, npix;
  long long base;
  segment_of(p, seg, l, b, base, npix);
  const int q = threadIdx.x & 63;
  const float mean = stats32[((long long)seg * GN32_G + (q >> 1)) * 2], rstd = stats32[((long long)seg * GN32_G + (q >> 1)) * 2 + 1];
  const float4 ga = ldg4(gamma + q * 4), be = ldg4(beta + q * 4);
  const long long n4 = (long long)npix * C / 4;
  const float4* xs = reinterpret_cast<const float4*>(x + base);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(xs + i);
    v.x = gn32_affine(v.x, mean, rstd, ga.x, be.x); v.y = gn32_affine(v.y, mean, rstd, ga.y, be.y);
    v.z = gn32_affine(v.z, mean, rstd, ga.z, be.z); v.w = gn32_affine(v.w, mean, rstd, ga.w, be.w);
    if (relu) { v.x = relu_keep_nan(v.x); v.y = relu_keep_nan(v.y); v.z = relu_keep_nan(v.z); v.w = relu_keep_nan(v.w); }
    if (y_half != nullptr) {
      const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
      uint2 hv;
      hv.x = *reinterpret_cast<const uint32_t*>(&h0);
      hv.y = *reinterpret_cast<const uint32_t*>(&h1);
      // the fp16 copy shows the activation pattern (nonzero = pass): a positive value never rounds to zero
      if (v.x > 0.f && (hv.x & 0xffffu) == 0) hv.x |= 1u;
      if (v.y > 0.f && (hv.x >> 16) == 0) hv.x |= 0x10000u;
      if (v.z > 0.f && (hv.y & 0xffffu) == 0) hv.y |= 1u;
      if (v.w > 0.f && (hv.y >> 16) == 0) hv.y |= 0x10000u;
      reinterpret_cast<uint2*>(y_half + base)[i] = hv;
    }
    if (y != nullptr) reinterpret_cast<float4*>(y + base)[i] = v;
  }
}

// stage 1 of the backward: per (segment, split, channel) sums of (g, g * xhat, g^2), g = gy [masked by y > 0]
__global__ void __launch_bounds__(256)
gn32_bwd_sums_kernel(Pyr p, const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ stats32,
                     const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                     float* __restrict__ partial /* [seg][NSPLIT][3][256] */) {
  __shared__ float4 sh[3][4][64];
  const int seg = blockIdx.y, split = blockIdx.x;
  int l, b, npix;
  long long base;
  segment_of(p, seg, l, b, base, npix);
  const int q = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const float mean = stats32[((long long)seg * GN32_G + (q >> 1)) * 2], rstd = stats32[((long long)seg * GN32_G + (q >> 1)) * 2 + 1];
  const float4 ga = ldg4(gamma + q * 4), be = ldg4(beta + q * 4);
  const int p_begin = (int)((long long)npix * split / NSPLIT), p_end = (int)((long long)npix * (split + 1) / NSPLIT);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a, e = a;
  const long long cbase = base + q * 4;
  auto add = [&](const float4& xv, float4 g) {
    const float h0 = (xv.x - mean) * rstd, h1 = (xv.y - mean) * rstd, h2 = (xv.z - mean) * rstd, h3 = (xv.w - mean) * rstd;
    if (relu) {
      g.x = gn32_affine(xv.x, mean, rstd, ga.x, be.x) > 0.f ? g.x : 0.f;
      g.y = gn32_affine(xv.y, mean, rstd, ga.y, be.y) > 0.f ? g.y : 0.f;
      g.z = gn32_affine(xv.z, mean, rstd, ga.z, be.z) > 0.f ? g.z : 0.f;
      g.w = gn32_affine(xv.w, mean, rstd, ga.w, be.w) > 0.f ? g.w : 0.f;
    }
    a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
    c.x += g.x * h0; c.y += g.y * h1; c.z += g.z * h2; c.w += g.w * h3;
    e.x += g.x * g.x; e.y += g.y * g.y; e.z += g.z * g.z; e.w += g.w * g.w;
  };
  int px = p_begin + sub;
  for (; px + 12 < p_end; px += 16) {   // four pixels in flight per thread
    float4 xv[4], gv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      xv[j] = ldg4(x + cbase + (long long)(px + 4 * j) * C);
      gv[j] = ldg4(gy + cbase + (long long)(px + 4 * j) * C);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) add(xv[j], gv[j]);
  }
  for (; px < p_end; px += 4) add(ldg4(x + cbase + (long long)px * C), ldg4(gy + cbase + (long long)px * C));
  sh[0][sub][q] = a;
  sh[1][sub][q] = c;
  sh[2][sub][q] = e;
  __syncthreads();
  if (sub < 3) {   // warp pair `sub` reduces output `sub` over the four pixel lanes
    float4 r = sh[sub][0][q];
#pragma unroll
    for (int j = 1; j < 4; ++j) { const float4 t = sh[sub][j][q]; r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w; }
    stg4(partial + (((long long)seg * NSPLIT + split) * 3 + sub) * C + q * 4, r);
  }
}

// stage 2: per segment. coef (seg, 32, 2) = group means of (gamma g, gamma g xhat); seg_ch (seg, 3, 256) = {sum g,
// sum g xhat, channel sum of gx}; u2 (seg) = bound of the squared norm of gx in this segment
__global__ void gn32_bwd_finalize_kernel(Pyr p, const float* __restrict__ partial, const float* __restrict__ stats32,
                                         const float* __restrict__ chsum, const float* __restrict__ gamma,
                                         float* __restrict__ coef, float* __restrict__ seg_ch, double* __restrict__ u2) {
  __shared__ double t1[C], t2[C], red[32];
  const int seg = blockIdx.x, c = threadIdx.x;
  int l, b, npix;
  long long base;
  segment_of(p, seg, l, b, base, npix);
  double sg = 0.0, sgx = 0.0, sgg = 0.0;
  for (int i = 0; i < NSPLIT; ++i) {
    const float* o = partial + ((long long)seg * NSPLIT + i) * 3 * C;
    sg += (double)o[c];
    sgx += (double)o[C + c];
    sgg += (double)o[2 * C + c];
  }
  const double ga = (double)gamma[c];
  const double mean = (double)stats32[((long long)seg * GN32_G + (c >> 3)) * 2];
  const double rstd = (double)stats32[((long long)seg * GN32_G + (c >> 3)) * 2 + 1];
  t1[c] = ga * sg;
  t2[c] = ga * sgx;
  __syncthreads();
  const int g0 = (c >> 3) * 8;
  double m1 = 0.0, m2 = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m1 += t1[g0 + j];
    m2 += t2[g0 + j];
  }
  const double n = (double)npix;
  m1 /= 8.0 * n;
  m2 /= 8.0 * n;
  if ((c & 7) == 0) {
    coef[((long long)seg * GN32_G + (c >> 3)) * 2 + 0] = (float)m1;
    coef[((long long)seg * GN32_G + (c >> 3)) * 2 + 1] = (float)m2;
  }
  const double sxhat = ((double)chsum[(long long)seg * C + c] - n * mean) * rstd;
  float* o = seg_ch + (long long)seg * 3 * C;
  o[c] = (float)sg;
  o[C + c] = (float)sgx;
  o[2 * C + c] = (float)(rstd * (ga * sg - n * m1 - sxhat * m2));   // sum over the pixels of gx
  const double tot = block_sum<double>(rstd * rstd * ga * ga * sgg, red);
  if (c == 0) u2[seg] = tot;
}

// stage 3: totals over the segments: dgamma, dbeta, bias gradient of the convolution in front, fp16 scale of gx
__global__ void gn32_bwd_total_kernel(int nseg, const float* __restrict__ seg_ch, const double* __restrict__ u2,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                                      float* __restrict__ scale3) {
  const int c = threadIdx.x;
  double a = 0.0, b = 0.0, d = 0.0;
  for (int s = 0; s < nseg; ++s) {
    const float* o = seg_ch + (long long)s * 3 * C;
    a += (double)o[c];
    b += (double)o[C + c];
    d += (double)o[2 * C + c];
  }
  if (dbeta) dbeta[c] = (float)a;
  if (dgamma) dgamma[c] = (float)b;
  if (dbias) dbias[c] = (float)d;
  if (c == 0 && scale3 != nullptr) {
    double u = 0.0;
    for (int s = 0; s < nseg; ++s) u += u2[s];
    write_scale(sqrt(u), scale3);
  }
}

__global__ void gn32_bwd_apply_kernel(Pyr p, const float* __restrict__ gy, const float* __restrict__ x,
                                      const float* __restrict__ stats32, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, int relu, const float* __restrict__ coef,
                                      __half* __restrict__ gx_half, const float* __restrict__ scale3,
                                      float* __restrict__ gx) {
  const int seg = blockIdx.y;
  int l, b, npix;
  long long base;
  segment_of(p, seg, l, b, base, npix);
  const int q = threadIdx.x & 63;
  const long long gi = (long long)seg * GN32_G + (q >> 1);
  const float mean = stats32[gi * 2], rstd = stats32[gi * 2 + 1], m1 = coef[gi * 2], m2 = coef[gi * 2 + 1];
  const float4 ga = ldg4(gamma + q * 4), be = ldg4(beta + q * 4);
  const float hs = gx_half != nullptr ? __ldg(scale3) : 1.f;
  const long long n4 = (long long)npix * C / 4;
  const float4* xs = reinterpret_cast<const float4*>(x + base);
  const float4* gs = reinterpret_cast<const float4*>(gy + base);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 xv = __ldg(xs + i);
    float4 g = __ldg(gs + i);
    const float h0 = (xv.x - mean) * rstd, h1 = (xv.y - mean) * rstd, h2 = (xv.z - mean) * rstd, h3 = (xv.w - mean) * rstd;
    if (relu) {
      g.x = gn32_affine(xv.x, mean, rstd, ga.x, be.x) > 0.f ? g.x : 0.f;
      g.y = gn32_affine(xv.y, mean, rstd, ga.y, be.y) > 0.f ? g.y : 0.f;
      g.z = gn32_affine(xv.z, mean, rstd, ga.z, be.z) > 0.f ? g.z : 0.f;
      g.w = gn32_affine(xv.w, mean, rstd, ga.w, be.w) > 0.f ? g.w : 0.f;
    }
    float4 o;
    o.x = rstd * (ga.x * g.x - m1 - h0 * m2); o.y = rstd * (ga.y * g.y - m1 - h1 * m2);
    o.z = rstd * (ga.z * g.z - m1 - h2 * m2); o.w = rstd * (ga.w * g.w - m1 - h3 * m2);
    if (gx_half != nullptr) reinterpret_cast<uint2*>(gx_half + base)[i] = half4_scaled_sat(o, hs);
    if (gx != nullptr) reinterpret_cast<float4*>(gx + base)[i] = o;
  }
}

}  // namespace lgd

using namespace lgd;

extern "C" int lgd_version(void) { return 100; }
extern "C" const char* lgd_last_error(void) { return g_err; }
extern "C" int64_t lgd_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

extern "C" int64_t lgd_pyramid_elems(const lgd_pyramid_t* pyr) {
  Pyr p;
  if (make_pyr(pyr, &p) != LGD_OK) return LGD_EINVAL;
  return p.off[LGD_MAX_LEVELS];
}

extern "C" int lgd_nchw_to_pyramid(const float* const* src_levels_host, const lgd_pyramid_t* pyr, float* dst,
                                   int round_tf32, void* dst_half, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(src_levels_host && (dst || dst_half), "lgd_nchw_to_pyramid: null pointer");
  MoverLevels m;
  int strips = 0;
  for (int l = 0; l < LGD_MAX_LEVELS; ++l) {
    m.strip_start[l] = strips;
    m.src[l] = nullptr;
    m.dst[l] = nullptr;
    if (l < p.num_levels) {
      LGD_CHECK_ARG(src_levels_host[l], "lgd_nchw_to_pyramid: null level pointer");
      m.src[l] = src_levels_host[l];
      strips += (p.h[l] * p.w[l] + 31) / 32;
    }
  }
  m.strip_start[LGD_MAX_LEVELS] = strips;
  LGD_CHECK_ARG(p.batch <= 65535, "lgd_nchw_to_pyramid: batch too large for one launch");
  nchw_to_nhwc_kernel<<<dim3(strips, p.batch), 256, 0, (cudaStream_t)stream>>>(m, p, dst, round_tf32,
                                                                             static_cast<__half*>(dst_half));
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

// channels_last maps -> pyramid buffer: one launch over all levels, 8 floats per thread
struct LevelPtrs {
  const float* p[LGD_MAX_LEVELS];
};
__global__ void __launch_bounds__(256)
nhwc_gather_kernel(LevelPtrs src, Pyr p, float* __restrict__ dst, __half* __restrict__ dst_half) {
  const long long n8 = p.off[p.num_levels] / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 8;
    int l = 0;
    while (l + 1 < p.num_levels && e >= p.off[l + 1]) ++l;
    const float4* s = reinterpret_cast<const float4*>(src.p[l] + (e - p.off[l]));
    const float4 a = __ldg(s), b = __ldg(s + 1);
    if (dst != nullptr) {
      stg4(dst + e, a);
      stg4(dst + e + 4, b);
    }
    if (dst_half != nullptr) {
      const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
      const __half2 h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
      uint4 o;
      o.x = *reinterpret_cast<const uint32_t*>(&h0); o.y = *reinterpret_cast<const uint32_t*>(&h1);
      o.z = *reinterpret_cast<const uint32_t*>(&h2); o.w = *reinterpret_cast<const uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(dst_half + e) = o;
    }
  }
}

extern "C" int lgd_nhwc_to_pyramid(const float* const* src_levels_host, const lgd_pyramid_t* pyr, float* dst,
                                   void* dst_half, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(src_levels_host && (dst || dst_half), "lgd_nhwc_to_pyramid: null pointer");
  LevelPtrs lp;
  for (int l = 0; l < LGD_MAX_LEVELS; ++l) {
    lp.p[l] = l < p.num_levels ? src_levels_host[l] : nullptr;
    LGD_CHECK_ARG(l >= p.num_levels || (lp.p[l] && (reinterpret_cast<uintptr_t>(lp.p[l]) & 15) == 0),
                  "lgd_nhwc_to_pyramid: null or misaligned level pointer");
  }
  const long long n8 = p.off[p.num_levels] / 8;
  const int blocks = (int)((n8 + 255) / 256 < 148 * 16 ? (n8 + 255) / 256 : 148 * 16);
  nhwc_gather_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(lp, p, dst, static_cast<__half*>(dst_half));
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_pyramid_to_nchw(const float* src, const lgd_pyramid_t* pyr, float* const* dst_levels_host,
                                   int accumulate, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(dst_levels_host && src, "lgd_pyramid_to_nchw: null pointer");
  MoverLevels m;
  int strips = 0;
  for (int l = 0; l < LGD_MAX_LEVELS; ++l) {
    m.strip_start[l] = strips;
    m.src[l] = nullptr;
    m.dst[l] = nullptr;
    if (l < p.num_levels) {
      LGD_CHECK_ARG(dst_levels_host[l], "lgd_pyramid_to_nchw: null level pointer");
      m.dst[l] = dst_levels_host[l];
      strips += (p.h[l] * p.w[l] + 31) / 32;
    }
  }
  m.strip_start[LGD_MAX_LEVELS] = strips;
  LGD_CHECK_ARG(p.batch <= 65535, "lgd_pyramid_to_nchw: batch too large for one launch");
  nhwc_to_nchw_kernel<<<dim3(strips, p.batch), 256, 0, (cudaStream_t)stream>>>(src, m, p, accumulate);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_gn_finalize(const lgd_pyramid_t* pyr, const float* tile_stats, float* stats, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(tile_stats && stats, "lgd_gn_finalize: null pointer");
  gn_finalize_kernel<<<p.num_levels * p.batch, 32, 0, (cudaStream_t)stream>>>(p, tile_stats, stats);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

static int seg_blocks(const Pyr& p) {
  // blocks per segment for the flat elementwise kernels: enough to fill the GPU for the largest level
  long long n4 = (long long)p.h[0] * p.w[0] * C / 4;
  long long g = (n4 + 256 * 8 - 1) / (256 * 8);
  if (g < 1) g = 1;
  if (g > 64) g = 64;
  return (int)g;
}

extern "C" size_t lgd_gn_apply_workspace(const lgd_pyramid_t* pyr) {
  return (size_t)pyr->num_levels * pyr->batch * 64 * 2 * C * sizeof(float);
}

extern "C" int lgd_gn_apply(const lgd_pyramid_t* pyr, const float* x, const float* stats, float* y, int relu,
                            int round_out, void* y_half, float* in_stats, void* workspace, size_t workspace_bytes,
                            void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(x && stats && (y || y_half), "lgd_gn_apply: null pointer");
  LGD_CHECK_ARG(y || in_stats == nullptr, "lgd_gn_apply: in_stats are statistics of the stored fp32 output");
  LGD_CHECK_ARG(in_stats == nullptr || (workspace != nullptr && workspace_bytes >= lgd_gn_apply_workspace(pyr)),
                "lgd_gn_apply: InstanceNorm statistics need lgd_gn_apply_workspace() bytes of workspace");
  const int nb = seg_blocks(p);
  dim3 grid(nb, p.num_levels * p.batch);
  float* partial = in_stats ? static_cast<float*>(workspace) : nullptr;
  __half* yh = static_cast<__half*>(y_half);
  if (in_stats)
    gn_apply_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(p, x, stats, y, relu, round_out, partial, yh);
  else
    gn_apply_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(p, x, stats, y, relu, round_out, nullptr, yh);
  LGD_LAUNCH_CHECK();
  if (in_stats) {
    in_stats_finalize_kernel<<<p.num_levels * p.batch, C, 0, (cudaStream_t)stream>>>(p, y, partial, nb, in_stats);
    LGD_LAUNCH_CHECK();
  }
  return LGD_OK;
}

static size_t gn_bwd_sums_bytes(const lgd_pyramid_t* pyr) {
  return (size_t)pyr->num_levels * pyr->batch * 64 * GN_BWD_PARTS * sizeof(double);
}
extern "C" size_t lgd_gn_bwd_workspace(const lgd_pyramid_t* pyr) {
  return gn_bwd_sums_bytes(pyr) + (size_t)pyr->num_levels * pyr->batch * (64 + 1) * C * sizeof(float);
}

static int gn_bwd_impl(const lgd_pyramid_t* pyr, const float* gy, const float* x, const float* stats, int relu,
                       const float* tile_gn, float* gx, int round_out, void* gx_half, float* scale3, float* chan_sums,
                       float* chan_total, void* workspace, size_t workspace_bytes, void* stream);

extern "C" int lgd_gn_bwd(const lgd_pyramid_t* pyr, const float* gy, const float* x, const float* stats, int relu,
                          float* gx, int round_out, void* gx_half, float* scale3, float* chan_sums, float* chan_total,
                          void* workspace, size_t workspace_bytes, void* stream) {
  return gn_bwd_impl(pyr, gy, x, stats, relu, nullptr, gx, round_out, gx_half, scale3, chan_sums, chan_total, workspace,
                     workspace_bytes, stream);
}

extern "C" int lgd_gn_bwd_tile_sums(const lgd_pyramid_t* pyr, const float* gy, const float* x, const float* stats,
                                    int relu, const float* tile_gn, float* gx, int round_out, void* gx_half,
                                    float* scale3, float* chan_sums, float* chan_total, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  LGD_CHECK_ARG(tile_gn != nullptr, "lgd_gn_bwd_tile_sums: null tile sums");
  return gn_bwd_impl(pyr, gy, x, stats, relu, tile_gn, gx, round_out, gx_half, scale3, chan_sums, chan_total, workspace,
                     workspace_bytes, stream);
}

static int gn_bwd_impl(const lgd_pyramid_t* pyr, const float* gy, const float* x, const float* stats, int relu,
                       const float* tile_gn, float* gx, int round_out, void* gx_half, float* scale3, float* chan_sums,
                       float* chan_total, void* workspace, size_t workspace_bytes, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(gy && x && stats && (gx || gx_half) && workspace, "lgd_gn_bwd: null pointer");
  LGD_CHECK_ARG(gx_half == nullptr || scale3 != nullptr, "lgd_gn_bwd: gx_half needs scale3");
  LGD_CHECK_ARG(workspace_bytes >= lgd_gn_bwd_workspace(pyr), "lgd_gn_bwd: workspace too small");
  const int nb = seg_blocks(p);
  dim3 grid(nb, p.num_levels * p.batch);
  double* partial = static_cast<double*>(workspace);
  int nparts = nb;
  if (tile_gn != nullptr) {   // the sums came out of the epilogue of the dgrad that produced gy: no pass over gy / x
    gn_bwd_tile_sums_kernel<<<p.num_levels * p.batch, 256, 0, (cudaStream_t)stream>>>(p, tile_gn, partial);
    nparts = 1;
  } else {
    gn_bwd_sums_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, gy, x, stats, relu, partial);
  }
  LGD_LAUNCH_CHECK();
  if (scale3 != nullptr) {
    gn_bwd_scale_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partial, nparts, stats, p.num_levels * p.batch, scale3);
    LGD_LAUNCH_CHECK();
  }
  const bool want_sums = chan_sums != nullptr || chan_total != nullptr;
  float* cpart = reinterpret_cast<float*>(static_cast<char*>(workspace) + gn_bwd_sums_bytes(pyr));
  gn_bwd_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, gy, x, stats, relu, partial, nparts, gx, round_out,
                                                              want_sums ? cpart : nullptr,
                                                              static_cast<__half*>(gx_half), scale3);
  LGD_LAUNCH_CHECK();
  if (want_sums) {
    const int nseg = p.num_levels * p.batch;
    return finalize_chan_partials(nseg, nb, C, cpart, chan_sums, chan_total, cpart + (size_t)nseg * 64 * C,
                                  (cudaStream_t)stream);
  }
  return LGD_OK;
}

static size_t in_partial_bytes(const lgd_pyramid_t* pyr) {
  return (size_t)pyr->num_levels * pyr->batch * NSPLIT * 2 * C * sizeof(float);
}
static size_t in_moments_bytes(const lgd_pyramid_t* pyr) {
  return (size_t)pyr->num_levels * pyr->batch * NSPLIT * 5 * C * sizeof(float);
}
extern "C" size_t lgd_in_workspace(const lgd_pyramid_t* pyr) {
  const size_t two_stage = in_partial_bytes(pyr) + (size_t)pyr->num_levels * pyr->batch * (NSPLIT + 1) * C * sizeof(float);
  const size_t moments = in_moments_bytes(pyr) + (size_t)pyr->num_levels * pyr->batch * sizeof(double);
  return two_stage > moments ? two_stage : moments;
}
extern "C" size_t lgd_channel_sums_workspace(const lgd_pyramid_t* pyr) { return lgd_in_workspace(pyr); }

extern "C" int lgd_in_stats(const lgd_pyramid_t* pyr, const float* x, float* stats, void* workspace,
                            size_t workspace_bytes, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(x && stats && workspace, "lgd_in_stats: null pointer");
  LGD_CHECK_ARG(workspace_bytes >= lgd_in_workspace(pyr), "lgd_in_stats: workspace too small");
  const int nseg = p.num_levels * p.batch;
  float* partial = static_cast<float*>(workspace);
  chan_sums_kernel<SumSqF><<<dim3(NSPLIT, nseg), 256, 0, (cudaStream_t)stream>>>(p, SumSqF{x}, partial);
  LGD_LAUNCH_CHECK();
  in_stats_finalize_kernel<<<nseg, C, 0, (cudaStream_t)stream>>>(p, x, partial, NSPLIT, stats);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_pyramid_channel_sums(const lgd_pyramid_t* pyr, const float* g, float* out, float* total,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(g && workspace && (out || total), "lgd_pyramid_channel_sums: null pointer");
  LGD_CHECK_ARG(workspace_bytes >= lgd_in_workspace(pyr), "lgd_pyramid_channel_sums: workspace too small");
  const int nseg = p.num_levels * p.batch;
  float* partial = static_cast<float*>(workspace);
  chan_sums_kernel<SumF><<<dim3(NSPLIT, nseg), 256, 0, (cudaStream_t)stream>>>(p, SumF{g}, partial);
  LGD_LAUNCH_CHECK();
  return finalize_chan_partials(nseg, NSPLIT, 2 * C, partial, out, total,
                                reinterpret_cast<float*>(static_cast<char*>(workspace) + in_partial_bytes(pyr)),
                                (cudaStream_t)stream);
}

extern "C" int lgd_in_mse_fwd(const lgd_pyramid_t* pyr, const float* s, const float* t, const float* stats_s,
                              const float* stats_t, float coef, float* loss, void* workspace, size_t workspace_bytes,
                              void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(s && t && stats_s && stats_t && loss && workspace, "lgd_in_mse_fwd: null pointer");
  LGD_CHECK_ARG(workspace_bytes >= lgd_in_workspace(pyr), "lgd_in_mse_fwd: workspace too small");
  const int nseg = p.num_levels * p.batch;
  float* partial = static_cast<float*>(workspace);
  chan_sums_kernel<MseDiffF><<<dim3(NSPLIT, nseg), 256, 0, (cudaStream_t)stream>>>(
      p, MseDiffF{s, t, stats_s, stats_t, 0}, partial);
  LGD_LAUNCH_CHECK();
  const double scale = (double)coef / (double)p.off[LGD_MAX_LEVELS];  // mean over B*256*P elements
  double* seg_sum = reinterpret_cast<double*>(static_cast<char*>(workspace) + in_partial_bytes(pyr));
  mse_seg_kernel<<<nseg, C, 0, (cudaStream_t)stream>>>(partial, seg_sum);
  LGD_LAUNCH_CHECK();
  mse_total_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(nseg, seg_sum, scale, loss);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_in_mse_moments_fwd(const lgd_pyramid_t* pyr, const float* s, const float* t, float coef,
                                      float* stats_s, float* stats_t, float* bwd_sums, float* gs_terms, float* loss,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(s && t && stats_s && stats_t && bwd_sums && loss && workspace, "lgd_in_mse_moments_fwd: null pointer");
  LGD_CHECK_ARG(workspace_bytes >= lgd_in_workspace(pyr), "lgd_in_mse_moments_fwd: workspace too small");
  const int nseg = p.num_levels * p.batch;
  float* partial = static_cast<float*>(workspace);
  in_moments_kernel<<<dim3(NSPLIT, nseg), 256, 0, (cudaStream_t)stream>>>(p, s, t, partial);
  LGD_LAUNCH_CHECK();
  double* seg_sum = reinterpret_cast<double*>(static_cast<char*>(workspace) + in_moments_bytes(pyr));
  in_moments_finalize_kernel<<<nseg, C, 0, (cudaStream_t)stream>>>(p, s, t, partial, stats_s, stats_t, bwd_sums, seg_sum,
                                                                   gs_terms);
  LGD_LAUNCH_CHECK();
  const double scale = (double)coef / (double)p.off[LGD_MAX_LEVELS];  // mean over B*256*P elements
  mse_total_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(nseg, seg_sum, scale, loss);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_in_mse_bwd(const lgd_pyramid_t* pyr, const float* s, const float* t, const float* stats_s,
                              const float* stats_t, const float* bwd_sums, float coef, const float* gloss, float* gs,
                              int round_out, const float* gs_terms, void* gs_half, float* scale3,
                              float* chan_sums, float* chan_total, void* workspace, size_t workspace_bytes,
                              void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(s && t && stats_s && stats_t && gloss && (gs || gs_half) && workspace, "lgd_in_mse_bwd: null pointer");
  LGD_CHECK_ARG(workspace_bytes >= lgd_in_workspace(pyr), "lgd_in_mse_bwd: workspace too small");
  const int nseg = p.num_levels * p.batch;
  float* partial = static_cast<float*>(workspace);
  if (bwd_sums == nullptr) {  // totals of (d, d*u_s) not delivered by lgd_in_mse_moments_fwd: reduction pass
    chan_sums_kernel<MseDiffF><<<dim3(NSPLIT, nseg), 256, 0, (cudaStream_t)stream>>>(
        p, MseDiffF{s, t, stats_s, stats_t, 1}, partial);
    LGD_LAUNCH_CHECK();
  }
  const float two_k = (float)(2.0 * (double)coef / (double)p.off[LGD_MAX_LEVELS]);
  LGD_CHECK_ARG(gs_half == nullptr || (gs_terms != nullptr && scale3 != nullptr),
                "lgd_in_mse_bwd: gs_half needs gs_terms (from lgd_in_mse_moments_fwd) and scale3");
  if (gs_half != nullptr) {
    grad_scale_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(gs_terms, nseg, 1, gloss, nullptr, two_k, scale3);
    LGD_LAUNCH_CHECK();
  }
  const bool want_sums = chan_sums != nullptr || chan_total != nullptr;
  float* cpart = reinterpret_cast<float*>(static_cast<char*>(workspace) + in_partial_bytes(pyr));
  in_mse_bwd_apply_kernel<<<dim3(NSPLIT, nseg), 256, 0, (cudaStream_t)stream>>>(
      p, s, t, stats_s, stats_t, bwd_sums ? bwd_sums : partial, bwd_sums ? 1 : NSPLIT, two_k, gloss, gs, round_out,
      want_sums ? cpart : nullptr, static_cast<__half*>(gs_half), scale3);
  LGD_LAUNCH_CHECK();
  if (want_sums)
    return finalize_chan_partials(nseg, NSPLIT, C, cpart, chan_sums, chan_total, cpart + (size_t)nseg * NSPLIT * C,
                                  (cudaStream_t)stream);
  return LGD_OK;
}

extern "C" int lgd_grad_scale(const float* terms, int n, int stride, const float* m1, const float* m2, float m3,
                              float* out3, void* stream) {
  LGD_CHECK_ARG(out3 && (terms == nullptr || (n > 0 && stride > 0)), "lgd_grad_scale: bad arguments");
  grad_scale_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(terms, n, stride, m1, m2, m3, out3);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_relu_bwd(const float* gy, const float* y, float* gx, int64_t n, int round_out, void* stream) {
  LGD_CHECK_ARG(gy && y && gx && n >= 0 && n % 4 == 0, "lgd_relu_bwd: bad arguments (n must be a multiple of 4)");
  if (n == 0) return LGD_OK;
  relu_bwd_kernel<<<grid_for(n / 4, 148 * 16), 256, 0, (cudaStream_t)stream>>>(gy, y, gx, n / 4, round_out);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

__global__ void axpy_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += x[i];
}
extern "C" int lgd_axpy(const float* x, float* y, int64_t n, void* stream) {
  LGD_CHECK_ARG(x && y && n > 0, "lgd_axpy: bad arguments");
  axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, (long long)n);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

// A few scalars (losses) written straight into pinned host memory by a kernel: the read-back stays off the copy engines,
// which work in order and may be busy with a long transfer (the input copies of the next steps).
__global__ void store_to_host_kernel(const float* __restrict__ src, volatile float* __restrict__ dst, int n) {
  const int i = threadIdx.x;
  if (i < n) dst[i] = src[i];
  __threadfence_system();
}
extern "C" int lgd_store_to_host(const float* src, float* pinned_host_dst, int n, void* stream) {
  LGD_CHECK_ARG(src && pinned_host_dst && n > 0 && n <= 1024, "lgd_store_to_host: bad arguments (1 <= n <= 1024)");
  void* dptr = nullptr;
  LGD_CUDA(cudaHostGetDevicePointer(&dptr, pinned_host_dst, 0));   // fails loudly for pageable memory
  store_to_host_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(src, static_cast<volatile float*>(dptr), n);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

// Small host->device uploads (the step's box table) pulled from pinned host memory by a kernel. Copy engines work in
// order: a 3 KB cudaMemcpyAsync queued behind another stream's 367 MB input copy waits 6.7 ms for it, and the compute
// stream with it (measured in the end-to-end loop; the token programs fetch their op lists the same way).
__global__ void upload_words_kernel(const volatile unsigned int* __restrict__ src, unsigned int* __restrict__ dst,
                                    long long nwords) {
  // 16-byte requests over PCIe for the aligned body (both buffers come 16-byte aligned from the allocators), words for
  // the tail
  const long long n4 = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 ? nwords / 4 : 0;
  const volatile uint4* s4 = reinterpret_cast<const volatile uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n4; i += nth) {
    uint4 v;   // one 16-byte volatile load (member-wise reads of a volatile uint4 compile to four requests)
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(s4 + i) : "memory");
    d4[i] = v;
  }
  for (long long i = 4 * n4 + tid; i < nwords; i += nth) dst[i] = src[i];
}
extern "C" int lgd_upload_from_host(void* dst, const void* pinned_host_src, int64_t nbytes, void* stream) {
  LGD_CHECK_ARG(dst && pinned_host_src && nbytes > 0 && nbytes % 4 == 0, "lgd_upload_from_host: bad arguments");
  void* dptr = nullptr;
  LGD_CUDA(cudaHostGetDevicePointer(&dptr, const_cast<void*>(pinned_host_src), 0));   // fails for pageable memory
  const long long nwords = nbytes / 4;
  const unsigned blocks = (unsigned)std::min<long long>((nwords / 4 + 31) / 32 + 1, 148 * 4);
  upload_words_kernel<<<blocks, 32, 0, (cudaStream_t)stream>>>(static_cast<const volatile unsigned int*>(dptr),
                                                               static_cast<unsigned int*>(dst), nwords);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_round_tf32(const float* x, float* y, int64_t n, void* stream) {
  LGD_CHECK_ARG(x && y && n >= 0, "lgd_round_tf32: bad arguments");
  if (n == 0) return LGD_OK;
  round_kernel<<<grid_for(n, 148 * 16), 256, 0, (cudaStream_t)stream>>>(x, y, n);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

// x = hi + lo with hi = rna_tf32(x) (written back in place) and lo = rna_tf32(x - hi): the two TF32 operands of the
// split-operand (fp32-accurate) forward convolution. x - hi is exact in fp32; what is dropped is below 2^-22 |x|.
__global__ void tf32_split_kernel(float* __restrict__ x, float* __restrict__ lo, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    const float hi = tf32_rna(v);
    x[i] = hi;
    lo[i] = tf32_rna(v - hi);
  }
}

extern "C" int lgd_tf32_split(float* x, float* lo, int64_t n, void* stream) {
  LGD_CHECK_ARG(x && lo && x != lo && n >= 0, "lgd_tf32_split: bad arguments");
  if (n == 0) return LGD_OK;
  tf32_split_kernel<<<grid_for(n, 148 * 16), 256, 0, (cudaStream_t)stream>>>(x, lo, n);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_ctx_bias_table(const float* ctx, const int32_t* ctx_row, const float* conv_bias, int F, int B, int T,
                                  float* out, void* stream) {
  LGD_CHECK_ARG(ctx_row && out && F > 0 && B > 0, "lgd_ctx_bias_table: bad arguments");
  ctx_bias_table_kernel<<<dim3(B, F), C, 0, (cudaStream_t)stream>>>(ctx, ctx_row, conv_bias, B, T, out);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_ctx_bias_table_bwd(const float* gtable, const int32_t* ctx_row, const int32_t* img_of, int F, int B,
                                      int T, float* gctx, void* stream) {
  LGD_CHECK_ARG(gtable && ctx_row && img_of && gctx && F > 0 && B > 0 && T > 0, "lgd_ctx_bias_table_bwd: bad arguments");
  ctx_bias_table_bwd_kernel<<<dim3(T, F), C, 0, (cudaStream_t)stream>>>(gtable, ctx_row, img_of, B, T, gctx);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}


// ==== GroupNorm(32, 256) with affine parameters: the towers of the FCOS-family heads (thirdparty_heads/fcos.py:455-476)
static size_t gn32_partial_bytes(const lgd_pyramid_t* pyr) {
  return (size_t)pyr->num_levels * pyr->batch * NSPLIT * 3 * C * sizeof(float);
}
extern "C" size_t lgd_gn32_workspace(const lgd_pyramid_t* pyr) {
  const size_t nseg = (size_t)pyr->num_levels * pyr->batch;
  return gn32_partial_bytes(pyr) + nseg * 3 * C * sizeof(float) + nseg * GN32_G * 2 * sizeof(float) + nseg * sizeof(double) + 256;
}

extern "C" int lgd_gn32_stats(const lgd_pyramid_t* pyr, const float* x, float* stats32, float* chsum, void* workspace,
                              size_t workspace_bytes, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(x && stats32 && chsum && workspace, "lgd_gn32_stats: null pointer");
  LGD_CHECK_ARG(workspace_bytes >= lgd_gn32_workspace(pyr), "lgd_gn32_stats: workspace too small");
  const int nseg = p.num_levels * p.batch;
  float* partial = static_cast<float*>(workspace);
  chan_sums_kernel<SumSqF><<<dim3(NSPLIT, nseg), 256, 0, (cudaStream_t)stream>>>(p, SumSqF{x}, partial);
  LGD_LAUNCH_CHECK();
  gn32_finalize_kernel<<<nseg, C, 0, (cudaStream_t)stream>>>(p, x, partial, stats32, chsum);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_gn32_apply(const lgd_pyramid_t* pyr, const float* x, const float* stats32, const float* gamma,
                              const float* beta, int relu, void* y_half, float* y, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(x && stats32 && gamma && beta && (y_half || y), "lgd_gn32_apply: null pointer");
  gn32_apply_kernel<<<dim3(seg_blocks(p), p.num_levels * p.batch), 256, 0, (cudaStream_t)stream>>>(
      p, x, stats32, gamma, beta, relu, static_cast<__half*>(y_half), y);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_gn32_bwd(const lgd_pyramid_t* pyr, const float* gy, const float* x, const float* stats32,
                            const float* chsum, const float* gamma, const float* beta, int relu, void* gx_half,
                            float* scale3, float* gx, float* dgamma, float* dbeta, float* dbias, void* workspace,
                            size_t workspace_bytes, void* stream) {
  Pyr p;
  int rc = make_pyr(pyr, &p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(gy && x && stats32 && chsum && gamma && beta && workspace && (gx || gx_half), "lgd_gn32_bwd: null pointer");
  LGD_CHECK_ARG(gx_half == nullptr || scale3 != nullptr, "lgd_gn32_bwd: gx_half needs scale3");
  LGD_CHECK_ARG(workspace_bytes >= lgd_gn32_workspace(pyr), "lgd_gn32_bwd: workspace too small");
  const int nseg = p.num_levels * p.batch;
  char* w = static_cast<char*>(workspace);
  float* partial = reinterpret_cast<float*>(w);
  float* seg_ch = reinterpret_cast<float*>(w + gn32_partial_bytes(pyr));
  float* coef = seg_ch + (size_t)nseg * 3 * C;
  double* u2 = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(coef + (size_t)nseg * GN32_G * 2) + 7) & ~uintptr_t(7));
  gn32_bwd_sums_kernel<<<dim3(NSPLIT, nseg), 256, 0, (cudaStream_t)stream>>>(p, gy, x, stats32, gamma, beta, relu, partial);
  LGD_LAUNCH_CHECK();
  gn32_bwd_finalize_kernel<<<nseg, C, 0, (cudaStream_t)stream>>>(p, partial, stats32, chsum, gamma, coef, seg_ch, u2);
  LGD_LAUNCH_CHECK();
  gn32_bwd_total_kernel<<<1, C, 0, (cudaStream_t)stream>>>(nseg, seg_ch, u2, dgamma, dbeta, dbias, scale3);
  LGD_LAUNCH_CHECK();
  gn32_bwd_apply_kernel<<<dim3(seg_blocks(p), nseg), 256, 0, (cudaStream_t)stream>>>(
      p, gy, x, stats32, gamma, beta, relu, coef, static_cast<__half*>(gx_half), scale3, gx);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}
