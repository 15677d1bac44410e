// Internal helpers shared by all translation units of liblgd_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/lgd_b200.h"

namespace lgd {

void set_error(const char* fmt, ...);

#define LGD_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      lgd::set_error(__VA_ARGS__);      \
      return LGD_EINVAL;                \
    }                                   \
  } while (0)

#define LGD_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      lgd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return LGD_ECUDA;                                                                     \
    }                                                                                       \
  } while (0)

// every kernel launch of the library goes through this macro; the counter only feeds lgd_launch_count()
// (diagnostics for bench.py's "gpu_launches"), it never influences results.
void count_launch();
#define LGD_LAUNCH_CHECK()          \
  do {                              \
    lgd::count_launch();            \
    LGD_CUDA(cudaGetLastError());   \
  } while (0)

constexpr int C = LGD_CHANNELS;  // 256 channels everywhere on this path (dynamic_teacher.py:28)
constexpr float EPS = 1e-5f;
// output tile of the tcgen05 convolution (conv3x3_tc.cu) = 128 consecutive SLOTS of ONE image of one level, a slot being a
// position of the zero-padded rows (w + 2 per row, row-major; the two pad positions of a row are computed and dropped).
// Every level holds an even number of tiles (CTA pairs never straddle levels): the last one may be a dummy whose per-tile
// by-products are written as zeros. Shared with the kernels that reduce per-tile by-products.
constexpr int TILE_PIX = 128;
__host__ __device__ inline int tiles_per_image(int h, int w) { return (h * (w + 2) + TILE_PIX - 1) / TILE_PIX; }
__host__ __device__ inline int tiles_per_level(int h, int w, int batch) { return (batch * tiles_per_image(h, w) + 1) & ~1; }
// pixel splits per (level, image) segment in the two-stage deterministic reductions
constexpr int NSPLIT = 32;

// Device-side copy of the pyramid geometry (passed by value as a kernel argument).
struct Pyr {
  int num_levels;
  int batch;
  int h[LGD_MAX_LEVELS];
  int w[LGD_MAX_LEVELS];
  long long off[LGD_MAX_LEVELS + 1];  // element offset of level l in a pyramid buffer (x256 channels included)
  int pix_start[LGD_MAX_LEVELS + 1];  // prefix sum of h*w (pixels of ONE image)
};

inline int make_pyr(const lgd_pyramid_t* p, Pyr* out) {
  if (p == nullptr || p->num_levels < 1 || p->num_levels > LGD_MAX_LEVELS || p->batch < 1) {
    set_error("bad pyramid descriptor");
    return LGD_EINVAL;
  }
  out->num_levels = p->num_levels;
  out->batch = p->batch;
  long long off = 0;
  int pix = 0;
  for (int l = 0; l < LGD_MAX_LEVELS; ++l) {
    if (l < p->num_levels) {
      if (p->h[l] < 1 || p->w[l] < 1) {
        set_error("bad level size at level %d", l);
        return LGD_EINVAL;
      }
      out->h[l] = p->h[l];
      out->w[l] = p->w[l];
      out->off[l] = off;
      out->pix_start[l] = pix;
      off += (long long)p->batch * p->h[l] * p->w[l] * C;
      pix += p->h[l] * p->w[l];
    } else {
      out->h[l] = out->w[l] = 0;
      out->off[l] = off;
      out->pix_start[l] = pix;
    }
  }
  out->off[LGD_MAX_LEVELS] = off;
  out->pix_start[LGD_MAX_LEVELS] = pix;
  return LGD_OK;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum (all threads get the result). blockDim.x <= 1024, multiple of 32.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  T r = (lane < nw) ? smem[lane] : T(0);
  r = warp_sum(r);
  return r;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stg4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ReLU that lets NaN through (fmaxf(NaN, 0) would return 0): a forward activation that overflowed the fp16 operand range
// (|x| > 65504 -> inf -> NaN after the next convolution / statistics) must reach the loss and the teacher pyramid as a
// non-finite value -- the training loop aborts on it (train.py:194) -- instead of being silently zeroed on the way.
__device__ __forceinline__ float relu_keep_nan(float x) { return x < 0.f ? 0.f : x; }

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

}  // namespace lgd
