// Bodies of the "small-T" kernels as device functions on VIRTUAL block indices: smallt.cu wraps each one in a __global__
// kernel of its own (one launch per op), tokenprog.cu runs whole sequences of them inside one persistent kernel
// with grid barriers in between. Both see the same arithmetic in the same order, so results are bit-identical.
// NC = true: inputs are immutable for the kernel's lifetime (read through the non-coherent path); NC = false: inputs
// may have been written earlier in the same kernel by another CTA (plain loads, coherent after the grid barrier).
#pragma once
#include "common.cuh"

namespace lgd {

template <bool NC>
__device__ __forceinline__ float ldf(const float* p) {
  if (NC) return __ldg(p);
  return *p;
}
template <bool NC>
__device__ __forceinline__ int ldi(const int* p) {
  if (NC) return __ldg(p);
  return *p;
}

// sum over the NT threads of a thread group (NT = 128 or 256, a whole CTA or one half of it); every thread gets the
// result. bar_id: named barrier of the group (0 = the CTA-wide barrier).
template <int NT>
__device__ __forceinline__ float group_sum(float v, float* smem /* >= 32 */, int tid, int bar_id) {
  const int lane = tid & 31, wid = tid >> 5;
  v = warp_sum(v);
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(NT) : "memory");
  if (lane == 0) smem[wid] = v;
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(NT) : "memory");
  float r = (lane < NT / 32) ? smem[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// ------------------------------------------------------------------------------------ generic small GEMM
// Cm[m*ldc + n] (+)= sum_k A(m,k) * B(k,n) + bias[n]
//   A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn]
// The matrices here have M = T ~ 10^2 rows, so a plain tiling leaves most SMs idle and serialises long
// contractions (fc3 of the STNs: K = 7056). The z index therefore splits K: split z accumulates its K range and writes
// a partial [z][M][N] to the workspace; splitk_reduce sums the partials in a fixed order (deterministic) and
// applies bias / accumulate. With one split the tile writes Cm directly. Global loads of tile i+1 are issued before
// the FMAs of tile i (register double buffering).
constexpr int GT = 64, GK = 16;
struct GemmSmem {
  float As[GK][GT + 4];
  float Bs[GK][GT + 4];
};
struct GemmArgs {
  const float* A;
  long long sam, sak;
  const float* B;
  long long sbk, sbn;
  const float* bias;
  float* Cm;
  int ldc, M, N, K, accumulate, k_per_split;
  float* partial;
};

template <bool NC>
__device__ __forceinline__ void gemm_body(GemmSmem& sm, const GemmArgs& g, int bx, int by, int bz) {
  const int tid = threadIdx.x;
  const int m0 = by * GT, n0 = bx * GT;
  const int k_begin = bz * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (g.sak == 1), b_kfast = (g.sbk == 1);
  int am[4], ak[4], bn[4], bk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int e = tid + 256 * j;
    if (a_kfast) { ak[j] = e & 15; am[j] = e >> 4; } else { am[j] = e & 63; ak[j] = e >> 6; }
    if (b_kfast) { bk[j] = e & 15; bn[j] = e >> 4; } else { bn[j] = e & 63; bk[j] = e >> 6; }
  }
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = m0 + am[j], gk = k0 + ak[j];
      ra[j] = (gm < g.M && gk < k_end) ? ldf<NC>(g.A + gm * g.sam + gk * g.sak) : 0.f;
      const int gn = n0 + bn[j], gk2 = k0 + bk[j];
      rb[j] = (gn < g.N && gk2 < k_end) ? ldf<NC>(g.B + gk2 * g.sbk + gn * g.sbn) : 0.f;
    }
  };
  fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += GK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sm.As[ak[j]][am[j]] = ra[j];
      sm.Bs[bk[j]][bn[j]] = rb[j];
    }
    __syncthreads();
    if (k0 + GK < k_end) fetch(k0 + GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&sm.As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sm.Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (g.partial != nullptr) {
    float* o = g.partial + (long long)bz * g.M * g.N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gm = m0 + ty * 4 + i;
      if (gm >= g.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gn = n0 + tx * 4 + j;
        if (gn < g.N) o[(long long)gm * g.N + gn] = acc[i][j];
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += __ldg(g.bias + gn);
      float* o = g.Cm + (long long)gm * g.ldc + gn;
      *o = g.accumulate ? *o + v : v;
    }
  }
}

// vb = virtual block of 256 elements
__device__ __forceinline__ void splitk_reduce_body(const float* partial, int splits, const float* bias, float* Cm, int ldc,
                                                   int M, int N, int accumulate, int vb) {
  const long long i = (long long)vb * 256 + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = (int)(i / N), n = (int)(i - (long long)m * N);
  float v = 0.f;
  for (int z = 0; z < splits; ++z) v += partial[(long long)z * M * N + i];
  if (bias) v += __ldg(bias + n);
  float* o = Cm + (long long)m * ldc + n;
  *o = accumulate ? *o + v : v;
}

// out[n] (+)= sum_m g[m*ld + n]. Virtual block = 32 columns x 8 row groups; fixed-order smem reduction (deterministic).
template <bool NC>
__device__ __forceinline__ void colsum_body(float (*sh)[33], const float* g, int ld, int M, int N, float* out,
                                            int accumulate, int vb) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = vb * 32 + tx;
  float s0 = 0.f, s1 = 0.f;
  if (n < N) {
    int m = ty;
    for (; m + 8 < M; m += 16) {
      s0 += ldf<NC>(g + (long long)m * ld + n);
      s1 += ldf<NC>(g + (long long)(m + 8) * ld + n);
    }
    if (m < M) s0 += ldf<NC>(g + (long long)m * ld + n);
  }
  sh[ty][tx] = s0 + s1;
  __syncthreads();
  if (ty == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += sh[j][tx];
    out[n] = accumulate ? out[n] + s : s;
  }
}

// ------------------------------------------------------------------------------------ LayerNorm (+ReLU)
// one row per group of 128 threads (tid = index inside the group)
__device__ __forceinline__ void layernorm_fwd_body(float* red, const float* x, float* y, float* mean_out,
                                                   float* rstd_out, int N, int relu, int row, int tid, int bar_id) {
  const float* xr = x + (long long)row * N;
  float* yr = y + (long long)row * N;
  float s = 0.f;
  for (int i = tid; i < N; i += 128) s += xr[i];
  const float mean = group_sum<128>(s, red, tid, bar_id) / (float)N;
  float v = 0.f;
  for (int i = tid; i < N; i += 128) {
    const float d = xr[i] - mean;
    v += d * d;
  }
  const float var = group_sum<128>(v, red, tid, bar_id) / (float)N;
  const float rstd = rsqrtf(var + EPS);
  for (int i = tid; i < N; i += 128) {
    float o = (xr[i] - mean) * rstd;
    if (relu) o = fmaxf(o, 0.f);
    yr[i] = o;
  }
  if (tid == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

// Split-K reduction + bias fused into the LayerNorm of the same rows (token programs): the row is summed from the
// partials in the order splitk_reduce_body uses, stored as the pre-activation (the backward needs it) and normalised
// from a shared-memory copy -- the same values in the same order as the two separate ops.
__device__ __forceinline__ void reduce_layernorm_fwd_body(float* red, float* rowbuf, const float* partial, int splits,
                                                          const float* bias, float* pre, float* y, float* mean_out,
                                                          float* rstd_out, int M, int N, int relu, int row, int tid,
                                                          int bar_id) {
  float s = 0.f;
  for (int i = tid; i < N; i += 128) {
    float v = 0.f;
    for (int z = 0; z < splits; ++z) v += partial[((long long)z * M + row) * N + i];
    if (bias) v += __ldg(bias + i);
    pre[(long long)row * N + i] = v;
    rowbuf[i] = v;   // each thread re-reads only what it wrote
    s += v;
  }
  const float mean = group_sum<128>(s, red, tid, bar_id) / (float)N;
  float q = 0.f;
  for (int i = tid; i < N; i += 128) {
    const float d = rowbuf[i] - mean;
    q += d * d;
  }
  const float var = group_sum<128>(q, red, tid, bar_id) / (float)N;
  const float rstd = rsqrtf(var + EPS);
  float* yr = y + (long long)row * N;
  for (int i = tid; i < N; i += 128) {
    float o = (rowbuf[i] - mean) * rstd;
    if (relu) o = fmaxf(o, 0.f);
    yr[i] = o;
  }
  if (tid == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

__device__ __forceinline__ void layernorm_bwd_body(float* red, const float* gy, const float* x, const float* mean,
                                                   const float* rstd, float* gx, int N, int relu, int row, int tid,
                                                   int bar_id) {
  const float* xr = x + (long long)row * N;
  const float* gr = gy + (long long)row * N;
  float* o = gx + (long long)row * N;
  const float mu = mean[row], rs = rstd[row];
  float s1 = 0.f, s2 = 0.f;
  for (int i = tid; i < N; i += 128) {
    const float h = (xr[i] - mu) * rs;
    const float g = (relu && h <= 0.f) ? 0.f : gr[i];
    s1 += g;
    s2 += g * h;
  }
  const float m1 = group_sum<128>(s1, red, tid, bar_id) / (float)N;
  const float m2 = group_sum<128>(s2, red, tid, bar_id) / (float)N;
  for (int i = tid; i < N; i += 128) {
    const float h = (xr[i] - mu) * rs;
    const float g = (relu && h <= 0.f) ? 0.f : gr[i];
    o[i] = rs * (g - m1 - h * m2);
  }
}

// ------------------------------------------------------------------------------------ row-vector x matrix
// sx: k floats of shared memory; nthreads = threads of the (whole) CTA
template <bool NC>
__device__ __forceinline__ void rowvec_fwd_body(float* sx, const float* x, const float* mats, float* y, int k, int t,
                                                int nthreads) {
  for (int i = threadIdx.x; i < k; i += nthreads) sx[i] = x[(long long)t * k + i];
  __syncthreads();
  const float* m = mats + (long long)t * k * k;
  for (int j = threadIdx.x; j < k; j += nthreads) {
    float s = 0.f;
    for (int i = 0; i < k; ++i) s = fmaf(sx[i], ldf<NC>(m + (long long)i * k + j), s);
    y[(long long)t * k + j] = s;
  }
}

// sm: 2*k floats of shared memory
template <bool NC>
__device__ __forceinline__ void rowvec_bwd_body(float* sm, const float* gy, const float* x, const float* mats, float* gx,
                                                float* gmats, int k, int t, int nthreads) {
  float* sg = sm;      // gy row
  float* sx = sm + k;  // x row
  for (int i = threadIdx.x; i < k; i += nthreads) {
    sg[i] = gy[(long long)t * k + i];
    sx[i] = x[(long long)t * k + i];
  }
  __syncthreads();
  const float* m = mats + (long long)t * k * k;
  float* gm = gmats + (long long)t * k * k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = nthreads >> 5;
  for (int i = warp; i < k; i += nw) {
    float s = 0.f;
    const float xi = sx[i];
    for (int j = lane; j < k; j += 32) {
      s = fmaf(sg[j], ldf<NC>(m + (long long)i * k + j), s);
      gm[(long long)i * k + j] = xi * sg[j];
    }
    s = warp_sum(s);
    if (lane == 0) gx[(long long)t * k + i] = s;
  }
}

// ------------------------------------------------------------------------------------ hier_pool + concat
__device__ __forceinline__ void segmax_fwd_body(const float* local, int c_local, const float* x, int Cx,
                                                const int* img_start, float* out, int* argmax, int b, int nthreads) {
  const int t0 = img_start[b], t1 = img_start[b + 1];
  const int ld = c_local + Cx;
  for (int c = threadIdx.x; c < Cx; c += nthreads) {
    float best = x[(long long)t0 * Cx + c];
    int bi = t0;
    for (int t = t0 + 1; t < t1; ++t) {
      const float v = x[(long long)t * Cx + c];
      if (v > best || (v != v && best == best)) {  // first maximum wins (torch.max semantics), NaN propagates
        best = v;
        bi = t;
      }
    }
    argmax[(long long)b * Cx + c] = bi;
    for (int t = t0; t < t1; ++t) out[(long long)t * ld + c_local + c] = best;
  }
  for (int i = threadIdx.x; i < (t1 - t0) * c_local; i += nthreads) {
    const int t = t0 + i / c_local, c = i % c_local;
    out[(long long)t * ld + c] = local[(long long)t * c_local + c];
  }
}

__device__ __forceinline__ void segmax_bwd_body(const float* gout, int c_local, int Cx, const int* img_start,
                                                const int* argmax, float* glocal, float* gx, int b, int nthreads) {
  const int t0 = img_start[b], t1 = img_start[b + 1];
  const int ld = c_local + Cx;
  for (int c = threadIdx.x; c < Cx; c += nthreads) {
    float s = 0.f;
    for (int t = t0; t < t1; ++t) s += gout[(long long)t * ld + c_local + c];
    const int bi = argmax[(long long)b * Cx + c];
    for (int t = t0; t < t1; ++t) gx[(long long)t * Cx + c] = (t == bi) ? s : 0.f;
  }
  for (int i = threadIdx.x; i < (t1 - t0) * c_local; i += nthreads) {
    const int t = t0 + i / c_local, c = i % c_local;
    glocal[(long long)t * c_local + c] = gout[(long long)t * ld + c];
  }
}

// host-side split-K plan of one GEMM (shared by the per-op launcher and the program builder): the number of k splits
// depends on the tile count, K and the workspace the partials may use
struct GemmPlan {
  int gx, gy, splits, kps;
};
inline GemmPlan plan_gemm(int M, int N, int K, bool have_ws, size_t ws_bytes) {
  GemmPlan p;
  p.gx = (N + GT - 1) / GT;
  p.gy = (M + GT - 1) / GT;
  const int tiles = p.gx * p.gy;
  // aim at ~2 CTAs per SM; never split below 64 k per CTA; stay inside the caller's workspace
  int splits = (2 * 148 + tiles - 1) / tiles;
  const int max_by_k = (K + 63) / 64;
  if (splits > max_by_k) splits = max_by_k;
  if (!have_ws) splits = 1;
  while (splits > 1 && (size_t)splits * M * N * sizeof(float) > ws_bytes) --splits;
  p.kps = K;
  if (splits > 1) {
    int kps = (K + splits - 1) / splits;
    kps = (kps + GK - 1) / GK * GK;
    splits = (K + kps - 1) / kps;
    p.kps = kps;
  }
  p.splits = splits < 1 ? 1 : splits;
  return p;
}

}  // namespace lgd
