// "Small-T" kernels: everything that runs on the T (~10^2) object tokens of a batch -- the PointNet-style
// label encoder + STNs (a2: label_encoder.py:216-276, spatial_transformer.py:30-47), the 1-D projections
// (a3/a7: dynamic_teacher.py:56-65) and the block-diagonal multi-head attention (a6: dynamic_teacher.py:255-275).
// < 0.2 % of the step's FLOPs and latency-bound: plain fp32 SIMT kernels (bit-for-bit fp32 semantics, no
// tensor cores), deterministic reductions.
#include "common.cuh"

namespace lgd {

// ------------------------------------------------------------------------------------ generic small GEMM
// Cm[m*ldc + n] (+)= sum_k A(m,k) * B(k,n) + bias[n]
//   A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn]
// The matrices here have M = T ~ 10^2 rows, so a plain tiling leaves most SMs idle and serialises long
// contractions (fc3 of the STNs: K = 7056). grid.z therefore splits K: split z accumulates its K range and writes a
// partial [z][M][N] to the workspace; splitk_reduce_kernel sums the partials in a fixed order (deterministic) and
// applies bias / accumulate. With one split the kernel writes Cm directly. Global loads of tile i+1 are issued before
// the FMAs of tile i (register double buffering).
constexpr int GT = 64, GK = 16;

__global__ void __launch_bounds__(256)
gemm_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B, long long sbk,
            long long sbn, const float* __restrict__ bias, float* __restrict__ Cm, int ldc, int M, int N, int K,
            int accumulate, int k_per_split, float* __restrict__ partial) {
  __shared__ float As[GK][GT + 4];
  __shared__ float Bs[GK][GT + 4];
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (sak == 1), b_kfast = (sbk == 1);
  int am[4], ak[4], bn[4], bk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int e = threadIdx.x + 256 * j;
    if (a_kfast) { ak[j] = e & 15; am[j] = e >> 4; } else { am[j] = e & 63; ak[j] = e >> 6; }
    if (b_kfast) { bk[j] = e & 15; bn[j] = e >> 4; } else { bn[j] = e & 63; bk[j] = e >> 6; }
  }
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = m0 + am[j], gk = k0 + ak[j];
      ra[j] = (gm < M && gk < k_end) ? __ldg(A + gm * sam + gk * sak) : 0.f;
      const int gn = n0 + bn[j], gk2 = k0 + bk[j];
      rb[j] = (gn < N && gk2 < k_end) ? __ldg(B + gk2 * sbk + gn * sbn) : 0.f;
    }
  };
  fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += GK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[ak[j]][am[j]] = ra[j];
      Bs[bk[j]][bn[j]] = rb[j];
    }
    __syncthreads();
    if (k0 + GK < k_end) fetch(k0 + GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (partial != nullptr) {
    float* o = partial + (long long)blockIdx.z * M * N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gm = m0 + ty * 4 + i;
      if (gm >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gn = n0 + tx * 4 + j;
        if (gn < N) o[(long long)gm * N + gn] = acc[i][j];
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + gn);
      float* o = Cm + (long long)gm * ldc + gn;
      *o = accumulate ? *o + v : v;
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, const float* __restrict__ bias,
                                     float* __restrict__ Cm, int ldc, int M, int N, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = (int)(i / N), n = (int)(i - (long long)m * N);
  float v = 0.f;
  for (int z = 0; z < splits; ++z) v += partial[(long long)z * M * N + i];
  if (bias) v += __ldg(bias + n);
  float* o = Cm + (long long)m * ldc + n;
  *o = accumulate ? *o + v : v;
}

// out[n] (+)= sum_m g[m*ld + n]. Block = 32 columns x 8 row groups; fixed-order smem reduction (deterministic).
__global__ void colsum_kernel(const float* __restrict__ g, int ld, int M, int N, float* __restrict__ out, int accumulate) {
  __shared__ float sh[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float s0 = 0.f, s1 = 0.f;
  if (n < N) {
    int m = threadIdx.y;
    for (; m + 8 < M; m += 16) {
      s0 += __ldg(g + (long long)m * ld + n);
      s1 += __ldg(g + (long long)(m + 8) * ld + n);
    }
    if (m < M) s0 += __ldg(g + (long long)m * ld + n);
  }
  sh[threadIdx.y][threadIdx.x] = s0 + s1;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += sh[j][threadIdx.x];
    out[n] = accumulate ? out[n] + s : s;
  }
}

// ------------------------------------------------------------------------------------ LayerNorm (+ReLU)
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ mean_out,
                                     float* __restrict__ rstd_out, int N, int relu) {
  __shared__ float red[32];
  const float* xr = x + (long long)blockIdx.x * N;
  float* yr = y + (long long)blockIdx.x * N;
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += xr[i];
  const float mean = block_sum<float>(s, red) / (float)N;
  float v = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float d = xr[i] - mean;
    v += d * d;
  }
  const float var = block_sum<float>(v, red) / (float)N;
  const float rstd = rsqrtf(var + EPS);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float o = (xr[i] - mean) * rstd;
    if (relu) o = fmaxf(o, 0.f);
    yr[i] = o;
  }
  if (threadIdx.x == 0) {
    mean_out[blockIdx.x] = mean;
    rstd_out[blockIdx.x] = rstd;
  }
}

__global__ void layernorm_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                     float* __restrict__ gx, int N, int relu) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const float* xr = x + row * N;
  const float* gr = gy + row * N;
  float* o = gx + row * N;
  const float mu = mean[row], rs = rstd[row];
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float h = (xr[i] - mu) * rs;
    const float g = (relu && h <= 0.f) ? 0.f : gr[i];
    s1 += g;
    s2 += g * h;
  }
  const float m1 = block_sum<float>(s1, red) / (float)N;
  const float m2 = block_sum<float>(s2, red) / (float)N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float h = (xr[i] - mu) * rs;
    const float g = (relu && h <= 0.f) ? 0.f : gr[i];
    o[i] = rs * (g - m1 - h * m2);
  }
}

// ------------------------------------------------------------------------------------ row-vector x matrix
__global__ void rowvec_matmul_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mats,
                                         float* __restrict__ y, int k) {
  extern __shared__ float sx[];
  const long long t = blockIdx.x;
  for (int i = threadIdx.x; i < k; i += blockDim.x) sx[i] = x[t * k + i];
  __syncthreads();
  const float* m = mats + t * k * k;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < k; ++i) s = fmaf(sx[i], __ldg(m + (long long)i * k + j), s);
    y[t * k + j] = s;
  }
}

__global__ void rowvec_matmul_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                         const float* __restrict__ mats, float* __restrict__ gx,
                                         float* __restrict__ gmats, int k) {
  extern __shared__ float sm[];
  float* sg = sm;      // gy row
  float* sx = sm + k;  // x row
  const long long t = blockIdx.x;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    sg[i] = gy[t * k + i];
    sx[i] = x[t * k + i];
  }
  __syncthreads();
  const float* m = mats + t * k * k;
  float* gm = gmats + t * k * k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = warp; i < k; i += nw) {
    float s = 0.f;
    const float xi = sx[i];
    for (int j = lane; j < k; j += 32) {
      s = fmaf(sg[j], __ldg(m + (long long)i * k + j), s);
      gm[(long long)i * k + j] = xi * sg[j];
    }
    s = warp_sum(s);
    if (lane == 0) gx[t * k + i] = s;
  }
}

// ------------------------------------------------------------------------------------ hier_pool + concat
__global__ void segmax_concat_fwd_kernel(const float* __restrict__ local, int c_local, const float* __restrict__ x,
                                         int Cx, const int* __restrict__ img_start, float* __restrict__ out,
                                         int* __restrict__ argmax) {
  const int b = blockIdx.x;
  const int t0 = img_start[b], t1 = img_start[b + 1];
  const int ld = c_local + Cx;
  for (int c = threadIdx.x; c < Cx; c += blockDim.x) {
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
  for (int i = threadIdx.x; i < (t1 - t0) * c_local; i += blockDim.x) {
    const int t = t0 + i / c_local, c = i % c_local;
    out[(long long)t * ld + c] = local[(long long)t * c_local + c];
  }
}

__global__ void segmax_concat_bwd_kernel(const float* __restrict__ gout, int c_local, int Cx,
                                         const int* __restrict__ img_start, const int* __restrict__ argmax,
                                         float* __restrict__ glocal, float* __restrict__ gx) {
  const int b = blockIdx.x;
  const int t0 = img_start[b], t1 = img_start[b + 1];
  const int ld = c_local + Cx;
  for (int c = threadIdx.x; c < Cx; c += blockDim.x) {
    float s = 0.f;
    for (int t = t0; t < t1; ++t) s += gout[(long long)t * ld + c_local + c];
    const int bi = argmax[(long long)b * Cx + c];
    for (int t = t0; t < t1; ++t) gx[(long long)t * Cx + c] = (t == bi) ? s : 0.f;
  }
  for (int i = threadIdx.x; i < (t1 - t0) * c_local; i += blockDim.x) {
    const int t = t0 + i / c_local, c = i % c_local;
    glocal[(long long)t * c_local + c] = gout[(long long)t * ld + c];
  }
}

// ------------------------------------------------------------------------------------ attention core
// one block per (query row t, level l); E threads (one per channel). scores / probs in dynamic smem.
__global__ void attention_fwd_kernel(const float* __restrict__ q, int nsets_q, const float* __restrict__ k,
                                     const float* __restrict__ v, int nsets_kv, int T, int heads, int E,
                                     const int* __restrict__ img_of, const int* __restrict__ img_start, int max_n,
                                     float* __restrict__ out, float* __restrict__ probs) {
  extern __shared__ float sm[];
  float* sq = sm;                // E
  float* sp = sm + E;            // heads * max_n
  const int t = blockIdx.x, l = blockIdx.y;
  const int hd = E / heads;
  const float scale = rsqrtf((float)hd);
  const int b = img_of[t];
  const int s0 = img_start[b], n = img_start[b + 1] - s0;
  const float* qrow = q + ((long long)(nsets_q > 1 ? l : 0) * T + t) * E;
  const float* kb = k + ((long long)(nsets_kv > 1 ? l : 0) * T + s0) * E;
  const float* vb = v + ((long long)(nsets_kv > 1 ? l : 0) * T + s0) * E;
  for (int c = threadIdx.x; c < E; c += blockDim.x) sq[c] = qrow[c] * scale;  // q * head_dim^-0.5 like torch
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int item = warp; item < heads * n; item += nw) {
    const int h = item / n, s = item - h * n;
    float d = 0.f;
    for (int c = lane; c < hd; c += 32) d = fmaf(sq[h * hd + c], __ldg(kb + (long long)s * E + h * hd + c), d);
    d = warp_sum(d);
    if (lane == 0) sp[h * max_n + s] = d;
  }
  __syncthreads();
  for (int h = warp; h < heads; h += nw) {
    float m = -INFINITY;
    for (int s = lane; s < n; s += 32) m = fmaxf(m, sp[h * max_n + s]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float z = 0.f;
    for (int s = lane; s < n; s += 32) {
      const float e = expf(sp[h * max_n + s] - m);
      sp[h * max_n + s] = e;
      z += e;
    }
    z = warp_sum(z);
    const float inv = 1.f / z;
    float* pr = probs + (((long long)l * heads + h) * T + t) * max_n;
    for (int s = lane; s < max_n; s += 32) {
      const float pv = s < n ? sp[h * max_n + s] * inv : 0.f;
      if (s < n) sp[h * max_n + s] = pv;
      pr[s] = pv;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    const int h = c / hd;
    float o = 0.f;
    for (int s = 0; s < n; ++s) o = fmaf(sp[h * max_n + s], __ldg(vb + (long long)s * E + c), o);
    out[((long long)l * T + t) * E + c] = o;
  }
}

// query-centric backward: gq and the score gradients gs (same layout as probs)
__global__ void attention_bwd_q_kernel(const float* __restrict__ gout, const float* __restrict__ k,
                                       const float* __restrict__ v, int nsets_kv, int T, int heads, int E,
                                       const int* __restrict__ img_of, const int* __restrict__ img_start, int max_n,
                                       const float* __restrict__ probs, float* __restrict__ gs_out,
                                       float* __restrict__ gq, int nsets_q) {
  extern __shared__ float sm[];
  float* sg = sm;      // E: gout row
  float* sp = sm + E;  // heads*max_n: gp then gs
  const int t = blockIdx.x, l = blockIdx.y;
  const int hd = E / heads;
  const float scale = rsqrtf((float)hd);
  const int b = img_of[t];
  const int s0 = img_start[b], n = img_start[b + 1] - s0;
  const float* kb = k + ((long long)(nsets_kv > 1 ? l : 0) * T + s0) * E;
  const float* vb = v + ((long long)(nsets_kv > 1 ? l : 0) * T + s0) * E;
  for (int c = threadIdx.x; c < E; c += blockDim.x) sg[c] = gout[((long long)l * T + t) * E + c];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int item = warp; item < heads * n; item += nw) {
    const int h = item / n, s = item - h * n;
    float d = 0.f;
    for (int c = lane; c < hd; c += 32) d = fmaf(sg[h * hd + c], __ldg(vb + (long long)s * E + h * hd + c), d);
    d = warp_sum(d);
    if (lane == 0) sp[h * max_n + s] = d;
  }
  __syncthreads();
  for (int h = warp; h < heads; h += nw) {
    const float* pr = probs + (((long long)l * heads + h) * T + t) * max_n;
    float dot = 0.f;
    for (int s = lane; s < n; s += 32) dot = fmaf(pr[s], sp[h * max_n + s], dot);
    dot = warp_sum(dot);
    float* go = gs_out + (((long long)l * heads + h) * T + t) * max_n;
    for (int s = lane; s < max_n; s += 32) {
      const float g = s < n ? pr[s] * (sp[h * max_n + s] - dot) : 0.f;
      if (s < n) sp[h * max_n + s] = g;
      go[s] = g;
    }
  }
  __syncthreads();
  // gq[c] = scale * sum_s gs[h][s] * k[s][c]; with a shared query set (nsets_q == 1) the per-level pieces are
  // written to a (F,T,E) scratch and summed by the caller-side reduction kernel below.
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    const int h = c / hd;
    float o = 0.f;
    for (int s = 0; s < n; ++s) o = fmaf(sp[h * max_n + s], __ldg(kb + (long long)s * E + c), o);
    gq[((long long)l * T + t) * E + c] = o * scale;
  }
  (void)nsets_q;
}

// key-centric backward: one block per (key row s, kv set). Sums over the query rows of the same image and, when the
// kv set is shared by all levels (nsets_kv == 1), over the levels as well -- in a fixed order.
__global__ void attention_bwd_kv_kernel(const float* __restrict__ gout, const float* __restrict__ q, int nsets_q,
                                        int nsets_kv, int F, int T, int heads, int E, const int* __restrict__ img_of,
                                        const int* __restrict__ img_start, int max_n, const float* __restrict__ probs,
                                        const float* __restrict__ gs, float* __restrict__ gk, float* __restrict__ gv) {
  const int s = blockIdx.x, set = blockIdx.y;
  const int hd = E / heads;
  const float scale = rsqrtf((float)hd);
  const int b = img_of[s];
  const int t0 = img_start[b], n = img_start[b + 1] - t0;
  const int si = s - t0;
  const int l_begin = nsets_kv > 1 ? set : 0, l_end = nsets_kv > 1 ? set + 1 : F;
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    const int h = c / hd;
    float ak = 0.f, av = 0.f;
    for (int l = l_begin; l < l_end; ++l) {
      for (int j = 0; j < n; ++j) {
        const int t = t0 + j;
        const long long pi = (((long long)l * heads + h) * T + t) * max_n + si;
        const float qv = __ldg(q + ((long long)(nsets_q > 1 ? l : 0) * T + t) * E + c) * scale;
        ak = fmaf(__ldg(gs + pi), qv, ak);
        av = fmaf(__ldg(probs + pi), __ldg(gout + ((long long)l * T + t) * E + c), av);
      }
    }
    gk[((long long)set * T + s) * E + c] = ak;
    gv[((long long)set * T + s) * E + c] = av;
  }
}

// out[t,c] = sum_l in[l,t,c]   (reduce the per-level query gradients when the query set is shared)
__global__ void sum_sets_kernel(const float* __restrict__ in, int F, long long n, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int l = 0; l < F; ++l) s += in[(long long)l * n + i];
  out[i] = s;
}

static int launch_gemm(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn,
                       const float* bias, float* Cm, int ldc, int M, int N, int K, int accumulate, void* workspace,
                       size_t workspace_bytes, void* stream) {
  dim3 grid((N + GT - 1) / GT, (M + GT - 1) / GT, 1);
  const int tiles = grid.x * grid.y;
  // aim at ~2 CTAs per SM; never split below 64 k per CTA; stay inside the caller's workspace
  int splits = (2 * 148 + tiles - 1) / tiles;
  const int max_by_k = (K + 63) / 64;
  if (splits > max_by_k) splits = max_by_k;
  if (workspace == nullptr) splits = 1;
  while (splits > 1 && (size_t)splits * M * N * sizeof(float) > workspace_bytes) --splits;
  if (splits <= 1) {
    gemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, sam, sak, B, sbk, sbn, bias, Cm, ldc, M, N, K, accumulate, K,
                                                        nullptr);
    LGD_LAUNCH_CHECK();
    return LGD_OK;
  }
  int kps = (K + splits - 1) / splits;
  kps = (kps + GK - 1) / GK * GK;
  splits = (K + kps - 1) / kps;
  grid.z = splits;
  float* partial = static_cast<float*>(workspace);
  gemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, sam, sak, B, sbk, sbn, nullptr, Cm, ldc, M, N, K, 0, kps,
                                                      partial);
  LGD_LAUNCH_CHECK();
  const long long n = (long long)M * N;
  splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(partial, splits, bias, Cm, ldc, M,
                                                                                      N, accumulate);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

}  // namespace lgd

using namespace lgd;

extern "C" size_t lgd_linear_workspace(int M, int N, int K) {
  // launch_gemm shrinks the split factor to fit whatever it is given; 32 partials of the largest of the three
  // products of one layer (capped at 64 MiB) never constrain it for the shapes of this path
  size_t out = (size_t)M * N;
  if ((size_t)M * K > out) out = (size_t)M * K;
  if ((size_t)N * K > out) out = (size_t)N * K;
  const size_t want = 32 * out * sizeof(float), cap = (size_t)64 << 20;
  return want < cap ? want : cap;
}

extern "C" int lgd_linear_fwd(const float* x, int ldx, const float* w, int ldw, const float* bias, float* y, int ldy,
                              int M, int N, int K, void* workspace, size_t workspace_bytes, void* stream) {
  LGD_CHECK_ARG(x && w && y && M > 0 && N > 0 && K > 0, "lgd_linear_fwd: bad arguments");
  // y = x * w^T: A = x (k fast), B(k,n) = w[n*ldw + k] (k fast)
  return launch_gemm(x, ldx, 1, w, 1, ldw, bias, y, ldy, M, N, K, 0, workspace, workspace_bytes, stream);
}

extern "C" int lgd_linear_bwd_input(const float* gy, int ldgy, const float* w, int ldw, float* gx, int ldgx, int M,
                                    int N, int K, int accumulate, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  LGD_CHECK_ARG(gy && w && gx && M > 0 && N > 0 && K > 0, "lgd_linear_bwd_input: bad arguments");
  // gx[M,K] = gy[M,N] * w[N,K]: contraction over N; B(n,k) = w[n*ldw + k] (output dim fast)
  return launch_gemm(gy, ldgy, 1, w, ldw, 1, nullptr, gx, ldgx, M, K, N, accumulate, workspace, workspace_bytes,
                     stream);
}

extern "C" int lgd_linear_bwd_weight(const float* gy, int ldgy, const float* x, int ldx, float* gw, int ldgw, float* gb,
                                     int M, int N, int K, int accumulate, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  LGD_CHECK_ARG(gy && x && gw && M > 0 && N > 0 && K > 0, "lgd_linear_bwd_weight: bad arguments");
  // gw[N,K] = gy^T[N,M] * x[M,K]: A(n,m) = gy[m*ldgy + n] (m slow -> "row" index fast), B(m,k) = x[m*ldx + k]
  int rc = launch_gemm(gy, 1, ldgy, x, ldx, 1, nullptr, gw, ldgw, N, K, M, accumulate, workspace, workspace_bytes,
                       stream);
  if (rc != LGD_OK) return rc;
  if (gb) {
    colsum_kernel<<<(N + 31) / 32, dim3(32, 8), 0, (cudaStream_t)stream>>>(gy, ldgy, M, N, gb, accumulate);
    LGD_LAUNCH_CHECK();
  }
  return LGD_OK;
}

extern "C" int lgd_layernorm_fwd(const float* x, float* y, float* mean, float* rstd, int M, int N, int relu,
                                 void* stream) {
  LGD_CHECK_ARG(x && y && mean && rstd && M > 0 && N > 0, "lgd_layernorm_fwd: bad arguments");
  layernorm_fwd_kernel<<<M, 128, 0, (cudaStream_t)stream>>>(x, y, mean, rstd, N, relu);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_layernorm_bwd(const float* gy, const float* x, const float* mean, const float* rstd, float* gx,
                                 int M, int N, int relu, void* stream) {
  LGD_CHECK_ARG(gy && x && mean && rstd && gx && M > 0 && N > 0, "lgd_layernorm_bwd: bad arguments");
  layernorm_bwd_kernel<<<M, 128, 0, (cudaStream_t)stream>>>(gy, x, mean, rstd, gx, N, relu);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_rowvec_matmul_fwd(const float* x, const float* mats, float* y, int T, int k, void* stream) {
  LGD_CHECK_ARG(x && mats && y && T > 0 && k > 0, "lgd_rowvec_matmul_fwd: bad arguments");
  rowvec_matmul_fwd_kernel<<<T, 128, k * sizeof(float), (cudaStream_t)stream>>>(x, mats, y, k);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_rowvec_matmul_bwd(const float* gy, const float* x, const float* mats, float* gx, float* gmats, int T,
                                     int k, void* stream) {
  LGD_CHECK_ARG(gy && x && mats && gx && gmats && T > 0 && k > 0, "lgd_rowvec_matmul_bwd: bad arguments");
  rowvec_matmul_bwd_kernel<<<T, 256, 2 * k * sizeof(float), (cudaStream_t)stream>>>(gy, x, mats, gx, gmats, k);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_segmax_concat_fwd(const float* local, int c_local, const float* x, int Cx, const int32_t* img_start,
                                     int B, float* out, int32_t* argmax, void* stream) {
  LGD_CHECK_ARG(local && x && img_start && out && argmax && B > 0 && Cx > 0 && c_local > 0,
                "lgd_segmax_concat_fwd: bad arguments");
  segmax_concat_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(local, c_local, x, Cx, img_start, out, argmax);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_segmax_concat_bwd(const float* gout, int c_local, int Cx, const int32_t* img_start, int B,
                                     const int32_t* argmax, float* glocal, float* gx, void* stream) {
  LGD_CHECK_ARG(gout && img_start && argmax && glocal && gx && B > 0, "lgd_segmax_concat_bwd: bad arguments");
  segmax_concat_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(gout, c_local, Cx, img_start, argmax, glocal, gx);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_attention_fwd(const float* q, int nsets_q, const float* k, const float* v, int nsets_kv, int F, int T,
                                 int heads, int E, const int32_t* img_of, const int32_t* img_start, int max_n,
                                 float* out, float* probs, void* stream) {
  LGD_CHECK_ARG(q && k && v && img_of && img_start && out && probs, "lgd_attention_fwd: null pointer");
  LGD_CHECK_ARG(F > 0 && T > 0 && heads > 0 && E > 0 && E % heads == 0 && max_n > 0 && E <= 1024,
                "lgd_attention_fwd: bad shape");
  const size_t smem = (size_t)(E + heads * max_n) * sizeof(float);
  LGD_CHECK_ARG(smem <= 48 * 1024, "lgd_attention_fwd: too many boxes per image for the score buffer");
  attention_fwd_kernel<<<dim3(T, F), 256, smem, (cudaStream_t)stream>>>(q, nsets_q, k, v, nsets_kv, T, heads, E, img_of,
                                                                        img_start, max_n, out, probs);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_attention_bwd(const float* gout, const float* q, int nsets_q, const float* k, const float* v,
                                 int nsets_kv, int F, int T, int heads, int E, const int32_t* img_of,
                                 const int32_t* img_start, int max_n, const float* probs, float* gs_scratch, float* gq,
                                 float* gk, float* gv, void* stream) {
  LGD_CHECK_ARG(gout && q && k && v && img_of && img_start && probs && gs_scratch && gq && gk && gv,
                "lgd_attention_bwd: null pointer");
  LGD_CHECK_ARG(F > 0 && T > 0 && heads > 0 && E % heads == 0 && max_n > 0, "lgd_attention_bwd: bad shape");
  const size_t smem = (size_t)(E + heads * max_n) * sizeof(float);
  LGD_CHECK_ARG(smem <= 48 * 1024, "lgd_attention_bwd: too many boxes per image for the score buffer");
  float* gs = gs_scratch;
  attention_bwd_q_kernel<<<dim3(T, F), 256, smem, (cudaStream_t)stream>>>(gout, k, v, nsets_kv, T, heads, E, img_of,
                                                                          img_start, max_n, probs, gs, gq, nsets_q);
  LGD_LAUNCH_CHECK();
  attention_bwd_kv_kernel<<<dim3(T, nsets_kv > 1 ? F : 1), 256, 0, (cudaStream_t)stream>>>(
      gout, q, nsets_q, nsets_kv, F, T, heads, E, img_of, img_start, max_n, probs, gs, gk, gv);
  LGD_LAUNCH_CHECK();
  if (nsets_q == 1 && F > 1) {
    // sum the per-level pieces into a temp (gk/gv are done) -- in place is safe: out[i] reads in[l*n+i] for all l
    // before writing out[i] (same thread), and out == in set 0.
    const long long n = (long long)T * E;
    sum_sets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gq, F, n, gq);
    LGD_LAUNCH_CHECK();
  }
  return LGD_OK;
}
