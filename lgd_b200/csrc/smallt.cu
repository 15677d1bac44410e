// "Small-T" kernels: everything that runs on the T (~10^2) object tokens of a batch -- the PointNet-style
// label encoder + STNs (a2: label_encoder.py:216-276, spatial_transformer.py:30-47), the 1-D projections
// (a3/a7: dynamic_teacher.py:56-65) and the block-diagonal multi-head attention (a6: dynamic_teacher.py:255-275).
// < 0.2 % of the step's FLOPs and latency-bound: plain fp32 SIMT kernels (bit-for-bit fp32 semantics, no
// tensor cores), deterministic reductions.
#include "common.cuh"
#include "smallt_ops.cuh"

namespace lgd {

// ------------------------------------------------------------------------------------ per-op kernels (bodies: smallt_ops.cuh)
__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs g) {
  __shared__ GemmSmem sm;
  gemm_body<true>(sm, g, blockIdx.x, blockIdx.y, blockIdx.z);
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, const float* __restrict__ bias,
                                     float* __restrict__ Cm, int ldc, int M, int N, int accumulate) {
  splitk_reduce_body(partial, splits, bias, Cm, ldc, M, N, accumulate, blockIdx.x);
}

__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ g, int ld, int M, int N,
                                                     float* __restrict__ out, int accumulate) {
  __shared__ float sh[8][33];
  colsum_body<true>(sh, g, ld, M, N, out, accumulate, blockIdx.x);
}

__global__ void __launch_bounds__(128) layernorm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            int N, int relu) {
  __shared__ float red[32];
  layernorm_fwd_body(red, x, y, mean_out, rstd_out, N, relu, blockIdx.x, threadIdx.x, 0);
}

__global__ void __launch_bounds__(128) layernorm_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, float* __restrict__ gx, int N,
                                                            int relu) {
  __shared__ float red[32];
  layernorm_bwd_body(red, gy, x, mean, rstd, gx, N, relu, blockIdx.x, threadIdx.x, 0);
}

__global__ void rowvec_matmul_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mats,
                                         float* __restrict__ y, int k) {
  extern __shared__ float sx[];
  rowvec_fwd_body<true>(sx, x, mats, y, k, blockIdx.x, blockDim.x);
}

__global__ void rowvec_matmul_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                         const float* __restrict__ mats, float* __restrict__ gx,
                                         float* __restrict__ gmats, int k) {
  extern __shared__ float sm[];
  rowvec_bwd_body<true>(sm, gy, x, mats, gx, gmats, k, blockIdx.x, blockDim.x);
}

__global__ void segmax_concat_fwd_kernel(const float* __restrict__ local, int c_local, const float* __restrict__ x,
                                         int Cx, const int* __restrict__ img_start, float* __restrict__ out,
                                         int* __restrict__ argmax) {
  segmax_fwd_body(local, c_local, x, Cx, img_start, out, argmax, blockIdx.x, blockDim.x);
}

__global__ void segmax_concat_bwd_kernel(const float* __restrict__ gout, int c_local, int Cx,
                                         const int* __restrict__ img_start, const int* __restrict__ argmax,
                                         float* __restrict__ glocal, float* __restrict__ gx) {
  segmax_bwd_body(gout, c_local, Cx, img_start, argmax, glocal, gx, blockIdx.x, blockDim.x);
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
  const GemmPlan p = plan_gemm(M, N, K, workspace != nullptr, workspace_bytes);
  GemmArgs g{A, sam, sak, B, sbk, sbn, bias, Cm, ldc, M, N, K, accumulate, K, nullptr};
  dim3 grid(p.gx, p.gy, 1);
  if (p.splits <= 1) {
    gemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g);
    LGD_LAUNCH_CHECK();
    return LGD_OK;
  }
  grid.z = p.splits;
  float* partial = static_cast<float*>(workspace);
  g.bias = nullptr;
  g.accumulate = 0;
  g.k_per_split = p.kps;
  g.partial = partial;
  gemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g);
  LGD_LAUNCH_CHECK();
  const long long n = (long long)M * N;
  splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(partial, p.splits, bias, Cm, ldc, M,
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
    colsum_kernel<<<(N + 31) / 32, 256, 0, (cudaStream_t)stream>>>(gy, ldgy, M, N, gb, accumulate);
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
