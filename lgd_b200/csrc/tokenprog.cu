// Token programs: whole sequences of the latency-bound "small-T" ops (label encoder + STNs, canonical projection:
// label_encoder.py:239-274, spatial_transformer.py:30-47, dynamic_teacher.py:229) executed by ONE persistent
// cooperative kernel instead of ~50 (forward) / ~100 (backward) dependent launches.
//
// A program is a list of ops sorted by STAGE. Ops of one stage are independent of each other; the virtual blocks of
// all of them are dealt out round-robin over the CTAs of the grid (one CTA per SM), then the grid meets at a barrier
// (monotonic counter in global memory, release/acquire fences) and moves on to the next stage. With T ~ 10^2 tokens
// every op is a few microseconds of work, so a stage costs about one barrier (~2 us) instead of a kernel boundary
// plus launch latency (~8-10 us measured per dependent launch). The op bodies are the very device functions the
// per-op kernels of smallt.cu wrap (smallt_ops.cuh): same arithmetic, same order, bit-identical results.
// Inputs written earlier in the same program are read with plain (coherent) loads, never through the non-coherent
// path; parameters (weights, biases) are immutable for the kernel's lifetime.
#include <cuda_runtime.h>

#include <vector>

#include "common.cuh"
#include "smallt_ops.cuh"
#include "tokenprog.h"

namespace lgd {

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();   // release: this CTA's writes of the stage
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

static_assert(sizeof(TokOp) % 4 == 0, "ops are copied word by word");

union ProgSmem {
  GemmSmem gemm;
  float colsum[8][33];
  float red[2][32];
  float vec[2 * 160];   // rowvec: up to 2k floats, k <= 160
  struct {
    float red[2][32];
    float rows[2][1024];   // fused reduce + LayerNorm: one row per half of the CTA (N <= 1024)
  } rln;
};

__device__ __forceinline__ void run_op(ProgSmem& sm, const TokOp& op, int vb) {
  switch (op.type) {
    case TOK_GEMM: {
      GemmArgs g;
      g.A = static_cast<const float*>(op.p0); g.sam = op.l0; g.sak = op.l1;
      g.B = static_cast<const float*>(op.p1); g.sbk = op.l2; g.sbn = op.l3;
      g.bias = static_cast<const float*>(op.p2);
      g.Cm = static_cast<float*>(op.p3);
      g.ldc = op.i0; g.M = op.i1; g.N = op.i2; g.K = op.i3; g.accumulate = op.i4; g.k_per_split = op.i5;
      g.partial = static_cast<float*>(op.p4);
      const int bx = vb % op.gx, by = (vb / op.gx) % op.gy, bz = vb / (op.gx * op.gy);
      gemm_body<false>(sm.gemm, g, bx, by, bz);
      break;
    }
    case TOK_REDUCE:
      splitk_reduce_body(static_cast<const float*>(op.p4), op.i5, static_cast<const float*>(op.p2),
                         static_cast<float*>(op.p3), op.i0, op.i1, op.i2, op.i4, vb);
      break;
    case TOK_COLSUM:
      colsum_body<false>(sm.colsum, static_cast<const float*>(op.p0), op.i0, op.i1, op.i2, static_cast<float*>(op.p3),
                         op.i4, vb);
      break;
    case TOK_LN_FWD:
    case TOK_LN_BWD: {
      // two rows per CTA pass: threads 0-127 and 128-255 are two independent 128-thread groups (named barriers 1, 2)
      const int half = threadIdx.x >> 7, tid = threadIdx.x & 127;
      const int row = 2 * vb + half;
      if (row < op.i1) {
        if (op.type == TOK_LN_FWD)
          layernorm_fwd_body(sm.red[half], static_cast<const float*>(op.p0), static_cast<float*>(op.p3),
                             static_cast<float*>(op.p4), static_cast<float*>(op.p5), op.i2, op.i4, row, tid, 1 + half);
        else
          layernorm_bwd_body(sm.red[half], static_cast<const float*>(op.p0), static_cast<const float*>(op.p1),
                             static_cast<const float*>(op.p4), static_cast<const float*>(op.p5),
                             static_cast<float*>(op.p3), op.i2, op.i4, row, tid, 1 + half);
      }
      break;
    }
    case TOK_REDUCE_LN: {
      const int half = threadIdx.x >> 7, tid = threadIdx.x & 127;
      const int row = 2 * vb + half;
      if (row < op.i1)
        reduce_layernorm_fwd_body(sm.rln.red[half], sm.rln.rows[half], static_cast<const float*>(op.p0), op.i5,
                                  static_cast<const float*>(op.p2), reinterpret_cast<float*>(op.l0),
                                  static_cast<float*>(op.p3), static_cast<float*>(op.p4), static_cast<float*>(op.p5),
                                  op.i1, op.i2, op.i4, row, tid, 1 + half);
      break;
    }
    case TOK_ROWVEC_FWD:
      rowvec_fwd_body<false>(sm.vec, static_cast<const float*>(op.p0), static_cast<const float*>(op.p1),
                             static_cast<float*>(op.p3), op.i0, vb, 256);
      break;
    case TOK_ROWVEC_BWD:
      rowvec_bwd_body<false>(sm.vec, static_cast<const float*>(op.p0), static_cast<const float*>(op.p1),
                             static_cast<const float*>(op.p2), static_cast<float*>(op.p3), static_cast<float*>(op.p4),
                             op.i0, vb, 256);
      break;
    case TOK_SEGMAX_FWD:
      segmax_fwd_body(static_cast<const float*>(op.p0), op.i0, static_cast<const float*>(op.p1), op.i1,
                      static_cast<const int*>(op.p2), static_cast<float*>(op.p3), static_cast<int*>(op.p4), vb, 256);
      break;
    case TOK_SEGMAX_BWD:
      segmax_bwd_body(static_cast<const float*>(op.p0), op.i0, op.i1, static_cast<const int*>(op.p2),
                      static_cast<const int*>(op.p1), static_cast<float*>(op.p3), static_cast<float*>(op.p4), vb, 256);
      break;
    case TOK_AXPY: {
      const long long i = (long long)vb * 256 + threadIdx.x;
      if (i < op.l0) static_cast<float*>(op.p3)[i] += static_cast<const float*>(op.p0)[i];
      break;
    }
    default:
      break;
  }
}

// host_ops: the op list in pinned host memory. The CTAs pull it over PCIe with plain loads (one op = 128 bytes per warp)
// into ops, then meet at the first barrier. A cudaMemcpyAsync would queue behind whatever bulk host->device transfer
// another stream has in flight on the same copy engine (measured: the 367 MB feature upload of the end-to-end loop
// delayed the 30 KB program by a whole step).
__global__ void __launch_bounds__(256, 1)
token_program_kernel(const TokOp* host_ops, TokOp* ops, int nops, unsigned int* barrier) {
  __shared__ ProgSmem sm;
  const int G = gridDim.x;
  unsigned int target = 0;
  {
    constexpr int WORDS = sizeof(TokOp) / 4;
    const volatile unsigned int* src = reinterpret_cast<const volatile unsigned int*>(host_ops);
    unsigned int* dst = reinterpret_cast<unsigned int*>(ops);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = blockIdx.x * 8 + warp; j < nops; j += G * 8)
      for (int w = lane; w < WORDS; w += 32) dst[(long long)j * WORDS + w] = src[(long long)j * WORDS + w];
    target += G;
    grid_barrier(barrier, target);
  }
  int i = 0;
  while (i < nops) {
    const int stage = ops[i].stage;
    int base = 0;
    int j = i;
    for (; j < nops && ops[j].stage == stage; ++j) {
      const int n = ops[j].nblocks;
      // global index of this op's virtual block v is base + v; CTA b takes the indices congruent to b modulo G
      int v = (int)blockIdx.x - base % G;
      if (v < 0) v += G;
      for (; v < n; v += G) {
        run_op(sm, ops[j], v);
        __syncthreads();   // shared memory is reused by the next virtual block
      }
      base += n;
    }
    i = j;
    if (i < nops) {
      target += G;
      grid_barrier(barrier, target);
    }
  }
}

// ------------------------------------------------------------------------------------------------ host-side builder
void TokenProgram::linear(const float* x, int ldx, const float* w, int ldw, const float* bias, float* y, int ldy, int M,
                          int N, int K) {
  gemm(x, ldx, 1, w, 1, ldw, bias, y, ldy, M, N, K, 0);
}
void TokenProgram::linear_bwd_input(const float* gy, int ldgy, const float* w, int ldw, float* gx, int ldgx, int M, int N,
                                    int K, int accumulate) {
  gemm(gy, ldgy, 1, w, ldw, 1, nullptr, gx, ldgx, M, K, N, accumulate);
}
void TokenProgram::linear_bwd_weight(const float* gy, int ldgy, const float* x, int ldx, float* gw, int ldgw, float* gb,
                                     int M, int N, int K) {
  gemm(gy, 1, ldgy, x, ldx, 1, nullptr, gw, ldgw, N, K, M, 0);
  if (gb != nullptr) {
    TokOp op{};
    op.type = TOK_COLSUM;
    op.stage = stage_;
    op.nblocks = (N + 31) / 32;
    op.p0 = gy; op.i0 = ldgy; op.i1 = M; op.i2 = N; op.p3 = gb; op.i4 = 0;
    ops_.push_back(op);
  }
}

// one GEMM in the current stage; when it is split over K, its reduction is queued for the next stage (flush_reduces)
void TokenProgram::gemm(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn,
                        const float* bias, float* Cm, int ldc, int M, int N, int K, int accumulate) {
  // every GEMM sees a workspace slice of the size the per-op path gives it, so the split decision is the same
  float* slice = nullptr;
  const size_t slice_bytes = slice_bytes_;
  const int arena = stage_ & 1;
  if (used_[arena] + slice_bytes <= arena_bytes_) {
    slice = reinterpret_cast<float*>(static_cast<char*>(arena_[arena]) + used_[arena]);
  }
  const GemmPlan p = plan_gemm(M, N, K, slice != nullptr, slice_bytes);
  TokOp op{};
  op.type = TOK_GEMM;
  op.stage = stage_;
  op.gx = p.gx; op.gy = p.gy; op.gz = p.splits;
  op.nblocks = p.gx * p.gy * p.splits;
  op.p0 = A; op.l0 = sam; op.l1 = sak; op.p1 = B; op.l2 = sbk; op.l3 = sbn;
  op.p3 = Cm; op.i0 = ldc; op.i1 = M; op.i2 = N; op.i3 = K;
  if (p.splits <= 1) {
    op.p2 = bias; op.i4 = accumulate; op.i5 = K; op.p4 = nullptr;
    ops_.push_back(op);
    return;
  }
  used_[arena] += slice_bytes;
  op.p2 = nullptr; op.i4 = 0; op.i5 = p.kps; op.p4 = slice;
  ops_.push_back(op);
  TokOp r{};
  r.type = TOK_REDUCE;
  r.nblocks = (int)(((long long)M * N + 255) / 256);
  r.p4 = slice; r.i5 = p.splits; r.p2 = bias; r.p3 = Cm; r.i0 = ldc; r.i1 = M; r.i2 = N; r.i4 = accumulate;
  pending_.push_back(r);
}

// ends the current stage; the reductions of its split GEMMs form the next stage (followed by another boundary)
void TokenProgram::next_stage() {
  ++stage_;
  used_[stage_ & 1] = 0;
  if (!pending_.empty()) {
    for (auto& r : pending_) {
      r.stage = stage_;
      ops_.push_back(r);
    }
    pending_.clear();
    ++stage_;
    used_[stage_ & 1] = 0;
  }
}

void TokenProgram::linear_layernorm(const float* x, int ldx, const float* w, int ldw, const float* bias, float* pre,
                                    float* y, float* mean, float* rstd, int M, int N, int K, int relu) {
  const size_t before = pending_.size();
  gemm(x, ldx, 1, w, 1, ldw, bias, pre, N, M, N, K, 0);
  if (pending_.size() == before || N > 1024) {   // not split over K (or a row does not fit): plain LayerNorm stage
    next_stage();
    layernorm_fwd(pre, y, mean, rstd, M, N, relu);
    next_stage();
    return;
  }
  // the queued reduction becomes a fused reduce + LayerNorm op
  TokOp r = pending_.back();
  pending_.pop_back();
  ++stage_;
  used_[stage_ & 1] = 0;
  for (auto& other : pending_) {   // reductions of other GEMMs of the closed stage run next to the fused op
    other.stage = stage_;
    ops_.push_back(other);
  }
  pending_.clear();
  TokOp op{};
  op.type = TOK_REDUCE_LN; op.stage = stage_; op.nblocks = (M + 1) / 2;
  op.p0 = r.p4; op.i5 = r.i5; op.p2 = bias; op.l0 = reinterpret_cast<long long>(pre);
  op.p3 = y; op.p4 = mean; op.p5 = rstd; op.i1 = M; op.i2 = N; op.i4 = relu;
  ops_.push_back(op);
  next_stage();
}

void TokenProgram::layernorm_fwd(const float* x, float* y, float* mean, float* rstd, int M, int N, int relu) {
  TokOp op{};
  op.type = TOK_LN_FWD; op.stage = stage_; op.nblocks = (M + 1) / 2;
  op.p0 = x; op.p3 = y; op.p4 = mean; op.p5 = rstd; op.i1 = M; op.i2 = N; op.i4 = relu;
  ops_.push_back(op);
}
void TokenProgram::layernorm_bwd(const float* gy, const float* x, const float* mean, const float* rstd, float* gx, int M,
                                 int N, int relu) {
  TokOp op{};
  op.type = TOK_LN_BWD; op.stage = stage_; op.nblocks = (M + 1) / 2;
  op.p0 = gy; op.p1 = x; op.p4 = const_cast<float*>(mean); op.p5 = const_cast<float*>(rstd); op.p3 = gx;
  op.i1 = M; op.i2 = N; op.i4 = relu;
  ops_.push_back(op);
}
void TokenProgram::rowvec_fwd(const float* x, const float* mats, float* y, int T, int k) {
  TokOp op{};
  op.type = TOK_ROWVEC_FWD; op.stage = stage_; op.nblocks = T;
  op.p0 = x; op.p1 = mats; op.p3 = y; op.i0 = k;
  ops_.push_back(op);
}
void TokenProgram::rowvec_bwd(const float* gy, const float* x, const float* mats, float* gx, float* gmats, int T, int k) {
  TokOp op{};
  op.type = TOK_ROWVEC_BWD; op.stage = stage_; op.nblocks = T;
  op.p0 = gy; op.p1 = x; op.p2 = mats; op.p3 = gx; op.p4 = gmats; op.i0 = k;
  ops_.push_back(op);
}
void TokenProgram::segmax_fwd(const float* local, int c_local, const float* x, int Cx, const int32_t* img_start, int B,
                              float* out, int32_t* argmax) {
  TokOp op{};
  op.type = TOK_SEGMAX_FWD; op.stage = stage_; op.nblocks = B;
  op.p0 = local; op.i0 = c_local; op.p1 = x; op.i1 = Cx; op.p2 = img_start; op.p3 = out; op.p4 = argmax;
  ops_.push_back(op);
}
void TokenProgram::segmax_bwd(const float* gout, int c_local, int Cx, const int32_t* img_start, int B,
                              const int32_t* argmax, float* glocal, float* gx) {
  TokOp op{};
  op.type = TOK_SEGMAX_BWD; op.stage = stage_; op.nblocks = B;
  op.p0 = gout; op.i0 = c_local; op.i1 = Cx; op.p2 = img_start; op.p1 = argmax; op.p3 = glocal; op.p4 = gx;
  ops_.push_back(op);
}
void TokenProgram::axpy(const float* x, float* y, long long n) {
  TokOp op{};
  op.type = TOK_AXPY; op.stage = stage_; op.nblocks = (int)((n + 255) / 256);
  op.p0 = x; op.p3 = y; op.l0 = n;
  ops_.push_back(op);
}

size_t TokenProgram::device_bytes(int max_ops) { return (size_t)max_ops * sizeof(TokOp) + 256; }

// uploads the program into `dev_prog` (device_bytes(ops) bytes; the first 256 bytes hold the barrier counter) and
// launches the executor cooperatively on `stream`
int TokenProgram::launch(void* dev_prog, size_t dev_prog_bytes, void* pinned_staging, void* stream) {
  if (!pending_.empty()) next_stage();
  if (ops_.empty()) return LGD_OK;
  LGD_CHECK_ARG(dev_prog != nullptr && dev_prog_bytes >= device_bytes((int)ops_.size()),
                "token program: device buffer too small (%zu ops)", ops_.size());
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    LGD_CUDA(cudaGetDevice(&dev));
    LGD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  cudaStream_t s = (cudaStream_t)stream;
  unsigned int* barrier = static_cast<unsigned int*>(dev_prog);
  TokOp* dev_ops = reinterpret_cast<TokOp*>(static_cast<char*>(dev_prog) + 256);
  LGD_CUDA(cudaMemsetAsync(barrier, 0, 256, s));
  LGD_CHECK_ARG(pinned_staging != nullptr, "token program: no pinned staging buffer");
  memcpy(pinned_staging, ops_.data(), ops_.size() * sizeof(TokOp));
  const TokOp* host_ops = static_cast<const TokOp*>(pinned_staging);   // read by the kernel itself (see above)
  int nops = (int)ops_.size();
  void* args[] = {&host_ops, &dev_ops, &nops, &barrier};
  LGD_CUDA(cudaLaunchCooperativeKernel((const void*)token_program_kernel, dim3(sms), dim3(256), args, 0, s));
  count_launch();
  return LGD_OK;
}

}  // namespace lgd
