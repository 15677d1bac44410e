// Host-side builder + device op format of the token programs (tokenprog.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace lgd {

enum TokOpType : int {
  TOK_GEMM = 1, TOK_REDUCE, TOK_COLSUM, TOK_LN_FWD, TOK_LN_BWD, TOK_ROWVEC_FWD, TOK_ROWVEC_BWD, TOK_SEGMAX_FWD,
  TOK_SEGMAX_BWD, TOK_AXPY, TOK_REDUCE_LN
};

struct TokOp {
  int type, stage, nblocks, gx, gy, gz;
  int i0, i1, i2, i3, i4, i5;
  const void *p0, *p1, *p2;
  void *p3, *p4, *p5;
  long long l0, l1, l2, l3;
};

// Ops are appended to the CURRENT stage; next_stage() closes it. A GEMM that the split-K plan cuts along K queues its
// reduction, which next_stage() places in a stage of its own right after (so a consumer added after next_stage() sees
// the reduced result). Split-K partials live in two arenas that alternate with the stage parity; every GEMM gets a
// slice of `slice_bytes`, the workspace size the per-op path (lgd_linear_*) hands its GEMMs, so both take the same
// split decision and produce the same bits.
class TokenProgram {
 public:
  TokenProgram(void* arena0, void* arena1, size_t arena_bytes, size_t slice_bytes)
      : arena_bytes_(arena_bytes), slice_bytes_(slice_bytes) {
    arena_[0] = arena0;
    arena_[1] = arena1;
    used_[0] = used_[1] = 0;
  }
  void linear(const float* x, int ldx, const float* w, int ldw, const float* bias, float* y, int ldy, int M, int N, int K);
  void linear_bwd_input(const float* gy, int ldgy, const float* w, int ldw, float* gx, int ldgx, int M, int N, int K,
                        int accumulate);
  void linear_bwd_weight(const float* gy, int ldgy, const float* x, int ldx, float* gw, int ldgw, float* gb, int M, int N,
                         int K);
  // y = relu?(LN(x w^T + b)), pre = x w^T + b: GEMM in the current stage, then ONE stage that reduces the split-K
  // partials and normalises (closes two stages: call nothing in between)
  void linear_layernorm(const float* x, int ldx, const float* w, int ldw, const float* bias, float* pre, float* y,
                        float* mean, float* rstd, int M, int N, int K, int relu);
  void layernorm_fwd(const float* x, float* y, float* mean, float* rstd, int M, int N, int relu);
  void layernorm_bwd(const float* gy, const float* x, const float* mean, const float* rstd, float* gx, int M, int N,
                     int relu);
  void rowvec_fwd(const float* x, const float* mats, float* y, int T, int k);
  void rowvec_bwd(const float* gy, const float* x, const float* mats, float* gx, float* gmats, int T, int k);
  void segmax_fwd(const float* local, int c_local, const float* x, int Cx, const int32_t* img_start, int B, float* out,
                  int32_t* argmax);
  void segmax_bwd(const float* gout, int c_local, int Cx, const int32_t* img_start, int B, const int32_t* argmax,
                  float* glocal, float* gx);
  void axpy(const float* x, float* y, long long n);
  void next_stage();
  size_t size() const { return ops_.size() + pending_.size(); }
  static size_t device_bytes(int max_ops);
  // pinned_staging: host-pinned buffer of device_bytes(size()) bytes that stays untouched until the copy has run
  int launch(void* dev_prog, size_t dev_prog_bytes, void* pinned_staging, void* stream);

 private:
  void gemm(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn,
            const float* bias, float* Cm, int ldc, int M, int N, int K, int accumulate);
  std::vector<TokOp> ops_, pending_;
  void* arena_[2];
  size_t used_[2];
  size_t arena_bytes_, slice_bytes_;
  int stage_ = 0;
};

}  // namespace lgd
