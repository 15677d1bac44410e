// Probe (development tool): issue rate of tcgen05.mma.cta_group::2.kind::f16 M256 N256 K16 on B200, alone and with TMA
// traffic of the convolution kernel's volume streaming into the same CTA's shared memory. Answers: what tensor-pipe
// ceiling can conv3x3_tc_kernel reach at a given clock, and how much of it do the operand loads cost?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../lgd_b200/csrc -o probe_mma_rate probe_mma_rate.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace lgd;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int STG = 12;            // barrier slots; the ring depth in use is a kernel argument
constexpr int STG_BYTES = 32768;   // A 16 KiB + B 16 KiB (depth <= 6), or B only with one static A tile (b_only)

// mode 0: MMAs only (operands resident). mode 1: + a producer streaming `bytes_per_group` per 4 MMAs through a ring the
// MMAs wait on (the kernel's real dependency structure). mode 2: producer streams but MMAs do not wait (interference only).
// CONVERGENT: the whole MMA warp runs the loop and one elected lane issues (operands on the uniform datapath) instead of
// issuing from a lane-divergent branch (ptxas then wraps every MMA in an ELECT / R2UR.BROADCAST loop)
template <bool CONVERGENT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
rate_kernel(const __grid_constant__ CUtensorMap tm, int groups, int mode, int tma_bytes, long long* clocks, int depth, int b_only, int reps) {
  const int sbytes = b_only ? 16384 : STG_BYTES;   // b_only: ring of weight tiles at base + 16 KiB, A static at base
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[STG], empty[STG], done;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < STG; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(&tmem_ptr, 512);
  for (int i = threadIdx.x; i < 6 * STG_BYTES / 4 + 4096; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  long long t0 = 0, t1 = 0;
  if (warp == 0 && lane == 0 && mode != 0) {
    int stage = 0; uint32_t phase = 0;
    for (int g = 0; g < groups; ++g) {
      if (mode != 2) mbar_wait(&empty[stage], phase ^ 1);
      if (mode == 2) {   // nobody consumes the loads: each CTA tracks its own (local barrier), re-arming only a completed phase
        if (g >= depth) mbar_wait(&full[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], tma_bytes);
        for (int off = 0; off < tma_bytes; off += 16384)
          tma_load_2d(base + (b_only ? 16384 : 0) + stage * sbytes + off, &tm, &full[stage], 0, ((g * 2 + off / 16384) * 148 + blockIdx.x) % 4096 * 128);
      } else {
        const uint32_t full_leader = mapa_shared(smem_u32(&full[stage]), 0);
        if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * tma_bytes);
        for (int off = 0; off < tma_bytes; off += 16384)
          tma_load_2d_2sm(base + (b_only ? 16384 : 0) + stage * sbytes + off, &tm, full_leader, 0, ((g * 2 + off / 16384) * 148 + blockIdx.x) % 4096 * 128);
      }
      if (++stage == depth) { stage = 0; phase ^= 1; }
    }
    if (mode == 2) {   // drain before the CTA may exit
      for (int g = groups; g < groups + depth && g >= depth; ++g) {
        mbar_wait(&full[stage], phase ^ 1);
        if (++stage == depth) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && (CONVERGENT || lane == 0) && rank == 0) {
    constexpr uint32_t idesc = make_idesc_f16(256, 256);
    int stage = 0; uint32_t phase = 0;
    t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (mode == 1) { mbar_wait(&full[stage], phase); tc_fence_after(); }
      const uint64_t ad = make_smem_desc_sw128(smem_u32(base + (b_only ? 0 : stage * sbytes)), 16, 1024);
      const uint64_t bd = make_smem_desc_sw128(smem_u32(base + 16384 + stage * sbytes), 16, 1024);
      if (!CONVERGENT || elect_one()) {
        for (int r = 0; r < reps; ++r) {   // reps x 4 MMAs per operand handshake
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_f16_ss_2sm(tmem + (g & 1) * 256, ad + 2 * k, bd + 2 * k, idesc, 1u);
        }
        if (mode == 1) mma_commit_2sm(&empty[stage], 3);
      }
      if (CONVERGENT) __syncwarp();
      if (++stage == depth) { stage = 0; phase ^= 1; }
    }
    if (!CONVERGENT || elect_one()) mma_commit_2sm(&done, 1);
    mbar_wait(&done, 0);
    t1 = clock64();
    if (lane == 0) clocks[blockIdx.x >> 1] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem, 512);
}

typedef CUresult (*EncTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  // a 64 MiB fp16 matrix [rows][64] to stream from (L2 resident after the first pass)
  const size_t rows = 4096 * 128;
  __half* d;
  CK(cudaMalloc(&d, rows * 128));
  CK(cudaMemset(d, 0, rows * 128));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  CUtensorMap tm;
  cuuint64_t dims[2] = {64, rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t e1[2] = {1, 1};
  if (((EncTiled)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, e1, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("encode failed\n");
    return 1;
  }
  long long* dclk;
  CK(cudaMalloc(&dclk, 74 * 8));
  const int smem = 6 * STG_BYTES + 16384 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1v;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1v));
  const int groups = 20000;   // 80000 MMAs per pair ~ 10 M clocks at 128 clk each ~ 6 ms
  struct Cfg { int mode, bytes, depth, b_only; const char* name; int reps = 1; int conv = 0; } cfgs[] = {
      {0, 0, 4, 0, "MMA only"}, {2, 32768, 4, 0, "MMA + 32 KiB/group TMA, no dependency"},
      {1, 32768, 4, 0, "MMA waits on 32 KiB/group TMA, ring of 4 (old kernel)"},
      {1, 32768, 6, 0, "MMA waits on 32 KiB/group TMA, ring of 6"},
      {1, 16384, 4, 1, "MMA waits on 16 KiB/group TMA (weights), ring of 4"},
      {1, 16384, 5, 1, "MMA waits on 16 KiB/group TMA (weights), ring of 5"},
      {1, 16384, 6, 1, "MMA waits on 16 KiB/group TMA (weights), ring of 6"},
      {1, 16384, 8, 1, "MMA waits on 16 KiB/group TMA (weights), ring of 8"},
      {1, 16384, 12, 1, "MMA waits on 16 KiB/group TMA (weights), ring of 12"},
      {1, 16384, 6, 1, "8 MMAs per handshake, 16 KiB/group, ring of 6", 2},
      {1, 32768, 6, 0, "8 MMAs per handshake, 32 KiB/group, ring of 6", 2},
      {1, 32768, 6, 0, "16 MMAs per handshake, 32 KiB/group, ring of 6", 4},
      {0, 0, 4, 0, "convergent warp + elect: MMA only", 1, 1},
      {1, 16384, 4, 1, "convergent warp + elect: waits on 16 KiB/group, ring of 4", 1, 1},
      {1, 32768, 4, 0, "convergent warp + elect: waits on 32 KiB/group, ring of 4", 1, 1},
      {1, 32768, 6, 0, "convergent warp + elect: waits on 32 KiB/group, ring of 6", 1, 1},
      {0, 0, 4, 0, "MMA only (again)"}};
  for (auto c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      if (c.conv)
        rate_kernel<true><<<148, 128, smem>>>(tm, groups, c.mode, c.bytes ? c.bytes : 16384, dclk, c.depth, c.b_only, c.reps);
      else
        rate_kernel<false><<<148, 128, smem>>>(tm, groups, c.mode, c.bytes ? c.bytes : 16384, dclk, c.depth, c.b_only, c.reps);
      CK(cudaEventRecord(e1v));
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel error: %s\n", cudaGetErrorString(e)); return 1; }
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1v));
      std::vector<long long> h(74);
      CK(cudaMemcpy(h.data(), dclk, 74 * 8, cudaMemcpyDeviceToHost));
      double avg = 0; long long mx = 0;
      for (auto v : h) { avg += (double)v; if (v > mx) mx = v; }
      avg /= 74;
      const double flops = 74.0 * groups * 4 * c.reps * 2.0 * 256 * 256 * 16;
      printf("%-58s rep %d: %.3f ms, %.1f clk/MMA (max pair %.1f), %.0f TFLOP/s, implied clock %.0f MHz\n", c.name, rep, ms,
             avg / (groups * 4.0 * c.reps), (double)mx / (groups * 4.0 * c.reps), flops / (ms * 1e-3) / 1e12, avg / (ms * 1e-3) / 1e6);
      fflush(stdout);
    }
  }
  return 0;
}
