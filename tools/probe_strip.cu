// Probe (development tool, not part of the library): can ONE shared-memory copy of an input strip serve all nine taps of
// a 3x3 convolution on tcgen05?
//   (1) im2col-mode TMA with bounding-box corners (-1, +1) and zero offsets walks the zero-PADDED image (W+2 positions
//       per row, H+2 rows) in flat order: S consecutive "padded-flat" pixels x 64 fp16 channels land K-major, SWIZZLE_128B.
//   (2) a K-major SWIZZLE_128B UMMA descriptor whose start address is shifted by an arbitrary number of 128-byte rows
//       (tap (dy,dx) = (dy*(W+2)+dx) rows) reads the rows it should -- with the descriptor's base-offset field left 0
//       or set to (start >> 7) & 7.
// D[128, 64] = A[128 rows from the shifted strip, K=64] * I (identity as the B operand), so D must equal the strip rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o probe_strip probe_strip.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int STRIP = 512;   // pixels per strip (64 KiB)

__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_eye, float* out, __half* strip_out,
      int w, int h, int n, int shift_rows, int use_base_offset, int npix_ops) {
  extern __shared__ uint8_t smem[];
  __shared__ uint64_t bar, mbar;
  __shared__ uint32_t tmem_ptr;
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  uint8_t* eye = base + STRIP * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < STRIP * 64; i += blockDim.x) ((__half*)base)[i] = __float2half(-7.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(STRIP * 128 + 64 * 128) : "memory");
    // the strip: npix_ops im2col operations of STRIP / npix_ops pixels each, continuing where the previous one stopped
    const int per = STRIP / npix_ops;
    int pw = w, ph = h, pn = n;   // padded-box coordinates of the first pixel (lower corner = -1)
    for (int o = 0; o < npix_ops; ++o) {
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
          ::"r"(smem_u32(base + o * per * 128)), "l"((uint64_t)&tm_act), "r"(smem_u32(&bar)), "r"(0), "r"(pw), "r"(ph), "r"(pn),
            "h"((uint16_t)0), "h"((uint16_t)0) : "memory");
      (void)per;
    }
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(eye)), "l"((uint64_t)&tm_eye), "r"(smem_u32(&bar)), "r"(0), "r"(0) : "memory");
    wait_bar(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t a_addr = smem_u32(base) + shift_rows * 128;
    uint64_t ad = (uint64_t)((a_addr & 0x3FFFF) >> 4) | ((uint64_t)(16 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
                  ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    if (use_base_offset) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t bd = (uint64_t)((smem_u32(eye) & 0x3FFFF) >> 4) | ((uint64_t)(16 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
                        ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t idesc = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);   // f16 x f16 -> f32, M=128, N=64
    for (int k = 0; k < 4; ++k) {
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                   ::"r"(tmem), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(k ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  wait_bar(&mbar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < 64; ++c) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[(warp * 32 + lane) * 64 + c] = __uint_as_float(v);
  }
  for (int i = threadIdx.x; i < STRIP * 64; i += blockDim.x) strip_out[i] = ((__half*)base)[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

typedef CUresult (*EncIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int C = 64, W = 21, H = 13, N = 4;
  std::vector<__half> h((size_t)N * H * W * C);
  for (size_t i = 0; i < h.size(); ++i) h[i] = __float2half((float)((long)(i * 7 % 2039) - 1000));
  std::vector<__half> eye(64 * 64, __float2half(0.f));
  for (int i = 0; i < 64; ++i) eye[i * 64 + i] = __float2half(1.f);
  __half *d, *deye, *dstrip;
  float* dout;
  CK(cudaMalloc(&d, h.size() * 2));
  CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&deye, eye.size() * 2));
  CK(cudaMemcpy(deye, eye.data(), eye.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dout, 128 * 64 * 4));
  CK(cudaMalloc(&dstrip, STRIP * 64 * 2));
  void *fn = nullptr, *fn2 = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn2, cudaEnableDefault, &q));
  CUtensorMap tm, tme;
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lower[2] = {-1, -1}, upper[2] = {1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int pix = STRIP; pix >= 128; pix >>= 1) {
    CUresult r = ((EncIm2col)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, strides, lower, upper, 64, pix, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode im2col, corners (-1,+1), pixelsPerColumn %d: result %d\n", pix, (int)r);
    if (r == CUDA_SUCCESS) {
      if (pix != STRIP) { printf("(only probing the encoder for smaller boxes)\n"); }
      if (pix == STRIP) break;
    }
    if (pix == 128 && r != CUDA_SUCCESS) return 1;
  }
  {
    CUresult r = ((EncIm2col)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, strides, lower, upper, 64, STRIP, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cannot encode a %d-pixel strip\n", STRIP); return 1; }
  }
  cuuint64_t edims[2] = {64, 64};
  cuuint64_t estrides[1] = {128};
  cuuint32_t ebox[2] = {64, 64};
  cuuint32_t e1[2] = {1, 1};
  CUresult r2 = ((EncTiled)fn2)(&tme, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, deye, edims, estrides, ebox, e1, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r2 != CUDA_SUCCESS) { printf("eye encode %d\n", (int)r2); return 1; }
  const int smem_bytes = STRIP * 128 + 64 * 128 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int PW = W + 2, PH = H + 2;
  // strip start positions in padded coordinates (x in [-1, W], y in [-1, H]) of image n
  struct Start { int x, y, n; } starts[] = {{-1, -1, 0}, {5, 3, 1}, {20, 12, 2}};
  const int shifts[] = {0, 1, 2, 7, 8, 9, PW - 1, PW, PW + 1, 2 * PW, 2 * PW + 2, 301, 383};
  int total_bad = 0;
  for (auto st : starts) {
    // expected strip in padded-flat order
    std::vector<float> want((size_t)STRIP * 64, 0.f);
    long f0 = ((long)st.n * PH + (st.y + 1)) * PW + (st.x + 1);
    for (int i = 0; i < STRIP; ++i) {
      long f = f0 + i;
      int n = (int)(f / (PH * PW)), py = (int)((f / PW) % PH) - 1, px = (int)(f % PW) - 1;
      if (n < N && py >= 0 && py < H && px >= 0 && px < W)
        for (int c = 0; c < 64; ++c) want[(size_t)i * 64 + c] = __half2float(h[(((size_t)n * H + py) * W + px) * C + c]);
    }
    for (int bo = 0; bo < 2; ++bo) {
      for (int sh : shifts) {
        if (sh + 128 > STRIP) continue;
        CK(cudaMemset(dout, 0, 128 * 64 * 4));
        probe<<<1, 128, smem_bytes>>>(tm, tme, dout, dstrip, st.x, st.y, st.n, sh, bo, 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> o(128 * 64);
        std::vector<__half> so((size_t)STRIP * 64);
        CK(cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(so.data(), dstrip, so.size() * 2, cudaMemcpyDeviceToHost));
        int bad_strip = 0, bad_mma = 0;
        for (int i = 0; i < STRIP; ++i)
          for (int c = 0; c < 64; ++c) {
            const int chunk = (c >> 3) ^ (i & 7);   // SWIZZLE_128B on a 1024-aligned base: 16-byte chunk XOR (row & 7)
            if (__half2float(so[(size_t)i * 64 + chunk * 8 + (c & 7)]) != want[(size_t)i * 64 + c]) ++bad_strip;
          }
        for (int m = 0; m < 128; ++m)
          for (int c = 0; c < 64; ++c)
            if (o[m * 64 + c] != want[(size_t)(m + sh) * 64 + c]) {
              if (!bad_mma) printf("    first mma mismatch row %d c %d: got %.0f want %.0f\n", m, c, o[m * 64 + c], want[(size_t)(m + sh) * 64 + c]);
              ++bad_mma;
            }
        printf("start (%d,%d,%d) shift %3d rows, base_offset %s: strip %s (%d bad), mma %s (%d bad)\n", st.x, st.y, st.n, sh,
               bo ? "set" : "0", bad_strip ? "FAIL" : "ok", bad_strip, bad_mma ? "FAIL" : "ok", bad_mma);
        total_bad += bad_mma + bad_strip;
      }
    }
  }
  printf("TOTAL bad %d\n", total_bad);
  return 0;
}
