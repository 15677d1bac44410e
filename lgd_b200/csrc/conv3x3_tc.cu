// K1: 3x3 / stride 1 / pad 1 / 256->256 convolution over a whole FPN pyramid as ONE persistent, warp-specialised
// tcgen05 kernel running on CTA pairs (replaces F.conv2d -> cuDNN at layers.py:25, dynamic_teacher.py:61,68-72 and
// adapters/sequential_convs.py:11-13 of the reference).
//
// Implicit GEMM, NHWC fp32 activations, TF32 operands (pre-rounded rna by the producer kernels), fp32 accumulation
// in TMEM, cta_group::2 (one MMA spans two SMs):
//   forward / dgrad : D[256 pixels, 256 co] += A[256 pixels, 32 ci] * B[256 co, 32 ci]^T per (tap, ci-chunk)
//       output tile = 128 CONSECUTIVE pixels (row-major) of one image of one level per CTA, so only the last tile of
//                 an image is partial (tile efficiency 98 % at 800x1344 instead of 87 % with 16x8 boxes);
//       A tile  = ONE im2col-mode TMA load per CTA: 128 pixels x 32 channels of the input shifted by the filter tap,
//                 wrapping over image rows exactly like the output tile; the halo and the zero padding come from the
//                 tensor map's bounding box / out-of-bounds zero fill, no im2col buffer exists;
//       B tile  = each CTA loads HALF (128 co) of the packed weights [tap][co][ci] box;
//       both land K-major with 128-byte swizzle, i.e. exactly the canonical UMMA SW128 layout.
//   wgrad           : D[256 co, 256 ci] += A[32 pixels, 256 co]^T * B[32 pixels, 256 ci]   (MN-major operands,
//                 SWIZZLE_128B_BASE32B, 5-D tiled TMA maps), co / ci halves split over the CTA pair.
// Why pairs: with fp32 operands a single-SM kernel is bound by shared-memory / L2->SM bytes per MMA (48 KiB per
// 128x256x32 block, measured 63-68 % tensor-pipe at full clock); splitting B over two SMs makes it 32 KiB.
// Roles per CTA: warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only) + TMEM owner, warps 2.. = epilogue
// (TMEM -> regs -> bias / ReLU / mask / TF32 rounding / GroupNorm partials / per-channel sums -> global).
// Barriers: the leader CTA (cluster rank 0) issues all MMAs; both CTAs' TMA loads complete on the leader's "full"
// barrier; tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs; the epilogue warps of both CTAs
// hand a TMEM accumulator stage back on the leader's "tempty" barrier.
#include <cuda.h>
#include <cuda_fp16.h>
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace lgd {

constexpr int TILE_M = TILE_PIX;  // 128 output slots per CTA tile
constexpr int BLOCK_K = 32;       // fp32 elements per K block = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;         // K per tcgen05.mma for 32-bit operands
constexpr int STAGES = 5;         // wgrad ring
constexpr int A_BYTES = TILE_M * BLOCK_K * 4;    // 16 KiB: wgrad co half
constexpr int B_BYTES = (C / 2) * BLOCK_K * 4;   // 16 KiB: this CTA's half of the weight tile (fwd) / ci half (wgrad)
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_KB = 9 * (C / BLOCK_K);  // 72 K blocks per output tile
constexpr int TMEM_COLS = 512;             // two 128x256 fp32 accumulators per CTA
constexpr int NUM_THREADS = 192;      // wgrad: producer warp, MMA warp, 4 epilogue warps
constexpr int FWD_EPI_WARPS = 8;      // forward / dgrad: two warps per TMEM lane quarter, 128 of the 256 columns each
constexpr int FWD_THREADS = 64 + 32 * FWD_EPI_WARPS;
constexpr int SMEM_EXTRA = 8192;  // barriers, tmem pointer, bias stage, reduction scratch, per-warp channel sums
// epilogue store staging: every epilogue warp transposes its {32 rows x 128 bytes} pieces through shared memory so that
// one store instruction writes four whole 128-byte segments instead of 32 scattered 16-byte ones (row stride 144
// bytes: conflict-free for the per-row writes and the per-segment reads)
constexpr int EPI_ROW_BYTES = 144;
constexpr int EPI_WARP_BYTES = 32 * EPI_ROW_BYTES;
constexpr int EPI_STAGE_BYTES = FWD_EPI_WARPS * EPI_WARP_BYTES;   // 36 KiB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + SMEM_EXTRA + EPI_STAGE_BYTES + 1024 /* alignment slack */;
// forward / dgrad: the input STRIP of a tile -- every input pixel any of the nine taps needs, for one 128-byte channel
// block -- is loaded ONCE and the nine taps read it through UMMA descriptors whose start address is shifted by whole
// 128-byte rows (conv3x3_tc_kernel below). Output slots are positions of the zero-padded image rows (W + 2 per row), so a
// tap is a constant row shift: (dy + 1) * pitch + (dx + 1).
//   narrow levels (W + 2 <= 130): one contiguous run of 130 + 2 (W + 2) padded positions, pitch = W + 2
//   wide levels: three runs of 130 positions (rows y-1, y, y+1 of the same columns) at a pitch of SEG_ROWS rows
constexpr int STRIP_RUN = TILE_M + 2;              // 130 positions: slots s0-1 .. s0+128
constexpr int SEG_ROWS = 136;                      // pitch of the three runs of a wide level (multiple of 8 rows)
constexpr int STRIP_ROWS = 3 * SEG_ROWS;           // 408 rows >= 130 + 2 * 130 = 390
constexpr int A_STRIP_BYTES = STRIP_ROWS * 128;    // 51 KiB per stage
constexpr int A_STAGES = 2;                        // one strip feeds 9 taps x 4 MMAs: two stages cover the next load
constexpr int B_STAGES = 5;                        // weight half tiles (16 KiB), one per (channel block, tap); 5 fill
                                                   // the 227 KiB exactly (4 -> 5: forward 0.280 -> 0.2755 ms)
constexpr int FWD_RING_BYTES = A_STAGES * A_STRIP_BYTES + B_STAGES * B_BYTES;
constexpr int FWD_SMEM_BYTES = FWD_RING_BYTES + SMEM_EXTRA + EPI_STAGE_BYTES + 1024 /* alignment slack */;
static_assert(A_STRIP_BYTES % 1024 == 0, "strip stages keep the 1024-byte swizzle alignment");
static_assert(FWD_SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct ConvTmaps {
  CUtensorMap act[LGD_MAX_LEVELS];
  CUtensorMap act2[LGD_MAX_LEVELS];  // wgrad: second activation operand
  CUtensorMap w;
};

struct ConvArgs {
  Pyr pyr;
  int tiles_img[LGD_MAX_LEVELS];  // tiles per image of the level
  int tile_start[LGD_MAX_LEVELS + 1];   // every level holds an EVEN number of tiles (a CTA pair never straddles levels)
  int total_tiles;
  const float* bias;
  int bias_lstride, bias_istride;
  float* out;
  const float* relu_mask;
  const __half* relu_mask_h;  // the same mask from the fp16 copy of the activation (nonzero = pass); one of the two
  const float* addend;  // ADD kernels: fp32 tensor (layout of out; may alias out) added to the accumulator
  float* tile_stats;
  float* tile_csum;  // [tile][256] per-channel sums of the stored (un-rounded) values, or nullptr
  __half* out_half;  // optional fp16 copy of the stored values (operand of the next convolution)
  const float* acc_scale;   // optional device scalar: accumulator *= acc_scale[0] (undo the power-of-two scale of an
                            // fp16 gradient operand)
  const float* half_scale;  // optional device scalar: out_half = fp16(stored value * half_scale[0]), saturated
  int relu, round_out;
  // fp32 output as a slice of a wider pixel-major matrix (detection heads: 720 / 36 output channels written by
  // 256-column launches): row stride in elements, first column, number of valid columns of this launch
  int out_ld, out_col0, out_cols;
  // MASK == 3 (dgrad whose output feeds a GroupNorm backward): the epilogue also reads the GroupNorm's input x at the
  // tile's pixels and emits per-tile sums of (g, g*xhat, g^2), g = out [masked by xhat > 0 when gn_relu] -- the first
  // pass of the GroupNorm backward (2 F1 of reads) disappears
  const float* gn_x;
  const float* gn_stats;   // (F,B,2) = {mean, rstd}
  float* tile_gn;          // [tile][4]
  int gn_relu;
};

// tile t -> level l, image b, first slot f0 of the tile inside that image (slot s = y * (W + 2) + x + 1: position of the
// zero-padded row). dummy: the padding tile that makes a level's tile count even; it repeats the last real tile's loads
// and stores nothing.
__device__ __forceinline__ void decode_tile(const ConvArgs& a, int t, int& l, int& b, int& f0, bool& dummy) {
  l = 0;
  while (l + 1 < a.pyr.num_levels && t >= a.tile_start[l + 1]) ++l;
  int r = t - a.tile_start[l];
  const int nreal = a.pyr.batch * a.tiles_img[l];
  dummy = r >= nreal;
  if (dummy) r = nreal - 1;
  b = r / a.tiles_img[l];
  f0 = (r - b * a.tiles_img[l]) * TILE_M;
}

struct SmemLayout {
  uint8_t* base;
  __device__ __forceinline__ uint8_t* a(int i) const { return base + i * STAGE_BYTES; }
  __device__ __forceinline__ uint8_t* b(int i) const { return base + i * STAGE_BYTES + A_BYTES; }
  uint64_t* full;
  uint64_t* empty;
  uint64_t* tfull;
  uint64_t* tempty;
  uint32_t* tmem_ptr;
  float* bias;
  float* red;
  float* csum;  // [4 epilogue warps][256]
  uint8_t* epi;  // [FWD_EPI_WARPS][32 rows][144 bytes] store staging
};

__device__ __forceinline__ SmemLayout carve(uint8_t* raw) {
  SmemLayout s;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  s.base = base;
  uint8_t* x = base + STAGES * STAGE_BYTES;
  s.full = reinterpret_cast<uint64_t*>(x);
  s.empty = s.full + STAGES;
  s.tfull = s.empty + STAGES;
  s.tempty = s.tfull + 2;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(s.tempty + 2);
  s.bias = reinterpret_cast<float*>(x + 256);
  s.red = reinterpret_cast<float*>(x + 256 + C * 4);
  s.csum = reinterpret_cast<float*>(x + 2048);
  s.epi = x + SMEM_EXTRA;
  return s;
}

// forward / dgrad: A strip ring, B (weight half tile) ring, then the same extras as above
struct FwdSmem {
  uint8_t* base;
  __device__ __forceinline__ uint8_t* a(int i) const { return base + i * A_STRIP_BYTES; }
  __device__ __forceinline__ uint8_t* b(int i) const { return base + A_STAGES * A_STRIP_BYTES + i * B_BYTES; }
  uint64_t* a_full;
  uint64_t* a_empty;
  uint64_t* b_full;
  uint64_t* b_empty;
  uint64_t* tfull;
  uint64_t* tempty;
  uint32_t* tmem_ptr;
  float* bias;
  float* red;
  float* csum;  // [4 epilogue warps][256]
  uint8_t* epi;  // [FWD_EPI_WARPS][32 rows][144 bytes] store staging
};

__device__ __forceinline__ FwdSmem carve_fwd(uint8_t* raw) {
  FwdSmem s;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  s.base = base;
  uint8_t* x = base + FWD_RING_BYTES;
  s.a_full = reinterpret_cast<uint64_t*>(x);
  s.a_empty = s.a_full + A_STAGES;
  s.b_full = s.a_empty + A_STAGES;
  s.b_empty = s.b_full + B_STAGES;
  s.tfull = s.b_empty + B_STAGES;
  s.tempty = s.tfull + 2;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(s.tempty + 2);
  s.bias = reinterpret_cast<float*>(x + 256);
  s.red = reinterpret_cast<float*>(x + 256 + C * 4);
  s.csum = reinterpret_cast<float*>(x + 2048);
  s.epi = x + SMEM_EXTRA;
  return s;
}

// A warp-convergent role (TMA producer, MMA issuer) waits with all of its lanes: measured on one box (forward / dgrad /
// wgrad per launch) 0.273 / 0.309 / 0.280 ms, against 0.300 / 0.336 / 0.306 ms when lane 0 alone polls and the warp
// reconverges behind it.
__device__ __forceinline__ void warp_wait(uint64_t* bar, uint32_t parity, int /*lane*/) { mbar_wait(bar, parity); }

// strip geometry of a level: rows between the three dy positions of a tap, and the pixels one TMA operation delivers
__device__ __host__ __forceinline__ bool strip_contiguous(int w) { return w + 2 <= STRIP_RUN; }
__device__ __host__ __forceinline__ int strip_pitch(int w) { return strip_contiguous(w) ? w + 2 : SEG_ROWS; }
__device__ __host__ __forceinline__ int strip_box_pixels(int w) { return strip_contiguous(w) ? STRIP_RUN + 2 * (w + 2) : STRIP_RUN; }
__device__ __host__ __forceinline__ int strip_bytes(int w) { return (strip_contiguous(w) ? STRIP_RUN + 2 * (w + 2) : 3 * STRIP_RUN) * 128; }

// Sum over the 32 lanes of a warp of 32 per-lane values each (a 32x32 transpose-reduce in 31 shuffles): the return
// value of lane L is sum over lanes of their v[L].
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < off) {
        const float send = hi ? v[i] : v[i + off];
        const float keep = hi ? v[i + off] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
  }
  return v[0];
}

// F16 = false: TF32 operands (fp32 bits in memory), 32 channels per 128-byte k-block  -> forward, dgrad
// F16 = true : fp16 operands,                       64 channels per 128-byte k-block  -> forward only (half the k-blocks,
//              twice the MACs per MMA; same 10-bit mantissa as TF32; activations are O(1) after the norms so the 5-bit
//              exponent is not a constraint in the forward direction)
// ADD = true  : the epilogue adds a previously computed partial result (a.addend) to the accumulator before bias /
//              ReLU / statistics: the split-operand fp32-accurate forward runs three TF32 launches (lo*hi, hi*lo, hi*hi)
//              that chain through it
// MASK = true : the epilogue can apply a ReLU mask (prefetched one chunk ahead) and emit per-tile channel sums -- the
//              dgrad-through-ReLU launches. The plain variants are compiled without those registers (the 320-thread
//              CTA then leaves enough of the register file for a block of an HBM-bound kernel of the other chain to
//              run next to it on the same SM).
// WIDE = true: the fp32 output is a column slice of a wider pixel-major matrix (a.out_ld / out_col0 / out_cols; detection
//              heads); false: the plain 256-channel pyramid (constants, no extra registers)
template <bool F16, bool ADD = false, int MASK = 0, bool WIDE = false>   // MASK: 0 none, 1 fp32 ReLU mask, 2 fp16 ReLU mask (+ channel sums), 3 / 4 GroupNorm-backward sums from x (fp32) / from y = relu(xhat) (fp16)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FWD_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ ConvTmaps tm, const __grid_constant__ ConvArgs a) {
  constexpr int KE = F16 ? 64 : 32;        // channels per k-block
  constexpr int KBLKS = C / KE;            // k-blocks (= input strips) per tile
  extern __shared__ uint8_t smem_raw[];
  FwdSmem s = carve_fwd(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs_grid = gridDim.x >> 1;
  const int npairs = a.total_tiles >> 1;   // every level holds an even number of tiles

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < a.pyr.num_levels; ++l) tma_prefetch_desc(&tm.act[l]);
    tma_prefetch_desc(&tm.w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < A_STAGES; ++i) {
        mbar_init(&s.a_full[i], 1);   // leader's: one arrive.expect_tx per phase, bytes of BOTH CTAs
        mbar_init(&s.a_empty[i], 1);  // multicast tcgen05.commit
      }
      for (int i = 0; i < B_STAGES; ++i) {
        mbar_init(&s.b_full[i], 1);
        mbar_init(&s.b_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s.tfull[i], 1);   // multicast tcgen05.commit
        mbar_init(&s.tempty[i], 2 * FWD_EPI_WARPS);  // leader's: epilogue warps of both CTAs
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2sm(s.tmem_ptr, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (each CTA: its strips, its B half)
    // One warp feeds two rings: per tile KBLKS strips (A ring) and KBLKS x 9 weight half tiles (B ring, channel block
    // major, tap minor -- the order the MMA warp consumes them): a strip that waits for its stage never holds back the
    // weight tiles in front of it. Like the MMA warp, the
    // whole warp runs the loop convergently and one elected lane issues, so that the TMA operands (coordinates, barrier
    // and shared-memory addresses) live on the uniform datapath instead of going through a broadcast loop per operation.
    {
      int a_stage = 0, b_stage = 0;
      uint32_t a_phase = 0, b_phase = 0;
      int ta = pair, tb = pair;     // tile pair the next A / B item belongs to
      int a_kc = 0, b_item = 0;     // next strip of ta / next (kc, tap) item of tb
      // geometry of the A cursor's tile
      int l = 0, b = 0, f0 = 0, W = 1;
      bool dummy = false;
      if (ta < npairs) {
        decode_tile(a, 2 * ta + (int)rank, l, b, f0, dummy);
        W = a.pyr.w[l];
      }
      const uint32_t a_ring = smem_u32(s.base), b_ring = a_ring + A_STAGES * A_STRIP_BYTES;
      while (tb < npairs) {
        // a strip whenever its stage is free (lane 0's view of the barrier decides for the warp) ...
        const bool a_free = ta < npairs && __shfl_sync(0xffffffffu, (int)mbar_test(&s.a_empty[a_stage], a_phase ^ 1), 0) != 0;
        if (a_free) {
          if (elect_one()) {
            const uint32_t full_leader = mapa_shared(smem_u32(&s.a_full[a_stage]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&s.a_full[a_stage], 2 * strip_bytes(W));
            // first position of the strip = slot f0 - 1 one padded row up, in box coordinates (x in [-1, W], y in [-2, H])
            const int PW = W + 2;
            const int q0 = f0 - 1 + PW;
            const int qy = q0 / PW, qx = q0 - qy * PW;
            const CUtensorMap* am = &tm.act[l];
            const uint32_t dst = a_ring + a_stage * A_STRIP_BYTES;
            if (strip_contiguous(W)) {
              tma_load_im2col_4d_2sm(dst, am, full_leader, a_kc * KE, qx - 1, qy - 2, b, 0, 0);
            } else {
#pragma unroll
              for (int d = 0; d < 3; ++d)
                tma_load_im2col_4d_2sm(dst + d * SEG_ROWS * 128, am, full_leader, a_kc * KE, qx - 1, qy - 2 + d, b, 0, 0);
            }
          }
          __syncwarp();
          if (++a_stage == A_STAGES) {
            a_stage = 0;
            a_phase ^= 1;
          }
          if (++a_kc == KBLKS) {
            a_kc = 0;
            ta += npairs_grid;
            if (ta < npairs) {
              decode_tile(a, 2 * ta + (int)rank, l, b, f0, dummy);
              W = a.pyr.w[l];
            }
          }
        }
        // ... then the next weight tile, sleeping on its stage (one per 512 clk of MMA work: the strip test above runs
        // often enough, and strip kc always becomes issuable before weight tile (kc, 0) does, so waiting here cannot
        // keep a strip the MMA warp needs from being issued)
        warp_wait(&s.b_empty[b_stage], b_phase ^ 1, lane);
        if (elect_one()) {
          const uint32_t full_leader = mapa_shared(smem_u32(&s.b_full[b_stage]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&s.b_full[b_stage], 2 * B_BYTES);
          const int kc = b_item / 9, tap = b_item - kc * 9;
          tma_load_2d_2sm(b_ring + b_stage * B_BYTES, &tm.w, full_leader, kc * KE, tap * C + (int)rank * (C / 2));
        }
        __syncwarp();
        if (++b_stage == B_STAGES) {
          b_stage = 0;
          b_phase ^= 1;
        }
        if (++b_item == 9 * KBLKS) {
          b_item = 0;
          tb += npairs_grid;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // The WHOLE warp runs this loop convergently and one elected lane issues: every operand of tcgen05.mma / commit is
    // then computed on the uniform datapath and the four MMAs of a weight tile issue back to back. Issued from a
    // lane-divergent branch (`if (lane == 0)`), ptxas wraps each MMA in an ELECT / R2UR.BROADCAST loop of ~100 clk --
    // as long as the MMA itself (128 clk), which made the issuing thread the bottleneck of the kernel
    // (tools/probe_mma_rate.cu).
    if (rank == 0) {
      constexpr uint32_t idesc = F16 ? make_idesc_f16(2 * TILE_M, C) : make_idesc_tf32(2 * TILE_M, C, 0, 0);
      int a_stage = 0, b_stage = 0;
      uint32_t a_phase = 0, b_phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t a_ring = smem_u32(s.base), b_ring = a_ring + A_STAGES * A_STRIP_BYTES;
      for (int tp = pair; tp < npairs; tp += npairs_grid) {
        int l, b, f0;
        bool dummy;
        decode_tile(a, 2 * tp, l, b, f0, dummy);   // both tiles of the pair lie in the same level
        const int pitch = strip_pitch(a.pyr.w[l]);
        warp_wait(&s.tempty[acc], acc_phase ^ 1, lane);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * C;
        for (int kc = 0; kc < KBLKS; ++kc) {
          warp_wait(&s.a_full[a_stage], a_phase, lane);
          const uint32_t a_base = a_ring + a_stage * A_STRIP_BYTES;
          for (int tap = 0; tap < 9; ++tap) {
            warp_wait(&s.b_full[b_stage], b_phase, lane);
            tc_fence_after();
            // tap (dy, dx) = the strip shifted by (dy + 1) * pitch + (dx + 1) rows: SWIZZLE_128B is a function of the
            // shared-memory address, so a descriptor may start at any 128-byte row (tools/probe_strip.cu)
            const int dy = tap / 3, dx = tap - 3 * dy;
            const uint64_t ad = make_smem_desc_sw128(a_base + (uint32_t)(dy * pitch + dx) * 128u, 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(b_ring + b_stage * B_BYTES, 16, 1024);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // advance 8 fp32 / 16 fp16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr>>4) field
                if (F16)
                  mma_f16_ss_2sm(d, ad + 2 * k, bd + 2 * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
                else
                  mma_tf32_ss_2sm(d, ad + 2 * k, bd + 2 * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
              }
              mma_commit_2sm(&s.b_empty[b_stage], 3);
              if (tap == 8) mma_commit_2sm(&s.a_empty[a_stage], 3);
              if (tap == 8 && kc == KBLKS - 1) mma_commit_2sm(&s.tfull[acc], 3);
            }
            __syncwarp();
            if (++b_stage == B_STAGES) {
              b_stage = 0;
              b_phase ^= 1;
            }
          }
          if (++a_stage == A_STAGES) {
            a_stage = 0;
            a_phase ^= 1;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (256 threads per CTA, own tile)
    // warp e = warp - 2: TMEM lane quarter (warp & 3) = 32 pixel rows, column half (e >> 2) = 4 of the 8 channel chunks
    constexpr int EPI_THREADS = 32 * FWD_EPI_WARPS;
    const int epi_tid = threadIdx.x - 64;
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int chunk_begin = (ew >> 2) * (C / 64), chunk_end = chunk_begin + C / 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tp = pair; tp < npairs; tp += npairs_grid) {
      const int t = 2 * tp + (int)rank;
      int l = 0, b = 0, f0 = 0;
      bool dummy = false;
      decode_tile(a, t, l, b, f0, dummy);
      named_bar_sync(1, EPI_THREADS);  // everyone is done with the previous tile's bias / scratch
      if (a.bias != nullptr && !dummy) {
        const float* bp = a.bias + (long long)l * a.bias_lstride + (long long)b * a.bias_istride;
        s.bias[epi_tid] = __ldg(bp + epi_tid);
      } else {
        s.bias[epi_tid] = 0.f;
      }
      named_bar_sync(1, EPI_THREADS);
      mbar_wait(&s.tfull[acc], acc_phase);
      tc_fence_after();
      // slot of this thread's row -> pixel of the image (the two pad positions of a padded row hold nothing)
      const int Wl = a.pyr.w[l], HW = a.pyr.h[l] * Wl;
      const int slot = f0 + row;
      const int sy = slot / (Wl + 2), sx = slot - sy * (Wl + 2);
      const bool valid = !dummy && sy < a.pyr.h[l] && sx >= 1 && sx <= Wl;
      const int pix = valid ? sy * Wl + sx - 1 : -1;
      const long long img_off = a.pyr.off[l] + (long long)b * HW * C;   // first pixel of the image
      const long long pix_off = img_off + (long long)(valid ? pix : 0) * C;
      // outputs go through the warp's staging buffer: lane = row while computing, then rows x 128-byte segments; the
      // pixel of each row this lane stores comes from the lane that owns the row
      // the fp32 copy is optional when an fp16 copy is written; it may be a column slice of a wider matrix
      const int out_ld = WIDE ? a.out_ld : C, out_cols = WIDE ? a.out_cols : C;
      float* obase = nullptr;   // first pixel of the image
      if (a.out != nullptr)
        obase = WIDE ? a.out + (a.pyr.off[l] / C + (long long)b * HW) * out_ld + a.out_col0 : a.out + img_off;
      __half* hbase = a.out_half ? a.out_half + img_off : nullptr;
      const int st_row = lane >> 3, st_seg = lane & 7;                       // store phase: 4 rows x 8 segments
      int pix4[8], pix8[4];   // pixel (or -1) of row 4 i + st_row / of row 8 i + (lane >> 2)
#pragma unroll
      for (int i = 0; i < 8; ++i) pix4[i] = __shfl_sync(0xffffffffu, pix, 4 * i + st_row);
#pragma unroll
      for (int i = 0; i < 4; ++i) pix8[i] = __shfl_sync(0xffffffffu, pix, 8 * i + (lane >> 2));
      uint8_t* stg = s.epi + ew * EPI_WARP_BYTES;
      const float* mptr = (MASK == 3) ? a.gn_x + pix_off : (a.relu_mask ? a.relu_mask + pix_off : nullptr);
      float gs1 = 0.f, gs2 = 0.f, gs3 = 0.f, gn_mean = 0.f, gn_rstd = 1.f;
      if (MASK == 3 && !dummy) {
        gn_mean = __ldg(a.gn_stats + 2 * (l * a.pyr.batch + b));
        gn_rstd = __ldg(a.gn_stats + 2 * (l * a.pyr.batch + b) + 1);
      }
      const __half* hmptr = ((MASK == 2 || MASK == 4) && a.relu_mask_h) ? a.relu_mask_h + pix_off : nullptr;
      const float* aptr = ADD ? a.addend + pix_off : nullptr;
      float sum = 0.f, sumsq = 0.f;
      const float asc = a.acc_scale ? __ldg(a.acc_scale) : 1.f;
      const float hsc = a.half_scale ? __ldg(a.half_scale) : 1.f;
      if (!dummy) {
        // ReLU mask of the layer below (dgrad): software-pipelined one 32-channel chunk ahead so that its DRAM latency
        // is not exposed once per chunk
        float4 mcur[(MASK == 1 || MASK == 3) ? 8 : 1];
        uint4 hcur[(MASK == 2 || MASK == 4) ? 4 : 1];   // fp16 mask: 32 halves of the chunk
        const bool use_mask = (MASK == 1 || MASK == 3) && mptr != nullptr && valid;
        const bool use_hmask = (MASK == 2 || MASK == 4) && hmptr != nullptr && valid;
        if ((MASK == 1 || MASK == 3) && use_mask) {
#pragma unroll
          for (int j = 0; j < 8; ++j) mcur[j] = ldg4(mptr + chunk_begin * 32 + j * 4);
        }
        if ((MASK == 2 || MASK == 4) && use_hmask) {
#pragma unroll
          for (int j = 0; j < 4; ++j) hcur[j] = __ldg(reinterpret_cast<const uint4*>(hmptr + chunk_begin * 32) + j);
        }
#pragma unroll 1
        for (int chunk = chunk_begin; chunk < chunk_end; ++chunk) {
          float4 mnext[(MASK == 1 || MASK == 3) ? 8 : 1];
          uint4 hnext[(MASK == 2 || MASK == 4) ? 4 : 1];
          if ((MASK == 1 || MASK == 3) && use_mask && chunk + 1 < chunk_end) {
#pragma unroll
            for (int j = 0; j < 8; ++j) mnext[j] = ldg4(mptr + (chunk + 1) * 32 + j * 4);
          }
          if ((MASK == 2 || MASK == 4) && use_hmask && chunk + 1 < chunk_end) {
#pragma unroll
            for (int j = 0; j < 4; ++j) hnext[j] = __ldg(reinterpret_cast<const uint4*>(hmptr + (chunk + 1) * 32) + j);
          }
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * C + chunk * 32), r);
          tmem_ld_wait();
          // every lane computes (the store staging needs the whole warp); rows outside the image are never stored
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v;
            v.x = fmaf(__uint_as_float(r[j + 0]), asc, s.bias[chunk * 32 + j + 0]);
            v.y = fmaf(__uint_as_float(r[j + 1]), asc, s.bias[chunk * 32 + j + 1]);
            v.z = fmaf(__uint_as_float(r[j + 2]), asc, s.bias[chunk * 32 + j + 2]);
            v.w = fmaf(__uint_as_float(r[j + 3]), asc, s.bias[chunk * 32 + j + 3]);
            if (ADD && valid) {  // plain (coherent) load: addend may be the tensor overwritten below
              const float4 ad = *reinterpret_cast<const float4*>(aptr + chunk * 32 + j);
              v.x += ad.x; v.y += ad.y; v.z += ad.z; v.w += ad.w;
            }
            if (valid) {
              sum += (v.x + v.y) + (v.z + v.w);
              sumsq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
            }
            if (a.relu) {
              v.x = relu_keep_nan(v.x); v.y = relu_keep_nan(v.y); v.z = relu_keep_nan(v.z); v.w = relu_keep_nan(v.w);
            }
            if (MASK == 1 && use_mask) {
              const float4 m = mcur[MASK == 1 ? (j >> 2) : 0];
              v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f;
              v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
            }
            if (MASK == 3 && use_mask) {   // sums of the GroupNorm backward; the stored gradient stays unmasked
              const float4 m = mcur[MASK == 3 ? (j >> 2) : 0];
              const float h0 = (m.x - gn_mean) * gn_rstd, h1 = (m.y - gn_mean) * gn_rstd;
              const float h2 = (m.z - gn_mean) * gn_rstd, h3 = (m.w - gn_mean) * gn_rstd;
              const float g0 = (a.gn_relu && !(h0 > 0.f)) ? 0.f : v.x, g1 = (a.gn_relu && !(h1 > 0.f)) ? 0.f : v.y;
              const float g2 = (a.gn_relu && !(h2 > 0.f)) ? 0.f : v.z, g3 = (a.gn_relu && !(h3 > 0.f)) ? 0.f : v.w;
              gs1 += (g0 + g1) + (g2 + g3);
              gs2 += (g0 * h0 + g1 * h1) + (g2 * h2 + g3 * h3);
              gs3 += (g0 * g0 + g1 * g1) + (g2 * g2 + g3 * g3);
            }
            if (MASK == 4 && use_hmask) {   // GroupNorm-ReLU sums from y = relu(xhat) (fp16 copy): y != 0 <=> xhat > 0, g * xhat = g * y
              const uint4 q = hcur[MASK == 4 ? (j >> 3) : 0];
              const uint32_t w0 = (j & 4) ? q.z : q.x, w1 = (j & 4) ? q.w : q.y;
              const float2 y01 = __half22float2(*reinterpret_cast<const __half2*>(&w0));
              const float2 y23 = __half22float2(*reinterpret_cast<const __half2*>(&w1));
              const float g0 = (w0 & 0xffffu) ? v.x : 0.f, g1 = (w0 >> 16) ? v.y : 0.f;
              const float g2 = (w1 & 0xffffu) ? v.z : 0.f, g3 = (w1 >> 16) ? v.w : 0.f;
              gs1 += (g0 + g1) + (g2 + g3);
              gs2 += (g0 * y01.x + g1 * y01.y) + (g2 * y23.x + g3 * y23.y);
              gs3 += (g0 * g0 + g1 * g1) + (g2 * g2 + g3 * g3);
            }
            if (MASK == 2 && use_hmask) {   // halves j..j+3 = two 32-bit words of the chunk's 64 bytes
              const uint4 q = hcur[MASK == 2 ? (j >> 3) : 0];
              const uint32_t w0 = (j & 4) ? q.z : q.x, w1 = (j & 4) ? q.w : q.y;
              v.x = (w0 & 0xffffu) ? v.x : 0.f; v.y = (w0 >> 16) ? v.y : 0.f;
              v.z = (w1 & 0xffffu) ? v.z : 0.f; v.w = (w1 >> 16) ? v.w : 0.f;
            }
            r[j + 0] = __float_as_uint(v.x); r[j + 1] = __float_as_uint(v.y);
            r[j + 2] = __float_as_uint(v.z); r[j + 3] = __float_as_uint(v.w);
            if (obase != nullptr) {
              if (a.round_out) {
                v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
              }
              *reinterpret_cast<float4*>(stg + lane * EPI_ROW_BYTES + j * 4) = v;   // own row, 16 bytes at a time
            }
          }
          if (obase != nullptr) {   // 32 rows x 128 bytes staged: store 4 rows x (8 x 16 bytes) per instruction
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + st_row;
              const float4 v = *reinterpret_cast<const float4*>(stg + rr * EPI_ROW_BYTES + st_seg * 16);
              if (pix4[i] >= 0 && (!WIDE || chunk * 32 + st_seg * 4 < out_cols))
                stg4(obase + (long long)pix4[i] * out_ld + chunk * 32 + st_seg * 4, v);
            }
            __syncwarp();
          }
          if (hbase != nullptr) {  // r[] holds the un-rounded stored values: fp16 copy
            const bool scaled = a.half_scale != nullptr;  // gradient operand: power-of-two scale, saturating
            auto h2 = [&](int k) {
              float x = __uint_as_float(r[k]), y = __uint_as_float(r[k + 1]);
              if (scaled) {
                x = fminf(fmaxf(x * hsc, -65504.f), 65504.f);
                y = fminf(fmaxf(y * hsc, -65504.f), 65504.f);
              }
              const __half2 t = __floats2half2_rn(x, y);
              uint32_t w = *reinterpret_cast<const uint32_t*>(&t);
              if (a.relu) {  // the copy doubles as the ReLU mask of the backward: a positive value never becomes 0
                if (x > 0.f && (w & 0xffffu) == 0) w |= 1u;
                if (y > 0.f && (w >> 16) == 0) w |= 0x10000u;
              }
              return w;
            };
            // two chunks (64 channels = 128 bytes per row) share one store pass when the staging buffer is not also
            // carrying the fp32 copy; otherwise each chunk (64 bytes per row) is flushed on its own
            const bool pairwise = obase == nullptr;
            const int hoff = pairwise ? (chunk & 1) * 64 : 0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 h;
              h.x = h2(j + 0); h.y = h2(j + 2); h.z = h2(j + 4); h.w = h2(j + 6);
              *reinterpret_cast<uint4*>(stg + lane * EPI_ROW_BYTES + hoff + j * 2) = h;
            }
            if (pairwise) {
              if (chunk & 1) {
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int rr = 4 * i + st_row;
                  const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * EPI_ROW_BYTES + st_seg * 16);
                  if (pix4[i] >= 0)
                    *reinterpret_cast<uint4*>(hbase + (long long)pix4[i] * C + (chunk - 1) * 32 + st_seg * 8) = v;
                }
                __syncwarp();
              }
            } else {
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rr = 8 * i + (lane >> 2);
                const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * EPI_ROW_BYTES + (lane & 3) * 16);
                if (pix8[i] >= 0)
                  *reinterpret_cast<uint4*>(hbase + (long long)pix8[i] * C + chunk * 32 + (lane & 3) * 8) = v;
              }
              __syncwarp();
            }
          }
          if ((MASK == 1 || MASK == 3) && use_mask) {
#pragma unroll
            for (int j = 0; j < ((MASK == 1 || MASK == 3) ? 8 : 1); ++j) mcur[j] = mnext[j];
          }
          if ((MASK == 2 || MASK == 4) && use_hmask) {
#pragma unroll
            for (int j = 0; j < ((MASK == 2 || MASK == 4) ? 4 : 1); ++j) hcur[j] = hnext[j];
          }
          if (MASK && MASK < 3 && a.tile_csum != nullptr) {  // warp-uniform: per-channel sums over this warp's 32 rows (un-rounded)
            float cv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) cv[j] = valid ? __uint_as_float(r[j]) : 0.f;
            const float cs = warp_transpose_sum(cv, lane);
            s.csum[(ew & 3) * C + chunk * 32 + lane] = cs;
          }
        }
      }
      // accumulator drained -> hand the TMEM stage back to the leader's MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(&s.tempty[acc]), 0));
      if (dummy && epi_tid == 0) {   // the padding tile of a level: its by-products are read by the reductions -> zeros
        if (a.tile_stats != nullptr) a.tile_stats[2 * t + 0] = a.tile_stats[2 * t + 1] = 0.f;
        if (MASK >= 3) a.tile_gn[4 * (long long)t + 0] = a.tile_gn[4 * (long long)t + 1] = a.tile_gn[4 * (long long)t + 2] = 0.f;
      }
      if (MASK && MASK < 3 && a.tile_csum != nullptr && dummy) a.tile_csum[(long long)t * C + epi_tid] = 0.f;
      if (a.tile_stats != nullptr && !dummy) {
        sum = warp_sum(sum);
        sumsq = warp_sum(sumsq);
        if (lane == 0) {
          s.red[ew * 2 + 0] = sum;
          s.red[ew * 2 + 1] = sumsq;
        }
        named_bar_sync(1, EPI_THREADS);
        if (epi_tid == 0) {
          a.tile_stats[2 * t + 0] = ((s.red[0] + s.red[2]) + (s.red[4] + s.red[6])) + ((s.red[8] + s.red[10]) + (s.red[12] + s.red[14]));
          a.tile_stats[2 * t + 1] = ((s.red[1] + s.red[3]) + (s.red[5] + s.red[7])) + ((s.red[9] + s.red[11]) + (s.red[13] + s.red[15]));
        }
      }
      if (MASK >= 3 && !dummy) {   // per-tile GroupNorm-backward sums: warps in fixed order
        gs1 = warp_sum(gs1);
        gs2 = warp_sum(gs2);
        gs3 = warp_sum(gs3);
        named_bar_sync(1, EPI_THREADS);   // s.red is free again (tile statistics above are done with it)
        if (lane == 0) {
          s.red[32 + ew * 3 + 0] = gs1;
          s.red[32 + ew * 3 + 1] = gs2;
          s.red[32 + ew * 3 + 2] = gs3;
        }
        named_bar_sync(1, EPI_THREADS);
        if (epi_tid < 3) {
          float tot = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < FWD_EPI_WARPS; ++w8) tot += s.red[32 + w8 * 3 + epi_tid];
          a.tile_gn[4 * (long long)t + epi_tid] = tot;
        }
      }
      if (MASK && MASK < 3 && a.tile_csum != nullptr && !dummy) {
        named_bar_sync(1, EPI_THREADS);
        const int c = epi_tid;  // one channel per epilogue thread
        a.tile_csum[(long long)t * C + c] = (s.csum[c] + s.csum[C + c]) + (s.csum[2 * C + c] + s.csum[3 * C + c]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may leave (or free TMEM) while its peer can still signal / read it
  if (warp == 1) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
}

// out[(l*B+b)][c] = sum over the tiles of image b of level l of tile_csum[tile][c]   (fixed order)
__global__ void tile_csum_finalize_kernel(Pyr p, const float* __restrict__ tile_csum, float* __restrict__ out) {
  const int seg = blockIdx.x, c = threadIdx.x;
  const int l = seg / p.batch, b = seg - l * p.batch;
  int tile_start = 0;
  for (int j = 0; j < l; ++j) tile_start += tiles_per_level(p.h[j], p.w[j], p.batch);
  const int per_img = tiles_per_image(p.h[l], p.w[l]);
  const float* ts = tile_csum + (long long)(tile_start + b * per_img) * C + c;
  double s0 = 0.0, s1 = 0.0;
  int i = 0;
  for (; i + 2 <= per_img; i += 2) {
    s0 += (double)ts[(long long)i * C];
    s1 += (double)ts[(long long)(i + 1) * C];
  }
  if (i < per_img) s0 += (double)ts[(long long)i * C];
  out[(long long)seg * C + c] = (float)(s0 + s1);
}
__global__ void seg_total_kernel(int nseg, const float* __restrict__ seg, float* __restrict__ total) {
  const int c = threadIdx.x;
  double s = 0.0;
  for (int i = 0; i < nseg; ++i) s += (double)seg[(long long)i * C + c];
  total[c] = (float)s;
}

// =====================================================================================================
// wgrad: D[256 co, 256 ci] += A[32 pixels, 256 co]^T * B[32 pixels, 256 ci] per filter tap, on CTA pairs.
// CTA r of a pair owns co half r (A rows + accumulator rows) and loads ci half r of the shifted input (B columns).
// A = gout chunk, B = input chunk shifted by the tap, both MN-major: 4 blocks of {32 px x 32 ch} = 4 KiB each,
// K block = 32 pixels = box {32 ch, 8 x, 4 y, 4 channel blocks, 1 img}.
// The kernel is bound by the TMA delivery rate (about 2 clk per 128-byte box row per SM), not by the tensor pipe, so
// most pairs accumulate TWO taps at once: the gout chunk is loaded once and multiplied with two differently shifted
// input chunks into two TMEM accumulators (2 x 256 columns) -- 384 instead of 512 box rows per two tap-chunks.
// Jobs (74 pairs = 148 CTAs): tap pairs (0,1) (2,3) (5,6) (7,8) x 16 pixel splits, centre tap 4 x 10 splits; a split
// takes chunks s, s+n, s+2n, ... so all pairs stream through one neighbourhood of the tensors at a time (L2 locality)
// and finish together. Deterministic second-stage reduction over the splits.
// =====================================================================================================
constexpr int WG_CX = 8, WG_CY = 4;  // pixel chunk = 8 x 4: 168x100 and 84x50 divide (almost) evenly -> 3.7 % padding
constexpr int WG_SPLITS2 = 16;       // pixel splits of a two-tap job
constexpr int WG_SPLITS1 = 10;       // pixel splits of the single (centre) tap
constexpr int WG_PAIRS = 4 * WG_SPLITS2 + WG_SPLITS1;  // 74
constexpr int WG_MAX_SPLITS = WG_SPLITS2;
constexpr int WG_BOX_BYTES = 32 * BLOCK_K * 4;  // one {32 ch x 32 px} block = 4 KiB
constexpr int WG_STAGES = 4;
constexpr int WG_STAGE_BYTES = 3 * A_BYTES;     // gout half + two shifted input halves
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + SMEM_EXTRA + 1024;
// fp16 operands (F16 = true): the same 8 x 4 pixel chunks, 64-channel MN blocks of {32 px x 128 B} = 4 KiB, two blocks
// per operand half (8 KiB), plain SWIZZLE_128B (LBO = block stride 4096, SBO = 8 pixel rows = 1024), K = 16 pixels per
// MMA -> half the TMA rows and half the MMAs per chunk. The gout copy carries a power-of-two scale (lgd_grad_scale)
// that the second-stage reduction divides out.
constexpr int WGH_CY = 8;                                  // chunk = 8 x WGH_CY pixels
constexpr int WGH_PX = WG_CX * WGH_CY;                     // K block in pixels
constexpr int WGH_BLOCK_BYTES = WGH_PX * 128;              // one {64 ch x K block} MN block
constexpr int WGH_OPERAND_BYTES = 2 * WGH_BLOCK_BYTES;     // 128-channel half
constexpr int WGH_STAGE_BYTES = 3 * WGH_OPERAND_BYTES;
constexpr int WGH_STAGES = WGH_CY == 4 ? 8 : 4;

struct WgradArgs {
  Pyr pyr;
  int chunks_x[LGD_MAX_LEVELS];
  int chunks_y[LGD_MAX_LEVELS];
  int chunk_start[LGD_MAX_LEVELS + 1];
  int total_chunks;
  int cy;          // chunk height in pixels
  float* partial;  // [9 taps][WG_MAX_SPLITS][256 co][256 ci]
};

__host__ __device__ inline int wg_splits_of_tap(int tap) { return tap == 4 ? WG_SPLITS1 : WG_SPLITS2; }

__device__ __forceinline__ void decode_chunk(const WgradArgs& a, int t, int& l, int& b, int& y0, int& x0) {
  l = 0;
  while (l + 1 < a.pyr.num_levels && t >= a.chunk_start[l + 1]) ++l;
  int r = t - a.chunk_start[l];
  const int per_img = a.chunks_x[l] * a.chunks_y[l];
  b = r / per_img;
  r -= b * per_img;
  const int cy = r / a.chunks_x[l];
  const int cx = r - cy * a.chunks_x[l];
  y0 = cy * a.cy;
  x0 = cx * WG_CX;
}

template <bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv3x3_wgrad_kernel(const __grid_constant__ ConvTmaps tm, const __grid_constant__ WgradArgs a) {
  constexpr int WG_STAGES = F16 ? lgd::WGH_STAGES : lgd::WG_STAGES;                 // shadow the tf32 constants
  constexpr int WG_STAGE_BYTES = F16 ? lgd::WGH_STAGE_BYTES : lgd::WG_STAGE_BYTES;
  constexpr int OPERAND_BYTES = F16 ? WGH_OPERAND_BYTES : A_BYTES;                   // one operand half of a chunk
  constexpr int HALF_BLOCKS = F16 ? 2 : 4;                                           // MN blocks per operand half
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(base + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* empty = full + WG_STAGES;
  uint64_t* tfull = empty + WG_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int job = blockIdx.x >> 1;  // 0 .. WG_PAIRS-1
  int tap0, ntaps, split, nsplit;
  if (job < 4 * WG_SPLITS2) {
    const int g = job / WG_SPLITS2;
    tap0 = g < 2 ? 2 * g : 2 * g + 1;  // (0,1) (2,3) (5,6) (7,8)
    ntaps = 2;
    split = job % WG_SPLITS2;
    nsplit = WG_SPLITS2;
  } else {
    tap0 = 4;
    ntaps = 1;
    split = job - 4 * WG_SPLITS2;
    nsplit = WG_SPLITS1;
  }
  const int c_begin = split, c_end = a.total_chunks;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < a.pyr.num_levels; ++l) {
      tma_prefetch_desc(&tm.act[l]);
      tma_prefetch_desc(&tm.act2[l]);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < WG_STAGES; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      mbar_init(&tfull[0], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2sm(tmem_ptr, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // whole warp, convergent, one elected lane issues (TMA operands on the uniform datapath, see conv3x3_tc_kernel)
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t ring = smem_u32(base);
      for (int t = c_begin; t < c_end; t += nsplit) {
        int l, b, y0, x0;
        decode_chunk(a, t, l, b, y0, x0);
        warp_wait(&empty[stage], phase ^ 1, lane);
        if (elect_one()) {
          const uint32_t full_leader = mapa_shared(smem_u32(&full[stage]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * (1 + ntaps) * OPERAND_BYTES);
          const uint32_t sa = ring + stage * WG_STAGE_BYTES;
          // 5-D maps {32 (64) ch, x, y, channel block, image}: ONE copy lands the MN blocks of an operand half back to
          // back ([block][pixel][128 B], 4 KiB per block)
          tma_load_5d_2sm(sa, &tm.act[l], full_leader, 0, x0, y0, (int)rank * HALF_BLOCKS, b);  // gout, co half
          for (int j = 0; j < ntaps; ++j) {                                            // input shifted by the tap, ci half
            const int tap = tap0 + j;
            tma_load_5d_2sm(sa + (1 + j) * OPERAND_BYTES, &tm.act2[l], full_leader, 0, x0 + tap % 3 - 1,
                            y0 + tap / 3 - 1, (int)rank * HALF_BLOCKS, b);
          }
        }
        __syncwarp();
        if (++stage == WG_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // whole warp, convergent, one elected lane issues: the MMA operands stay on the uniform datapath (see
    // conv3x3_tc_kernel)
    if (rank == 0) {
      constexpr uint32_t idesc = F16 ? make_idesc_f16(256, C, 1, 1) : make_idesc_tf32(256, C, 1, 1);  // MN-major A, B
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t ring = smem_u32(base);
      for (int t = c_begin; t < c_end; t += nsplit) {
        warp_wait(&full[stage], phase, lane);
        tc_fence_after();
        // MN-major tf32 = SWIZZLE_128B_BASE32B: LBO = stride between 32-element MN blocks (one 4 KiB block),
        // SBO = stride between groups of 4 K rows (512 B). MN-major fp16 = SWIZZLE_128B: 64-element MN blocks (4 KiB),
        // SBO = stride between groups of 8 K rows (1024 B).
        const uint32_t sa = ring + stage * WG_STAGE_BYTES;
        const uint64_t ad = F16 ? make_smem_desc_sw128(sa, WGH_BLOCK_BYTES, 1024) : make_smem_desc_sw128_32b(sa, WG_BOX_BYTES, 512);
        const uint32_t acc = (t > c_begin) ? 1u : 0u;
        if (elect_one()) {
          for (int j = 0; j < ntaps; ++j) {
            const uint32_t sb = sa + (1 + j) * OPERAND_BYTES;
            const uint64_t bd = F16 ? make_smem_desc_sw128(sb, WGH_BLOCK_BYTES, 1024) : make_smem_desc_sw128_32b(sb, WG_BOX_BYTES, 512);
            if (F16) {
#pragma unroll
              for (int k = 0; k < WGH_PX / 16; ++k)  // 16 pixels per MMA = two groups of 8 K rows: +128 in the (addr>>4) field
                mma_f16_ss_2sm(tmem_base + j * C, ad + 128 * k, bd + 128 * k, idesc, (acc | k) != 0 ? 1u : 0u);
            } else {
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                // next 8 pixels = next two 512 B atoms: +64 in the (addr>>4) field
                mma_tf32_ss_2sm(tmem_base + j * C, ad + 64 * k, bd + 64 * k, idesc, (acc | k) != 0 ? 1u : 0u);
              }
            }
          }
          mma_commit_2sm(&empty[stage], 3);
          if (t + nsplit >= c_end) mma_commit_2sm(&tfull[0], 3);
        }
        __syncwarp();
        if (++stage == WG_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // co within this CTA's half
    const bool any = c_end > c_begin;
    if (any) {
      mbar_wait(&tfull[0], 0);
      tc_fence_after();
    }
    for (int j = 0; j < ntaps; ++j) {
      float* optr = a.partial + (((long long)(tap0 + j) * WG_MAX_SPLITS + split) * C + (int)rank * 128 + row) * C;
      if (any) {
#pragma unroll 1
        for (int chunk = 0; chunk < C / 32; ++chunk) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(j * C + chunk * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; q += 4)
            stg4(optr + chunk * 32 + q, make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]),
                                                     __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3])));
        }
      } else {
        for (int q = 0; q < C; q += 4) stg4(optr + q, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
}

// packed_grad[tap][co][ci] = sum over the tap's splits of partial[tap][split][co][ci]   (fixed order)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                    const float* __restrict__ inv_scale) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= 9 * C * C) return;
  const float sc = inv_scale ? __ldg(inv_scale) : 1.f;
  const int tap = i / (C * C);
  const int within = i - tap * C * C;
  const float* p = partial + (long long)tap * WG_MAX_SPLITS * C * C + within;
  const int n = wg_splits_of_tap(tap);
  float4 acc = ldg4(p);
  for (int s = 1; s < n; ++s) {
    const float4 v = ldg4(p + (long long)s * C * C);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  stg4(out + i, make_float4(acc.x * sc, acc.y * sc, acc.z * sc, acc.w * sc));
}

// ----------------------------------------------------------------------------------------- weight packing
__global__ void pack_weight_kernel(const float* __restrict__ w, float* __restrict__ packed, int mode, int lo) {
  // one thread per packed element; packed index = (tap*256 + r)*256 + k
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 9 * C * C) return;
  const int k = idx & 255, r = (idx >> 8) & 255, tap = idx >> 16;
  int co, ci, src_tap;
  if (mode == 0) {
    co = r; ci = k; src_tap = tap;
  } else {  // dgrad: rows = ci, K = co, taps flipped
    ci = r; co = k; src_tap = 8 - tap;
  }
  const float v = __ldg(w + ((long long)co * C + ci) * 9 + src_tap);
  const float hi = tf32_rna(v);
  packed[idx] = lo ? tf32_rna(v - hi) : hi;  // v - hi is exact in fp32
}

// weights as fp16: mode 0 (forward) packed[tap][co][ci], mode 1 (dgrad) packed[tap][ci][co] with flipped taps.
// tap_sumsq (optional, [9*256]): sum of squares of each packed row (one block = one (tap, row)); the sum of the nine
// per-tap Frobenius norms bounds the operator norm of the convolution (scale of an fp16 dgrad output, see
// lgd_conv3x3_dgrad_f16).
__global__ void pack_weight_f16_kernel(const float* __restrict__ w, __half* __restrict__ packed, int mode,
                                       float* __restrict__ tap_sumsq) {
  __shared__ float red[32];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // grid = 9*256 blocks of 256 threads: exact cover
  const int k = idx & 255, r = (idx >> 8) & 255, tap = idx >> 16;
  int co, ci, src_tap;
  if (mode == 0) {
    co = r; ci = k; src_tap = tap;
  } else {
    ci = r; co = k; src_tap = 8 - tap;
  }
  const float v = __ldg(w + ((long long)co * C + ci) * 9 + src_tap);
  packed[idx] = __float2half_rn(v);
  if (tap_sumsq != nullptr) {
    const float t = block_sum<float>(v * v, red);
    if (threadIdx.x == 0) tap_sumsq[blockIdx.x] = t;
  }
}
// gain[0] = sum over taps of ||W_tap||_F  (>= the l2 operator norm of the 3x3 convolution and of its transpose)
__global__ void weight_gain_kernel(const float* __restrict__ tap_sumsq, float* __restrict__ gain) {
  __shared__ float red[32];
  float g = 0.f;
  for (int tap = 0; tap < 9; ++tap) {
    const float t = block_sum<float>(tap_sumsq[tap * C + threadIdx.x], red);
    g += sqrtf(t);
    __syncthreads();
  }
  if (threadIdx.x == 0) gain[0] = g;
}

// All convolution weights of a chain in ONE launch: block = (32 co x 32 ci tile, convolution). The 32 x 288 contiguous
// floats of a tile are read coalesced into shared memory, then written as fp16 in the forward layout [tap][co][ci]
// and / or the dgrad layout [8-tap][ci][co] (64-byte segments both ways). tile_sumsq (optional,
// [conv][tap][64 tiles]): per-tile sum of squares per tap for the gain (see weight_gain_multi_kernel).
struct PackJobs {
  const float* w[8];
  __half* fwd[8];    // may be nullptr
  __half* dgrad[8];  // may be nullptr
};
__global__ void __launch_bounds__(256)
pack_weights_multi_kernel(PackJobs jobs, float* __restrict__ tile_sumsq) {
  __shared__ float tile[32][289];   // [co][ci*9 + tap], pitch 289: conflict-free for both access patterns
  __shared__ float red[8][9];
  const int conv = blockIdx.y, t = blockIdx.x, co0 = (t >> 3) * 32, ci0 = (t & 7) * 32;
  const float* w = jobs.w[conv];
  for (int i = threadIdx.x; i < 32 * 288; i += 256) {
    const int r = i / 288, c = i - r * 288;
    tile[r][c] = __ldg(w + ((long long)(co0 + r) * C + ci0) * 9 + c);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __half* pf = jobs.fwd[conv];
  __half* pd = jobs.dgrad[conv];
  float ss[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) ss[tap] = 0.f;
  for (int r = warp; r < 32; r += 8) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      // forward layout: row = co0 + r, 32 consecutive ci
      const float vf = tile[r][lane * 9 + tap];
      if (pf) pf[((long long)tap * C + co0 + r) * C + ci0 + lane] = __float2half_rn(vf);
      ss[tap] += vf * vf;
      // dgrad layout: row = ci0 + r, 32 consecutive co, taps flipped
      if (pd) pd[((long long)(8 - tap) * C + ci0 + r) * C + co0 + lane] = __float2half_rn(tile[lane][r * 9 + tap]);
    }
  }
  if (tile_sumsq != nullptr) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float v = warp_sum(ss[tap]);
      if (lane == 0) red[warp][tap] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) v += red[j][threadIdx.x];
      tile_sumsq[((long long)conv * 9 + threadIdx.x) * 64 + t] = v;
    }
  }
}
// gain[conv] = sum over taps of ||W_tap||_F
__global__ void weight_gain_multi_kernel(const float* __restrict__ tile_sumsq, float* __restrict__ gains) {
  const int conv = blockIdx.x;
  if (threadIdx.x == 0) {
    float g = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      float t = 0.f;
      for (int i = 0; i < 64; ++i) t += tile_sumsq[((long long)conv * 9 + tap) * 64 + i];
      g += sqrtf(t);
    }
    gains[conv] = g;
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ packed, float* __restrict__ gw, int accumulate) {
  // gw[co][ci][tap] (+)= packed[tap][co][ci]
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over gw layout
  if (idx >= 9 * C * C) return;
  const int tap = idx % 9;
  const int ci = (idx / 9) & 255;
  const int co = idx / (9 * 256);
  const float v = __ldg(packed + ((long long)tap * C + co) * C + ci);
  gw[idx] = accumulate ? gw[idx] + v : v;
}

// ----------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

static void* driver_entry_point(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    return p;
  return nullptr;
}

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() { fn = reinterpret_cast<EncodeTiledFn>(driver_entry_point("cuTensorMapEncodeTiled")); });
  return fn;
}

static EncodeIm2colFn get_encode_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() { fn = reinterpret_cast<EncodeIm2colFn>(driver_entry_point("cuTensorMapEncodeIm2col")); });
  return fn;
}

// Memo of encoded tensor maps. An encoding is a pure function of (kind, base pointer, shape, dtype): caching it keeps
// ~350 cuTensorMapEncode* calls per step off the launch path (the stream-ordered allocator hands the same blocks back
// step after step, so nearly every lookup hits). Fixed-size open-addressed table, cleared when full.
struct MapKey {
  const void* base;
  int kind, B, H, W, p0, p1, p2;
  bool operator==(const MapKey& o) const {
    return base == o.base && kind == o.kind && B == o.B && H == o.H && W == o.W && p0 == o.p0 && p1 == o.p1 && p2 == o.p2;
  }
};
struct MapMemo {
  static constexpr int SLOTS = 1024;
  MapKey keys[SLOTS];
  CUtensorMap maps[SLOTS];
  bool used[SLOTS];
  int count = 0;
  std::mutex mu;
  MapMemo() { memset(used, 0, sizeof(used)); }
  static size_t hash(const MapKey& k) {
    size_t h = reinterpret_cast<size_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= ((size_t)k.kind << 40) ^ ((size_t)k.B << 32) ^ ((size_t)k.H << 16) ^ (size_t)k.W ^ ((size_t)k.p0 << 48) ^
         ((size_t)k.p1 << 52) ^ ((size_t)k.p2 << 56);
    return (h ^ (h >> 29)) * 0xBF58476D1CE4E5B9ull;
  }
  bool find(const MapKey& k, CUtensorMap* out) {
    std::lock_guard<std::mutex> lk(mu);
    size_t i = hash(k) % SLOTS;
    for (int probe = 0; probe < 16; ++probe, i = (i + 1) % SLOTS) {
      if (!used[i]) return false;
      if (keys[i] == k) {
        *out = maps[i];
        return true;
      }
    }
    return false;
  }
  void put(const MapKey& k, const CUtensorMap& m) {
    std::lock_guard<std::mutex> lk(mu);
    if (count > SLOTS / 2) {
      memset(used, 0, sizeof(used));
      count = 0;
    }
    size_t i = hash(k) % SLOTS;
    for (int probe = 0; probe < 16; ++probe, i = (i + 1) % SLOTS) {
      if (!used[i] || keys[i] == k) {
        if (!used[i]) ++count;
        used[i] = true;
        keys[i] = k;
        maps[i] = m;
        return;
      }
    }
  }
};
static MapMemo& map_memo() {
  static MapMemo memo;
  return memo;
}

// forward / dgrad A operand: (C, W, H, N) activation of one level read as the input strip of a tile (see STRIP_RUN):
// im2col mode over the bounding box of the zero-padded image -- lower corner (-1, -2), upper corner (+1, +1): W + 2
// positions per row, rows -2 .. H -- with zero tap offsets, so that one operation delivers consecutive padded positions
// (out-of-bounds positions zero filled) K-major in SWIZZLE_128B rows. Verified on B200 with tools/probe_strip.cu.
static int encode_act_map_strip(CUtensorMap* m, const void* base, int B, int H, int W, bool f16 = false) {
  const int es = f16 ? 2 : 4;
  const MapKey key{base, 4, B, H, W, (int)f16, 0, 0};
  if (map_memo().find(key, m)) return LGD_OK;
  EncodeIm2colFn enc = get_encode_im2col_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeIm2col entry point not available");
    return LGD_ECUDA;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
  int lower[2] = {-1, -2}, upper[2] = {1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base),
                   dims, strides, lower, upper, (cuuint32_t)(128 / es), (cuuint32_t)strip_box_pixels(W), estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col(activation strip %dx%dx%d) failed with CUresult %d", B, H, W, (int)r);
    return LGD_ECUDA;
  }
  // drivers up to CUDA 13.1 mis-encode im2col maps of tensors smaller than 128 KiB (same fix-up CUTLASS applies)
  int drv = 0;
  if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && (size_t)B * H * W * C * es < 131072)
    reinterpret_cast<uint64_t*>(m)[1] &= ~(1llu << 21);
  map_memo().put(key, *m);
  return LGD_OK;
}

// wgrad operand map: the 256 channels are split into (32 inner, 8 blocks) and the block index is made the 4th
// dimension, so that a box {32, box_x, box_y, nblk, 1} arrives in shared memory as [block][y][x][32 ch] -- the
// MN-major SWIZZLE_128B_BASE32B operand layout (LBO = one block = box_x*box_y*128 bytes).
static int encode_act_map_blocked(CUtensorMap* m, const void* base, int B, int H, int W, int box_x, int box_y,
                                  int nblk, bool f16 = false) {
  const MapKey key{base, 2, B, H, W, (int)f16, box_x * 64 + box_y, nblk};
  if (map_memo().find(key, m)) return LGD_OK;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return LGD_ECUDA;
  }
  const cuuint64_t es = f16 ? 2 : 4, blk = f16 ? 64 : 32;   // 128-byte channel blocks
  cuuint64_t dims[5] = {blk, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / blk), (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, 128, (cuuint64_t)H * W * C * es};
  cuuint32_t box[5] = {(cuuint32_t)blk, (cuuint32_t)box_x, (cuuint32_t)box_y, (cuuint32_t)nblk, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   f16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(blocked activation %dx%dx%d) failed with CUresult %d", B, H, W, (int)r);
    return LGD_ECUDA;
  }
  map_memo().put(key, *m);
  return LGD_OK;
}

// packed weights [9*256 rows][256 k]: box = {32 k, 128 rows} = the half of a (tap, k-chunk) tile one CTA of a pair loads
static int encode_weight_map(CUtensorMap* m, const void* packed, bool f16 = false) {
  const int es = f16 ? 2 : 4;
  const MapKey key{packed, 3, 0, 0, 0, (int)f16, 0, 0};
  if (map_memo().find(key, m)) return LGD_OK;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return LGD_ECUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)9 * C};
  cuuint64_t strides[1] = {(cuuint64_t)C * es};
  cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)(C / 2)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(packed), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r);
    return LGD_ECUDA;
  }
  map_memo().put(key, *m);
  return LGD_OK;
}

static int device_sm_count(int* sms) {
  int dev = 0;
  LGD_CUDA(cudaGetDevice(&dev));
  static int cached[64] = {0};   // SM count per device ordinal (an immutable device property)
  if (dev >= 0 && dev < 64 && cached[dev] > 0) {
    *sms = cached[dev];
    return LGD_OK;
  }
  int major = 0;
  LGD_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    set_error("liblgd_b200 needs an sm_100 device (found compute capability major %d)", major);
    return LGD_ENOSUP;
  }
  LGD_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  if (dev >= 0 && dev < 64) cached[dev] = *sms;
  return LGD_OK;
}

static void fill_tiles(const Pyr& p, ConvArgs* a) {
  int acc = 0;
  for (int l = 0; l < LGD_MAX_LEVELS; ++l) {
    a->tile_start[l] = acc;
    if (l < p.num_levels) {
      a->tiles_img[l] = tiles_per_image(p.h[l], p.w[l]);
      acc += tiles_per_level(p.h[l], p.w[l], p.batch);
    } else {
      a->tiles_img[l] = 0;
    }
  }
  a->tile_start[LGD_MAX_LEVELS] = acc;
  a->total_tiles = acc;
}

}  // namespace lgd

using namespace lgd;

extern "C" int lgd_conv3x3_num_tiles(const lgd_pyramid_t* pyr) {
  Pyr p;
  if (make_pyr(pyr, &p) != LGD_OK) return LGD_EINVAL;
  ConvArgs a;
  fill_tiles(p, &a);
  return a.total_tiles;
}

extern "C" int lgd_pack_conv_weight(const float* w, float* packed, int mode, void* stream) {
  LGD_CHECK_ARG(w && packed && mode >= 0 && mode <= 3, "lgd_pack_conv_weight: bad arguments");
  pack_weight_kernel<<<(9 * C * C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, packed, mode & 1, mode >> 1);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_unpack_conv_wgrad(const float* packed_grad, float* gw, int accumulate, void* stream) {
  LGD_CHECK_ARG(packed_grad && gw, "lgd_unpack_conv_wgrad: null pointer");
  unpack_wgrad_kernel<<<(9 * C * C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(packed_grad, gw, accumulate);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" size_t lgd_conv3x3_fwd_workspace(const lgd_pyramid_t* pyr) {
  Pyr p;
  if (make_pyr(pyr, &p) != LGD_OK) return 0;
  ConvArgs a;
  fill_tiles(p, &a);
  return ((size_t)a.total_tiles + (size_t)p.num_levels * p.batch) * C * sizeof(float);
}

// shared launcher of the two operand precisions
template <bool F16, bool ADD = false, int MASK = 0, bool WIDE = false>
static int launch_conv_t(const lgd_pyramid_t* pyr, const void* in, const void* packed_w, const float* bias,
                       int bias_level_stride, int bias_image_stride, float* out, void* out_half, int relu,
                       int round_out, const float* relu_mask, float* tile_stats, float* chan_sums, float* chan_total,
                       void* workspace, size_t workspace_bytes, void* stream, const float* addend = nullptr,
                       const float* acc_scale = nullptr, const float* half_scale = nullptr,
                       const void* relu_mask_half = nullptr, int out_ld = C, int out_col0 = 0, int out_cols = C,
                       const float* gn_x = nullptr, const float* gn_stats = nullptr, float* tile_gn = nullptr,
                       int gn_relu = 0) {
  const bool want_csum = chan_sums != nullptr || chan_total != nullptr;
  LGD_CHECK_ARG(!want_csum || (workspace != nullptr && workspace_bytes >= lgd_conv3x3_fwd_workspace(pyr)),
                "lgd_conv3x3_fwd: channel sums need lgd_conv3x3_fwd_workspace() bytes of workspace");
  ConvArgs a;
  int rc = make_pyr(pyr, &a.pyr);
  if (rc != LGD_OK) return rc;
  fill_tiles(a.pyr, &a);
  int sms = 0;
  rc = device_sm_count(&sms);
  if (rc != LGD_OK) return rc;
  ConvTmaps tm;
  memset(&tm, 0, sizeof(tm));
  for (int l = 0; l < a.pyr.num_levels; ++l) {
    const char* lvl = static_cast<const char*>(in) + a.pyr.off[l] * (F16 ? 2 : 4);
    rc = encode_act_map_strip(&tm.act[l], lvl, a.pyr.batch, a.pyr.h[l], a.pyr.w[l], F16);
    if (rc != LGD_OK) return rc;
  }
  rc = encode_weight_map(&tm.w, packed_w, F16);
  if (rc != LGD_OK) return rc;
  a.bias = bias;
  a.bias_lstride = bias_level_stride;
  a.bias_istride = bias_image_stride;
  a.out = out;
  a.out_half = static_cast<__half*>(out_half);
  a.relu_mask = relu_mask;
  a.relu_mask_h = static_cast<const __half*>(relu_mask_half);
  a.addend = addend;
  a.acc_scale = acc_scale;
  a.half_scale = half_scale;
  a.tile_stats = tile_stats;
  a.tile_csum = want_csum ? static_cast<float*>(workspace) : nullptr;
  a.relu = relu;
  a.round_out = round_out;
  a.out_ld = out_ld;
  a.out_col0 = out_col0;
  a.out_cols = out_cols;
  a.gn_x = gn_x;
  a.gn_stats = gn_stats;
  a.tile_gn = tile_gn;
  a.gn_relu = gn_relu;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, []() {
    attr_err = cudaFuncSetAttribute(conv3x3_tc_kernel<F16, ADD, MASK, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    FWD_SMEM_BYTES);
  });
  LGD_CUDA(attr_err);
  int grid = a.total_tiles;  // persistent: one CTA per SM, whole pairs only (total_tiles is even)
  if (grid > (sms & ~1)) grid = sms & ~1;
  conv3x3_tc_kernel<F16, ADD, MASK, WIDE><<<grid, FWD_THREADS, FWD_SMEM_BYTES, (cudaStream_t)stream>>>(tm, a);
  LGD_LAUNCH_CHECK();
  if (want_csum) {
    const int nseg = a.pyr.num_levels * a.pyr.batch;
    float* seg = chan_sums ? chan_sums : a.tile_csum + (size_t)a.total_tiles * C;
    tile_csum_finalize_kernel<<<nseg, C, 0, (cudaStream_t)stream>>>(a.pyr, a.tile_csum, seg);
    LGD_LAUNCH_CHECK();
    if (chan_total) {
      seg_total_kernel<<<1, C, 0, (cudaStream_t)stream>>>(nseg, seg, chan_total);
      LGD_LAUNCH_CHECK();
    }
  }
  return LGD_OK;
}

// masked / channel-sum launches take the MASK instantiation (TF32 operands only: the fp16 forward never masks)
template <bool F16, bool ADD = false>
static int launch_conv(const lgd_pyramid_t* pyr, const void* in, const void* packed_w, const float* bias,
                       int bias_level_stride, int bias_image_stride, float* out, void* out_half, int relu,
                       int round_out, const float* relu_mask, float* tile_stats, float* chan_sums, float* chan_total,
                       void* workspace, size_t workspace_bytes, void* stream, const float* addend = nullptr,
                       const float* acc_scale = nullptr, const float* half_scale = nullptr,
                       const void* relu_mask_half = nullptr, int out_ld = C, int out_col0 = 0, int out_cols = C) {
  if (out_ld != C || out_col0 != 0 || out_cols != C) {
    if (F16 && !ADD)
      return launch_conv_t<true, false, 0, true>(pyr, in, packed_w, bias, bias_level_stride, bias_image_stride, out,
                                                 out_half, relu, round_out, nullptr, tile_stats, nullptr, nullptr,
                                                 workspace, workspace_bytes, stream, addend, acc_scale, half_scale,
                                                 nullptr, out_ld, out_col0, out_cols);
    set_error("column-sliced outputs are available for the fp16 forward convolution only");
    return LGD_EINVAL;
  }
  if (relu_mask_half != nullptr)
    return launch_conv_t<F16, ADD, 2>(pyr, in, packed_w, bias, bias_level_stride, bias_image_stride, out, out_half,
                                      relu, round_out, relu_mask, tile_stats, chan_sums, chan_total, workspace,
                                      workspace_bytes, stream, addend, acc_scale, half_scale, relu_mask_half);
  if (relu_mask != nullptr || chan_sums != nullptr || chan_total != nullptr)
    return launch_conv_t<F16, ADD, 1>(pyr, in, packed_w, bias, bias_level_stride, bias_image_stride, out, out_half,
                                      relu, round_out, relu_mask, tile_stats, chan_sums, chan_total, workspace,
                                      workspace_bytes, stream, addend, acc_scale, half_scale, relu_mask_half);
  return launch_conv_t<F16, ADD, 0>(pyr, in, packed_w, bias, bias_level_stride, bias_image_stride, out, out_half,
                                        relu, round_out, relu_mask, tile_stats, chan_sums, chan_total, workspace,
                                        workspace_bytes, stream, addend, acc_scale, half_scale);
}

extern "C" int lgd_conv3x3_fwd(const lgd_pyramid_t* pyr, const float* in, const float* packed_w, const float* bias,
                               int bias_level_stride, int bias_image_stride, float* out, int relu, int round_out,
                               const float* relu_mask, float* tile_stats, float* chan_sums, float* chan_total,
                               void* workspace, size_t workspace_bytes, void* stream) {
  LGD_CHECK_ARG(in && packed_w && out, "lgd_conv3x3_fwd: null pointer");
  LGD_CHECK_ARG(in != out, "lgd_conv3x3_fwd: in-place convolution is not supported");
  return launch_conv<false>(pyr, in, packed_w, bias, bias_level_stride, bias_image_stride, out, nullptr, relu, round_out,
                            relu_mask, tile_stats, chan_sums, chan_total, workspace, workspace_bytes, stream);
}

extern "C" int lgd_conv3x3_fwd_addend(const lgd_pyramid_t* pyr, const float* in, const float* packed_w,
                                      const float* addend, const float* bias, int bias_level_stride,
                                      int bias_image_stride, float* out, int relu, int round_out,
                                      const float* relu_mask, float* tile_stats, float* chan_sums, float* chan_total,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  LGD_CHECK_ARG(in && packed_w && out && addend, "lgd_conv3x3_fwd_addend: null pointer");
  LGD_CHECK_ARG(in != out, "lgd_conv3x3_fwd_addend: in-place convolution is not supported");
  return launch_conv<false, true>(pyr, in, packed_w, bias, bias_level_stride, bias_image_stride, out, nullptr, relu,
                                  round_out, relu_mask, tile_stats, chan_sums, chan_total, workspace, workspace_bytes,
                                  stream, addend);
}

extern "C" int lgd_pack_conv_weight_f16(const float* w, void* packed_half, int mode, float* gain, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  LGD_CHECK_ARG(w && packed_half && (mode == 0 || mode == 1), "lgd_pack_conv_weight_f16: bad arguments");
  LGD_CHECK_ARG(gain == nullptr || (workspace != nullptr && workspace_bytes >= 9 * C * sizeof(float)),
                "lgd_pack_conv_weight_f16: the gain needs 9*256 floats of workspace");
  float* tap_sumsq = gain ? static_cast<float*>(workspace) : nullptr;
  pack_weight_f16_kernel<<<9 * C, C, 0, (cudaStream_t)stream>>>(w, static_cast<__half*>(packed_half), mode, tap_sumsq);
  LGD_LAUNCH_CHECK();
  if (gain) {
    weight_gain_kernel<<<1, C, 0, (cudaStream_t)stream>>>(tap_sumsq, gain);
    LGD_LAUNCH_CHECK();
  }
  return LGD_OK;
}

extern "C" int lgd_pack_conv_weights_f16_multi(const float* const* w_host, int n, void* const* fwd_host,
                                               void* const* dgrad_host, float* gains, void* workspace,
                                               size_t workspace_bytes, void* stream) {
  LGD_CHECK_ARG(w_host && n >= 1 && n <= 8, "lgd_pack_conv_weights_f16_multi: between 1 and 8 convolutions per call");
  LGD_CHECK_ARG(gains == nullptr || (workspace != nullptr && workspace_bytes >= (size_t)n * 9 * 64 * sizeof(float)),
                "lgd_pack_conv_weights_f16_multi: the gains need n*9*64 floats of workspace");
  PackJobs jobs;
  memset(&jobs, 0, sizeof(jobs));
  for (int i = 0; i < n; ++i) {
    LGD_CHECK_ARG(w_host[i] != nullptr, "lgd_pack_conv_weights_f16_multi: null weight pointer");
    jobs.w[i] = w_host[i];
    jobs.fwd[i] = fwd_host ? static_cast<__half*>(fwd_host[i]) : nullptr;
    jobs.dgrad[i] = dgrad_host ? static_cast<__half*>(dgrad_host[i]) : nullptr;
  }
  float* tss = gains ? static_cast<float*>(workspace) : nullptr;
  pack_weights_multi_kernel<<<dim3(64, n), 256, 0, (cudaStream_t)stream>>>(jobs, tss);
  LGD_LAUNCH_CHECK();
  if (gains) {
    weight_gain_multi_kernel<<<n, 32, 0, (cudaStream_t)stream>>>(tss, gains);
    LGD_LAUNCH_CHECK();
  }
  return LGD_OK;
}

extern "C" int lgd_conv3x3_fwd_f16(const lgd_pyramid_t* pyr, const void* in_half, const void* packed_w_half,
                                   const float* bias, int bias_level_stride, int bias_image_stride, float* out,
                                   void* out_half, int relu, int round_out, float* tile_stats, void* stream) {
  LGD_CHECK_ARG(in_half && packed_w_half && (out || out_half), "lgd_conv3x3_fwd_f16: null pointer");
  return launch_conv<true>(pyr, in_half, packed_w_half, bias, bias_level_stride, bias_image_stride, out, out_half, relu,
                           round_out, nullptr, tile_stats, nullptr, nullptr, nullptr, 0, stream);
}

extern "C" int lgd_conv3x3_dgrad_f16(const lgd_pyramid_t* pyr, const void* gout_half, const void* packed_w_half,
                                     const float* acc_scale, float* out, int round_out, const float* relu_mask,
                                     const void* relu_mask_half, void* out_half, const float* half_scale,
                                     float* tile_stats, float* chan_sums, float* chan_total, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  LGD_CHECK_ARG(!(relu_mask && relu_mask_half), "lgd_conv3x3_dgrad_f16: give the ReLU mask as fp32 OR as fp16");
  LGD_CHECK_ARG(gout_half && packed_w_half && acc_scale && (out || out_half), "lgd_conv3x3_dgrad_f16: null pointer");
  LGD_CHECK_ARG(out_half == nullptr || half_scale != nullptr, "lgd_conv3x3_dgrad_f16: out_half needs half_scale");
  return launch_conv<true>(pyr, gout_half, packed_w_half, nullptr, 0, 0, out, out_half, 0, round_out, relu_mask,
                           tile_stats, chan_sums, chan_total, workspace, workspace_bytes, stream, nullptr, acc_scale,
                           half_scale, relu_mask_half);
}

extern "C" int lgd_conv3x3_dgrad_f16_gnsums(const lgd_pyramid_t* pyr, const void* gout_half, const void* packed_w_half,
                                           const float* acc_scale, float* out, const float* gn_x,
                                           const float* gn_stats, int gn_relu, float* tile_gn, void* stream) {
  LGD_CHECK_ARG(gout_half && packed_w_half && acc_scale && out && gn_x && gn_stats && tile_gn,
                "lgd_conv3x3_dgrad_f16_gnsums: null pointer");
  return launch_conv_t<true, false, 3>(pyr, gout_half, packed_w_half, nullptr, 0, 0, out, nullptr, 0, 0, nullptr, nullptr,
                                       nullptr, nullptr, nullptr, 0, stream, nullptr, acc_scale, nullptr, nullptr, C, 0, C,
                                       gn_x, gn_stats, tile_gn, gn_relu);
}

extern "C" int lgd_conv3x3_dgrad_f16_gnsums_y(const lgd_pyramid_t* pyr, const void* gout_half, const void* packed_w_half,
                                             const float* acc_scale, float* out, const void* gn_y_half, float* tile_gn,
                                             void* stream) {
  LGD_CHECK_ARG(gout_half && packed_w_half && acc_scale && out && gn_y_half && tile_gn,
                "lgd_conv3x3_dgrad_f16_gnsums_y: null pointer");
  return launch_conv_t<true, false, 4>(pyr, gout_half, packed_w_half, nullptr, 0, 0, out, nullptr, 0, 0, nullptr, nullptr,
                                       nullptr, nullptr, nullptr, 0, stream, nullptr, acc_scale, nullptr, gn_y_half, C, 0,
                                       C, nullptr, nullptr, tile_gn, 1);
}

extern "C" int lgd_conv3x3_fwd_f16_cols(const lgd_pyramid_t* pyr, const void* in_half, const void* packed_w_half,
                                        const float* bias256, float* out, int out_ld, int out_col0, int out_cols,
                                        int relu, void* stream) {
  LGD_CHECK_ARG(in_half && packed_w_half && out, "lgd_conv3x3_fwd_f16_cols: null pointer");
  LGD_CHECK_ARG(out_cols > 0 && out_cols <= C && out_cols % 4 == 0 && out_col0 >= 0 && out_col0 % 4 == 0 &&
                    out_ld % 4 == 0 && out_col0 + out_cols <= out_ld,
                "lgd_conv3x3_fwd_f16_cols: columns must be multiples of 4 inside the row (ld %d, col0 %d, cols %d)", out_ld,
                out_col0, out_cols);
  return launch_conv<true>(pyr, in_half, packed_w_half, bias256, 0, 0, out, nullptr, relu, 0, nullptr, nullptr, nullptr,
                           nullptr, nullptr, 0, stream, nullptr, nullptr, nullptr, nullptr, out_ld, out_col0, out_cols);
}

extern "C" int lgd_conv3x3_dgrad_f16_addend(const lgd_pyramid_t* pyr, const void* gout_half, const void* packed_w_half,
                                            const float* acc_scale, const float* addend, float* out,
                                            const void* relu_mask_half, void* out_half, const float* half_scale,
                                            float* tile_stats, float* chan_sums, float* chan_total, void* workspace,
                                            size_t workspace_bytes, void* stream) {
  LGD_CHECK_ARG(gout_half && packed_w_half && acc_scale && addend && (out || out_half),
                "lgd_conv3x3_dgrad_f16_addend: null pointer");
  LGD_CHECK_ARG(out_half == nullptr || half_scale != nullptr, "lgd_conv3x3_dgrad_f16_addend: out_half needs half_scale");
  return launch_conv<true, true>(pyr, gout_half, packed_w_half, nullptr, 0, 0, out, out_half, 0, 0, nullptr, tile_stats,
                                 chan_sums, chan_total, workspace, workspace_bytes, stream, addend, acc_scale, half_scale,
                                 relu_mask_half);
}

// packed[tap][row][k] for the 256 output channels [co0, co0 + 256) of a (co_total, 256, 3, 3) weight, rows beyond
// co_total zero: mode 0 forward layout (row = co - co0, k = ci), mode 1 dgrad layout (row = ci, k = co - co0, taps
// flipped). One block per (tap, row).
__global__ void pack_weight_rows_f16_kernel(const float* __restrict__ w, int co_total, int co0, __half* __restrict__ packed,
                                            int mode, float* __restrict__ tap_sumsq) {
  __shared__ float red[32];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = idx & 255, r = (idx >> 8) & 255, tap = idx >> 16;
  int co, ci, src_tap;
  if (mode == 0) {
    co = co0 + r; ci = k; src_tap = tap;
  } else {
    ci = r; co = co0 + k; src_tap = 8 - tap;
  }
  const float v = co < co_total ? __ldg(w + ((long long)co * C + ci) * 9 + src_tap) : 0.f;
  packed[idx] = __float2half_rn(v);
  if (tap_sumsq != nullptr) {
    const float t = block_sum<float>(v * v, red);
    if (threadIdx.x == 0) tap_sumsq[blockIdx.x] = t;
  }
}
__global__ void pad_bias_kernel(const float* __restrict__ bias, int co_total, int co0, float* __restrict__ out) {
  const int c = threadIdx.x;
  out[c] = (co0 + c < co_total) ? bias[co0 + c] : 0.f;
}

extern "C" int lgd_pack_conv_weight_f16_rows(const float* w, const float* bias, int co_total, int co0, void* fwd_half,
                                             void* dgrad_half, float* bias256, float* gain, void* workspace,
                                             size_t workspace_bytes, void* stream) {
  LGD_CHECK_ARG(w && co_total > 0 && co0 >= 0 && co0 < co_total, "lgd_pack_conv_weight_f16_rows: bad arguments");
  LGD_CHECK_ARG(gain == nullptr || (dgrad_half && workspace && workspace_bytes >= 9 * C * sizeof(float)),
                "lgd_pack_conv_weight_f16_rows: the gain comes with the dgrad layout and needs 9*256 floats of workspace");
  cudaStream_t s = (cudaStream_t)stream;
  if (fwd_half) {
    pack_weight_rows_f16_kernel<<<9 * C, C, 0, s>>>(w, co_total, co0, static_cast<__half*>(fwd_half), 0, nullptr);
    LGD_LAUNCH_CHECK();
  }
  if (dgrad_half) {
    float* tss = gain ? static_cast<float*>(workspace) : nullptr;
    pack_weight_rows_f16_kernel<<<9 * C, C, 0, s>>>(w, co_total, co0, static_cast<__half*>(dgrad_half), 1, tss);
    LGD_LAUNCH_CHECK();
    if (gain) {
      weight_gain_kernel<<<1, C, 0, s>>>(tss, gain);
      LGD_LAUNCH_CHECK();
    }
  }
  if (bias256) {
    LGD_CHECK_ARG(bias != nullptr, "lgd_pack_conv_weight_f16_rows: bias256 needs bias");
    pad_bias_kernel<<<1, C, 0, s>>>(bias, co_total, co0, bias256);
    LGD_LAUNCH_CHECK();
  }
  return LGD_OK;
}

// gw[(co0 + co)][ci][tap] = packed[tap][co][ci] for co < co_count (a 256-row chunk of a wider weight gradient)
__global__ void unpack_wgrad_rows_kernel(const float* __restrict__ packed, float* __restrict__ gw, int co0, int co_count) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over (co_count, 256, 9)
  if (idx >= co_count * C * 9) return;
  const int tap = idx % 9;
  const int ci = (idx / 9) & 255;
  const int co = idx / (9 * 256);
  gw[(long long)co0 * C * 9 + idx] = __ldg(packed + ((long long)tap * C + co) * C + ci);
}
extern "C" int lgd_unpack_conv_wgrad_rows(const float* packed_grad, float* gw, int co0, int co_count, void* stream) {
  LGD_CHECK_ARG(packed_grad && gw && co0 >= 0 && co_count > 0 && co_count <= C, "lgd_unpack_conv_wgrad_rows: bad arguments");
  unpack_wgrad_rows_kernel<<<(co_count * C * 9 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(packed_grad, gw, co0, co_count);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" size_t lgd_conv3x3_wgrad_workspace(const lgd_pyramid_t* pyr) {
  (void)pyr;
  return (size_t)WG_MAX_SPLITS * 9 * C * C * sizeof(float);
}

template <bool F16>
static int launch_wgrad(const lgd_pyramid_t* pyr, const void* in, const void* gout, const float* inv_scale,
                        float* packed_grad, void* workspace, size_t workspace_bytes, void* stream) {
  LGD_CHECK_ARG(in && gout && packed_grad && workspace, "lgd_conv3x3_wgrad: null pointer");
  LGD_CHECK_ARG(workspace_bytes >= lgd_conv3x3_wgrad_workspace(pyr), "lgd_conv3x3_wgrad: workspace too small");
  WgradArgs a;
  int rc = make_pyr(pyr, &a.pyr);
  if (rc != LGD_OK) return rc;
  int sms = 0;
  rc = device_sm_count(&sms);
  if (rc != LGD_OK) return rc;
  int acc = 0;
  constexpr int cy = F16 ? WGH_CY : WG_CY;
  a.cy = cy;
  for (int l = 0; l < LGD_MAX_LEVELS; ++l) {
    a.chunk_start[l] = acc;
    if (l < a.pyr.num_levels) {
      a.chunks_x[l] = (a.pyr.w[l] + WG_CX - 1) / WG_CX;
      a.chunks_y[l] = (a.pyr.h[l] + cy - 1) / cy;
      acc += a.pyr.batch * a.chunks_x[l] * a.chunks_y[l];
    } else {
      a.chunks_x[l] = a.chunks_y[l] = 0;
    }
  }
  a.chunk_start[LGD_MAX_LEVELS] = acc;
  a.total_chunks = acc;
  a.partial = static_cast<float*>(workspace);
  ConvTmaps tm;
  memset(&tm, 0, sizeof(tm));
  constexpr int es = F16 ? 2 : 4, half_blocks = F16 ? 2 : 4;
  for (int l = 0; l < a.pyr.num_levels; ++l) {
    rc = encode_act_map_blocked(&tm.act[l], static_cast<const char*>(gout) + a.pyr.off[l] * es, a.pyr.batch,
                                a.pyr.h[l], a.pyr.w[l], WG_CX, cy, half_blocks, F16);
    if (rc != LGD_OK) return rc;
    rc = encode_act_map_blocked(&tm.act2[l], static_cast<const char*>(in) + a.pyr.off[l] * es, a.pyr.batch,
                                a.pyr.h[l], a.pyr.w[l], WG_CX, cy, half_blocks, F16);
    if (rc != LGD_OK) return rc;
  }
  constexpr int smem = (F16 ? WGH_STAGES * WGH_STAGE_BYTES : WG_STAGES * WG_STAGE_BYTES) + SMEM_EXTRA + 1024;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, []() {
    attr_err = cudaFuncSetAttribute(conv3x3_wgrad_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  });
  LGD_CUDA(attr_err);
  conv3x3_wgrad_kernel<F16><<<2 * WG_PAIRS, NUM_THREADS, smem, (cudaStream_t)stream>>>(tm, a);
  LGD_LAUNCH_CHECK();
  wgrad_reduce_kernel<<<(9 * C * C / 4 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a.partial, packed_grad, inv_scale);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_conv3x3_wgrad(const lgd_pyramid_t* pyr, const float* in, const float* gout, float* packed_grad,
                                 float* gbias, void* workspace, size_t workspace_bytes, void* stream) {
  (void)gbias;  // bias gradients are by-products of the kernel that produced gout (or lgd_pyramid_channel_sums)
  return launch_wgrad<false>(pyr, in, gout, nullptr, packed_grad, workspace, workspace_bytes, stream);
}

extern "C" int lgd_conv3x3_wgrad_f16(const lgd_pyramid_t* pyr, const void* in_half, const void* gout_half,
                                     const float* inv_scale, float* packed_grad, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  LGD_CHECK_ARG(inv_scale != nullptr, "lgd_conv3x3_wgrad_f16: inv_scale (device scalar) is required");
  return launch_wgrad<true>(pyr, in_half, gout_half, inv_scale, packed_grad, workspace, workspace_bytes, stream);
}
