// Persistent warp-specialised 16-bit GEMM for sm_100a: C[M,N] = A[M,K] * W[N,K]^T, fp32 accumulation in
// TMEM, operands staged by TMA into 128B-swizzled shared memory, tcgen05.mma issued by one thread.
// Operands are fp16 or bf16 (template FMT; tcgen05 kind::f16 runs both at the same rate).
//
// Split-operand ("x2") mode for the decision re-check path: A and W are given as TWO fp16 planes each (x = hi + lo,
// plane p at columns [p K, (p+1) K) of the row) and the kernel accumulates the three products A_lo W_hi + A_hi W_lo +
// A_hi W_hi into the same TMEM accumulator -- the k-loop simply runs over 3 K/64 k-blocks whose TMA coordinates come
// from a product table -- which carries 22 significant bits per operand (fp32-class results) at 3x the tensor time.
// Weights are pre-scaled by a power of two (acc_scale undoes it exactly) so that their lo plane stays in fp16's
// normal range.  The ..._SPLIT epilogues write their 16-bit output as hi / lo planes again.
//
//   warp 0      TMA producer   (4-stage ring of {A 128x64, W 256x64} bf16 tiles, 48 KiB / stage)
//   warp 1      TMEM allocator + MMA issuer (UMMA 128x256x16, 4 per stage)
//   warps 2..9  epilogue: TMEM -> registers -> fused bias / GELU -> 128B-swizzled smem staging -> TMA tile store
//               (bf16 outputs) or TMA tile REDUCE-ADD into the fp32 residual stream (x += acc + bias without the
//               SMs ever reading x); the patch-embedding variant scatters rows directly (+position table)
//
// The accumulator is double buffered in TMEM (2 x 256 columns) so the epilogue of tile t overlaps the
// MMAs of tile t+1.  Replaces the cuBLAS calls behind nn.Linear on the reference path
// (HF:modeling_audio_spectrogram_transformer.py:146-148,197,230,243) and the patch-embedding conv (:88-96).
#include <stdlib.h>

#include "zk_b200.h"
#include "zk_common.cuh"
#include "zk_internal.cuh"

namespace zk {

namespace gemm {
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int STG_BYTES = 32 * 128;  // per epilogue warp: 32 rows x 128 B (64 bf16 / 32 fp32 columns)
constexpr int OFF_STG = STAGES * STAGE_BYTES, OFF_BAR = OFF_STG + EPI_WARPS * STG_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;  // + barriers + alignment slack
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int MAX_PRODUCTS = 3;

struct Params {
  const float* bias;
  void* out;
  const float* aux;
  long long M;
  int N, K, aux_rows;
  int num_m_tiles, num_n_tiles;
  float acc_scale;             // result = acc * acc_scale + bias (undoes the power-of-two weight scale)
  int nprod;                   // 1, or 3 in split-operand mode
  int a_off[MAX_PRODUCTS];     // column (element) offset of the A plane of product i
  int w_off[MAX_PRODUCTS];     // same for W
  int out_plane_stride;        // ..._SPLIT epilogues: columns between the hi and lo output planes (= N)
  int seg_kb;                  // SEG kernels: k-blocks per accumulator chain
  // LayerNorm tail (LNT kernels, N = 768 residual GEMMs): rows of the residual stream that all their column tiles have
  // been reduce-added into are normalised by extra warps of the CTA that ran the LAST column tile, while they are
  // still in L2, and written as the next GEMM's 16-bit A operand
  const float* ln_w;
  const float* ln_b;
  uint16_t* ln_out;            // [M][768]
  int* ln_count;               // [num_m_tiles][2]: epilogue-warp arrivals per (256-row tile, CTA of the pair); zeroed by the host
  float ln_eps;
};

// PATCH epilogue only: one thread owns 32 consecutive columns [col0, col0+32) of row `row`; rows are scattered to
// token rows 2.. of their window and the position table is added (0.2 % of the FLOPs, direct stores are fine).
__device__ __forceinline__ void patch_store(const Params& p, long long row, int col0, const uint32_t (&r)[32]) {
  if (row >= p.M) return;
  const float4* bias4 = reinterpret_cast<const float4*>(p.bias + col0);
  const long long w = row / p.aux_rows;
  const int pr = (int)(row - w * p.aux_rows);
  const long long orow = w * (p.aux_rows + 2) + 2 + pr;
  const float4* pos4 = reinterpret_cast<const float4*>(p.aux + (long long)(2 + pr) * p.N + col0);
  float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.N + col0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 b = __ldg(bias4 + i);
    float4 o = __ldg(pos4 + i);
    o.x += fmaf(__uint_as_float(r[i * 4 + 0]), p.acc_scale, b.x);
    o.y += fmaf(__uint_as_float(r[i * 4 + 1]), p.acc_scale, b.y);
    o.z += fmaf(__uint_as_float(r[i * 4 + 2]), p.acc_scale, b.z);
    o.w += fmaf(__uint_as_float(r[i * 4 + 3]), p.acc_scale, b.w);
    dst[i] = o;
  }
}

// erf-GELU (HF "gelu", activations.py) for two values at once with ONE MUFU op per value: with a = |x|,
//     erfc(a / sqrt2) / 2 = 2^P(a),  P = degree-5 weighted least-squares fit of log2(erfc(a / sqrt2) / 2) on [0, 7.07]
// (beyond 7.07 the term is < 1e-12), and gelu(x) = max(x, 0) - a * 2^P(a), which covers both signs because
// x Phi(x) = x - x erfc(x / sqrt2) / 2 for x >= 0 and = -a erfc(a / sqrt2) / 2 for x < 0.  The fit is weighted by
// a exp(-0.55 a^2), i.e. by how much an error of P moves the result; against the exact function, evaluated in fp32:
// max abs error 6.6e-7, max error relative to max(|gelu|, 1e-3) 4.7e-4 -- an order below the bf16 rounding of the
// result (3.9e-3).  The Horner chain runs on the packed pipe (5 FFMA2 per pair); 6.5 instructions and one ex2 per value
// instead of 18 and two (rcp + ex2) for Abramowitz-Stegun 7.1.26: the fc1 epilogue is 32 768 values per tile, and under
// the power cap every instruction it issues is clock taken from the tensor pipe.
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 a = make_float2(fminf(fabsf(x.x), 7.0710678f), fminf(fabsf(x.y), 7.0710678f));
  float2 p = ffma2(a, make_float2(-0.00047368594096042216f, -0.00047368594096042216f),
                   make_float2(0.007084728218615055f, 0.007084728218615055f));
  p = ffma2(p, a, make_float2(-0.05182575806975365f, -0.05182575806975365f));
  p = ffma2(p, a, make_float2(-0.4599885642528534f, -0.4599885642528534f));
  p = ffma2(p, a, make_float2(-1.150799036026001f, -1.150799036026001f));
  p = ffma2(p, a, make_float2(-1.0000325441360474f, -1.0000325441360474f));
  return make_float2(fmaf(-fabsf(x.x), fast_exp2(p.x), fmaxf(x.x, 0.f)), fmaf(-fabsf(x.y), fast_exp2(p.y), fmaxf(x.y, 0.f)));
}

// 32 accumulator columns (* scale + bias, optional GELU) -> 16 packed 16-bit pairs
template <bool GELU, int FMT>
__device__ __forceinline__ void bias_act_pack(const uint32_t (&r)[32], const float* bias, float scale, uint32_t (&pk)[16]) {
  const float4* bias4 = reinterpret_cast<const float4*>(bias);
  const float2 sc2 = make_float2(scale, scale);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 b = __ldg(bias4 + i);
    float2 v01 = ffma2(make_float2(__uint_as_float(r[i * 4 + 0]), __uint_as_float(r[i * 4 + 1])), sc2, make_float2(b.x, b.y));
    float2 v23 = ffma2(make_float2(__uint_as_float(r[i * 4 + 2]), __uint_as_float(r[i * 4 + 3])), sc2, make_float2(b.z, b.w));
    if (GELU) {
      v01 = gelu_erf2(v01);
      v23 = gelu_erf2(v23);
    }
    pk[i * 2 + 0] = pack16<FMT>(v01.x, v01.y);
    pk[i * 2 + 1] = pack16<FMT>(v23.x, v23.y);
  }
}

// erf-GELU as the reference evaluates it (HF activations "gelu" = torch.nn.functional.gelu, erf form), libdevice erff
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// split-operand epilogues: 32 accumulator columns (* scale + bias, optional exact GELU) -> fp16 hi and lo planes
template <bool GELU>
__device__ __forceinline__ void bias_act_split(const uint32_t (&r)[32], const float* bias, float scale, uint32_t (&hi)[16],
                                               uint32_t (&lo)[16]) {
  const float4* bias4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 b = __ldg(bias4 + i);
    float v0 = fmaf(__uint_as_float(r[i * 4 + 0]), scale, b.x), v1 = fmaf(__uint_as_float(r[i * 4 + 1]), scale, b.y);
    float v2 = fmaf(__uint_as_float(r[i * 4 + 2]), scale, b.z), v3 = fmaf(__uint_as_float(r[i * 4 + 3]), scale, b.w);
    if (GELU) {
      v0 = gelu_exact(v0);
      v1 = gelu_exact(v1);
      v2 = gelu_exact(v2);
      v3 = gelu_exact(v3);
    }
    split_f16_pair(v0, v1, hi[i * 2 + 0], lo[i * 2 + 0]);
    split_f16_pair(v2, v3, hi[i * 2 + 1], lo[i * 2 + 1]);
  }
}

constexpr bool epi_is_split(int epi) { return epi == ZK_EPI_BIAS_SPLIT || epi == ZK_EPI_BIAS_GELU_SPLIT; }

// Segmented accumulation (SEG kernels: the split-operand residual GEMMs, K = 768 / 3072).  The tensor core TRUNCATES on
// every accumulate -- one ulp of the accumulator per tcgen05.mma step, always towards zero; scripts/accum_probe.py
// measures a relative shrink of -4e-6 after the 48 steps of K = 768 and -1.8e-5 after the 192 of K = 3072 on
// same-signed data (torch's fp32 FFMA matmul: 1e-9) -- so a long K chain is cut into segments of SEG_KB k-blocks: 16
// hi x hi steps, preceded by the 32 steps of the two small products while the accumulator is still ~2^-11 of its final
// size.  Every segment is its own accumulator use (the two TMEM buffers alternate between segments) and its own
// epilogue pass, joined to the others by the rounded fp32 add of the TMA reduce into the residual stream.
// K = 768 runs 3 segments of 4 k-blocks; K = 3072 runs 6 of 8 (32 hi x hi steps: every segment is another 128 KiB
// reduce-add per CTA into L2, and at 12 segments that traffic, not the tensor pipe, set the pace: 0.31 against 0.22 ms).
// Measured against float64 (tests/test_gpu_gemm.py): rms error 0.4-0.7x that of torch's fp32 matmul.
constexpr int SEG_KB = 4, SEG_KB_LONG = 8, SEG_LONG_K = 2048;

// Epilogue of one accumulator tile for one warp (32 rows x 128 columns at t_acc), shared by the single-CTA and the
// CTA-pair kernels: TMEM -> registers -> fused scale / bias / GELU -> 128B-swizzled staging tile -> TMA store (16-bit
// outputs, one or two planes) or TMA reduce-add into the fp32 residual stream.
template <int EPI, int FMT>
__device__ __forceinline__ void epilogue_rows(const Params& p, const CUtensorMap* tmC, uint8_t* stg, uint32_t stg_row,
                                              uint32_t t_acc, int row0, int col_base, int lane, int seg = 0) {
  if constexpr (EPI == ZK_EPI_BIAS_RESID_F32) {
    // 4 chunks of 32 fp32 columns: (acc * scale + bias) -> staging -> TMA reduce-add into the residual stream.
    // Segmented chains (SEG kernels): every segment is reduce-added on its own (a rounded fp32 add in L2), the bias goes
    // with the first; this warp's previous segment must have LANDED before the next one is issued, so that the adds to an
    // element always happen in the same order (bit-reproducible results).
    if (seg > 0) {
      if (lane == 0) bulk_wait0();
      __syncwarp();
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld32(t_acc + c * 32, r);
      tmem_ld_wait();
      const float4* bias4 = reinterpret_cast<const float4*>(p.bias + col_base + c * 32);
      if (lane == 0) bulk_wait_read0();  // the previous TMA store has finished reading the staging tile
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 b = __ldg(bias4 + i);
        if (seg > 0) b = make_float4(0.f, 0.f, 0.f, 0.f);
        st_shared_v4(stg_row + ((uint32_t)(i ^ (lane & 7)) << 4),
                     __float_as_uint(fmaf(__uint_as_float(r[i * 4 + 0]), p.acc_scale, b.x)),
                     __float_as_uint(fmaf(__uint_as_float(r[i * 4 + 1]), p.acc_scale, b.y)),
                     __float_as_uint(fmaf(__uint_as_float(r[i * 4 + 2]), p.acc_scale, b.z)),
                     __float_as_uint(fmaf(__uint_as_float(r[i * 4 + 3]), p.acc_scale, b.w)));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_2d(tmC, stg, col_base + c * 32, row0);
        bulk_commit();
      }
    }
  } else if constexpr (epi_is_split(EPI)) {
    // 2 chunks of 64 columns, each stored twice: hi plane at column c, lo plane at column out_plane_stride + c
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t r0[32], r1[32];
      tmem_ld32(t_acc + c * 64, r0);
      tmem_ld32(t_acc + c * 64 + 32, r1);
      tmem_ld_wait();
      uint32_t hi[32], lo[32];
      bias_act_split<EPI == ZK_EPI_BIAS_GELU_SPLIT>(r0, p.bias + col_base + c * 64, p.acc_scale,
                                                    *reinterpret_cast<uint32_t(*)[16]>(&hi[0]),
                                                    *reinterpret_cast<uint32_t(*)[16]>(&lo[0]));
      bias_act_split<EPI == ZK_EPI_BIAS_GELU_SPLIT>(r1, p.bias + col_base + c * 64 + 32, p.acc_scale,
                                                    *reinterpret_cast<uint32_t(*)[16]>(&hi[16]),
                                                    *reinterpret_cast<uint32_t(*)[16]>(&lo[16]));
#pragma unroll
      for (int plane = 0; plane < 2; ++plane) {
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (plane == 0)
            st_shared_v4(stg_row + ((uint32_t)(i ^ (lane & 7)) << 4), hi[i * 4 + 0], hi[i * 4 + 1], hi[i * 4 + 2], hi[i * 4 + 3]);
          else
            st_shared_v4(stg_row + ((uint32_t)(i ^ (lane & 7)) << 4), lo[i * 4 + 0], lo[i * 4 + 1], lo[i * 4 + 2], lo[i * 4 + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tmC, stg, plane * p.out_plane_stride + col_base + c * 64, row0);
          bulk_commit();
        }
      }
    }
  } else {
    // 2 chunks of 64 16-bit columns: (acc * scale + bias [, GELU]) -> staging -> TMA store
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t r0[32], r1[32];
      tmem_ld32(t_acc + c * 64, r0);
      tmem_ld32(t_acc + c * 64 + 32, r1);
      tmem_ld_wait();
      uint32_t pk[32];
      bias_act_pack<EPI == ZK_EPI_BIAS_GELU_BF16, FMT>(r0, p.bias + col_base + c * 64, p.acc_scale,
                                                       *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
      bias_act_pack<EPI == ZK_EPI_BIAS_GELU_BF16, FMT>(r1, p.bias + col_base + c * 64 + 32, p.acc_scale,
                                                       *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i)
        st_shared_v4(stg_row + ((uint32_t)(i ^ (lane & 7)) << 4), pk[i * 4 + 0], pk[i * 4 + 1], pk[i * 4 + 2], pk[i * 4 + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmC, stg, col_base + c * 64, row0);
        bulk_commit();
      }
    }
  }
}

template <int EPI, int FMT, bool SEG>
__global__ void __launch_bounds__(THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const Params p) {
  static_assert(!SEG || EPI == ZK_EPI_BIAS_RESID_F32, "segments are joined by the reduce-add of the residual epilogue");
  constexpr uint32_t IDESC = umma_idesc_16(FMT, BM, BN, 0, 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int kblocks = p.K / BK;
  // accumulator chains per tile and k-blocks (all products) per chain
  const int nseg = SEG ? kblocks / p.seg_kb : 1;
  const int kb_per_seg = SEG ? p.seg_kb * p.nprod : kblocks * p.nprod;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.num_n_tiles, n_blk = tile - m_blk * p.num_n_tiles;
        for (int seg = 0; seg < nseg; ++seg) {
          const int kb0 = SEG ? seg * p.seg_kb : 0, kb1 = SEG ? kb0 + p.seg_kb : kblocks;
          for (int pr = 0; pr < p.nprod; ++pr) {
            const int a0 = p.a_off[pr], w0 = p.w_off[pr];
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait(&empty[stage], phase ^ 1);
              mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
              uint8_t* sa = smem + stage * STAGE_BYTES;
              tma_load_2d(sa, &tmA, &full[stage], a0 + kb * BK, m_blk * BM);
              tma_load_2d(sa + A_BYTES, &tmB, &full[stage], w0 + kb * BK, n_blk * BN);
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // The whole warp walks the pipeline on warp-uniform state and one elected lane issues: under a divergent
    // `lane == 0` ptxas wraps every UTCHMMA in an ELECT / BRA.U.ANY loop (~90 clk per instruction).
    int stage = 0;
    uint32_t phase = 0;
    int t = 0;  // accumulator uses so far: one per tile, or one per segment of a tile
    for (int u = blockIdx.x * nseg, uend = num_tiles * nseg; u < uend; u += ((u + 1) % nseg ? 1 : (gridDim.x - 1) * nseg + 1), ++t) {
      const int acc = t & 1;
      const uint32_t acc_phase = (t >> 1) & 1;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < kb_per_seg; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_desc = umma_desc_sw128(sa, 16, 1024);
          const uint64_t b_desc = umma_desc_sw128(sa + A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, IDESC, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);
          if (kb == kb_per_seg - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    const int quarter = warp & 3;         // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;     // which 128 accumulator columns
    uint8_t* stg = smem + OFF_STG + (warp - 2) * STG_BYTES;  // this warp's 32 x 128 B staging tile (1024-B aligned)
    const uint32_t stg_row = smem_u32(stg) + lane * 128;
    int t = 0;
    for (int u = blockIdx.x * nseg, uend = num_tiles * nseg; u < uend; u += ((u + 1) % nseg ? 1 : (gridDim.x - 1) * nseg + 1), ++t) {
      const int tile = u / nseg, seg = u - tile * nseg;
      const int m_blk = tile / p.num_n_tiles, n_blk = tile - m_blk * p.num_n_tiles;
      const int acc = t & 1;
      const uint32_t acc_phase = (t >> 1) & 1;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row0 = m_blk * BM + quarter * 32;
      const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + half * 128;
      const int col_base = n_blk * BN + half * 128;
      if constexpr (EPI == ZK_EPI_PATCH_F32) {
        // rows of this warp: patches pr0 .. pr0 + 31 of window w0 unless the group straddles a window (or the end of
        // the matrix); the common case goes through the staging tile and ONE TMA store per 32 x 32 block (the 3-D
        // map addresses x as [window][token][768], token = 2 + patch), the straddling groups store row by row
        const long long grow = (long long)row0;
        const long long w0 = grow / p.aux_rows;
        const int pr0 = (int)(grow - w0 * p.aux_rows);
        const bool whole = pr0 + 32 <= p.aux_rows && grow + 32 <= p.M;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(t_acc + c * 32, r);
          tmem_ld_wait();
          if (!whole) {
            patch_store(p, grow + lane, col_base + c * 32, r);
            continue;
          }
          const float4* bias4 = reinterpret_cast<const float4*>(p.bias + col_base + c * 32);
          const float4* pos4 = reinterpret_cast<const float4*>(p.aux + (long long)(2 + pr0 + lane) * p.N + col_base + c * 32);
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b = __ldg(bias4 + i);
            const float4 o = __ldg(pos4 + i);
            st_shared_v4(stg_row + ((uint32_t)(i ^ (lane & 7)) << 4),
                         __float_as_uint(fmaf(__uint_as_float(r[i * 4 + 0]), p.acc_scale, b.x) + o.x),
                         __float_as_uint(fmaf(__uint_as_float(r[i * 4 + 1]), p.acc_scale, b.y) + o.y),
                         __float_as_uint(fmaf(__uint_as_float(r[i * 4 + 2]), p.acc_scale, b.z) + o.z),
                         __float_as_uint(fmaf(__uint_as_float(r[i * 4 + 3]), p.acc_scale, b.w) + o.w));
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmC, stg, col_base + c * 32, 2 + pr0, (int)w0);
            bulk_commit();
          }
        }
      } else {
        epilogue_rows<EPI, FMT>(p, &tmC, stg, stg_row, t_acc, row0, col_base, lane, seg);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
    if (lane == 0) bulk_wait0();  // all tile stores of this warp have landed before the CTA retires
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ CTA-pair variant
// The same pipeline with tcgen05.mma.cta_group::2: a cluster of two CTAs (the two SMs of a TPC) computes a 256 x 256
// output tile.  Each CTA stages its own 128 rows of A and only HALF of the W tile (128 of the 256 output columns); the
// tensor cores of the pair exchange the W halves, so per flop each SM fills and reads a third less shared memory and
// pulls a third less from L2 than with 128 x 256 single-CTA tiles (32 KiB instead of 48 KiB per k-block) -- under the
// 1 kW power cap that is what separates this GEMM from the cuBLAS number.  The leader CTA (cluster rank 0) issues
// every MMA; TMA loads of both CTAs complete on the leader's "full" barrier; tcgen05.commit multicasts the "stage
// free" / "accumulator ready" arrivals to both CTAs; the follower's epilogue warps release the accumulator on the
// leader's barrier through the cluster shared window.  Epilogues are the single-CTA ones (each CTA owns 128 rows).
namespace pair {
constexpr int STAGES2 = 5;
constexpr int HALF_B_BYTES = 128 * BK * 2, STAGE2_BYTES = A_BYTES + HALF_B_BYTES;
constexpr int OFF_STG2 = STAGES2 * STAGE2_BYTES, OFF_BAR2 = OFF_STG2 + EPI_WARPS * STG_BYTES;
constexpr int OFF_LN2 = OFF_BAR2 + 256;                    // LNT kernels: gamma | beta (2 x 768 fp32)
constexpr int SMEM2_BYTES = OFF_BAR2 + 256 + 1024;
constexpr int SMEM2_LN_BYTES = OFF_LN2 + 2 * 768 * 4 + 1024;
static_assert(SMEM2_LN_BYTES <= 227 * 1024, "shared memory budget");
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-pair bit of a shared::cluster address -> even CTA

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_slot, uint32_t ncols) {  // whole warp, both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA tile load into THIS CTA's shared memory whose bytes are accounted on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once all previously issued MMAs of the pair have completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_on_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}

constexpr int LN_WARPS = 4, LN_ROW = 768;

// One row of the residual stream -> LayerNorm -> 16-bit operand row.  The arithmetic is layernorm_kernel's (zk_ops.cu),
// operation for operation, so the fused and the separate path give bit-identical results; x comes from L2 (ld.global.cg:
// the TMA reduce-adds that produced it never touched this SM's L1), gamma / beta from shared memory.
template <int FMT>
__device__ __forceinline__ void ln_tail_rows2(const float* __restrict__ x, const float4* __restrict__ gw, const float4* __restrict__ gb,
                                              float eps, uint16_t* __restrict__ out, long long row_a, long long row_b, bool has_b,
                                              int lane) {
  // two rows in flight per warp: all twelve 16-byte loads are issued before the first use
  const float4* xa = reinterpret_cast<const float4*>(x + row_a * LN_ROW);
  const float4* xb = reinterpret_cast<const float4*>(x + (has_b ? row_b : row_a) * LN_ROW);
  float4 va[6], vb[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) va[i] = __ldcg(xa + lane + 32 * i);
#pragma unroll
  for (int i = 0; i < 6; ++i) vb[i] = __ldcg(xb + lane + 32 * i);
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    if (which == 1 && !has_b) break;
    float4* v = which ? vb : va;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / LN_ROW);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      v[i].x -= mean;
      v[i].y -= mean;
      v[i].z -= mean;
      v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / LN_ROW) + eps);
    uint2* orow = reinterpret_cast<uint2*>(out + (which ? row_b : row_a) * LN_ROW);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float4 g = gw[lane + 32 * i];
      const float4 bb = gb[lane + 32 * i];
      const float y0 = v[i].x * rstd * g.x + bb.x, y1 = v[i].y * rstd * g.y + bb.y;
      const float y2 = v[i].z * rstd * g.z + bb.z, y3 = v[i].w * rstd * g.w + bb.w;
      uint2 o;
      o.x = pack16<FMT>(y0, y1);
      o.y = pack16<FMT>(y2, y3);
      orow[lane + 32 * i] = o;
    }
  }
}

template <int EPI, int FMT, bool SEG, bool LNT = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS + (LNT ? LN_WARPS * 32 : 0), 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const Params p) {
  static_assert(EPI != ZK_EPI_PATCH_F32, "the patch-embedding epilogue stays on the single-CTA kernel");
  static_assert(!SEG || EPI == ZK_EPI_BIAS_RESID_F32, "segments are joined by the reduce-add of the residual epilogue");
  static_assert(!LNT || (EPI == ZK_EPI_BIAS_RESID_F32 && !SEG), "the LayerNorm tail follows the one-product residual epilogue");
  constexpr uint32_t IDESC2 = umma_idesc_16(FMT, 2 * BM, BN, 0, 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR2);  // leader's copy is the live one
  uint64_t* empty = full + STAGES2;                                // one per CTA
  uint64_t* tfull = empty + STAGES2;                               // one per CTA
  uint64_t* tempty = tfull + 2;                                    // leader's copy is the live one (2 x 8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float4* ln_gw = reinterpret_cast<float4*>(smem + OFF_LN2);        // LNT: gamma | beta, 2 x 768 floats
  float4* ln_gb = ln_gw + LN_ROW / 4;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  if (LNT) {
    for (int i = threadIdx.x; i < LN_ROW / 4; i += blockDim.x) {
      ln_gw[i] = __ldg(reinterpret_cast<const float4*>(p.ln_w) + i);
      ln_gb[i] = __ldg(reinterpret_cast<const float4*>(p.ln_b) + i);
    }
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int i = 0; i < STAGES2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // both CTAs' barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;  // num_m_tiles counts 256-row tiles here
  const int kblocks = p.K / BK;
  const int nseg = SEG ? kblocks / p.seg_kb : 1;
  const int kb_per_seg = SEG ? p.seg_kb * p.nprod : kblocks * p.nprod;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_blk = tile / p.num_n_tiles, n_blk = tile - m_blk * p.num_n_tiles;
        for (int seg = 0; seg < nseg; ++seg) {
          const int kb0 = SEG ? seg * p.seg_kb : 0, kb1 = SEG ? kb0 + p.seg_kb : kblocks;
          for (int pr = 0; pr < p.nprod; ++pr) {
            const int a0 = p.a_off[pr], w0 = p.w_off[pr];
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait(&empty[stage], phase ^ 1);
              if (leader) mbar_arrive_expect_tx(&full[stage], 2 * STAGE2_BYTES);  // both CTAs' tiles land on this barrier
              uint8_t* sa = smem + stage * STAGE2_BYTES;
              tma_load_2d_pair(sa, &tmA, &full[stage], a0 + kb * BK, m_blk * 2 * BM + (int)rank * BM);
              tma_load_2d_pair(sa + A_BYTES, &tmB, &full[stage], w0 + kb * BK, n_blk * BN + (int)rank * 128);
              if (++stage == STAGES2) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int t = 0;  // accumulator uses so far: one per tile, or one per segment of a tile
      for (int u = cluster_id * nseg, uend = num_tiles * nseg; u < uend; u += ((u + 1) % nseg ? 1 : (num_clusters - 1) * nseg + 1), ++t) {
        const int acc = t & 1;
        const uint32_t acc_phase = (t >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kb_per_seg; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * STAGE2_BYTES);
            const uint64_t a_desc = umma_desc_sw128(sa, 16, 1024);
            const uint64_t b_desc = umma_desc_sw128(sa + A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_bf16_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, IDESC2, (kb | k) != 0);
            umma_commit_pair(&empty[stage]);
            if (kb == kb_per_seg - 1) umma_commit_pair(&tfull[acc]);
          }
          __syncwarp();
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp < 2 + EPI_WARPS) {
    const int quarter = warp & 3;         // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;     // which 128 accumulator columns
    uint8_t* stg = smem + OFF_STG2 + (warp - 2) * STG_BYTES;  // this warp's 32 x 128 B staging tile (1024-B aligned)
    const uint32_t stg_row = smem_u32(stg) + lane * 128;
    int t = 0;
    for (int u = cluster_id * nseg, uend = num_tiles * nseg; u < uend; u += ((u + 1) % nseg ? 1 : (num_clusters - 1) * nseg + 1), ++t) {
      const int tile = u / nseg, seg = u - tile * nseg;
      const int m_blk = tile / p.num_n_tiles, n_blk = tile - m_blk * p.num_n_tiles;
      const int acc = t & 1;
      const uint32_t acc_phase = (t >> 1) & 1;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row0 = m_blk * 2 * BM + (int)rank * BM + quarter * 32;
      const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + half * 128;
      const int col_base = n_blk * BN + half * 128;
      epilogue_rows<EPI, FMT>(p, &tmC, stg, stg_row, t_acc, row0, col_base, lane, seg);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_on_leader(&tempty[acc]);
      if (LNT && lane == 0) {
        // the accumulator is already released; now wait until this warp's reduce-adds of the tile have LANDED (not
        // just been read out of shared memory) and count the warp in for its 128 rows of the 256-row tile
        bulk_wait0();
        asm volatile("fence.proxy.async;" ::: "memory");  // async-proxy (TMA) writes before the generic-proxy flag
        __threadfence();
        atomicAdd(p.ln_count + (m_blk * 2 + (int)rank), 1);
      }
    }
    if (lane == 0) bulk_wait0();
  } else if (LNT) {
    // ------------------------------------------------------------------ LayerNorm tail (4 warps)
    // The CTA that runs the last column tile of a 256-row tile normalises its own 128 rows once all 3 x 8 epilogue
    // warps that add into them (this CTA's and, for the other column tiles, the same-rank CTAs of other clusters,
    // which are resident and make progress on their own) have been counted in.
    const int lw = warp - (2 + EPI_WARPS);
    const int target = p.num_n_tiles * EPI_WARPS;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m_blk = tile / p.num_n_tiles, n_blk = tile - m_blk * p.num_n_tiles;
      if (n_blk != p.num_n_tiles - 1) continue;
      if (lane == 0) {
        const int* cnt = p.ln_count + (m_blk * 2 + (int)rank);
        int seen;
        for (;;) {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
          if (seen >= target) break;
          __nanosleep(256);
        }
      }
      __syncwarp();
      const long long row0 = (long long)m_blk * 2 * BM + (long long)rank * BM;
      for (int r = 2 * lw; r < BM; r += 2 * LN_WARPS) {
        const long long ra = row0 + r, rb = ra + 1;
        if (ra >= p.M) break;
        ln_tail_rows2<FMT>(reinterpret_cast<const float*>(p.out), ln_gw, ln_gb, p.ln_eps, p.ln_out, ra, rb, rb < p.M, lane);
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // neither CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

template <int EPI, int FMT, bool SEG = false, bool LNT = false>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const Params& p, int prof_cls,
                       cudaStream_t stream) {
  constexpr int NTHREADS = THREADS + (LNT ? LN_WARPS * 32 : 0);
  constexpr int NSMEM = LNT ? SMEM2_LN_BYTES : SMEM2_BYTES;
  static unsigned long long attr_done = 0;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(gemm_pair_kernel<EPI, FMT, SEG, LNT>), NSMEM, &attr_done)) return rc;
  // persistent grid = the number of CTA pairs the device can hold at once (a TPC with one usable SM cannot host a
  // pair, so this may be less than num_sms / 2); asked from the runtime once per device
  static int resident[64] = {0};
  int dev = 0;
  ZK_CUDA(cudaGetDevice(&dev));
  int cap = __atomic_load_n(&resident[dev & 63], __ATOMIC_ACQUIRE);
  if (cap == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (num_sms() / 2));
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = NSMEM;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2;
    at.val.clusterDim.y = 1;
    at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_pair_kernel<EPI, FMT, SEG, LNT>, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = num_sms() / 2;
    }
    cap = n < num_sms() / 2 ? n : num_sms() / 2;
    __atomic_store_n(&resident[dev & 63], cap, __ATOMIC_RELEASE);
    if (getenv("ZK_DEBUG")) fprintf(stderr, "zk: %d resident CTA pairs for gemm_pair_kernel<%d, %d>\n", cap, EPI, FMT);
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int clusters = tiles < cap ? tiles : cap;
  ProfScope prof(prof_cls, stream);
  if (LNT) ZK_CUDA(cudaMemsetAsync(p.ln_count, 0, (size_t)p.num_m_tiles * 2 * sizeof(int), stream));
  gemm_pair_kernel<EPI, FMT, SEG, LNT><<<2 * clusters, NTHREADS, NSMEM, stream>>>(tmA, tmB, tmC, p);
  ZK_LAUNCH_CHECK("gemm_pair_kernel");
  return 0;
}
}  // namespace pair

template <int EPI, int FMT, bool SEG = false>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const Params& p, int prof_cls,
                  cudaStream_t stream) {
  static unsigned long long attr_done = 0;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(gemm_kernel<EPI, FMT, SEG>), SMEM_BYTES, &attr_done)) return rc;
  int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = tiles < num_sms() ? tiles : num_sms();
  ProfScope prof(prof_cls, stream);
  gemm_kernel<EPI, FMT, SEG><<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, tmC, p);
  ZK_LAUNCH_CHECK("gemm_kernel");
  return 0;
}
}  // namespace gemm

bool gemm_ln_tail_ok(long long M) {
  static const int use_pair = getenv("ZK_GEMM_PAIR") ? atoi(getenv("ZK_GEMM_PAIR")) : 2;
  return use_pair && M >= 4 * gemm::BM;
}

int gemm16(const GemmArgs& g, cudaStream_t stream) {
  using namespace gemm;
  int rc = device_check();
  if (rc) return rc;
  const long long M = g.M;
  const int N = g.N, K = g.K, epilogue = g.epilogue;
  if (!g.a || !g.w || !g.bias || !g.out || M <= 0) {
    set_error("gemm16: null pointer or M <= 0");
    return ZK_ERR_ARG;
  }
  if (N % BN || K % BK || N <= 0 || K <= 0) {
    set_error("gemm16: N (%d) must be a multiple of %d and K (%d) of %d", N, BN, K, BK);
    return ZK_ERR_SHAPE;
  }
  if (epilogue < 0 || epilogue > ZK_EPI_BIAS_GELU_SPLIT) {
    set_error("gemm16: unknown epilogue %d", epilogue);
    return ZK_ERR_ARG;
  }
  if (g.fmt != FMT_BF16 && g.fmt != FMT_F16) {
    set_error("gemm16: unknown operand format %d", g.fmt);
    return ZK_ERR_ARG;
  }
  if (g.products != 1 && g.products != 3) {
    set_error("gemm16: products must be 1 or 3 (got %d)", g.products);
    return ZK_ERR_ARG;
  }
  const bool split_out = epi_is_split(epilogue);
  if ((g.products == 3 || split_out) && g.fmt != FMT_F16) {
    set_error("gemm16: split operands / split outputs are fp16 planes (operand format %d given)", g.fmt);
    return ZK_ERR_ARG;
  }
  if (epilogue == ZK_EPI_PATCH_F32 && (!g.aux || g.aux_rows <= 0)) {
    set_error("gemm16: ZK_EPI_PATCH_F32 needs the position table and patches per window");
    return ZK_ERR_ARG;
  }
  const int planes_in = g.products == 3 ? 2 : 1;
  const long long lda = g.lda > 0 ? g.lda : (long long)planes_in * K;
  const long long ldw = g.ldw > 0 ? g.ldw : (long long)planes_in * K;
  const long long out_cols = split_out ? 2LL * N : N;
  const long long ldo = g.ldo > 0 ? g.ldo : out_cols;
  if (lda < (long long)planes_in * K || ldw < (long long)planes_in * K || ldo < out_cols ||
      (epilogue == ZK_EPI_PATCH_F32 && ldo != N)) {
    set_error("gemm16: row pitches (a %lld, w %lld, out %lld) too small for K %d, N %d, %d operand plane(s)", lda, ldw, ldo, K,
              N, planes_in);
    return ZK_ERR_SHAPE;
  }
  const int aux_rows = g.aux_rows;
  CUtensorMap tmA, tmB, tmC;
  if ((rc = make_tmap_bf16_2d(&tmA, g.a, (uint64_t)M, (uint64_t)planes_in * K, (uint64_t)lda, BM, BK))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmB, g.w, (uint64_t)N, (uint64_t)planes_in * K, (uint64_t)ldw, BN, BK))) return rc;
  if (epilogue == ZK_EPI_BIAS_BF16 || epilogue == ZK_EPI_BIAS_GELU_BF16 || split_out) {
    if ((rc = make_tmap_bf16_2d(&tmC, g.out, (uint64_t)M, (uint64_t)out_cols, (uint64_t)ldo, 32, 64))) return rc;
  } else if (epilogue == ZK_EPI_BIAS_RESID_F32) {
    if ((rc = make_tmap_f32_2d(&tmC, g.out, (uint64_t)M, (uint64_t)N, (uint64_t)ldo, 32, 32))) return rc;
  } else {
    // patch embedding: x viewed as [windows][patches + 2 tokens][N]
    const uint64_t windows = (uint64_t)((M + aux_rows - 1) / aux_rows);
    if ((rc = make_tmap_f32_3d(&tmC, g.out, windows, (uint64_t)aux_rows + 2, (uint64_t)N, (uint64_t)N,
                               (uint64_t)(aux_rows + 2) * N, 32, 32)))
      return rc;
  }
  Params p;
  p.bias = g.bias;
  p.out = g.out;
  p.aux = g.aux;
  p.M = M;
  p.N = N;
  p.K = K;
  p.aux_rows = aux_rows;
  p.num_m_tiles = (int)((M + BM - 1) / BM);
  p.num_n_tiles = N / BN;
  p.acc_scale = g.acc_scale;
  p.nprod = g.products;
  p.out_plane_stride = N;
  p.ln_w = g.ln_w, p.ln_b = g.ln_b, p.ln_out = reinterpret_cast<uint16_t*>(g.ln_out), p.ln_count = g.ln_count, p.ln_eps = g.ln_eps;
  if (g.products == 3) {  // smallest terms first: A_lo W_hi, A_hi W_lo, A_hi W_hi
    p.a_off[0] = K, p.w_off[0] = 0;
    p.a_off[1] = 0, p.w_off[1] = K;
    p.a_off[2] = 0, p.w_off[2] = 0;
  } else {
    for (int i = 0; i < MAX_PRODUCTS; ++i) p.a_off[i] = p.w_off[i] = 0;
  }
  const int prof_cls = g.prof_cls;
  const auto cls = [&](int by_epilogue) { return prof_cls >= 0 ? prof_cls : by_epilogue; };
  const int cls_resid = K > 768 ? ZK_K_GEMM_FC2 : ZK_K_GEMM_OUT;
  // CTA-pair tiles for the large GEMMs (ZK_GEMM_PAIR: 0 = never, 1 = all but fc1, 2 = all, the default).  With the
  // two-MUFU GELU the fc1 epilogue paced its tile and pair tiles did not pay; with the one-MUFU form they do
  // (0.586 against 0.606 ms at M = 155 392 on the same box).
  static const int use_pair = getenv("ZK_GEMM_PAIR") ? atoi(getenv("ZK_GEMM_PAIR")) : 2;
  const bool f16 = g.fmt == FMT_F16;
  // split-operand residual GEMMs: cut the K chain into segments joined by rounded adds (see SEG_KB)
  p.seg_kb = K >= SEG_LONG_K ? SEG_KB_LONG : SEG_KB;
  const bool seg = g.products == 3 && epilogue == ZK_EPI_BIAS_RESID_F32 && K % (BK * p.seg_kb) == 0 && K > BK * p.seg_kb;
  const bool pair_ok = use_pair && epilogue != ZK_EPI_PATCH_F32 && M >= 4 * BM && (epilogue != ZK_EPI_BIAS_GELU_BF16 || use_pair >= 2);
  // LayerNorm tail asked for: only the one-product residual GEMM over full 768-wide rows on CTA-pair tiles has it
  const bool ln_tail = g.ln_out != nullptr;
  if (ln_tail && !(pair_ok && epilogue == ZK_EPI_BIAS_RESID_F32 && !seg && g.products == 1 && N == 768 && ldo == N && g.ln_w &&
                   g.ln_b && g.ln_count)) {
    set_error("gemm16: the LayerNorm tail needs the CTA-pair residual GEMM with N = 768, one product and M >= %d", 4 * BM);
    return ZK_ERR_ARG;
  }
  if (pair_ok) {
    if ((rc = make_tmap_bf16_2d(&tmB, g.w, (uint64_t)N, (uint64_t)planes_in * K, (uint64_t)ldw, 128, BK))) return rc;  // half W tiles
    p.num_m_tiles = (int)((M + 2 * BM - 1) / (2 * BM));
    switch (epilogue) {
      case ZK_EPI_BIAS_BF16:
        return f16 ? pair::launch_pair<ZK_EPI_BIAS_BF16, FMT_F16>(tmA, tmB, tmC, p, cls(ZK_K_GEMM_QKV), stream)
                   : pair::launch_pair<ZK_EPI_BIAS_BF16, FMT_BF16>(tmA, tmB, tmC, p, cls(ZK_K_GEMM_QKV), stream);
      case ZK_EPI_BIAS_GELU_BF16:
        return f16 ? pair::launch_pair<ZK_EPI_BIAS_GELU_BF16, FMT_F16>(tmA, tmB, tmC, p, cls(ZK_K_GEMM_FC1), stream)
                   : pair::launch_pair<ZK_EPI_BIAS_GELU_BF16, FMT_BF16>(tmA, tmB, tmC, p, cls(ZK_K_GEMM_FC1), stream);
      case ZK_EPI_BIAS_RESID_F32:
        if (seg) return pair::launch_pair<ZK_EPI_BIAS_RESID_F32, FMT_F16, true>(tmA, tmB, tmC, p, cls(cls_resid), stream);
        if (ln_tail)
          return f16 ? pair::launch_pair<ZK_EPI_BIAS_RESID_F32, FMT_F16, false, true>(tmA, tmB, tmC, p, cls(cls_resid), stream)
                     : pair::launch_pair<ZK_EPI_BIAS_RESID_F32, FMT_BF16, false, true>(tmA, tmB, tmC, p, cls(cls_resid), stream);
        return f16 ? pair::launch_pair<ZK_EPI_BIAS_RESID_F32, FMT_F16>(tmA, tmB, tmC, p, cls(cls_resid), stream)
                   : pair::launch_pair<ZK_EPI_BIAS_RESID_F32, FMT_BF16>(tmA, tmB, tmC, p, cls(cls_resid), stream);
      case ZK_EPI_BIAS_SPLIT: return pair::launch_pair<ZK_EPI_BIAS_SPLIT, FMT_F16>(tmA, tmB, tmC, p, cls(ZK_K_RECHECK), stream);
      case ZK_EPI_BIAS_GELU_SPLIT:
        return pair::launch_pair<ZK_EPI_BIAS_GELU_SPLIT, FMT_F16>(tmA, tmB, tmC, p, cls(ZK_K_RECHECK), stream);
    }
  }
  switch (epilogue) {
    case ZK_EPI_BIAS_BF16:
      return f16 ? launch<ZK_EPI_BIAS_BF16, FMT_F16>(tmA, tmB, tmC, p, cls(ZK_K_GEMM_QKV), stream)
                 : launch<ZK_EPI_BIAS_BF16, FMT_BF16>(tmA, tmB, tmC, p, cls(ZK_K_GEMM_QKV), stream);
    case ZK_EPI_BIAS_GELU_BF16:
      return f16 ? launch<ZK_EPI_BIAS_GELU_BF16, FMT_F16>(tmA, tmB, tmC, p, cls(ZK_K_GEMM_FC1), stream)
                 : launch<ZK_EPI_BIAS_GELU_BF16, FMT_BF16>(tmA, tmB, tmC, p, cls(ZK_K_GEMM_FC1), stream);
    case ZK_EPI_BIAS_RESID_F32:
      if (seg) return launch<ZK_EPI_BIAS_RESID_F32, FMT_F16, true>(tmA, tmB, tmC, p, cls(cls_resid), stream);
      return f16 ? launch<ZK_EPI_BIAS_RESID_F32, FMT_F16>(tmA, tmB, tmC, p, cls(cls_resid), stream)
                 : launch<ZK_EPI_BIAS_RESID_F32, FMT_BF16>(tmA, tmB, tmC, p, cls(cls_resid), stream);
    case ZK_EPI_PATCH_F32:
      return f16 ? launch<ZK_EPI_PATCH_F32, FMT_F16>(tmA, tmB, tmC, p, cls(ZK_K_GEMM_PATCH), stream)
                 : launch<ZK_EPI_PATCH_F32, FMT_BF16>(tmA, tmB, tmC, p, cls(ZK_K_GEMM_PATCH), stream);
    case ZK_EPI_BIAS_SPLIT: return launch<ZK_EPI_BIAS_SPLIT, FMT_F16>(tmA, tmB, tmC, p, cls(ZK_K_RECHECK), stream);
    case ZK_EPI_BIAS_GELU_SPLIT: return launch<ZK_EPI_BIAS_GELU_SPLIT, FMT_F16>(tmA, tmB, tmC, p, cls(ZK_K_RECHECK), stream);
  }
  set_error("gemm16: unknown epilogue %d", epilogue);
  return ZK_ERR_ARG;
}

}  // namespace zk

extern "C" int zk_gemm16(const void* d_a, int64_t lda, const void* d_w, int64_t ldw, const float* d_bias, void* d_out,
                         int64_t ldo, int64_t M, int N, int K, int epilogue, int operand_format, int products,
                         float acc_scale, const float* d_aux, int aux_rows, zk_stream_t stream) {
  zk::GemmArgs g;
  g.a = d_a, g.lda = lda, g.w = d_w, g.ldw = ldw, g.bias = d_bias, g.out = d_out, g.ldo = ldo;
  g.M = M, g.N = N, g.K = K, g.epilogue = epilogue, g.fmt = operand_format, g.products = products;
  g.acc_scale = acc_scale, g.aux = d_aux, g.aux_rows = aux_rows, g.prof_cls = -1;
  return zk::gemm16(g, (cudaStream_t)stream);
}

extern "C" int zk_gemm_bf16(const void* d_a, const void* d_w, const float* d_bias, void* d_out, int64_t M, int N, int K,
                            int epilogue, const float* d_aux, int aux_rows, zk_stream_t stream) {
  return zk_gemm16(d_a, 0, d_w, 0, d_bias, d_out, 0, M, N, K, epilogue, ZK_FMT_BF16, 1, 1.0f, d_aux, aux_rows, stream);
}
