// Attention at re-check precision: O = softmax(Q K^T / 8) V per (window, head) with every operand carried as two
// fp16 planes (x = hi + lo, 22 significant bits) and both contractions evaluated as three-product sums
//     Q K^T = Q_lo K_hi^T + Q_hi K_lo^T + Q_hi K_hi^T          P V = P_lo V_hi + P_hi V_lo + P_hi V_hi
// with fp32 accumulation (short tensor-core chains joined by rounded fp32 adds, because the tensor core truncates), an
// fp32 online softmax against the TRUE running maximum and exp2f (no polynomial, no stale maximum): the result is
// within a few 2^-22 of an fp32 evaluation (HF:modeling_audio_spectrogram_transformer.py:
// 162-176 on the CPU), which is what the decision re-check needs.
//
// This kernel only ever sees the few windows whose fast logits are within eps of a threshold, so it is built for
// exactness and robustness rather than peak rate: register-resident FlashAttention-2 dataflow on warp-level
// mma.sync.m16n8k16 (the probabilities never leave registers between the two contractions, so P is split into hi / lo
// in place), K / V planes staged through a cp.async double buffer, ldmatrix fragments.  One CTA = 64 queries of one
// (window, head), 4 warps x 16 rows; key blocks of 64.  The throughput path is zk_attn.cu (tcgen05 / TMEM).
//
// Layouts: qkv fp16 [batch*tokens][2*2304] = hi plane (q | k | v, head h at columns 64 h) followed by the lo plane;
//          out fp16 [batch*tokens][2*768]  = hi | lo planes of the attention output.
#include <stdlib.h>
#include <string.h>

#include "zk_b200.h"
#include "zk_common.cuh"
#include "zk_internal.cuh"

namespace zk {
namespace attn_split {

constexpr int D = 64, HEADS = 12, HID = 768, QKV = 3 * HID;
constexpr int LDQ = 2 * QKV, LDO = 2 * HID;      // row pitches (halves) of the two-plane buffers
constexpr int BQ = 64, BKV = 64, THREADS = 128;
constexpr int ROW_H = D + 8;                     // smem row pitch in halves (144 B): conflict-free ldmatrix
constexpr int TILE_H = BKV * ROW_H;              // one 64 x 64 plane tile
constexpr int STAGE_H = 4 * TILE_H;              // K_hi, K_lo, V_hi, V_lo
constexpr int SMEM_BYTES = 2 * STAGE_H * 2;      // double buffered: 73 728 B
constexpr float SCALE_LOG2E = 0.125f * 1.44269504088896340736f;
constexpr float P_SHIFT = 14.0f;  // P is carried as p * 2^14 (<= 16384) so that its lo plane stays in fp16's normal range

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int bytes = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// D (16x8, fp32) += A (16x16 fp16, row) * B (16x8 fp16, col)
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(THREADS, 3) attn_split_kernel(const __half* __restrict__ qkv, __half* __restrict__ out,
                                                             int tokens) {
  extern __shared__ __align__(16) __half sm[];
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const __half* base = qkv + (long long)b * tokens * LDQ;
  const int nkv = (tokens + BKV - 1) / BKV;

  // K / V planes of key block j -> stage (j & 1): 4 tiles x 64 rows x 8 chunks of 16 B = 2048 chunks, 16 per thread
  auto load_stage = [&](int j) {
    __half* st = sm + (j & 1) * STAGE_H;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = tid + i * THREADS;
      const int tile = c >> 9, row = (c >> 3) & 63, chunk = c & 7;  // tile: 0 K_hi, 1 K_lo, 2 V_hi, 3 V_lo
      const int key = j * BKV + row;
      const bool valid = key < tokens;
      const int col = (tile & 1) * QKV + (tile >> 1 ? 2 * HID : HID) + h * D + chunk * 8;
      const __half* src = base + (long long)(valid ? key : 0) * LDQ + col;
      cp_async16(smem_u32(st + tile * TILE_H + row * ROW_H + chunk * 8), src, valid);
    }
    cp_async_commit();
  };
  load_stage(0);

  // Q fragments (A operand, m16k16 row-major): rows g and g + 8 of this warp's 16 queries, both planes, 4 k-steps
  const int row_a = qt * BQ + warp * 16 + g, row_b = row_a + 8;
  const int ra = row_a < tokens ? row_a : tokens - 1, rb = row_b < tokens ? row_b : tokens - 1;
  uint32_t qf[2][4][4];
#pragma unroll
  for (int pl = 0; pl < 2; ++pl)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int col = pl * QKV + h * D + kk * 16 + 2 * t;
      qf[pl][kk][0] = __ldg(reinterpret_cast<const uint32_t*>(base + (long long)ra * LDQ + col));
      qf[pl][kk][1] = __ldg(reinterpret_cast<const uint32_t*>(base + (long long)rb * LDQ + col));
      qf[pl][kk][2] = __ldg(reinterpret_cast<const uint32_t*>(base + (long long)ra * LDQ + col + 8));
      qf[pl][kk][3] = __ldg(reinterpret_cast<const uint32_t*>(base + (long long)rb * LDQ + col + 8));
    }

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;  // running max / (per-thread partial) sum, rows g, g + 8

  // ldmatrix lane addressing (halves, relative to a tile):
  //   K (B operand of Q K^T, non-transposed): matrix i = lane / 8 -> key block (i >> 1), dim half (i & 1)
  //   V (B operand of P V, transposed):       matrix i = lane / 8 -> key half (i & 1), dim block (i >> 1)
  const int lm = lane >> 3, lr = lane & 7;
  const int k_off = ((lm >> 1) * 8 + lr) * ROW_H + (lm & 1) * 8;
  const int v_off = ((lm & 1) * 8 + lr) * ROW_H + (lm >> 1) * 8;

  for (int j = 0; j < nkv; ++j) {
    if (j + 1 < nkv) {
      load_stage(j + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __half* st = sm + (j & 1) * STAGE_H;
    const uint32_t k_hi = smem_u32(st), k_lo = smem_u32(st + TILE_H), v_hi = smem_u32(st + 2 * TILE_H),
                   v_lo = smem_u32(st + 3 * TILE_H);

    // ---- S = Q K^T (16 x 64 per warp).  The tensor core TRUNCATES when it adds into its accumulator (probed:
    // scripts/accum_probe.py), i.e. every mma step costs up to one ulp of the accumulator, always towards zero.  So the
    // large products (hi x hi) get a chain of their own that is as short as possible (4 steps) and the two small
    // products, 2^-11 of the result, another one; the chains are joined by an ordinary rounded fp32 add.
    float s[8][4];
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {  // pairs of 8-key n-tiles
      float h0[4] = {0.f, 0.f, 0.f, 0.f}, h1[4] = {0.f, 0.f, 0.f, 0.f}, l0[4] = {0.f, 0.f, 0.f, 0.f}, l1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t bh[4], bl[4];
        const uint32_t off = (uint32_t)((jp * 16 * ROW_H + kk * 16 + k_off) * 2);
        ldsm_x4(k_hi + off, bh);
        ldsm_x4(k_lo + off, bl);
        mma16816(l0, qf[1][kk], bh[0], bh[1]);      // Q_lo K_hi
        mma16816(l0, qf[0][kk], bl[0], bl[1]);      // Q_hi K_lo
        mma16816(h0, qf[0][kk], bh[0], bh[1]);      // Q_hi K_hi
        mma16816(l1, qf[1][kk], bh[2], bh[3]);
        mma16816(l1, qf[0][kk], bl[2], bl[3]);
        mma16816(h1, qf[0][kk], bh[2], bh[3]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s[2 * jp][e] = h0[e] + l0[e];
        s[2 * jp + 1][e] = h1[e] + l1[e];
      }
    }

    // ---- online softmax in the log2 domain (true running maximum)
    const int kbase = j * BKV + 2 * t;
    float mx_a = -INFINITY, mx_b = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kbase + i * 8 + (e & 1);
        s[i][e] = key < tokens ? s[i][e] * SCALE_LOG2E : -INFINITY;
      }
      mx_a = fmaxf(mx_a, fmaxf(s[i][0], s[i][1]));
      mx_b = fmaxf(mx_b, fmaxf(s[i][2], s[i][3]));
    }
    mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
    mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
    mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
    mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
    const float mn_a = fmaxf(m_a, mx_a), mn_b = fmaxf(m_b, mx_b);  // finite: every key block holds at least one valid key
    const float al_a = exp2f(m_a - mn_a), al_b = exp2f(m_b - mn_b);  // exp2f(-inf) = 0 on the first block
    m_a = mn_a;
    m_b = mn_b;
    l_a *= al_a;
    l_b *= al_b;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i][0] *= al_a;
      o[i][1] *= al_a;
      o[i][2] *= al_b;
      o[i][3] *= al_b;
    }
    const float sh_a = P_SHIFT - m_a, sh_b = P_SHIFT - m_b;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = exp2f(s[i][0] + sh_a);
      s[i][1] = exp2f(s[i][1] + sh_a);
      s[i][2] = exp2f(s[i][2] + sh_b);
      s[i][3] = exp2f(s[i][3] + sh_b);
      l_a += s[i][0] + s[i][1];
      l_b += s[i][2] + s[i][3];
    }

    // ---- O += P V: the accumulator layout of two adjacent n-tiles IS the A-fragment layout of one 16-key k-step.
    // Per key block the product is formed in fresh short chains (see above) and added to the running O with a rounded
    // fp32 add: accumulating all 19 blocks inside the tensor core would be a 228-step truncating chain (~1e-5).
    uint32_t ph[4][4], pl[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      split_f16_pair(s[2 * kk][0], s[2 * kk][1], ph[kk][0], pl[kk][0]);          // row g,     keys 16 kk + 2t, +1
      split_f16_pair(s[2 * kk][2], s[2 * kk][3], ph[kk][1], pl[kk][1]);          // row g + 8
      split_f16_pair(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[kk][2], pl[kk][2]);  // row g,     keys 16 kk + 8 + 2t, +1
      split_f16_pair(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[kk][3], pl[kk][3]);  // row g + 8
    }
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {  // pairs of 8-dim n-tiles
      float h0[4] = {0.f, 0.f, 0.f, 0.f}, h1[4] = {0.f, 0.f, 0.f, 0.f}, l0[4] = {0.f, 0.f, 0.f, 0.f}, l1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t bh[4], bl[4];
        const uint32_t off = (uint32_t)((kk * 16 * ROW_H + jp * 16 + v_off) * 2);
        ldsm_x4_trans(v_hi + off, bh);
        ldsm_x4_trans(v_lo + off, bl);
        mma16816(l0, pl[kk], bh[0], bh[1]);      // P_lo V_hi
        mma16816(l0, ph[kk], bl[0], bl[1]);      // P_hi V_lo
        mma16816(h0, ph[kk], bh[0], bh[1]);      // P_hi V_hi
        mma16816(l1, pl[kk], bh[2], bh[3]);
        mma16816(l1, ph[kk], bl[2], bl[3]);
        mma16816(h1, ph[kk], bh[2], bh[3]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[2 * jp][e] += h0[e] + l0[e];
        o[2 * jp + 1][e] += h1[e] + l1[e];
      }
    }
    __syncthreads();  // everyone is done with stage (j & 1) before the next iteration's prefetch overwrites it
  }

  // ---- epilogue: O / l (the 2^14 of P cancels), split into hi | lo planes
  l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
  l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
  l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
  l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
  const float inv_a = 1.0f / l_a, inv_b = 1.0f / l_b;
  __half* ob = out + (long long)b * tokens * LDO + h * D + 2 * t;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint32_t hi, lo;
    if (row_a < tokens) {
      split_f16_pair(o[i][0] * inv_a, o[i][1] * inv_a, hi, lo);
      *reinterpret_cast<uint32_t*>(ob + (long long)row_a * LDO + i * 8) = hi;
      *reinterpret_cast<uint32_t*>(ob + (long long)row_a * LDO + HID + i * 8) = lo;
    }
    if (row_b < tokens) {
      split_f16_pair(o[i][2] * inv_b, o[i][3] * inv_b, hi, lo);
      *reinterpret_cast<uint32_t*>(ob + (long long)row_b * LDO + i * 8) = hi;
      *reinterpret_cast<uint32_t*>(ob + (long long)row_b * LDO + HID + i * 8) = lo;
    }
  }
}

// ================================================================================================ tcgen05 variant
// The same arithmetic on the 5th-generation tensor cores (4x the rate of the warp-level mma.sync path above, which
// stays as the cross-check and as ZK_SPLIT_ATTN=mma).  One CTA = 128 queries of one (window, head); two CTAs per SM
// (96 KiB of shared memory, 256 TMEM columns each) so that one CTA's tensor work runs under the other's softmax.
//
//   warp 0 (lane 0)  TMA producer: Q_hi | Q_lo once, then {K_hi, K_lo} and {V_hi, V_lo} of every 64-key block (single
//                    stage: K is free as soon as the block's score MMAs have run, V once its P V MMAs have)
//   warp 1           MMA issuer (elect-one): per key block j
//                       S_hi = Q_hi K_hi^T            (4 x UMMA 128x64x16, fresh accumulator: a 4-step chain)
//                       S_lo = Q_lo K_hi^T + Q_hi K_lo^T   (8 steps on a 2^-11 accumulator)
//                       O_hi = P_hi V_hi,  O_lo = P_lo V_hi + P_hi V_lo   (same, once the softmax warps published P)
//   warps 2..5       one query row per thread: S_hi + S_lo out of TMEM, rounded fp32 add, true running maximum, exp2f,
//                    P' = p 2^14 split into fp16 hi | lo and written to shared memory in the 128B-swizzled K-major
//                    layout the tensor core reads; the block's O_hi + O_lo is pulled out of TMEM and added to the
//                    running O IN REGISTERS (rounded), so no accumulator chain is longer than 4 large steps.
namespace tc {
constexpr int BQ = 128, BKV = 64;
constexpr int THREADS = 192;
constexpr int Q_TILE = BQ * D * 2, KV_TILE = BKV * D * 2, P_TILE = BQ * BKV * 2;  // bytes: 16 KiB, 8 KiB, 16 KiB
constexpr int OFF_Q = 0;                       // Q_hi, Q_lo
constexpr int OFF_K = OFF_Q + 2 * Q_TILE;      // K_hi, K_lo
constexpr int OFF_V = OFF_K + 2 * KV_TILE;     // V_hi, V_lo
constexpr int OFF_P = OFF_V + 2 * KV_TILE;     // P_hi, P_lo
constexpr int OFF_BAR = OFF_P + 2 * P_TILE;
constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;  // + barriers + alignment slack (dynamic smem is only 16-byte aligned)
static_assert(2 * SMEM_BYTES <= 227 * 1024, "two CTAs per SM");
constexpr uint32_t TM_COLS = 256, TM_SHI = 0, TM_SLO = 64, TM_OHI = 128, TM_OLO = 192;
constexpr uint32_t IDESC_S = umma_idesc_16(FMT_F16, BQ, BKV, 0, 0);
constexpr uint32_t IDESC_O = umma_idesc_16(FMT_F16, BQ, D, 0, 1);  // B (= V) is MN-major

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// v[c0 .. c0 + 16) (+)= hi + lo of 16 accumulator columns (two TMEM loads, one rounded add per element)
template <bool ADD>
__device__ __forceinline__ void pull16(uint32_t t_hi, uint32_t t_lo, float* v) {
  uint32_t a[16], b[16];
  tmem_ld16(t_hi, a);
  tmem_ld16(t_lo, b);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float x = __uint_as_float(a[i]) + __uint_as_float(b[i]);
    v[i] = ADD ? v[i] + x : x;
  }
}

__global__ void __launch_bounds__(THREADS, 2)
attn_split_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     __half* __restrict__ out, int tokens) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t *q_full = bars, *k_full = bars + 1, *v_full = bars + 2, *s_full = bars + 3, *s_free = bars + 4,
           *p_full = bars + 5, *o_full = bars + 6, *o_free = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkv = (tokens + BKV - 1) / BKV;
  const int row_base = b * tokens;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(o_free, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      mbar_arrive_expect_tx(q_full, 2 * Q_TILE);
      tma_load_2d(smem + OFF_Q, &tm_q, q_full, h * D, row_base + qt * BQ);
      tma_load_2d(smem + OFF_Q + Q_TILE, &tm_q, q_full, QKV + h * D, row_base + qt * BQ);
      for (int j = 0; j < nkv; ++j) {
        if (j > 0) mbar_wait(s_full, (j - 1) & 1);  // the score MMAs of block j-1 have read K
        mbar_arrive_expect_tx(k_full, 2 * KV_TILE);
        tma_load_2d(smem + OFF_K, &tm_kv, k_full, HID + h * D, row_base + j * BKV);
        tma_load_2d(smem + OFF_K + KV_TILE, &tm_kv, k_full, QKV + HID + h * D, row_base + j * BKV);
        if (j > 0) mbar_wait(o_full, (j - 1) & 1);  // the P V MMAs of block j-1 have read V
        mbar_arrive_expect_tx(v_full, 2 * KV_TILE);
        tma_load_2d(smem + OFF_V, &tm_kv, v_full, 2 * HID + h * D, row_base + j * BKV);
        tma_load_2d(smem + OFF_V + KV_TILE, &tm_kv, v_full, QKV + 2 * HID + h * D, row_base + j * BKV);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp walks, one lane issues)
    const uint32_t q_hi = smem_u32(smem + OFF_Q), q_lo = q_hi + Q_TILE;
    const uint32_t k_hi = smem_u32(smem + OFF_K), k_lo = k_hi + KV_TILE;
    const uint32_t v_hi = smem_u32(smem + OFF_V), v_lo = v_hi + KV_TILE;
    const uint32_t p_hi = smem_u32(smem + OFF_P), p_lo = p_hi + P_TILE;
    // S(j+1) is issued as soon as the softmax warps have pulled S(j) out of TMEM, i.e. it runs underneath their
    // exponentials; P V of block j follows when they publish P(j)
    auto issue_s = [&]() {
      if (elect_one()) {
        const uint64_t dqh = umma_desc_sw128(q_hi, 16, 1024), dql = umma_desc_sw128(q_lo, 16, 1024);
        const uint64_t dkh = umma_desc_sw128(k_hi, 16, 1024), dkl = umma_desc_sw128(k_lo, 16, 1024);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_bf16_ss(tmem_base + TM_SLO, dql + 2 * k, dkh + 2 * k, IDESC_S, k != 0);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_bf16_ss(tmem_base + TM_SLO, dqh + 2 * k, dkl + 2 * k, IDESC_S, 1);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_bf16_ss(tmem_base + TM_SHI, dqh + 2 * k, dkh + 2 * k, IDESC_S, k != 0);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(k_full, 0);
    tc_fence_after();
    issue_s();
    for (int j = 0; j < nkv; ++j) {
      if (j + 1 < nkv) {
        mbar_wait(k_full, (j + 1) & 1);
        mbar_wait(s_free, j & 1);  // the softmax warps hold S(j) in registers
        tc_fence_after();
        issue_s();
      }
      mbar_wait(p_full, j & 1);  // P(j) is in shared memory (written through the generic proxy + fence.proxy.async)
      mbar_wait(v_full, j & 1);
      if (j > 0) mbar_wait(o_free, (j - 1) & 1);  // O_blk(j-1) has been added to the running O
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dph = umma_desc_sw128(p_hi, 16, 1024), dpl = umma_desc_sw128(p_lo, 16, 1024);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)  // A = P: 16 keys = 32 B along the 128-B row; B = V: 16 keys = 16 rows (MN-major)
          umma_bf16_ss(tmem_base + TM_OLO, dpl + 2 * k, umma_desc_sw128(v_hi + k * 16 * 128, 1024, 1024), IDESC_O, k != 0);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          umma_bf16_ss(tmem_base + TM_OLO, dph + 2 * k, umma_desc_sw128(v_lo + k * 16 * 128, 1024, 1024), IDESC_O, 1);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          umma_bf16_ss(tmem_base + TM_OHI, dph + 2 * k, umma_desc_sw128(v_hi + k * 16 * 128, 1024, 1024), IDESC_O, k != 0);
        umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax: one query row per thread
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t p_row_hi = smem_u32(smem + OFF_P) + r * 128, p_row_lo = p_row_hi + P_TILE;
    float o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nkv; ++j) {
      float sc[BKV];
      mbar_wait(s_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < BKV / 16; ++c) pull16<false>(t_lane + TM_SHI + c * 16, t_lane + TM_SLO + c * 16, sc + c * 16);
      tc_fence_before();
      mbar_arrive(s_free);
      const int kmax = tokens - j * BKV;  // keys [0, kmax) of this block exist
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < BKV; ++i) {
        sc[i] = i < kmax ? sc[i] * SCALE_LOG2E : -INFINITY;
        mx = fmaxf(mx, sc[i]);
      }
      const float mn = fmaxf(m, mx);        // finite: every key block holds at least one real key
      const float alpha = fast_exp2(m - mn);  // 0 on the first block (m = -inf)
      if (j > 0) {                          // O += O_blk(j-1), which was formed against the OLD maximum
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < D / 16; ++c) pull16<true>(t_lane + TM_OHI + c * 16, t_lane + TM_OLO + c * 16, o + c * 16);
        tc_fence_before();
        mbar_arrive(o_free);
      }
      m = mn;
      l *= alpha;
#pragma unroll
      for (int i = 0; i < D; ++i) o[i] *= alpha;
      // P' = exp2(s - m + 14) -> fp16 hi | lo -> 128B-swizzled K-major rows (chunk c of row r at chunk c ^ (r & 7));
      // the P buffers are free: P V of block j-1 completed (o_full above)
      const float sh = P_SHIFT - m;
#pragma unroll
      for (int c = 0; c < BKV / 8; ++c) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // ex2.approx: 2 ulp, like exp2f, without its subnormal-range fix-up (P' below 2^-126 is zero either way)
          const float p0 = fast_exp2(sc[c * 8 + 2 * e] + sh), p1 = fast_exp2(sc[c * 8 + 2 * e + 1] + sh);
          l += p0 + p1;
          split_f16_pair(p0, p1, hi[e], lo[e]);
        }
        const uint32_t off = (uint32_t)(c ^ (r & 7)) << 4;
        st_shared_v4(p_row_hi + off, hi[0], hi[1], hi[2], hi[3]);
        st_shared_v4(p_row_lo + off, lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();
      mbar_arrive(p_full);
    }
    mbar_wait(o_full, (nkv - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < D / 16; ++c) pull16<true>(t_lane + TM_OHI + c * 16, t_lane + TM_OLO + c * 16, o + c * 16);
    tc_fence_before();
    const int row = qt * BQ + r;
    if (row < tokens) {  // O / l (the 2^14 of P cancels), split into hi | lo planes: two 128-byte rows per thread
      const float inv = 1.0f / l;
      __half* dst = out + (long long)(row_base + row) * LDO + h * D;
#pragma unroll
      for (int c = 0; c < D / 8; ++c) {
        uint4 hi, lo;
        split_f16_pair(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv, hi.x, lo.x);
        split_f16_pair(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv, hi.y, lo.y);
        split_f16_pair(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv, hi.z, lo.z);
        split_f16_pair(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv, hi.w, lo.w);
        *reinterpret_cast<uint4*>(dst + c * 8) = hi;
        *reinterpret_cast<uint4*>(dst + HID + c * 8) = lo;
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TM_COLS);
}
}  // namespace tc

}  // namespace attn_split

int attention_split(const void* qkv, void* out, int batch, int tokens, cudaStream_t stream) {
  using namespace attn_split;
  int rc = device_check();
  if (rc) return rc;
  if (!qkv || !out || batch <= 0 || tokens <= 0) {
    set_error("attention_split: null pointer or empty shape");
    return ZK_ERR_ARG;
  }
  if (batch > 65535) {
    set_error("attention_split: batch %d exceeds the grid limit", batch);
    return ZK_ERR_SHAPE;
  }
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
    set_error("attention_split: buffers must be 16-byte aligned");
    return ZK_ERR_ARG;
  }
  // ZK_SPLIT_ATTN=mma selects the warp-level mma.sync kernel (the cross-check of the tcgen05 one in the tests)
  static const bool use_mma = getenv("ZK_SPLIT_ATTN") && !strcmp(getenv("ZK_SPLIT_ATTN"), "mma");
  if (!use_mma) {
    static unsigned long long attr_tc = 0;
    if ((rc = ensure_dynamic_smem(reinterpret_cast<const void*>(tc::attn_split_tc_kernel), tc::SMEM_BYTES, &attr_tc))) return rc;
    CUtensorMap tq, tkv;
    const uint64_t rows = (uint64_t)batch * tokens;
    if ((rc = make_tmap_bf16_2d(&tq, qkv, rows, LDQ, LDQ, tc::BQ, D))) return rc;
    if ((rc = make_tmap_bf16_2d(&tkv, qkv, rows, LDQ, LDQ, tc::BKV, D))) return rc;
    ProfScope prof(ZK_K_ATTENTION, stream);
    tc::attn_split_tc_kernel<<<dim3((tokens + tc::BQ - 1) / tc::BQ, HEADS, batch), tc::THREADS, tc::SMEM_BYTES, stream>>>(
        tq, tkv, reinterpret_cast<__half*>(out), tokens);
    ZK_LAUNCH_CHECK("attn_split_tc_kernel");
    return 0;
  }
  static unsigned long long attr_done = 0;
  if ((rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_split_kernel), SMEM_BYTES, &attr_done))) return rc;
  ProfScope prof(ZK_K_ATTENTION, stream);
  attn_split_kernel<<<dim3((tokens + BQ - 1) / BQ, HEADS, batch), THREADS, SMEM_BYTES, stream>>>(
      reinterpret_cast<const __half*>(qkv), reinterpret_cast<__half*>(out), tokens);
  ZK_LAUNCH_CHECK("attn_split_kernel");
  return 0;
}

}  // namespace zk

extern "C" int zk_attention_split(const void* d_qkv, void* d_out, int batch, int tokens, zk_stream_t stream) {
  return zk::attention_split(d_qkv, d_out, batch, tokens, (cudaStream_t)stream);
}
