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
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 3)) {
    set_error("attention_split: qkv must be 16-byte aligned");
    return ZK_ERR_ARG;
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
