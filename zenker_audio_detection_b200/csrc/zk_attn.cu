// Fused non-causal attention for sm_100a: O = softmax(Q K^T / sqrt(64)) V per (window, head), head_dim 64.
// One persistent-size CTA per SM handles 256 queries of one (window, head) as TWO 128-row tiles (A, B) that share the
// K/V stream and run their softmax in anti-phase: while tile A's softmax warps own the MUFU pipe, the tensor core
// produces tile B's next scores, and vice versa (at head_dim 64 the 16-lane MUFU pipe, not the tensor pipe, bounds
// attention: 128x128 exp2 per block = 1024 cycles vs 512 cycles of MMA).
//
//   warp 0       TMA producer: both Q tiles once, then a 3-stage ring of {K_j, V_j} 128x64 bf16 tiles
//   warp 1       TMEM allocator + MMA issuer, per tile t in {A, B}:
//                  S_t = Q_t K_j^T   (UMMA 128x128x16, TMEM cols [128 t, 128 t + 128))
//                  O_t += P_t V_j    (UMMA 128x64x16, V as MN-major operand, TMEM cols [256 + 64 t, +64))
//   warps 4..7   softmax of tile A, warps 8..11 of tile B: one query row per thread, S pulled out of TMEM in
//                double-buffered 32-column chunks, fp32 max (FMNMX3) / exp2 (FFMA2 + MUFU, optionally part on the
//                FMA pipe) / sum (FADD2), P written to 128B-swizzled shared memory as the bf16 A operand of the
//                second MMA.  O accumulates in TMEM across key blocks; the running maximum is only advanced (and O
//                rescaled in TMEM, a rare tcgen05.ld/st round trip) when a block maximum exceeds it by more than 2^8.
//
// Q, K and V are read in place from the fused QKV projection output [batch*tokens][2304] (q | k | v, head h
// at columns 64h), keys beyond `tokens` are masked to -inf (1214 = 9*128 + 62).
// Replaces F.scaled_dot_product_attention on the reference path (HF:modeling_audio_spectrogram_transformer.py:162-176).
#include <stdlib.h>

#include "zk_b200.h"
#include "zk_common.cuh"
#include "zk_internal.cuh"

namespace zk {
namespace attn {

constexpr int BQ = 128, BKV = 128, D = 64, HEADS = 12, HID = HEADS * D, KV_STAGES = 3, QTILES = 2;
constexpr int TILE_BYTES = 128 * 64 * 2;  // one 128 x 64 bf16 tile (Q_t, K_j or V_j)
constexpr int P_BYTES = BQ * BKV * 2;
constexpr int OFF_Q = 0, OFF_KV = QTILES * TILE_BYTES, OFF_P = OFF_KV + KV_STAGES * 2 * TILE_BYTES,
              OFF_BAR = OFF_P + QTILES * P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int THREADS = 384;  // warps 0-3: producer, MMA, 2 spare; 4-7: softmax A; 8-11: softmax B
constexpr uint32_t TMEM_COLS = 512, TM_S = 0, TM_O = 256;  // S_t at 128 t, O_t at 256 + 64 t
constexpr float RESCALE_TAU = 8.0f;  // in log2 units: p <= 2^8 with a stale maximum
constexpr uint32_t IDESC_S = umma_idesc_bf16(BQ, BKV, 0, 0);
constexpr uint32_t IDESC_O = umma_idesc_bf16(BQ, D, 0, 1);  // B (= V) is MN-major
constexpr float SCALE_LOG2E = 0.125f * 1.44269504088896340736f;

// exp2 on the FMA pipe for a pair of arguments (the MUFU unit does 16 ex2 / clk / SM and is the attention bottleneck at
// head_dim 64): round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-3 minimax 2^f (max rel. error 7.5e-5, far
// below the bf16 rounding of P), exponent patched in with one integer multiply-add.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f);
  const float2 t = fadd2(x, magic);
  const float2 n = fadd2(t, nmagic);
  const float2 f = ffma2(n, make_float2(-1.f, -1.f), x);
  float2 p = ffma2(make_float2(0.055171459913253784f, 0.055171459913253784f), f,
                   make_float2(0.2426108568906784f, 0.2426108568906784f));
  p = ffma2(p, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  p = ffma2(p, f, make_float2(0.9999281167984009f, 0.9999281167984009f));
  float2 r;
  r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return r;
}

// One key block of the online softmax for one query row (see the kernel comment).  RAGGED = the last key block,
// whose keys >= kmax are masked to -inf; the common instantiation carries no masking instructions at all.
// POLY = how many of every four element pairs take the FMA-pipe exp2 instead of MUFU.EX2.
template <bool RAGGED, int POLY>
__device__ __forceinline__ void softmax_block(int j, int kmax, uint32_t t_s, uint32_t t_o, uint32_t sp_row, int row,
                                              uint64_t* pv_done, float& m, float2& l2a, float2& l2b, long long* trj) {
  uint32_t buf[2][32];
  // ---- pass 1: row maximum (TMEM reads are double buffered against the FMNMX3 chains)
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
  tmem_ld32(t_s, buf[0]);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t(&cur)[32] = buf[c & 1];
    if (c < 3) tmem_ld32(t_s + (c + 1) * 32, buf[(c + 1) & 1]);
    if (RAGGED) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c * 32 + i >= kmax) cur[i] = 0xff800000u;  // -inf
    }
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      mx0 = fmax3(mx0, __uint_as_float(cur[i + 0]), __uint_as_float(cur[i + 1]));
      mx1 = fmax3(mx1, __uint_as_float(cur[i + 2]), __uint_as_float(cur[i + 3]));
      mx2 = fmax3(mx2, __uint_as_float(cur[i + 4]), __uint_as_float(cur[i + 5]));
      mx3 = fmax3(mx3, __uint_as_float(cur[i + 6]), __uint_as_float(cur[i + 7]));
    }
    if (c < 3) tmem_ld_wait();
  }
  const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
  if (trj) trj[3] = clock64();
  tmem_ld32(t_s, buf[0]);  // first chunk of pass 2, in flight across the (rare) rescale
  bool waited_pv = false;
  if (j == 0) {
    m = mx;
  } else if (__any_sync(0xffffffffu, (mx - m) * SCALE_LOG2E > RESCALE_TAU)) {
    // rare: advance the running maximum and rescale the accumulator in TMEM (whole warp, tcgen05 is collective)
    const float mn = fmaxf(m, mx);
    const float alpha = fast_exp2((m - mn) * SCALE_LOG2E);
    m = mn;
    l2a.x *= alpha; l2a.y *= alpha; l2b.x *= alpha; l2b.y *= alpha;
    mbar_wait(pv_done, (j - 1) & 1);  // O holds every block < j
    waited_pv = true;
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      tmem_ld32(t_o + c * 32, buf[1]);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) buf[1][i] = __float_as_uint(__uint_as_float(buf[1][i]) * alpha);
      tmem_st32(t_o + c * 32, buf[1]);
    }
    tmem_st_wait();
  }
  // ---- pass 2: p = exp2(s * c - m * c), row sum, bf16 pack, swizzled store of the A operand of P V
  const float2 sc2 = make_float2(SCALE_LOG2E, SCALE_LOG2E);
  const float2 mb2 = make_float2(-m * SCALE_LOG2E, -m * SCALE_LOG2E);
  tmem_ld_wait();
  if (j > 0 && !waited_pv) mbar_wait(pv_done, (j - 1) & 1);  // P buffer is free once P_{j-1} V_{j-1} has completed
  if (trj) trj[7] = clock64();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t(&cur)[32] = buf[c & 1];
    if (c < 3) tmem_ld32(t_s + (c + 1) * 32, buf[(c + 1) & 1]);
    if (RAGGED) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c * 32 + i >= kmax) cur[i] = 0xff800000u;
    }
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float2 xa = ffma2(make_float2(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])), sc2, mb2);
      const float2 xb = ffma2(make_float2(__uint_as_float(cur[i + 2]), __uint_as_float(cur[i + 3])), sc2, mb2);
      // pairs are numbered i/2; out of every four, the first POLY go to the FMA pipe
      const float2 pa = (((i >> 1) & 3) < POLY) ? exp2_poly2(xa) : make_float2(fast_exp2(xa.x), fast_exp2(xa.y));
      const float2 pb = ((((i >> 1) + 1) & 3) < POLY) ? exp2_poly2(xb) : make_float2(fast_exp2(xb.x), fast_exp2(xb.y));
      l2a = fadd2(l2a, pa);
      l2b = fadd2(l2b, pb);
      pk[i >> 1] = pack_bf16(pa.x, pa.y);
      pk[(i >> 1) + 1] = pack_bf16(pb.x, pb.y);
    }
    // keys [32c, 32c+32) = 64 B = four 16-B chunks of swizzle atom (c>>1)
    const uint32_t atom = sp_row + (c >> 1) * (BQ * 128);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t chunk = (uint32_t)(((c & 1) * 4 + q) ^ (row & 7));
      st_shared_v4(atom + chunk * 16, pk[q * 4 + 0], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
    }
    if (c < 3) tmem_ld_wait();
  }
}

template <int POLY>
__global__ void __launch_bounds__(THREADS, 1)
attn_kernel(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* __restrict__ out, int tokens, int stagger,
            long long* trace) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;    // [3]
  uint64_t* kv_empty = bars + 4;   // [3]
  uint64_t* s_full = bars + 7;     // [2] per tile
  uint64_t* p_full = bars + 9;     // [2]
  uint64_t* pv_done = bars + 11;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (tokens + BKV - 1) / BKV;
  const int row_base = b * tokens;  // first row of this window in the [batch*tokens] matrices
  // optional timeline capture (zk_attention_trace): 128 slots per CTA for the first 512 CTAs of the grid
  const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  long long* tr = (trace && cta_lin < 512) ? trace + (long long)cta_lin * 128 : nullptr;
  if (tr && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tr[0] = smid;
    tr[1] = clock64();
  }

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("zk attn: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tm);
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int t = 0; t < QTILES; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&pv_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, QTILES * TILE_BYTES);
      for (int t = 0; t < QTILES; ++t)
        tma_load_2d(smem + OFF_Q + t * TILE_BYTES, &tm, q_full, h * D, row_base + (qb * QTILES + t) * BQ);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * TILE_BYTES);
        uint8_t* dst = smem + OFF_KV + st * 2 * TILE_BYTES;
        tma_load_2d(dst, &tm, &kv_full[st], HID + h * D, row_base + j * BKV);
        tma_load_2d(dst + TILE_BYTES, &tm, &kv_full[st], 2 * HID + h * D, row_base + j * BKV);
        if (++st == KV_STAGES) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t sp = smem_u32(smem + OFF_P);
      auto kv_addr = [&](int j) { return smem_u32(smem + OFF_KV + (j % KV_STAGES) * 2 * TILE_BYTES); };
      auto issue_s = [&](int t, int j) {  // S_t = Q_t K_j^T
        const uint64_t q_desc = umma_desc_sw128(smem_u32(smem + OFF_Q + t * TILE_BYTES), 16, 1024);
        const uint64_t k_desc = umma_desc_sw128(kv_addr(j), 16, 1024);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_bf16_ss(tmem_base + TM_S + t * BKV, q_desc + 2 * k, k_desc + 2 * k, IDESC_S, k != 0);
        umma_commit(&s_full[t]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      // Tile B starts one softmax period after tile A (stagger): two softmax groups that start together slow each
      // other down symmetrically on the shared MUFU pipe and then idle together while the tensor core produces
      // their next scores; half a period apart, one group's MUFU phase covers the other's wait for S.
      issue_s(0, 0);
      if (!stagger) issue_s(1, 0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) {  // K_{j+1} / V_{j+1} must have landed before the first S_t(j+1)
          mbar_wait(&kv_full[(j + 1) % KV_STAGES], ((j + 1) / KV_STAGES) & 1);
          tc_fence_after();
        }
        const uint32_t sv = kv_addr(j) + TILE_BYTES;
        for (int t = 0; t < QTILES; ++t) {
          mbar_wait(&p_full[t], j & 1);  // P_t(j) is in smem and S_t(j) has been read out of TMEM
          tc_fence_after();
          if (tr && t == 0) tr[8 + j * 8 + 4] = clock64();
          if (j + 1 < nkv) issue_s(t, j + 1);
          if (tr && t == 0) tr[8 + j * 8 + 5] = clock64();
          const uint32_t d_o = tmem_base + TM_O + t * D;
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {
            // A = P_t: two 64-key swizzle atoms of 16 KiB; B = V_j: 16 keys = 16 rows of 128 B (MN-major)
            const uint64_t p_desc = umma_desc_sw128(sp + t * P_BYTES + (k >> 2) * (BQ * 128) + (k & 3) * 32, 16, 1024);
            const uint64_t v_desc = umma_desc_sw128(sv + k * 16 * 128, 1024, 1024);
            umma_bf16_ss(d_o, p_desc, v_desc, IDESC_O, (j | k) != 0);
          }
          umma_commit(&pv_done[t]);
          if (tr && t == 0) tr[8 + j * 8 + 6] = clock64();
          if (stagger && j == 0 && t == 0) issue_s(1, 0);
        }
        umma_commit(&kv_empty[j % KV_STAGES]);  // K_j and V_j are dead once both tiles' P V products have executed
      }
    }
  } else if (warp >= 4) {
    const int t = (warp - 4) >> 2;       // query tile of this warp group
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t t_s = t_lane + TM_S + t * BKV, t_o = t_lane + TM_O + t * D;
    const uint32_t sp_row = smem_u32(smem + OFF_P + t * P_BYTES) + row * 128;
    const bool tracer = tr && warp == 4 && lane == 0;
    float m = -INFINITY;
    float2 l2a = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);

    for (int j = 0; j < nkv; ++j) {
      if (tracer) tr[8 + j * 8 + 0] = clock64();
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      if (tracer) tr[8 + j * 8 + 1] = clock64();
      const int kmax = tokens - j * BKV;  // keys [0, kmax) of this block are valid; only the last block is ragged
      if (kmax < BKV)
        softmax_block<true, POLY>(j, kmax, t_s, t_o, sp_row, row, &pv_done[t], m, l2a, l2b, tracer ? tr + 8 + j * 8 : nullptr);
      else
        softmax_block<false, POLY>(j, kmax, t_s, t_o, sp_row, row, &pv_done[t], m, l2a, l2b, tracer ? tr + 8 + j * 8 : nullptr);
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(&p_full[t]);
      if (tracer) tr[8 + j * 8 + 2] = clock64();
    }
    if (tracer) tr[2] = clock64();
    {
      const int jl = nkv - 1;
      mbar_wait(&pv_done[t], jl & 1);
      tc_fence_after();
      const float inv = 1.0f / ((l2a.x + l2a.y) + (l2b.x + l2b.y));
      const int q = (qb * QTILES + t) * BQ + row;
      __nv_bfloat16* dst = out + (long long)(row_base + q) * HID + h * D;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(t_o + c * 32, r);
        tmem_ld_wait();
        if (q < tokens) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(r[g * 8 + 0]) * inv, __uint_as_float(r[g * 8 + 1]) * inv);
            v.y = pack_bf16(__uint_as_float(r[g * 8 + 2]) * inv, __uint_as_float(r[g * 8 + 3]) * inv);
            v.z = pack_bf16(__uint_as_float(r[g * 8 + 4]) * inv, __uint_as_float(r[g * 8 + 5]) * inv);
            v.w = pack_bf16(__uint_as_float(r[g * 8 + 6]) * inv, __uint_as_float(r[g * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = v;
          }
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace attn

int attention_bf16_impl(const void* qkv, void* out, int batch, int tokens, long long* trace, cudaStream_t stream) {
  using namespace attn;
  int rc = device_check();
  if (rc) return rc;
  if (!qkv || !out || batch <= 0 || tokens <= 0) {
    set_error("attention_bf16: null pointer or empty shape");
    return ZK_ERR_ARG;
  }
  if (batch > 65535) {
    set_error("attention_bf16: batch %d > 65535", batch);
    return ZK_ERR_SHAPE;
  }
  static int poly = -1;  // share of exp2 evaluated on the FMA pipe: 0, 1 or 2 of every 4 pairs (ZK_ATTN_POLY overrides)
  static int stagger = 1;
  if (poly < 0) {
    const char* sg = getenv("ZK_ATTN_STAGGER");
    if (sg) stagger = atoi(sg) != 0;
    ZK_CUDA(cudaFuncSetAttribute(attn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    ZK_CUDA(cudaFuncSetAttribute(attn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    ZK_CUDA(cudaFuncSetAttribute(attn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const char* e = getenv("ZK_ATTN_POLY");
    poly = e ? atoi(e) : 1;
    if (poly < 0 || poly > 2) poly = 1;
  }
  CUtensorMap tm;
  if ((rc = make_tmap_bf16_2d(&tm, qkv, (uint64_t)batch * tokens, 3 * HID, 3 * HID, 128, 64))) return rc;
  dim3 grid((tokens + QTILES * BQ - 1) / (QTILES * BQ), HEADS, batch);
  ProfScope prof(ZK_K_ATTENTION, stream);
  if (poly == 0)
    attn_kernel<0><<<grid, THREADS, SMEM_BYTES, stream>>>(tm, reinterpret_cast<__nv_bfloat16*>(out), tokens, stagger, trace);
  else if (poly == 1)
    attn_kernel<1><<<grid, THREADS, SMEM_BYTES, stream>>>(tm, reinterpret_cast<__nv_bfloat16*>(out), tokens, stagger, trace);
  else
    attn_kernel<2><<<grid, THREADS, SMEM_BYTES, stream>>>(tm, reinterpret_cast<__nv_bfloat16*>(out), tokens, stagger, trace);
  ZK_LAUNCH_CHECK("attn_kernel");
  return 0;
}

int attention_bf16(const void* qkv, void* out, int batch, int tokens, cudaStream_t stream) {
  return attention_bf16_impl(qkv, out, batch, tokens, nullptr, stream);
}

}  // namespace zk

extern "C" int zk_attention_trace(const void* d_qkv, void* d_out, int batch, int tokens, int64_t* d_trace, zk_stream_t stream) {
  return zk::attention_bf16_impl(d_qkv, d_out, batch, tokens, reinterpret_cast<long long*>(d_trace), (cudaStream_t)stream);
}

extern "C" int zk_attention_bf16(const void* d_qkv, void* d_out, int batch, int tokens, zk_stream_t stream) {
  return zk::attention_bf16(d_qkv, d_out, batch, tokens, (cudaStream_t)stream);
}
