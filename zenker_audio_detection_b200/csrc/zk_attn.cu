// Fused non-causal attention for sm_100a: O = softmax(Q K^T / sqrt(64)) V per (window, head), head_dim 64.
//
// Persistent kernel, one CTA per SM.  A work item is 256 queries of one (window, head), processed as TWO 128-row tiles
// (A, B) that share the K/V stream; the CTA walks its items back to back so the K/V ring, the tensor pipe and the
// softmax warps never drain between items.
//
//   warp 0       TMA producer: {Q_A, Q_B} of the next item (double buffered), 4-stage ring of {K_j, V_j} 128x64 tiles
//   warps 1, 2   MMA issuers of tile A / tile B (warp 1 also owns the TMEM allocation), per tile t:
//                  S_t  = Q_t K_j^T   (UMMA 128x128x16 SS, TMEM cols [128 t, +128))
//                  O_t += P_t V_j     (UMMA 128x64x16  TS: P is read from TMEM, V_j is the MN-major smem operand)
//   warps 4..7   softmax of tile A, warps 8..11 of tile B (216 registers each via setmaxnreg): one query row per
//                thread.  The whole 128-key score row is pulled out of TMEM ONCE into registers and the S buffer is
//                handed back to the tensor core at that point, so S_t(j+1) is computed underneath the exponentials of
//                block j.  fp32 max (FMNMX3) / exp2 (FFMA2 + MUFU, 1/4 on the FMA pipe) / sum (FADD2); P goes back
//                to TMEM as packed bf16 pairs (tcgen05.st), not through shared memory: at head_dim 64 the 128 B/clk
//                shared-memory port is as scarce as the 16-lane MUFU pipe, and the smem round trip of P was 57 % of
//                the operand traffic.  O accumulates in TMEM across key blocks; the running maximum is only advanced
//                (and O rescaled in TMEM, a rare tcgen05.ld/st round trip) when a block maximum exceeds it by 2^8.
//
// Q, K and V are read in place from the fused QKV projection output [batch*tokens][2304] (q | k | v, head h
// at columns 64h); keys beyond `tokens` are masked to -inf (1214 = 9*128 + 62).
// Replaces F.scaled_dot_product_attention on the reference path (HF:modeling_audio_spectrogram_transformer.py:162-176).
#include <stdlib.h>

#include "zk_b200.h"
#include "zk_common.cuh"
#include "zk_internal.cuh"

namespace zk {
namespace attn {

constexpr int BQ = 128, BKV = 128, D = 64, HEADS = 12, HID = HEADS * D, KV_STAGES = 4, QTILES = 2;
constexpr int TILE_BYTES = 128 * 64 * 2;  // one 128 x 64 bf16 tile (Q_t, K_j or V_j)
constexpr int OFF_Q = 0;                                   // [2 item parities][2 tiles]
constexpr int OFF_KV = 2 * QTILES * TILE_BYTES;            // [KV_STAGES]{K, V}
constexpr int OFF_STG = OFF_KV + KV_STAGES * 2 * TILE_BYTES;  // [2 tiles] 128 x 64 bf16 output staging (TMA store)
constexpr int OFF_BAR = OFF_STG + QTILES * TILE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int THREADS = 384;  // warps 0-3: producer, MMA A, MMA B, spare; 4-7: softmax A; 8-11: softmax B
// TMEM columns: S_t at 128 t (fp32), O_t at 256 + 64 t (fp32), P_t at 384 + 64 t (bf16 pairs, 128 keys)
constexpr uint32_t TMEM_COLS = 512, TM_S = 0, TM_O = 256, TM_P = 384;
// log2 units: p <= 2^tau against a stale maximum.  bf16 P keeps its relative precision over the whole fp32 exponent
// range (tau 24: 1214 keys * 2^24 * |v| is nowhere near fp32 range); fp16 P must stay below 65504 (tau 15), and its
// values below 2^-14 relative to a row sum >= 1 are subnormal, i.e. still good to 2^-25 of the sum.
template <int FMT>
constexpr float rescale_tau() { return FMT == FMT_F16 ? 15.0f : 24.0f; }
constexpr float SCALE_LOG2E = 0.125f * 1.44269504088896340736f;
// the rescale trigger of the online softmax: true = the block's row sum (free), false = the block's maximum (FMNMX3)
constexpr bool SUM_TEST = true;
constexpr int REGS_SOFTMAX = 216, REGS_OTHER = 56;
static_assert(128 * REGS_OTHER + 256 * REGS_SOFTMAX <= 65536, "register file");

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

// A D-operand MMA with A in tensor memory (P as packed bf16 pairs: lane = query row, 8 columns per 16 keys).
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// 32 consecutive columns into r[OFF .. OFF+32) of a larger register array
template <int OFF, int N>
__device__ __forceinline__ void tmem_ld32_at(uint32_t taddr, uint32_t (&r)[N]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[OFF + 0]), "=r"(r[OFF + 1]), "=r"(r[OFF + 2]), "=r"(r[OFF + 3]), "=r"(r[OFF + 4]), "=r"(r[OFF + 5]),
        "=r"(r[OFF + 6]), "=r"(r[OFF + 7]), "=r"(r[OFF + 8]), "=r"(r[OFF + 9]), "=r"(r[OFF + 10]), "=r"(r[OFF + 11]),
        "=r"(r[OFF + 12]), "=r"(r[OFF + 13]), "=r"(r[OFF + 14]), "=r"(r[OFF + 15]), "=r"(r[OFF + 16]),
        "=r"(r[OFF + 17]), "=r"(r[OFF + 18]), "=r"(r[OFF + 19]), "=r"(r[OFF + 20]), "=r"(r[OFF + 21]),
        "=r"(r[OFF + 22]), "=r"(r[OFF + 23]), "=r"(r[OFF + 24]), "=r"(r[OFF + 25]), "=r"(r[OFF + 26]),
        "=r"(r[OFF + 27]), "=r"(r[OFF + 28]), "=r"(r[OFF + 29]), "=r"(r[OFF + 30]), "=r"(r[OFF + 31])
      : "r"(taddr)
      : "memory");
}

struct SoftmaxState {
  float m;
  float2 l2a, l2b;
};

// One key block of the online softmax for one query row.  FIRST = first key block of the work item, `n` = running
// block number of this tile across items (mbarrier parities), RAGGED = the last key block, whose keys >= kmax are
// masked to -inf.  POLY = how many of every four element pairs take the FMA-pipe exp2 instead of MUFU.EX2.
//
// Only the first block of an item needs its row maximum before the exponentials.  Every later block exponentiates
// against the running (possibly stale) maximum straight away and computes its own maximum alongside (FMNMX3 on the
// ALU pipe, independent of the MUFU stream); only if some row of the warp then turns out to exceed the running maximum
// by more than 2^8 is the accumulator rescaled and the block redone from the scores still held in registers.  That
// keeps a warp's MUFU demand uniform over the block instead of 0 % during a max phase and 100 % after it, which is
// what lets the two softmax warps of a scheduler share the MUFU pipe without phase locking.
//
// W = how many key columns of the block are processed at all: a ragged last block with kmax <= 64 valid keys (tokens =
// 1214 leaves 62) only takes the lower half of S, writes the lower half of P, and the issuer shortens P V to W keys.
template <bool RAGGED, bool FIRST, int POLY, int FMT, int W = 128>
__device__ __forceinline__ void softmax_block(uint32_t n, int kmax, bool row_valid, uint32_t t_s, uint32_t t_o, uint32_t t_p,
                                              uint32_t s_free, uint32_t pv_done, SoftmaxState& st, long long* trj) {
  uint32_t s[128];
  static_assert(W == 64 || W == 128, "key columns per block");
  tmem_ld32_at<0>(t_s, s);
  tmem_ld32_at<32>(t_s + 32, s);
  if (W == 128) {
    tmem_ld32_at<64>(t_s + 64, s);
    tmem_ld32_at<96>(t_s + 96, s);
  }
  tmem_ld_wait();
  tc_fence_before();
  mbar_arrive_a(s_free);  // the score buffer may be overwritten by S(j+1) from here on
  if (trj) trj[3] = clock64();
  if (RAGGED) {  // keys >= kmax -> -inf; kmax is uniform, so only the group of 8 that straddles it pays the compares
#pragma unroll
    for (int i0 = 0; i0 < W; i0 += 8) {
      if (i0 + 8 > kmax) {
#pragma unroll
        for (int i = i0; i < i0 + 8; ++i)
          if (i >= kmax) s[i] = 0xff800000u;
      }
    }
  }
  if (FIRST) {
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
    for (int i = 0; i < W; i += 8) {
      mx0 = fmax3(mx0, __uint_as_float(s[i + 0]), __uint_as_float(s[i + 1]));
      mx1 = fmax3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
      mx2 = fmax3(mx2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
      mx3 = fmax3(mx3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
    }
    st.m = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    st.l2a = make_float2(0.f, 0.f);
    st.l2b = make_float2(0.f, 0.f);
  }
  if (trj) trj[7] = clock64();
  const float2 sc2 = make_float2(SCALE_LOG2E, SCALE_LOG2E);
  float2 lba, lbb;  // row sum of this block
#pragma unroll 1
  for (int pass = 0;; ++pass) {
    // p = exp2(s * c - m * c), row sum, bf16 pack; P goes to TMEM as the A operand of P V (column i = keys 2i, 2i+1)
    const float2 mb2 = make_float2(-st.m * SCALE_LOG2E, -st.m * SCALE_LOG2E);
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY, mxp = -INFINITY;
    lba = make_float2(0.f, 0.f);
    lbb = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < W / 32; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const int e = c * 32 + i;
        const float2 xa = ffma2(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), sc2, mb2);
        const float2 xb = ffma2(make_float2(__uint_as_float(s[e + 2]), __uint_as_float(s[e + 3])), sc2, mb2);
        // pairs are numbered i/2; out of every four, the first POLY go to the FMA pipe
        const float2 pa = (((i >> 1) & 3) < POLY) ? exp2_poly2(xa) : make_float2(fast_exp2(xa.x), fast_exp2(xa.y));
        const float2 pb = ((((i >> 1) + 1) & 3) < POLY) ? exp2_poly2(xb) : make_float2(fast_exp2(xb.x), fast_exp2(xb.y));
        // the FMA-pipe exp2 wraps around for arguments past 2^7 (its exponent patch is an integer add) instead of
        // saturating like MUFU.EX2, so the row-sum test below cannot see such an element: track their arguments
        if (SUM_TEST && !FIRST && (((i >> 1) & 3) < POLY)) mxp = fmax3(mxp, xa.x, xa.y);
        if (SUM_TEST && !FIRST && ((((i >> 1) + 1) & 3) < POLY)) mxp = fmax3(mxp, xb.x, xb.y);
        lba = fadd2(lba, pa);
        lbb = fadd2(lbb, pb);
        pk[i >> 1] = pack16<FMT>(pa.x, pa.y);
        pk[(i >> 1) + 1] = pack16<FMT>(pb.x, pb.y);
        if (!FIRST && !SUM_TEST) {
          if (i & 4) {
            mx2 = fmax3(mx2, __uint_as_float(s[e]), __uint_as_float(s[e + 1]));
            mx3 = fmax3(mx3, __uint_as_float(s[e + 2]), __uint_as_float(s[e + 3]));
          } else {
            mx0 = fmax3(mx0, __uint_as_float(s[e]), __uint_as_float(s[e + 1]));
            mx1 = fmax3(mx1, __uint_as_float(s[e + 2]), __uint_as_float(s[e + 3]));
          }
        }
      }
      if (c == 0 && pass == 0 && n > 0) {
        mbar_wait_a(pv_done, (n - 1) & 1);  // the P buffer is free (and O final up to block j-1) once P(j-1) V(j-1) is done
        tc_fence_after();
      }
      tmem_st16(t_p + c * 16, pk);
    }
    if (FIRST || pass == 1) break;
    // Only rows that are real queries of this window vote, and only rows that exceed the threshold themselves move
    // their maximum: a row's result must not depend on what else shares its warp (the rows past the last token of a
    // window hold the NEXT window's queries, i.e. they change with the batch composition).
    bool exceed;
    if (SUM_TEST) {
      // The block's row sum is computed anyway and bounds every p of the row from above: as long as it stays below
      // 2^tau no element left the operand format's range and the block maximum is never needed (64 FMNMX3 per block
      // that only the rare path below pays; the FMA-pipe elements keep a maximum of their own, see above).  A NaN /
      // inf sum fails the comparison and takes the rare path too.
      const float lsum = (lba.x + lba.y) + (lbb.x + lbb.y);
      exceed = row_valid && (!(lsum <= exp2f(rescale_tau<FMT>())) || mxp > rescale_tau<FMT>());
      if (!__any_sync(0xffffffffu, exceed)) break;
#pragma unroll
      for (int i = 0; i < W; i += 8) {
        mx0 = fmax3(mx0, __uint_as_float(s[i + 0]), __uint_as_float(s[i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        mx2 = fmax3(mx2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
        mx3 = fmax3(mx3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
      }
    }
    const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    if (!SUM_TEST) {
      exceed = row_valid && (mx - st.m) * SCALE_LOG2E > rescale_tau<FMT>();
      if (!__any_sync(0xffffffffu, exceed)) break;
    }
    // rare: advance the running maximum, rescale the accumulator in TMEM (whole warp, tcgen05 is collective), redo
    const float mn = exceed ? mx : st.m;
    const float alpha = fast_exp2((st.m - mn) * SCALE_LOG2E);
    st.m = mn;
    st.l2a.x *= alpha; st.l2a.y *= alpha; st.l2b.x *= alpha; st.l2b.y *= alpha;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(t_o + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
      tmem_st32(t_o + c * 32, o);
    }
  }
  st.l2a = fadd2(st.l2a, lba);
  st.l2b = fadd2(st.l2b, lbb);
  tmem_st_wait();
}

// TRACE = false is the production instantiation: the timeline hooks (and the pointer arithmetic / flag registers they
// keep alive across every key block: ~60 of the ~100 instructions of block entry + exit) are compiled out.
template <int POLY, int FMT, bool TRACE>
__global__ void __launch_bounds__(THREADS, 1)
attn_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_out, int tokens,
            int num_items, int qpairs, int stagger, int half_keys, long long* trace, int trace_items) {
  constexpr uint32_t IDESC_S = umma_idesc_16(FMT, BQ, BKV, 0, 0);
  constexpr uint32_t IDESC_O = umma_idesc_16(FMT, BQ, D, 0, 1);  // B (= V) is MN-major
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;          // [2] item parity
  uint64_t* q_empty = bars + 2;     // [2]
  uint64_t* kv_full = bars + 4;     // [KV_STAGES]
  uint64_t* kv_empty = bars + 8;    // [KV_STAGES]
  uint64_t* s_full = bars + 12;     // [2] per tile
  uint64_t* s_free = bars + 14;     // [2]
  uint64_t* p_full = bars + 16;     // [2]
  uint64_t* pv_done = bars + 18;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkv = (tokens + BKV - 1) / BKV;
  const int my_items = ((int)blockIdx.x < num_items) ? (num_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  // optional timeline capture (zk_attention_trace): 128 slots per CTA, first item of each CTA
  // trace_items: record only the end of every work item (slots 8 + it for tile A, 68 + it for tile B, it < 60)
  long long* tr_it = (TRACE && trace && trace_items && blockIdx.x < 512) ? trace + (long long)blockIdx.x * 128 : nullptr;
  long long* tr = (TRACE && trace && !trace_items && blockIdx.x < 512) ? trace + (long long)blockIdx.x * 128 : nullptr;
  if (TRACE && tr_it && threadIdx.x == 0) tr_it[1] = clock64();
  if (TRACE && tr && threadIdx.x == 0) {
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
    tma_prefetch_desc(&tm_out);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 2);
    }
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);
    }
    for (int t = 0; t < QTILES; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 128);
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

  // item i of this CTA -> (window b, head h, query pair qb); consecutive items share K/V through L2
  auto item_coords = [&](int it, int& b, int& h, int& qb) {
    const int item = (int)blockIdx.x + it * (int)gridDim.x;
    qb = item % qpairs;
    const int bh = item / qpairs;
    h = bh % HEADS;
    b = bh / HEADS;
  };

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_OTHER));
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      uint32_t g = 0;  // running K/V block number
      for (int it = 0; it < my_items; ++it) {
        int b, h, qb;
        item_coords(it, b, h, qb);
        const int row_base = b * tokens;
        const int qs = it & 1;
        if (it >= 2) mbar_wait(&q_empty[qs], ((it >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&q_full[qs], QTILES * TILE_BYTES);
        for (int t = 0; t < QTILES; ++t)
          tma_load_2d(smem + OFF_Q + (qs * QTILES + t) * TILE_BYTES, &tm, &q_full[qs], h * D,
                      row_base + (qb * QTILES + t) * BQ);
        for (int j = 0; j < nkv; ++j, ++g) {
          const uint32_t stg = g % KV_STAGES;
          if (g >= (uint32_t)KV_STAGES) mbar_wait(&kv_empty[stg], ((g / KV_STAGES) - 1) & 1);
          mbar_arrive_expect_tx(&kv_full[stg], 2 * TILE_BYTES);
          uint8_t* dst = smem + OFF_KV + stg * 2 * TILE_BYTES;
          tma_load_2d(dst, &tm, &kv_full[stg], HID + h * D, row_base + j * BKV);
          tma_load_2d(dst + TILE_BYTES, &tm, &kv_full[stg], 2 * HID + h * D, row_base + j * BKV);
        }
      }
    } else if (warp == 1 || warp == 2) {
      // ------------------------------------------------------------------ MMA issuers: warp 1 = tile A, warp 2 = tile B
      // Per tile the events arrive in a fixed order (S buffer read -> P published), so each issuer simply blocks on
      // its next mbarrier (a suspended try_wait costs no issue slots; a polling loop over both tiles' barriers
      // starved the two softmax warps that share its scheduler).  The whole warp walks the loop on warp-uniform
      // state and one elected lane issues: under a divergent `lane == 0` ptxas wraps every UTCHMMA in an
      // ELECT / BRA.U.ANY loop, ~90 clk per instruction against 32-64 clk of tensor work.
      const int t = warp - 1;
      const uint32_t G = (uint32_t)my_items * (uint32_t)nkv;  // key blocks of this tile over all items of this CTA
      const uint32_t d_s = tmem_base + TM_S + t * BKV, d_o = tmem_base + TM_O + t * D, a_p = tmem_base + TM_P + t * 64;
      auto issue_s = [&](uint32_t g) {  // S_t(g) = Q_t K_g^T; K/V stage g and the Q buffer of its item are full
        const uint32_t it = g / (uint32_t)nkv, j = g - it * (uint32_t)nkv;
        mbar_wait(&kv_full[g % KV_STAGES], (g / KV_STAGES) & 1);
        if (j == 0) mbar_wait(&q_full[it & 1], (it >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t q_desc = umma_desc_sw128(smem_u32(smem + OFF_Q + ((it & 1) * QTILES + t) * TILE_BYTES), 16, 1024);
          const uint64_t k_desc = umma_desc_sw128(smem_u32(smem + OFF_KV + (g % KV_STAGES) * 2 * TILE_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < D / 16; ++k) umma_bf16_ss(d_s, q_desc + 2 * k, k_desc + 2 * k, IDESC_S, k != 0);
          umma_commit(&s_full[t]);
          // the Q buffer of the item is dead once the last S of both tiles has executed (barrier count 2)
          if (j == (uint32_t)nkv - 1) umma_commit(&q_empty[it & 1]);
          if (TRACE && tr && t == 0 && g < 15) tr[8 + g * 8 + 5] = clock64();
        }
        __syncwarp();
      };
      // Tile B starts a fraction of a block period after tile A so that the two softmax groups do not need the MUFU
      // pipe at the same time (stagger 1: once A has pulled its first scores out of TMEM, 2: once A published P(0)).
      if (G > 0 && t == 1 && stagger == 1) mbar_wait(&s_free[0], 0);
      if (G > 0 && t == 1 && stagger == 2) mbar_wait(&p_full[0], 0);
      if (G > 0) issue_s(0);
      for (uint32_t g = 0; g < G; ++g) {
        if (g + 1 < G) {
          mbar_wait(&s_free[t], g & 1);  // the softmax warps hold S(g) in registers: the buffer can take S(g+1)
          issue_s(g + 1);
        }
        mbar_wait(&p_full[t], g & 1);    // P(g) is in TMEM
        tc_fence_after();
        if (elect_one()) {
          const uint32_t j = g % (uint32_t)nkv;
          const uint32_t sv = smem_u32(smem + OFF_KV + (g % KV_STAGES) * 2 * TILE_BYTES + TILE_BYTES);
          // a last block with at most 64 valid keys only has the lower half of P written (softmax_block W = 64)
          const int ksteps = (tokens - (int)j * BKV <= half_keys) ? BKV / 32 : BKV / 16;
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {
            // A = P_t: 16 keys = 8 TMEM columns; B = V_j: 16 keys = 16 rows of 128 B (MN-major)
            const uint64_t v_desc = umma_desc_sw128(sv + k * 16 * 128, 1024, 1024);
            if (k < ksteps) umma_bf16_ts(d_o, a_p + k * 8, v_desc, IDESC_O, (j | (uint32_t)k) != 0);
          }
          umma_commit(&pv_done[t]);
          umma_commit(&kv_empty[g % KV_STAGES]);  // K_g / V_g are dead once both tiles' P V have executed (count 2)
          if (TRACE && tr && t == 0 && g < 15) tr[8 + g * 8 + 6] = clock64();
        }
        __syncwarp();
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOFTMAX));
    // ------------------------------------------------------------------ softmax + epilogue
    const int t = (warp - 4) >> 2;       // query tile of this warp group
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t t_s = t_lane + TM_S + t * BKV, t_o = t_lane + TM_O + t * D, t_p = t_lane + TM_P + t * 64;
    const bool tracer = TRACE && tr && warp == 4 && lane == 0;
    // this tile's four barriers as 32-bit shared addresses held in registers (see smem_addr_keep)
    const uint32_t a_s_full = smem_addr_keep(&s_full[t]), a_s_free = smem_addr_keep(&s_free[t]);
    const uint32_t a_p_full = smem_addr_keep(&p_full[t]), a_pv_done = smem_addr_keep(&pv_done[t]);
    SoftmaxState st;
    st.m = -INFINITY;
    st.l2a = make_float2(0.f, 0.f);
    st.l2b = make_float2(0.f, 0.f);
    uint32_t n = 0;  // running key-block number of this tile
    for (int it = 0; it < my_items; ++it) {
      int b, h, qb;
      item_coords(it, b, h, qb);
      const bool row_valid = (qb * QTILES + t) * BQ + row < tokens;
      for (int j = 0; j < nkv; ++j, ++n) {
        long long* trj = (TRACE && tracer && n < 15) ? tr + 8 + n * 8 : nullptr;
        if (trj) trj[0] = clock64();
        mbar_wait_a(a_s_full, n & 1);
        tc_fence_after();
        if (trj) trj[1] = clock64();
        const int kmax = tokens - j * BKV;  // keys [0, kmax) of this block are valid; only the last block is ragged
        if (kmax <= half_keys) {
          if (j == 0)
            softmax_block<true, true, POLY, FMT, 64>(n, kmax, row_valid, t_s, t_o, t_p, a_s_free, a_pv_done, st, trj);
          else
            softmax_block<true, false, POLY, FMT, 64>(n, kmax, row_valid, t_s, t_o, t_p, a_s_free, a_pv_done, st, trj);
        } else if (kmax < BKV) {
          if (j == 0)
            softmax_block<true, true, POLY, FMT>(n, kmax, row_valid, t_s, t_o, t_p, a_s_free, a_pv_done, st, trj);
          else
            softmax_block<true, false, POLY, FMT>(n, kmax, row_valid, t_s, t_o, t_p, a_s_free, a_pv_done, st, trj);
        } else {
          if (j == 0)
            softmax_block<false, true, POLY, FMT>(n, kmax, row_valid, t_s, t_o, t_p, a_s_free, a_pv_done, st, trj);
          else
            softmax_block<false, false, POLY, FMT>(n, kmax, row_valid, t_s, t_o, t_p, a_s_free, a_pv_done, st, trj);
        }
        tc_fence_before();
        mbar_arrive_a(a_p_full);
        if (trj) trj[2] = clock64();
        if (TRACE && tr && warp == 8 && lane == 0 && n < 15) tr[8 + n * 8 + 4] = clock64();  // tile B: P published
      }
      // epilogue of the item: O / l -> bf16 -> 128B-swizzled staging tile -> one TMA store per tile (rows beyond the
      // window's last token are clipped by the 3-D tensor map).  The next item's first P V (accumulate = 0) is only
      // issued after this thread has published its next P, i.e. after these TMEM reads.
      mbar_wait_a(a_pv_done, (n - 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / ((st.l2a.x + st.l2a.y) + (st.l2b.x + st.l2b.y));
      uint32_t r[64];
      tmem_ld32_at<0>(t_o, r);
      tmem_ld32_at<32>(t_o + 32, r);
      tmem_ld_wait();
      const bool leader = (warp & 3) == 0 && lane == 0;
      if (leader) bulk_wait_read0();  // the previous item's store has finished reading the staging tile
      named_bar_sync(1 + t, 128);
      uint8_t* stg = smem + OFF_STG + t * TILE_BYTES;
      const uint32_t stg_row = smem_u32(stg) + row * 128;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        st_shared_v4(stg_row + ((uint32_t)(g ^ (row & 7)) << 4),
                     pack16<FMT>(__uint_as_float(r[g * 8 + 0]) * inv, __uint_as_float(r[g * 8 + 1]) * inv),
                     pack16<FMT>(__uint_as_float(r[g * 8 + 2]) * inv, __uint_as_float(r[g * 8 + 3]) * inv),
                     pack16<FMT>(__uint_as_float(r[g * 8 + 4]) * inv, __uint_as_float(r[g * 8 + 5]) * inv),
                     pack16<FMT>(__uint_as_float(r[g * 8 + 6]) * inv, __uint_as_float(r[g * 8 + 7]) * inv));
      fence_proxy_async();
      named_bar_sync(1 + t, 128);
      if (leader) {
        tma_store_3d(&tm_out, stg, h * D, (qb * QTILES + t) * BQ, b);
        bulk_commit();
      }
      if (TRACE && tracer && it == 0) tr[2] = clock64();
      if (TRACE && tr_it && quarter == 0 && lane == 0 && it < 60) tr_it[(t ? 68 : 8) + it] = clock64();
    }
    if ((warp & 3) == 0 && lane == 0) bulk_wait0();  // every output tile has landed before the CTA retires
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace attn

template <int POLY, int FMT, bool TRACE>
static int launch_attn_t(const CUtensorMap& tm, const CUtensorMap& o, int grid, int tokens, int items, int qpairs, int stagger,
                       int half_keys, long long* trace, int trace_items, cudaStream_t stream) {
  static unsigned long long attr_done = 0;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn::attn_kernel<POLY, FMT, TRACE>), attn::SMEM_BYTES, &attr_done))
    return rc;
  ProfScope prof(ZK_K_ATTENTION, stream);
  attn::attn_kernel<POLY, FMT, TRACE><<<grid, attn::THREADS, attn::SMEM_BYTES, stream>>>(tm, o, tokens, items, qpairs, stagger,
                                                                               half_keys, trace, trace_items);
  ZK_LAUNCH_CHECK("attn_kernel");
  return 0;
}

template <int POLY, int FMT>
static int launch_attn(const CUtensorMap& tm, const CUtensorMap& o, int grid, int tokens, int items, int qpairs, int stagger,
                       int half_keys, long long* trace, int trace_items, cudaStream_t stream) {
  // the traced build only exists for the default exp2 split (scripts/attn_trace.py, scripts/attn_items.py)
  if (trace && POLY == 1)
    return launch_attn_t<1, FMT, true>(tm, o, grid, tokens, items, qpairs, stagger, half_keys, trace, trace_items, stream);
  return launch_attn_t<POLY, FMT, false>(tm, o, grid, tokens, items, qpairs, stagger, half_keys, nullptr, 0, stream);
}

int attention16_impl(const void* qkv, void* out, int batch, int tokens, int fmt, long long* trace, cudaStream_t stream) {
  using namespace attn;
  int rc = device_check();
  if (rc) return rc;
  if (!qkv || !out || batch <= 0 || tokens <= 0) {
    set_error("attention16: null pointer or empty shape");
    return ZK_ERR_ARG;
  }
  if (fmt != FMT_F16 && fmt != FMT_BF16) {
    set_error("attention16: unknown operand format %d", fmt);
    return ZK_ERR_ARG;
  }
  const int qpairs = (tokens + QTILES * BQ - 1) / (QTILES * BQ);
  const long long items = (long long)qpairs * HEADS * batch;
  if (items > 0x7fffffffLL || (long long)batch * tokens > 0x7fffffffLL) {
    set_error("attention16: batch %d x tokens %d is too large", batch, tokens);
    return ZK_ERR_SHAPE;
  }
  // share of exp2 evaluated on the FMA pipe: 0, 1 or 2 of every 4 pairs (ZK_ATTN_POLY), start offset of tile B
  // (ZK_ATTN_STAGGER); read once (thread-safe function-local statics)
  static const int poly = [] {
    const char* e = getenv("ZK_ATTN_POLY");
    const int v = e ? atoi(e) : 1;
    return (v < 0 || v > 2) ? 1 : v;
  }();
  static const int stagger = [] {
    const char* e = getenv("ZK_ATTN_STAGGER");
    return e ? atoi(e) : 1;
  }();
  static const int half_keys = (getenv("ZK_ATTN_HALF") && atoi(getenv("ZK_ATTN_HALF")) == 0) ? 0 : BKV / 2;
  static const int trace_items = getenv("ZK_ATTN_TRACE_ITEMS") ? atoi(getenv("ZK_ATTN_TRACE_ITEMS")) : 0;
  CUtensorMap tm;
  if ((rc = make_tmap_bf16_2d(&tm, qkv, (uint64_t)batch * tokens, 3 * HID, 3 * HID, 128, 64))) return rc;
  const int grid = items < num_sms() ? (int)items : num_sms();
  CUtensorMap o;
  if ((rc = make_tmap_bf16_3d(&o, out, (uint64_t)batch, (uint64_t)tokens, HID, HID, (uint64_t)tokens * HID, 128, 64))) return rc;
  const bool f16 = fmt == FMT_F16;
#define ZK_ATTN_LAUNCH(P)                                                                                                     \
  return f16 ? launch_attn<P, FMT_F16>(tm, o, grid, tokens, (int)items, qpairs, stagger, half_keys, trace, trace_items, stream) \
             : launch_attn<P, FMT_BF16>(tm, o, grid, tokens, (int)items, qpairs, stagger, half_keys, trace, trace_items, stream)
  if (poly == 0) { ZK_ATTN_LAUNCH(0); }
  if (poly == 2) { ZK_ATTN_LAUNCH(2); }
  ZK_ATTN_LAUNCH(1);
#undef ZK_ATTN_LAUNCH
}

int attention16(const void* qkv, void* out, int batch, int tokens, int fmt, cudaStream_t stream) {
  return attention16_impl(qkv, out, batch, tokens, fmt, nullptr, stream);
}

}  // namespace zk

extern "C" int zk_attention_trace(const void* d_qkv, void* d_out, int batch, int tokens, int64_t* d_trace, zk_stream_t stream) {
  return zk::attention16_impl(d_qkv, d_out, batch, tokens, ZK_FMT_BF16, reinterpret_cast<long long*>(d_trace),
                              (cudaStream_t)stream);
}

extern "C" int zk_attention16(const void* d_qkv, void* d_out, int batch, int tokens, int operand_format, zk_stream_t stream) {
  return zk::attention16(d_qkv, d_out, batch, tokens, operand_format, (cudaStream_t)stream);
}

extern "C" int zk_attention_bf16(const void* d_qkv, void* d_out, int batch, int tokens, zk_stream_t stream) {
  return zk::attention16(d_qkv, d_out, batch, tokens, ZK_FMT_BF16, (cudaStream_t)stream);
}
