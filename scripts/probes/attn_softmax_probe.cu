// Tuning probe (not part of the product library): the softmax instruction stream of zk_attn.cu in isolation.
// NSOFT softmax warps per SM run softmax_block<...> back to back on scores that sit in tensor memory; there is no
// MMA, no TMA and nobody to wait for, so the result is the ceiling of exp2 / clk / SM that this instruction stream can
// reach with that many warps per scheduler.  Built by scripts/attn_probe.py into build/libzk_attn_probe.so.
#include "../../zenker_audio_detection_b200/csrc/zk_attn.cu"

namespace zk {
namespace attn {

template <int POLY, int FMT, int W, int NSOFT>
__global__ void __launch_bounds__(128 + NSOFT * 32, 1) softmax_probe_kernel(int iters, long long* clocks, int prefetch, int mma_mode, int mma_idle) {
  constexpr int REGS_S = NSOFT == 8 ? 216 : (NSOFT == 12 ? 152 : 112);
  constexpr int REGS_O = NSOFT == 16 ? 24 : 56;
  // mma_mode: background tensor work issued by warp 1 while the softmax warps run (0 none, 1 SS-form 128x128x64 like
  // S = Q K^T, 2 TS-form 128x64x128 like O += P V with P read from tensor memory, 3 both in turn); after every batch
  // the issuer waits for completion and then idles `mma_idle` ns
  extern __shared__ __align__(1024) uint8_t opnd[];  // two 128 x 64 16-bit tiles (contents do not matter)
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 256);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[3], 1);
    stop_flag = 0;
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_O));
    if (warp == 1 && mma_mode) {
      constexpr uint32_t IDESC_S = umma_idesc_16(FMT, 128, 128, 0, 0), IDESC_O = umma_idesc_16(FMT, 128, 64, 0, 1);
      const uint32_t d = tmem_base + 384;  // columns 384..511 are not used by the softmax groups (W = 128: 2 x 192)
      const uint32_t a_p = tmem_base + 128;  // the P region of group 0
      uint32_t ph = 0;
      long long batches = 0;
      while (!stop_flag) {
        if (elect_one()) {
          const uint64_t qd = umma_desc_sw128(smem_u32(opnd), 16, 1024), kd = umma_desc_sw128(smem_u32(opnd + 16384), 16, 1024);
          if (mma_mode & 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ss(d, qd + 2 * k, kd + 2 * k, IDESC_S, k != 0);
          }
          if (mma_mode & 2) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16_ts(d, a_p + k * 8, umma_desc_sw128(smem_u32(opnd + 16384) + k * 16 * 128, 1024, 1024), IDESC_O, k != 0);
          }
          umma_commit(&bars[3]);
        }
        __syncwarp();
        mbar_wait(&bars[3], ph & 1);
        ++ph;
        ++batches;
        if (mma_idle) __nanosleep(mma_idle);
      }
      if (lane == 0) clocks[(long long)gridDim.x * 32 + blockIdx.x] = batches;
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_S));
    const int g = (warp - 4) >> 2, quarter = warp & 3;
    constexpr int GROUP_COLS = W + W / 2;  // S | P (O only matters on the rescale path, which the probe never takes)
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + g * GROUP_COLS;
    const uint32_t t_s = t_lane, t_p = t_lane + W, t_o = t_lane + W + W / 2;
    {  // finite scores below the running maximum: the rescale path is never taken
      uint32_t v[32];
#pragma unroll
      for (int c = 0; c < W / 32; ++c) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(-0.25f * (float)((lane + i + c) & 31));
        tmem_st32(t_s + c * 32, v);
      }
      tmem_st_wait();
    }
    SoftmaxState st;
    st.m = 1.0f;
    st.l2a = make_float2(0.f, 0.f);
    st.l2b = make_float2(0.f, 0.f);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      softmax_block<false, false, POLY, FMT, W>(0, W, true, t_s, t_o, t_p, smem_u32(&bars[0]), smem_u32(&bars[1]), st, nullptr);
      tc_fence_before();
      mbar_arrive(&bars[0]);
    }
    tmem_ld_wait();
    const long long t1 = clock64();
    stop_flag = 1;
    if (lane == 0) {
      clocks[((long long)blockIdx.x * 16 + (warp - 4)) * 2] = t0;
      clocks[((long long)blockIdx.x * 16 + (warp - 4)) * 2 + 1] = t1;
    }
    if (st.l2a.x + st.l2a.y + st.l2b.x + st.l2b.y == 12345.f) clocks[0] = 0;  // keep the sums alive
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// Raw throughput of the exponential instruction forms: MODE 0 ex2.approx.ftz.f32, 1 ex2.approx.ftz.f16x2, 2 ex2.approx.ftz.bf16x2.
// 8 independent chains per thread, `warps` warps per CTA, one CTA per SM.
template <int MODE>
__global__ void mufu_probe_kernel(int iters, long long* clocks, uint32_t seed) {
  uint32_t v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + threadIdx.x * 8 + i;
  if (MODE == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(-1.0f / (float)(1 + (v[i] & 1023)));
  }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(v[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[i]));
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= v[i];
  if ((threadIdx.x & 31) == 0) {
    clocks[((long long)blockIdx.x * 32 + (threadIdx.x >> 5)) * 2] = t0;
    clocks[((long long)blockIdx.x * 32 + (threadIdx.x >> 5)) * 2 + 1] = t1 + (acc == 0x12345u ? 1 : 0);
  }
}

}  // namespace attn
}  // namespace zk

extern "C" int zk_mufu_probe(int mode, int warps, int iters, long long* d_clocks, int grid, void* stream) {
  using namespace zk::attn;
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == 0) mufu_probe_kernel<0><<<grid, warps * 32, 0, s>>>(iters, d_clocks, 7u);
  if (mode == 1) mufu_probe_kernel<1><<<grid, warps * 32, 0, s>>>(iters, d_clocks, 0x30003000u);
  if (mode == 2) mufu_probe_kernel<2><<<grid, warps * 32, 0, s>>>(iters, d_clocks, 0x30003000u);
  return (int)cudaGetLastError();
}

// variant = POLY (0..2); w = 64 | 128 key columns per block; nsoft = 8 | 12 | 16 softmax warps; fp16 P
extern "C" int zk_attn_softmax_probe(int poly, int w, int nsoft, int iters, long long* d_clocks, int grid, int prefetch,
                                     int mma_mode, int mma_idle, void* stream) {
  using namespace zk::attn;
  cudaStream_t s = (cudaStream_t)stream;
#define PROBE(P, W_, N_)                                                                             \
  if (poly == P && w == W_ && nsoft == N_) {                                                         \
    cudaFuncSetAttribute(softmax_probe_kernel<P, zk::FMT_F16, W_, N_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 33792); \
    softmax_probe_kernel<P, zk::FMT_F16, W_, N_><<<grid, 128 + N_ * 32, 33792, s>>>(iters, d_clocks, prefetch, mma_mode, mma_idle);      \
    return (int)cudaGetLastError();                                                                  \
  }
  PROBE(0, 128, 8) PROBE(1, 128, 8) PROBE(2, 128, 8)
  PROBE(0, 64, 8) PROBE(1, 64, 8) PROBE(2, 64, 8)
  PROBE(0, 64, 12) PROBE(1, 64, 12) PROBE(2, 64, 12)
#undef PROBE
  return -1;
}
