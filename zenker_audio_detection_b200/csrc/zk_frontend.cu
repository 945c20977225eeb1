// Audio front end for sm_100a: polyphase resampler and the Kaldi-compatible log-mel filterbank.
//
// fbank kernel: one persistent CTA per SM (14 warps) walks tiles of 56 consecutive frames.  The 9200 samples a tile
// needs are staged into shared memory with ONE 1-D bulk async copy (cp.async.bulk -> UBLKCP, the TMA engine) that
// completes on an mbarrier; the copy of tile i+1 is issued before tile i is processed (double buffer), so HBM reads are
// contiguous 36 KiB bursts and each sample is fetched from HBM once although frames overlap 2.5x.  A warp owns four
// frames: each 16-lane half processes TWO frames at once, packed in the two halves of f32x2 registers, so every fp32
// operation of the pipeline is a packed FADD2 / FMUL2 / FFMA2 (the fp32 peak of sm_100 is only reachable through the
// packed forms, and at ~14 flop/B this kernel sits at the fp32 ridge, not the HBM one).  DC removal, pre-emphasis and
// the window are applied in registers; the 512-point real FFT is a 256-point complex FFT (two in-lane DFT-16 passes
// around one padded shared-memory transpose); then the real-FFT split, the power spectrum, the mel bank in segment
// form (255 taps per frame instead of the reference's dense 257x128 matmul), log, and the optional (x-mean)/(2 std).
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "zk_b200.h"
#include "zk_common.cuh"
#include "zk_fbank_math.cuh"
#include "zk_internal.cuh"

struct zk_fbank_plan {
  float2* d_win2;      // [400] window, each value duplicated for the two packed frames
  float4* d_tw;        // [16 k1][16 lanes] W256^(n2 k1) as (re, re, im, im)
  float4* d_w512;      // [8 j][16 lanes]   W512^(L + 16 j) as (re, re, im, im)
  float4* d_segw;      // [taps][16 lanes]  (lo, lo, hi, hi) mel weights x 1/4
  int* d_seg_start;    // [128]
  int glen[zk::fb::SEG_GROUPS], goff[zk::fb::SEG_GROUPS + 1];
  float preemph, log_floor, log_of_floor;
  int device;
};

namespace zk {
namespace fbk {
using namespace fb;
typedef cpxv<float2> cpx2;
static_assert(sizeof(cpx2) == 16, "packed complex pair is one 16-byte shared-memory element");

constexpr int WARPS = 14, THREADS = WARPS * 32, TILE_FRAMES = 4 * WARPS;
constexpr int TILE_SAMPLES = (TILE_FRAMES - 1) * SHIFT + FRAME;  // 9200
constexpr int UNIT_SCRATCH = 16 * TPITCH * 4;                    // floats per 16-lane half: the transpose buffer
static_assert(UNIT_SCRATCH >= 2 * 512, "exchange + power buffers alias the transpose buffer");
// shared memory carve-up (floats)
constexpr int OFF_SAMPLES = 0;                               // [2][TILE_SAMPLES] landing buffers of the bulk copies
constexpr int OFF_WIN = OFF_SAMPLES + 2 * TILE_SAMPLES;      // window taps, each duplicated (f32x2)
constexpr int OFF_TW = OFF_WIN + 2 * FRAME;
constexpr int OFF_W512 = OFF_TW + 16 * 16 * 4;
constexpr int OFF_SEGW = OFF_W512 + 8 * 16 * 4;
constexpr int OFF_SEGSTART = OFF_SEGW + SEG_TAPS_MAX * 16 * 4;
constexpr int OFF_SCRATCH = OFF_SEGSTART + NMEL;
constexpr int OFF_BAR = OFF_SCRATCH + 2 * WARPS * UNIT_SCRATCH;
constexpr int SMEM_BYTES = (OFF_BAR + 4) * 4;
static_assert(OFF_WIN % 4 == 0 && OFF_TW % 4 == 0 && OFF_W512 % 4 == 0 && OFF_SEGW % 4 == 0 && OFF_SCRATCH % 4 == 0 &&
                  UNIT_SCRATCH % 4 == 0 && OFF_BAR % 2 == 0,
              "16-byte alignment of the vector tables");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct Job {
  const float* wave;     // segment s starts at wave + s*src_pitch
  float* out;            // segment s, frame f -> out + (s*out_pitch + f)*128
  long long src_pitch, out_pitch;
  int seg_frames;        // frames computed per segment
  int tiles_per_seg;
  long long num_tiles;
  int normalize;
  float mean, std2;
  double* stats;         // optional: stats[0] += sum(v), stats[1] += sum(v^2) over every value computed (fp64)
  int store;             // 0: statistics only, nothing is written to `out`
};

// Scalar shared-memory load that the compiler cannot fuse with its neighbours: nvcc 12.9 turns the loads of
// x[m - 1], x[m], x[m + 1] (pre-emphasis neighbour and the sample pair) into LDS.64 / LDS.128 although m - 1 is odd,
// which faults with "misaligned address".
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
struct PairLoad {  // samples of two frames of the staged tile, packed; pair index i = samples 2 i, 2 i + 1
  uint32_t a;      // shared address of the first frame
  uint32_t db;     // byte offset of the second frame relative to the first (0 when it is a clamped duplicate)
  __device__ __forceinline__ float2 even(int i) const { return make_float2(lds_f32(a + 8 * i), lds_f32(a + 8 * i + db)); }
  __device__ __forceinline__ float2 odd(int i) const {
    return make_float2(lds_f32(a + 8 * i + 4), lds_f32(a + 8 * i + 4 + db));
  }
};
struct WindowLoad {  // window taps, each duplicated for the two packed frames
  const float2* w;
  __device__ __forceinline__ float2 even(int i) const { return w[2 * i]; }
  __device__ __forceinline__ float2 odd(int i) const { return w[2 * i + 1]; }
};

template <bool BULK>
__global__ void __launch_bounds__(THREADS, 1) fbank_kernel(const zk_fbank_plan plan, const Job job) {
  extern __shared__ __align__(16) float sm[];
  float* samples = sm + OFF_SAMPLES;
  float2* win2 = reinterpret_cast<float2*>(sm + OFF_WIN);
  cpx2* tw = reinterpret_cast<cpx2*>(sm + OFF_TW);
  cpx2* w512 = reinterpret_cast<cpx2*>(sm + OFF_W512);
  float4* segw = reinterpret_cast<float4*>(sm + OFF_SEGW);
  int* seg_start = reinterpret_cast<int*>(sm + OFF_SEGSTART);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + OFF_BAR);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = lane & 15, unit = lane >> 4;
  float* scratch = sm + OFF_SCRATCH + (warp * 2 + unit) * UNIT_SCRATCH;
  cpx2* tbuf = reinterpret_cast<cpx2*>(scratch);              // [16 k1][TPITCH]   transpose
  cpx2* xbuf = reinterpret_cast<cpx2*>(scratch);              // [8][16]           upper half of Z for the partner lane
  float2* pbuf = reinterpret_cast<float2*>(scratch + 512);    // [256]             power spectrum
  float2* hbuf = reinterpret_cast<float2*>(scratch);          // [128]             rising-slope sums for the next filter

  for (int i = tid; i < FRAME; i += THREADS) win2[i] = plan.d_win2[i];
  for (int i = tid; i < 16 * 16; i += THREADS) reinterpret_cast<float4*>(tw)[i] = plan.d_tw[i];
  for (int i = tid; i < 8 * 16; i += THREADS) reinterpret_cast<float4*>(w512)[i] = plan.d_w512[i];
  for (int i = tid; i < SEG_TAPS_MAX * 16; i += THREADS) segw[i] = plan.d_segw[i];
  for (int i = tid; i < NMEL; i += THREADS) seg_start[i] = plan.d_seg_start[i];
  if (BULK && tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();

  auto tile_info = [&](long long tile, const float*& src, int& nf, long long& out_row) {
    const long long seg = tile / job.tiles_per_seg;
    const int f0 = (int)(tile - seg * job.tiles_per_seg) * TILE_FRAMES;
    nf = min(TILE_FRAMES, job.seg_frames - f0);
    src = job.wave + seg * job.src_pitch + (long long)f0 * SHIFT;
    out_row = seg * job.out_pitch + f0;
  };

  if (BULK && tid == 0 && blockIdx.x < job.num_tiles) {
    const float* src;
    int nf;
    long long orow;
    tile_info(blockIdx.x, src, nf, orow);
    const uint32_t bytes = (uint32_t)((nf - 1) * SHIFT + FRAME) * 4u;
    mbar_arrive_expect_tx(&bar[0], bytes);
    bulk_load_1d(samples, src, bytes, &bar[0]);
  }

  double st_sum = 0.0, st_sq = 0.0;  // this thread's share of the dataset statistics (job.stats)
  int it = 0;
  for (long long tile = blockIdx.x; tile < job.num_tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    float ts = 0.f, tq = 0.f;  // fp32 within a tile (a thread adds at most a few hundred values), fp64 across tiles
    const float* src;
    int nf;
    long long out_row;
    tile_info(tile, src, nf, out_row);
    float* xs_tile = samples + buf * TILE_SAMPLES;
    if (BULK) {
      if (tid == 0 && tile + gridDim.x < job.num_tiles) {
        const float* nsrc;
        int nnf;
        long long norow;
        tile_info(tile + gridDim.x, nsrc, nnf, norow);
        const uint32_t bytes = (uint32_t)((nnf - 1) * SHIFT + FRAME) * 4u;
        mbar_arrive_expect_tx(&bar[buf ^ 1], bytes);
        bulk_load_1d(samples + (buf ^ 1) * TILE_SAMPLES, nsrc, bytes, &bar[buf ^ 1]);
      }
      mbar_wait(&bar[buf], (it >> 1) & 1);
    } else {
      const int ns = (nf - 1) * SHIFT + FRAME;
      for (int i = tid; i < ns; i += THREADS) xs_tile[i] = __ldg(src + i);
      __syncthreads();
    }

    // this half-warp's two frames (clamped to the last frame of a partial tile; clamped frames are not stored)
    const int fa = warp * 4 + unit * 2, fb = fa + 1;
    const int ca = min(fa, nf - 1), cb = min(fb, nf - 1);
    if (warp * 4 < nf) {  // warp-uniform: at least one live frame in this warp
      PairLoad ld{smem_u32(xs_tile + ca * SHIFT), (uint32_t)((cb - ca) * SHIFT * 4)};
      cpx2 z[16];
      {
        float2 x0[13], x1[13], xp[13];
        float2 s = lane_load<float2>(ld, L, x0, x1, xp);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
          s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
        }
        const float2 mean = make_float2(__fdiv_rn(s.x, (float)FRAME), __fdiv_rn(s.y, (float)FRAME));
        lane_stage1<float2>(x0, x1, xp, mean, plan.preemph, WindowLoad{win2}, L, tw + L, 16, z);
      }
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) tbuf[k1 * TPITCH + L] = z[k1];
      __syncwarp();
#pragma unroll
      for (int n2 = 0; n2 < 16; ++n2) z[n2] = tbuf[L * TPITCH + n2];
      dft16(z);  // z[k2] = Z[L + 16 k2]
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) xbuf[i * 16 + L] = z[8 + i];
      __syncwarp();
      {
        const int partner = (16 - L) & 15;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const cpx2 other = xbuf[(7 - j) * 16 + partner];      // Z[256 - k] for L != 0
          const cpx2 own = z[(16 - j) & 15];                    // lane 0: Z[256 - 16 j] (j = 0: Z[0] itself)
          const cpx2 b = (L == 0) ? own : other;
          const cpx2 w = w512[j * 16 + L];
          float2 pk, pnk;
          split_power<float2>(z[j], b, w.re, w.im, pk, pnk);
          pbuf[L + 16 * j] = pk;
          if (L + 16 * j != 0) pbuf[NZ - L - 16 * j] = pnk;
        }
        if (L == 0) {
          const float2 q = vfma(z[8].re, z[8].re, vmul(z[8].im, z[8].im));
          pbuf[128] = make_float2(4.0f * q.x, 4.0f * q.y);
        }
      }
      __syncwarp();
      float2 lo[SEG_GROUPS], hi[SEG_GROUPS];
#pragma unroll
      for (int i = 0; i < SEG_GROUPS; ++i) {
        const int st = seg_start[L + 16 * i];
        float2 al = make_float2(0.f, 0.f), ah = make_float2(0.f, 0.f);
        const int g0 = plan.goff[i], gl = plan.glen[i];
        for (int t = 0; t < gl; ++t) {
          const float2 p = pbuf[min(st + t, NZ - 1)];
          const float4 w = segw[(g0 + t) * 16 + L];
          al = vfma(p, make_float2(w.x, w.y), al);
          ah = vfma(p, make_float2(w.z, w.w), ah);
        }
        lo[i] = al;
        hi[i] = ah;
      }
      // xbuf (aliased by hbuf) was last read before the previous __syncwarp
#pragma unroll
      for (int i = 0; i < SEG_GROUPS; ++i) hbuf[L + 16 * i] = hi[i];
      __syncwarp();
      float* dst_a = job.out + (out_row + fa) * NMEL + L;
      float* dst_b = job.out + (out_row + fb) * NMEL + L;
      const bool live_a = fa < nf, live_b = fb < nf;
#pragma unroll
      for (int i = 0; i < SEG_GROUPS; ++i) {
        const int r = L + 16 * i;
        float2 e = lo[i];
        if (r > 0) e = vadd(e, hbuf[r - 1]);
        float va = e.x > plan.log_floor ? __logf(e.x) : plan.log_of_floor;
        float vb = e.y > plan.log_floor ? __logf(e.y) : plan.log_of_floor;
        if (job.normalize) {
          va = __fdiv_rn(__fsub_rn(va, job.mean), job.std2);
          vb = __fdiv_rn(__fsub_rn(vb, job.mean), job.std2);
        }
        if (live_a && job.store) dst_a[16 * i] = va;
        if (live_b && job.store) dst_b[16 * i] = vb;
        if (job.stats) {
          if (live_a) {
            ts += va;
            tq = fmaf(va, va, tq);
          }
          if (live_b) {
            ts += vb;
            tq = fmaf(vb, vb, tq);
          }
        }
      }
    }
    st_sum += (double)ts;
    st_sq += (double)tq;
    __syncthreads();  // every warp is done with xs_tile (and its scratch) before the tile buffer is refilled
  }
  if (job.stats) {  // utils/compute_ast_normalization_stats.py:77-80 as an epilogue: one fp64 atomic pair per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      st_sum += __shfl_xor_sync(0xffffffffu, st_sum, o);
      st_sq += __shfl_xor_sync(0xffffffffu, st_sq, o);
    }
    if (lane == 0) {
      atomicAdd(job.stats, st_sum);
      atomicAdd(job.stats + 1, st_sq);
    }
  }
}

__global__ void fill_pad_kernel(float* __restrict__ out, int batch, int first_row, int max_length, float value) {
  const long long per = (long long)(max_length - first_row) * (NMEL / 4);
  const long long total = per * batch;
  const float4 v = make_float4(value, value, value, value);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per, r = i - b * per;
    reinterpret_cast<float4*>(out + (b * max_length + first_row) * NMEL)[r] = v;
  }
}

static int launch_fbank(const zk_fbank_plan* plan, Job job, cudaStream_t stream) {
  static unsigned long long attr_bulk = 0, attr_plain = 0;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(fbank_kernel<true>), SMEM_BYTES, &attr_bulk)) return rc;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(fbank_kernel<false>), SMEM_BYTES, &attr_plain)) return rc;
  if (job.num_tiles <= 0) return 0;
  const bool bulk = (reinterpret_cast<uintptr_t>(job.wave) % 16 == 0) && (job.src_pitch % 4 == 0);
  long long grid = job.num_tiles < (long long)num_sms() ? job.num_tiles : (long long)num_sms();
  ProfScope prof(ZK_K_FBANK, stream);
  if (bulk)
    fbank_kernel<true><<<(int)grid, THREADS, SMEM_BYTES, stream>>>(*plan, job);
  else
    fbank_kernel<false><<<(int)grid, THREADS, SMEM_BYTES, stream>>>(*plan, job);
  ZK_LAUNCH_CHECK("fbank_kernel");
  return 0;
}
}  // namespace fbk

// ------------------------------------------------------------------------------------------------ resample
namespace rs {
// Decimating fast path (new == 1): out[i] = sum_k taps[k] * x[i*ORIG - WIDTH + k].  Each thread produces R
// consecutive outputs from a register window of (R-1)*ORIG + KLEN inputs; inputs (channel mean already applied)
// and outputs are staged in shared memory so global traffic is fully coalesced.
template <typename T>
__device__ __forceinline__ float load_mean(const T* in, long long g, int channels, long long ch_pitch);
template <>
__device__ __forceinline__ float load_mean<float>(const float* in, long long g, int channels, long long ch_pitch) {
  float a = __ldg(in + g);
  if (channels == 1) return a;
  for (int c = 1; c < channels; ++c) a += __ldg(in + c * ch_pitch + g);
  return __fdiv_rn(a, (float)channels);
}
template <>
__device__ __forceinline__ float load_mean<int16_t>(const int16_t* in, long long g, int channels, long long) {
  float a = (float)in[g * channels] * (1.0f / 32768.0f);
  if (channels == 1) return a;
  for (int c = 1; c < channels; ++c) a += (float)in[g * channels + c] * (1.0f / 32768.0f);
  return __fdiv_rn(a, (float)channels);
}

// Persistent CTAs walk tiles of OB = 256 R outputs.  The ORIG*OB + KLEN inputs of a tile are staged in shared memory
// by ONE 1-D bulk async copy (UBLKCP, the TMA engine) on an mbarrier, issued one tile ahead into the other buffer, so
// HBM reads are 21 KiB bursts that overlap the FIR of the current tile; boundary tiles, multi-channel and PCM16
// sources take a guarded coalesced load instead (channel mean / 2^-15 scaling fused there).  Each thread produces R
// consecutive outputs from a register window of (R-1) ORIG + KLEN inputs (thread stride ORIG R floats: odd for
// 48 kHz -> conflict-free LDS); results go back through shared memory and leave as one bulk store per tile.
constexpr int DEC_THREADS = 256;
template <typename T, int ORIG, int WIDTH, int R>
__global__ void __launch_bounds__(DEC_THREADS, 2) decimate_kernel(const T* __restrict__ in, long long n_in, int channels,
                                                                  long long ch_pitch, const float* __restrict__ taps,
                                                                  float* __restrict__ out, long long n_out, int bulk_ok) {
  constexpr int KLEN = 2 * WIDTH + ORIG, OB = DEC_THREADS * R, WIN = (R - 1) * ORIG + KLEN;
  constexpr int PAD = (4 - WIDTH % 4) % 4;                       // tile source starts at a multiple of 4 floats
  constexpr int NIN = (ORIG * (OB - 1) + PAD + KLEN + 3) / 4 * 4;  // floats staged per tile
  static_assert((ORIG * OB) % 4 == 0 && OB % 4 == 0, "16-byte aligned tiles");
  extern __shared__ __align__(16) float dsm[];
  float* xs0 = dsm;                 // [2][NIN]
  float* ys = dsm + 2 * NIN;        // [OB]
  float* tp = ys + OB;              // [KLEN]
  uint64_t* bar = reinterpret_cast<uint64_t*>(tp + (KLEN + 1) / 2 * 2);  // [2]
  const int tid = threadIdx.x;
  for (int i = tid; i < KLEN; i += DEC_THREADS) tp[i] = taps[i];
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const long long num_tiles = (n_out + OB - 1) / OB;
  // stage tile `tile` into buffer `b`: returns true when it was issued as an asynchronous bulk copy
  auto stage = [&](long long tile, int b) -> bool {
    float* xs = xs0 + b * NIN;
    const long long g0 = tile * OB * ORIG - WIDTH - PAD;
    if (bulk_ok && g0 >= 0 && g0 + NIN <= n_in) {
      if (tid == 0) {
        mbar_arrive_expect_tx(&bar[b], NIN * 4);
        bulk_load_1d(xs, reinterpret_cast<const float*>(in) + g0, NIN * 4, &bar[b]);
      }
      return true;
    }
    for (int i = tid; i < NIN; i += DEC_THREADS) {
      const long long g = g0 + i;
      xs[i] = (g >= 0 && g < n_in) ? load_mean<T>(in, g, channels, ch_pitch) : 0.f;
    }
    return false;
  };
  uint32_t phase[2] = {0, 0};
  bool async_cur = false;
  if ((long long)blockIdx.x < num_tiles) async_cur = stage(blockIdx.x, 0);
  int it = 0;
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const int b = it & 1;
    bool async_next = false;
    if (tile + gridDim.x < num_tiles) async_next = stage(tile + gridDim.x, b ^ 1);  // buffer b^1 was released by the
                                                                                  // __syncthreads of the last round
    if (async_cur) {
      mbar_wait(&bar[b], phase[b]);
      phase[b] ^= 1;
    } else {
      __syncthreads();
    }
    const float* xs = xs0 + b * NIN + tid * (R * ORIG) + PAD;
    float xv[WIN];
#pragma unroll
    for (int i = 0; i < WIN; ++i) xv[i] = xs[i];
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll
    for (int k = 0; k < KLEN; ++k) {
      const float t = tp[k];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fmaf(t, xv[r * ORIG + k], acc[r]);
    }
    if (tid == 0) bulk_wait_read0();  // the previous tile's bulk store has finished reading ys
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) ys[tid * R + r] = acc[r];
    const long long o0 = tile * OB;
    if (o0 + OB <= n_out && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + o0), "r"(smem_u32(ys)),
                     "r"(OB * 4)
                     : "memory");
        bulk_commit();
      }
    } else {
      __syncthreads();
      for (int i = tid; i < OB; i += DEC_THREADS)
        if (o0 + i < n_out) out[o0 + i] = ys[i];
    }
    async_cur = async_next;
  }
  if (tid == 0) bulk_wait0();
}

// 48 kHz -> 16 kHz, mono float32, 16-byte aligned source: the configuration the path runs on every recording.
// Same tiling as decimate_kernel (2048 outputs per tile, 1-D bulk copies one tile ahead, bulk store), but the FIR runs
// on packed pairs: each thread pulls its 64-sample window out of shared memory with 16 LDS.128 and produces 8 outputs;
// an output's 41 products are 20 FFMA2 over aligned register pairs (x[2i], x[2i+1]) against pre-paired taps plus one
// scalar FFMA -- outputs whose window starts on an odd sample use the tap pairs shifted by one.  22 FMA-class
// instructions per output instead of 41, which moves the kernel from the fp32 issue limit back to the HBM roofline.
constexpr int D3_R = 8, D3_OB = DEC_THREADS * D3_R;      // 2048 outputs per tile
constexpr int D3_NIN = 3 * D3_OB + 44;                   // staged inputs per tile: starts 20 samples before 3 o0
constexpr int D3_STAGES = 3;                             // input tiles in flight per CTA (two ahead)
constexpr int D3_SMEM = (D3_STAGES * D3_NIN + D3_OB + 2 * 42) * 4 + 8 * D3_STAGES;
static_assert(D3_NIN % 4 == 0, "bulk copy size");
__global__ void __launch_bounds__(DEC_THREADS, 2) decimate3_kernel(const float* __restrict__ in, long long n_in,
                                                                   const float* __restrict__ taps,
                                                                   float* __restrict__ out, long long n_out) {
  constexpr int KLEN = 41, WIDTH = 19;
  extern __shared__ __align__(16) float dsm[];
  float* xs0 = dsm;                       // [D3_STAGES][D3_NIN]
  float* ys = dsm + D3_STAGES * D3_NIN;   // [D3_OB]
  float2* tp_a = reinterpret_cast<float2*>(ys + D3_OB);  // [21] (t0,t1) (t2,t3) .. (t38,t39) (t40,0)
  float2* tp_b = tp_a + 21;                              // [21] (0,t0) (t1,t2) .. (t39,t40)
  uint64_t* bar = reinterpret_cast<uint64_t*>(tp_b + 21);
  const int tid = threadIdx.x;
  if (tid < 21) {
    tp_a[tid] = make_float2(taps[2 * tid], 2 * tid + 1 < KLEN ? taps[2 * tid + 1] : 0.f);
    tp_b[tid] = make_float2(tid ? taps[2 * tid - 1] : 0.f, taps[2 * tid]);
  }
  if (tid == 0) {
    for (int i = 0; i < D3_STAGES; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const long long num_tiles = (n_out + D3_OB - 1) / D3_OB;
  auto stage = [&](long long tile, int b) -> bool {
    float* xs = xs0 + b * D3_NIN;
    const long long g0 = tile * D3_OB * 3 - WIDTH - 1;  // multiple of 4 samples
    if (g0 >= 0 && g0 + D3_NIN <= n_in) {
      if (tid == 0) {
        mbar_arrive_expect_tx(&bar[b], D3_NIN * 4);
        bulk_load_1d(xs, in + g0, D3_NIN * 4, &bar[b]);
      }
      return true;
    }
    for (int i = tid; i < D3_NIN; i += DEC_THREADS) {
      const long long g = g0 + i;
      xs[i] = (g >= 0 && g < n_in) ? __ldg(in + g) : 0.f;
    }
    return false;
  };
  uint32_t phase = 0;        // bit b: parity of the next completion of bar[b]
  uint32_t is_async = 0;     // bit b: the tile in buffer b arrives by bulk copy
  for (int p = 0; p < D3_STAGES - 1; ++p)
    if (blockIdx.x + (long long)p * gridDim.x < num_tiles && stage(blockIdx.x + (long long)p * gridDim.x, p)) is_async |= 1u << p;
  int it = 0;
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const int b = it % D3_STAGES;
    {  // buffer (it + 2) % 3 was released by the __syncthreads of the previous round
      const int nb = (it + D3_STAGES - 1) % D3_STAGES;
      const long long nt = tile + (long long)(D3_STAGES - 1) * gridDim.x;
      is_async &= ~(1u << nb);
      if (nt < num_tiles && stage(nt, nb)) is_async |= 1u << nb;
    }
    if (is_async >> b & 1) {
      mbar_wait(&bar[b], phase >> b & 1);
      phase ^= 1u << b;
    } else {
      __syncthreads();
    }
    // window of this thread: staged samples [24 tid, 24 tid + 64); output r uses samples 1 + 3 r + k, k < 41
    float2 xv[32];
    {
      const float4* src = reinterpret_cast<const float4*>(xs0 + b * D3_NIN + tid * (3 * D3_R));
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 v = src[i];
        xv[2 * i] = make_float2(v.x, v.y);
        xv[2 * i + 1] = make_float2(v.z, v.w);
      }
    }
    float2 acc[D3_R];
#pragma unroll
    for (int r = 0; r < D3_R; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 21; ++j) {
      const float2 ta = tp_a[j], tb = tp_b[j];
#pragma unroll
      for (int r = 0; r < D3_R; ++r) {
        // first sample of output r: 1 + 3 r.  Odd (r even): pair index (1 + 3 r - 1) / 2 + j against (0,t0),(t1,t2)..
        // Even (r odd): pair index (1 + 3 r) / 2 + j against (t0,t1),(t2,t3)..,(t40,0)
        if (r & 1)
          acc[r] = ffma2(xv[(1 + 3 * r) / 2 + j], ta, acc[r]);
        else
          acc[r] = ffma2(xv[(3 * r) / 2 + j], tb, acc[r]);
      }
    }
    if (tid == 0) bulk_wait_read0();  // the previous tile's bulk store has finished reading ys
    __syncthreads();
#pragma unroll
    for (int r = 0; r < D3_R; ++r) ys[tid * D3_R + r] = acc[r].x + acc[r].y;
    const long long o0 = tile * D3_OB;
    if (o0 + D3_OB <= n_out && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + o0), "r"(smem_u32(ys)),
                     "r"(D3_OB * 4)
                     : "memory");
        bulk_commit();
      }
    } else {
      __syncthreads();
      for (int i = tid; i < D3_OB; i += DEC_THREADS)
        if (o0 + i < n_out) out[o0 + i] = ys[i];
    }
  }
  if (tid == 0) bulk_wait0();
}

// Generic ratio: one thread per output sample (any orig/new, e.g. 44.1 kHz -> 16 kHz = 441/160).
template <typename T>
__global__ void generic_kernel(const T* __restrict__ in, long long n_in, int channels, long long ch_pitch,
                               const float* __restrict__ taps, int orig, int new_, int width, float* __restrict__ out,
                               long long n_out) {
  const int klen = 2 * width + orig;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += (long long)gridDim.x * blockDim.x) {
    const long long i = o / new_;
    const int j = (int)(o - i * new_);
    const long long g0 = i * orig - width;
    const float* tj = taps + (long long)j * klen;
    float acc = 0.f;
    for (int k = 0; k < klen; ++k) {
      const long long g = g0 + k;
      if (g >= 0 && g < n_in) acc = fmaf(__ldg(tj + k), load_mean<T>(in, g, channels, ch_pitch), acc);
    }
    out[o] = acc;
  }
}

template <typename T>
static int run(const T* in, long long n_in, int channels, long long ch_pitch, const float* taps, int orig, int new_,
               int width, float* out, long long n_out, cudaStream_t stream) {
  int rc = device_check();
  if (rc) return rc;
  if (n_in < 0 || n_out < 0 || channels < 1 || orig < 1 || new_ < 1 || width < 0 || (n_out > 0 && (!in || !out || !taps))) {
    set_error("resample: bad arguments");
    return ZK_ERR_ARG;
  }
  const long long expect = (n_in * new_ + orig - 1) / orig;
  if (n_out > expect) {
    set_error("resample: n_out %lld exceeds ceil(new*n_in/orig) = %lld", n_out, expect);
    return ZK_ERR_SHAPE;
  }
  if (n_out == 0) return 0;
  const int sms = num_sms();
#define ZK_DECIMATE(O, W, R)                                                                                   \
  if (new_ == 1 && orig == O && width == W) {                                                                  \
    constexpr int KLEN = 2 * W + O, OB = DEC_THREADS * R, PAD = (4 - W % 4) % 4;                                \
    constexpr int NIN = (O * (OB - 1) + PAD + KLEN + 3) / 4 * 4;                                                \
    constexpr int SMEM = (2 * NIN + OB + (KLEN + 1) / 2 * 2) * 4 + 16;                                          \
    static unsigned long long attr_done = 0;                                                                    \
    if (int rc_ = ensure_dynamic_smem(reinterpret_cast<const void*>(decimate_kernel<T, O, W, R>), SMEM, &attr_done)) \
      return rc_;                                                                                               \
    long long blocks = (n_out + OB - 1) / OB;                                                                   \
    if (blocks > 2LL * sms) blocks = 2LL * sms;                                                                 \
    const int bulk_ok = std::is_same<T, float>::value && channels == 1 && (reinterpret_cast<uintptr_t>(in) & 15) == 0; \
    ProfScope prof(ZK_K_RESAMPLE, stream);                                                                      \
    decimate_kernel<T, O, W, R><<<(int)blocks, DEC_THREADS, SMEM, stream>>>(in, n_in, channels, ch_pitch, taps, out, \
                                                                            n_out, bulk_ok);                   \
    ZK_LAUNCH_CHECK("decimate_kernel");                                                                        \
    return 0;                                                                                                  \
  }
  if (new_ == 1 && orig == 3 && width == 19 && std::is_same<T, float>::value && channels == 1 &&
      (reinterpret_cast<uintptr_t>(in) & 15) == 0) {  // 48 kHz mono float32: the packed-FIR kernel
    static unsigned long long attr3_done = 0;
    if ((rc = ensure_dynamic_smem(reinterpret_cast<const void*>(decimate3_kernel), D3_SMEM, &attr3_done))) return rc;
    long long blocks = (n_out + D3_OB - 1) / D3_OB;
    if (blocks > 2LL * sms) blocks = 2LL * sms;
    ProfScope prof(ZK_K_RESAMPLE, stream);
    decimate3_kernel<<<(int)blocks, DEC_THREADS, D3_SMEM, stream>>>(reinterpret_cast<const float*>(in), n_in, taps, out, n_out);
    ZK_LAUNCH_CHECK("decimate3_kernel");
    return 0;
  }
  ZK_DECIMATE(3, 19, 7)   // 48 kHz -> 16 kHz (stereo, PCM16, unaligned)
  ZK_DECIMATE(2, 13, 7)   // 32 kHz -> 16 kHz
  ZK_DECIMATE(6, 37, 5)   // 96 kHz -> 16 kHz
#undef ZK_DECIMATE
  long long blocks = (n_out + 255) / 256;
  if (blocks > 16LL * sms) blocks = 16LL * sms;
  ProfScope prof(ZK_K_RESAMPLE, stream);
  generic_kernel<T><<<(int)blocks, 256, 0, stream>>>(in, n_in, channels, ch_pitch, taps, orig, new_, width, out, n_out);
  ZK_LAUNCH_CHECK("resample generic_kernel");
  return 0;
}
}  // namespace rs
}  // namespace zk

extern "C" {

int zk_resample_f32(const float* d_in, int64_t n_in, int channels, int64_t ch_pitch, const float* d_taps, int orig,
                    int new_, int width, float* d_out, int64_t n_out, zk_stream_t stream) {
  return zk::rs::run<float>(d_in, n_in, channels, ch_pitch, d_taps, orig, new_, width, d_out, n_out, (cudaStream_t)stream);
}

int zk_resample_pcm16(const int16_t* d_in, int64_t n_in, int channels, const float* d_taps, int orig, int new_,
                      int width, float* d_out, int64_t n_out, zk_stream_t stream) {
  return zk::rs::run<int16_t>(d_in, n_in, channels, 0, d_taps, orig, new_, width, d_out, n_out, (cudaStream_t)stream);
}

int64_t zk_fbank_num_frames(int64_t n) { return n < zk::fb::FRAME ? 0 : 1 + (n - zk::fb::FRAME) / zk::fb::SHIFT; }

int zk_fbank_plan_create(const float* h_window, const float* h_mel, int num_mel, float preemph, float log_floor,
                         zk_fbank_plan** out) {
  using namespace zk::fb;
  if (!h_window || !h_mel || !out) {
    zk::set_error("zk_fbank_plan_create: null pointer");
    return ZK_ERR_ARG;
  }
  if (num_mel != NMEL) {
    zk::set_error("zk_fbank_plan_create: num_mel must be %d (got %d)", NMEL, num_mel);
    return ZK_ERR_SHAPE;
  }
  int rc = zk::device_check();
  if (rc) return rc;
  HostTables* t = new HostTables();
  const int bad = build_host_tables(h_mel, *t);
  if (bad) {
    delete t;
    zk::set_error("zk_fbank_plan_create: the mel bank is not triangular (FFT bin %d does not feed two adjacent filters)",
                  bad - 1);
    return ZK_ERR_SHAPE;
  }
  zk_fbank_plan* p = new zk_fbank_plan();
  memset(p, 0, sizeof(*p));
  p->preemph = preemph;
  p->log_floor = log_floor;
  p->log_of_floor = logf(log_floor);
  for (int i = 0; i < SEG_GROUPS; ++i) p->glen[i] = t->glen[i];
  for (int i = 0; i <= SEG_GROUPS; ++i) p->goff[i] = t->goff[i];
  cudaGetDevice(&p->device);
  // device layouts: every value duplicated for the two frames packed in f32x2
  float* win2 = new float[2 * FRAME];
  for (int i = 0; i < FRAME; ++i) win2[2 * i] = win2[2 * i + 1] = h_window[i];
  float* tw4 = new float[16 * 16 * 4];
  for (int i = 0; i < 16 * 16; ++i) {
    tw4[4 * i] = tw4[4 * i + 1] = t->tw[2 * i];
    tw4[4 * i + 2] = tw4[4 * i + 3] = t->tw[2 * i + 1];
  }
  float* w4 = new float[8 * 16 * 4];
  for (int i = 0; i < 8 * 16; ++i) {
    w4[4 * i] = w4[4 * i + 1] = t->w512[2 * i];
    w4[4 * i + 2] = w4[4 * i + 3] = t->w512[2 * i + 1];
  }
  float* s4 = new float[SEG_TAPS_MAX * 16 * 4];
  for (int i = 0; i < SEG_TAPS_MAX * 16; ++i) {
    s4[4 * i] = s4[4 * i + 1] = t->seg_w[2 * i];
    s4[4 * i + 2] = s4[4 * i + 3] = t->seg_w[2 * i + 1];
  }
  cudaError_t e = cudaSuccess;
  auto up = [&](void** dst, const void* src, size_t bytes) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(dst, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  up((void**)&p->d_win2, win2, 2 * FRAME * 4);
  up((void**)&p->d_tw, tw4, 16 * 16 * 16);
  up((void**)&p->d_w512, w4, 8 * 16 * 16);
  up((void**)&p->d_segw, s4, SEG_TAPS_MAX * 16 * 16);
  up((void**)&p->d_seg_start, t->seg_start, sizeof(t->seg_start));
  delete[] win2;
  delete[] tw4;
  delete[] w4;
  delete[] s4;
  delete t;
  if (e != cudaSuccess) {
    zk_fbank_plan_destroy(p);
    return zk::cuda_fail(e, "zk_fbank_plan_create upload");
  }
  *out = p;
  return 0;
}

void zk_fbank_plan_destroy(zk_fbank_plan* p) {
  if (!p) return;
  cudaFree(p->d_win2);
  cudaFree(p->d_tw);
  cudaFree(p->d_w512);
  cudaFree(p->d_segw);
  cudaFree(p->d_seg_start);
  delete p;
}

int zk_fbank_f32(const zk_fbank_plan* plan, const float* d_wave, int64_t n, float* d_out, int64_t m, zk_stream_t stream) {
  using namespace zk::fbk;
  int rc = zk::device_check();
  if (rc) return rc;
  if (!plan || n < 0 || m < 0 || (m > 0 && (!d_wave || !d_out))) {
    zk::set_error("zk_fbank_f32: bad arguments");
    return ZK_ERR_ARG;
  }
  {
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != plan->device) {  // the tables live on the device the plan was created on
      zk::set_error("%s: the plan belongs to device %d but the current device is %d", "zk_fbank_f32", plan->device, cur);
      return ZK_ERR_ARG;
    }
  }
  if (m > zk_fbank_num_frames(n)) {
    zk::set_error("zk_fbank_f32: m = %lld frames requested but %lld samples hold only %lld", (long long)m, (long long)n,
                  (long long)zk_fbank_num_frames(n));
    return ZK_ERR_SHAPE;
  }
  if (m > 2000000000LL) {
    zk::set_error("zk_fbank_f32: more than 2e9 frames per call");
    return ZK_ERR_SHAPE;
  }
  Job job;
  job.wave = d_wave;
  job.out = d_out;
  job.src_pitch = 0;
  job.out_pitch = 0;
  job.seg_frames = (int)m;
  job.tiles_per_seg = (int)((m + TILE_FRAMES - 1) / TILE_FRAMES);
  job.num_tiles = job.tiles_per_seg;
  job.normalize = 0;
  job.mean = 0.f;
  job.std2 = 1.f;
  job.stats = nullptr;
  job.store = 1;
  return launch_fbank(plan, job, (cudaStream_t)stream);
}

int zk_fx_contract_f32(const zk_fbank_plan* plan, const float* d_windows, int batch, int64_t win_len, int64_t win_pitch,
                       int do_normalize, float mean, float std, int max_length, float* d_out, zk_stream_t stream) {
  using namespace zk::fbk;
  int rc = zk::device_check();
  if (rc) return rc;
  if (!plan || batch < 0 || win_len < 0 || win_pitch < win_len || max_length <= 0 || (batch > 0 && (!d_windows || !d_out))) {
    zk::set_error("zk_fx_contract_f32: bad arguments");
    return ZK_ERR_ARG;
  }
  {
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != plan->device) {  // the tables live on the device the plan was created on
      zk::set_error("%s: the plan belongs to device %d but the current device is %d", "zk_fx_contract_f32", plan->device, cur);
      return ZK_ERR_ARG;
    }
  }
  if (batch == 0) return 0;
  const int64_t m = zk_fbank_num_frames(win_len);
  const int rows = (int)(m < max_length ? m : max_length);
  const float std2 = std * 2.0f;
  if (rows < max_length) {
    const float pad = do_normalize ? (0.0f - mean) / std2 : 0.0f;
    long long total = (long long)batch * (max_length - rows) * (NMEL / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 8LL * zk::num_sms()) blocks = 8LL * zk::num_sms();
    zk::ProfScope prof(ZK_K_MISC, (cudaStream_t)stream);
    fill_pad_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(d_out, batch, rows, max_length, pad);
    ZK_LAUNCH_CHECK("fill_pad_kernel");
  }
  if (rows == 0) return 0;
  Job job;
  job.wave = d_windows;
  job.out = d_out;
  job.src_pitch = win_pitch;
  job.out_pitch = max_length;
  job.seg_frames = rows;
  job.tiles_per_seg = (rows + TILE_FRAMES - 1) / TILE_FRAMES;
  job.num_tiles = (long long)job.tiles_per_seg * batch;
  job.normalize = do_normalize;
  job.mean = mean;
  job.std2 = std2;
  job.stats = nullptr;
  job.store = 1;
  return launch_fbank(plan, job, (cudaStream_t)stream);
}

int zk_fx_stats_f32(const zk_fbank_plan* plan, const float* d_windows, int batch, int64_t win_len, int64_t win_pitch,
                    int max_length, float* d_out, double* d_acc, zk_stream_t stream) {
  using namespace zk::fbk;
  int rc = zk::device_check();
  if (rc) return rc;
  if (!plan || batch < 0 || win_len < 0 || win_pitch < win_len || max_length <= 0 || !d_acc || (batch > 0 && !d_windows)) {
    zk::set_error("zk_fx_stats_f32: bad arguments");
    return ZK_ERR_ARG;
  }
  {
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != plan->device) {
      zk::set_error("zk_fx_stats_f32: the plan belongs to device %d but the current device is %d", plan->device, cur);
      return ZK_ERR_ARG;
    }
  }
  if (batch == 0) return 0;
  const int64_t m = zk_fbank_num_frames(win_len);
  const int rows = (int)(m < max_length ? m : max_length);
  if (d_out && rows < max_length) {  // zero padding (do_normalize = False): adds nothing to the sums
    long long total = (long long)batch * (max_length - rows) * (NMEL / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 8LL * zk::num_sms()) blocks = 8LL * zk::num_sms();
    zk::ProfScope prof(ZK_K_MISC, (cudaStream_t)stream);
    fill_pad_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(d_out, batch, rows, max_length, 0.0f);
    ZK_LAUNCH_CHECK("fill_pad_kernel");
  }
  if (rows == 0) return 0;
  Job job;
  job.wave = d_windows;
  job.out = d_out;
  job.src_pitch = win_pitch;
  job.out_pitch = max_length;
  job.seg_frames = rows;
  job.tiles_per_seg = (rows + TILE_FRAMES - 1) / TILE_FRAMES;
  job.num_tiles = (long long)job.tiles_per_seg * batch;
  job.normalize = 0;
  job.mean = 0.f;
  job.std2 = 1.f;
  job.stats = d_acc;
  job.store = d_out ? 1 : 0;
  return launch_fbank(plan, job, (cudaStream_t)stream);
}

}  // extern "C"
