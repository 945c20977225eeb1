// Audio front end for sm_100a: polyphase resampler and the Kaldi-compatible log-mel filterbank.
//
// fbank kernel: persistent CTAs walk tiles of 32 consecutive frames.  The 5360 samples a tile needs are
// staged into shared memory with ONE 1-D bulk async copy (cp.async.bulk -> UBLKCP, the TMA engine) that
// completes on an mbarrier; the copy of tile i+1 is issued before tile i is processed (double buffer), so
// HBM reads are contiguous 21 KiB bursts and each sample is fetched from HBM once although frames overlap
// 2.5x.  16 lanes own one frame (two frames per warp): DC removal, pre-emphasis and the window are applied
// in registers, the 512-point real FFT is a 256-point complex FFT (two in-lane DFT-16 passes around one
// padded shared-memory transpose), followed by the power spectrum, the sparse mel filterbank (<= 16 taps per
// bin instead of the reference's dense 257x128 matmul), log, and the optional (x-mean)/(2 std).
#include <string.h>

#include <type_traits>

#include "zk_b200.h"
#include "zk_common.cuh"
#include "zk_fbank_math.cuh"
#include "zk_internal.cuh"

struct zk_fbank_plan {
  float* d_win;       // [400]
  float* d_tw;        // [16 lanes][16] complex: W256^(n2*k1)
  float* d_w512;      // [128] complex
  int* d_mel_start;   // [128]
  float* d_mel_w;     // [MELW][128]
  int* d_group_len;   // [8]
  float preemph, log_floor;
  int device;
};

namespace zk {
namespace fbk {
using namespace fb;

constexpr int TILE_FRAMES = 32, WARPS = 8, THREADS = WARPS * 32;
constexpr int TILE_SAMPLES = (TILE_FRAMES - 1) * SHIFT + FRAME;  // 5360
constexpr int SCRATCH = ZBUF + PBUF;                             // floats per half-warp
// shared memory carve-up (floats)
constexpr int OFF_SAMPLES = 0;
constexpr int OFF_WIN = OFF_SAMPLES + 2 * TILE_SAMPLES;
constexpr int OFF_W512 = OFF_WIN + FRAME;
constexpr int OFF_MELW = OFF_W512 + 256;
constexpr int OFF_MELSTART = OFF_MELW + MELW * NMEL;
constexpr int OFF_GROUP = OFF_MELSTART + NMEL;
constexpr int OFF_SCRATCH = OFF_GROUP + 8;
constexpr int OFF_BAR = OFF_SCRATCH + 2 * WARPS * SCRATCH;
constexpr int SMEM_BYTES = (OFF_BAR + 4) * 4;
static_assert((OFF_SCRATCH % 2) == 0 && (SCRATCH % 2) == 0 && (ZBUF % 2) == 0, "float2 alignment");
static_assert((OFF_BAR % 2) == 0, "mbarrier alignment");
static_assert(SMEM_BYTES <= 113 * 1024, "two CTAs per SM");

struct Job {
  const float* wave;     // segment s starts at wave + s*src_pitch
  float* out;            // segment s, frame f -> out + (s*out_pitch + f)*128
  long long src_pitch, out_pitch;
  int seg_frames;        // frames computed per segment
  int tiles_per_seg;
  long long num_tiles;
  int normalize;
  float mean, std2;
};

template <bool BULK>
__global__ void __launch_bounds__(THREADS, 2) fbank_kernel(const zk_fbank_plan plan, const Job job) {
  extern __shared__ __align__(16) float sm[];
  float* samples = sm + OFF_SAMPLES;
  float* win = sm + OFF_WIN;
  float* w512 = sm + OFF_W512;
  float* mel_w = sm + OFF_MELW;
  int* mel_start = reinterpret_cast<int*>(sm + OFF_MELSTART);
  int* group_len = reinterpret_cast<int*>(sm + OFF_GROUP);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + OFF_BAR);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = lane & 15, half = lane >> 4;
  float* tbuf = sm + OFF_SCRATCH + (warp * 2 + half) * SCRATCH;
  float* pbuf = tbuf + ZBUF;

  for (int i = tid; i < FRAME; i += THREADS) win[i] = plan.d_win[i];
  for (int i = tid; i < 256; i += THREADS) w512[i] = plan.d_w512[i];
  for (int i = tid; i < MELW * NMEL; i += THREADS) mel_w[i] = plan.d_mel_w[i];
  for (int i = tid; i < NMEL; i += THREADS) mel_start[i] = plan.d_mel_start[i];
  if (tid < 8) group_len[tid] = plan.d_group_len[tid];
  for (int i = L; i < PBUF; i += 16) pbuf[i] = 0.f;
  cpx tw[16];
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) tw[k1] = {plan.d_tw[(L * 16 + k1) * 2], plan.d_tw[(L * 16 + k1) * 2 + 1]};
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

  int it = 0;
  for (long long tile = blockIdx.x; tile < job.num_tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
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

#pragma unroll 1
    for (int round = 0; round < TILE_FRAMES / (2 * WARPS); ++round) {
      const int fi = round * 2 * WARPS + warp * 2 + half;
      const bool live = fi < nf;
      const float* xs = xs_tile + (live ? fi : nf - 1) * SHIFT;
      float x[13][2];
      float s = lane_load(xs, L, x);
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      const float mean = __fdiv_rn(s, (float)FRAME);
      lane_stage1(xs, win, L, x, mean, plan.preemph, tw, tbuf);
      __syncwarp();
      cpx z[16];
      lane_stage2(tbuf, L, z);
      __syncwarp();
      lane_store_z(z, L, tbuf);
      __syncwarp();
      lane_power(z, tbuf, w512, L, pbuf);
      __syncwarp();
      float o[8];
      lane_mel(pbuf, mel_start, mel_w, group_len, L, plan.log_floor, o);
      if (live) {
        float* dst = job.out + (out_row + fi) * NMEL + L;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float v = o[i];
          if (job.normalize) v = __fdiv_rn(__fsub_rn(v, job.mean), job.std2);
          dst[16 * i] = v;
        }
      }
      __syncwarp();
    }
    __syncthreads();  // every warp is done with xs_tile before it is refilled
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
  static bool attr_done = false;
  if (!attr_done) {
    ZK_CUDA(cudaFuncSetAttribute(fbank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    ZK_CUDA(cudaFuncSetAttribute(fbank_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_done = true;
  }
  if (job.num_tiles <= 0) return 0;
  const bool bulk = (reinterpret_cast<uintptr_t>(job.wave) % 16 == 0) && (job.src_pitch % 4 == 0);
  long long grid = job.num_tiles < 2LL * num_sms() ? job.num_tiles : 2LL * num_sms();
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
    static bool attr_done = false;                                                                              \
    if (!attr_done) {                                                                                           \
      ZK_CUDA(cudaFuncSetAttribute(decimate_kernel<T, O, W, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); \
      attr_done = true;                                                                                         \
    }                                                                                                           \
    long long blocks = (n_out + OB - 1) / OB;                                                                   \
    if (blocks > 2LL * sms) blocks = 2LL * sms;                                                                 \
    const int bulk_ok = std::is_same<T, float>::value && channels == 1 && (reinterpret_cast<uintptr_t>(in) & 15) == 0; \
    ProfScope prof(ZK_K_RESAMPLE, stream);                                                                      \
    decimate_kernel<T, O, W, R><<<(int)blocks, DEC_THREADS, SMEM, stream>>>(in, n_in, channels, ch_pitch, taps, out, \
                                                                            n_out, bulk_ok);                   \
    ZK_LAUNCH_CHECK("decimate_kernel");                                                                        \
    return 0;                                                                                                  \
  }
  ZK_DECIMATE(3, 19, 7)   // 48 kHz -> 16 kHz
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
    zk::set_error("zk_fbank_plan_create: mel filter %d spans more than %d FFT bins", bad - 1, MELW);
    return ZK_ERR_SHAPE;
  }
  zk_fbank_plan* p = new zk_fbank_plan();
  memset(p, 0, sizeof(*p));
  p->preemph = preemph;
  p->log_floor = log_floor;
  cudaGetDevice(&p->device);
  cudaError_t e = cudaSuccess;
  auto up = [&](void** dst, const void* src, size_t bytes) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(dst, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  up((void**)&p->d_win, h_window, FRAME * 4);
  up((void**)&p->d_tw, t->tw, sizeof(t->tw));
  up((void**)&p->d_w512, t->w512, sizeof(t->w512));
  up((void**)&p->d_mel_start, t->start, sizeof(t->start));
  up((void**)&p->d_mel_w, t->melw, sizeof(t->melw));
  up((void**)&p->d_group_len, t->glen, sizeof(t->glen));
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
  cudaFree(p->d_win);
  cudaFree(p->d_tw);
  cudaFree(p->d_w512);
  cudaFree(p->d_mel_start);
  cudaFree(p->d_mel_w);
  cudaFree(p->d_group_len);
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
  return launch_fbank(plan, job, (cudaStream_t)stream);
}

}  // extern "C"
