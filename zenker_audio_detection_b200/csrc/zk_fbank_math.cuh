// Per-lane arithmetic of the Kaldi fbank kernel, templated on the value type so that the exact same index math runs
//   * inside the CUDA kernel with V = f32x2 (TWO frames per 16-lane group, every fp32 operation a packed
//     FADD2 / FMUL2 / FFMA2: sm_100 only reaches its fp32 peak through the packed forms), and
//   * inside the CPU emulation harness (tests/native/fbank_emulate.cpp) with V = float (one frame, 16 "lanes" in a
//     loop, plain arrays standing in for shared memory), which pins it against the oracle without a GPU.
//
// Frame pipeline (TA:compliance/kaldi.py:177-211, 616-633):
//   400 samples -> DC removal -> pre-emphasis (replicate pad) -> window -> zero-pad to 512
//   -> 512-point real FFT computed as a 256-point complex FFT of z[n] = y[2n] + i y[2n+1]
//      (256 = 16 x 16: in-lane DFT-16, twiddle, transpose through shared memory, in-lane DFT-16)
//   -> split into the real spectrum, power |X[k]|^2 for k < 256
//   -> mel bank in its sparse "segment" form: every FFT bin k feeds at most two ADJACENT filters (the falling slope of
//      filter r_k and the rising slope of filter r_k + 1), so with seg(r) = {k : r_k = r}
//          mel[r] = sum_{k in seg(r)} lo_k P_k + sum_{k in seg(r-1)} hi_k P_k
//      and each power value is read once (255 taps instead of the 896 of the padded per-filter gather, and instead of
//      the 32 768 of the reference's dense matmul); lo_k / hi_k are the dense bank's own entries, so the sum is exact
//   -> log.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define ZK_HD __host__ __device__ __forceinline__
#else
#define ZK_HD inline
#endif

namespace zk {
namespace fb {

constexpr int FRAME = 400, SHIFT = 160, NFFT = 512, NZ = 256, NMEL = 128;
constexpr int TPITCH = 17;          // row pitch (complex elements) of the 16 x 16 transpose buffer
constexpr int SEG_GROUPS = 8;       // lane L owns the mel segments r = L + 16 i, i < SEG_GROUPS
constexpr int SEG_TAPS_MAX = 64;    // sum over groups of the longest segment (19 for the 128-bin 16 kHz bank)

// ---- value-type operations: float (host / emulation) ------------------------------------------------------------
ZK_HD float vadd(float a, float b) { return a + b; }
ZK_HD float vsub(float a, float b) { return a - b; }
ZK_HD float vmul(float a, float b) { return a * b; }
ZK_HD float vfma(float a, float b, float c) { return a * b + c; }
#ifdef __CUDACC__
// ---- value-type operations: f32x2 (device): two frames per register pair, one packed instruction for both ---------
__device__ __forceinline__ float2 vadd(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return reinterpret_cast<float2&>(d);
}
__device__ __forceinline__ float2 vsub(float2 a, float2 b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return reinterpret_cast<float2&>(d);
}
__device__ __forceinline__ float2 vmul(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return reinterpret_cast<float2&>(d);
}
__device__ __forceinline__ float2 vfma(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return reinterpret_cast<float2&>(d);
}
#endif
template <typename V>
ZK_HD V vbc(float a);  // broadcast a scalar into the value type
#ifdef __CUDACC__
template <>
__device__ __forceinline__ float2 vbc<float2>(float a) {
  return make_float2(a, a);
}
#endif
template <>
ZK_HD float vbc<float>(float a) {
  return a;
}
template <typename V>
struct cpxv {
  V re, im;
};

// forward 4-point DFT (W4 = -i), in place on (a,b,c,d) -> (X0,X1,X2,X3)
template <typename V>
ZK_HD void dft4(cpxv<V>& a, cpxv<V>& b, cpxv<V>& c, cpxv<V>& d) {
  const cpxv<V> s0 = {vadd(a.re, c.re), vadd(a.im, c.im)}, d0 = {vsub(a.re, c.re), vsub(a.im, c.im)};
  const cpxv<V> s1 = {vadd(b.re, d.re), vadd(b.im, d.im)}, d1 = {vsub(b.re, d.re), vsub(b.im, d.im)};
  a = {vadd(s0.re, s1.re), vadd(s0.im, s1.im)};
  c = {vsub(s0.re, s1.re), vsub(s0.im, s1.im)};
  b = {vadd(d0.re, d1.im), vsub(d0.im, d1.re)};  // d0 - i*d1
  d = {vsub(d0.re, d1.im), vadd(d0.im, d1.re)};  // d0 + i*d1
}
// a * (wr + i wi) with scalar constants
template <typename V>
ZK_HD cpxv<V> cmulc(cpxv<V> a, float wr, float wi) {
  const V r = vbc<V>(wr), i = vbc<V>(wi);
  return {vsub(vmul(a.re, r), vmul(a.im, i)), vfma(a.re, i, vmul(a.im, r))};
}
template <typename V>
ZK_HD cpxv<V> cmulv(cpxv<V> a, V wr, V wi) {
  return {vsub(vmul(a.re, wr), vmul(a.im, wi)), vfma(a.re, wi, vmul(a.im, wr))};
}

// forward 16-point DFT, natural order in and out: X[k] = sum_n x[n] exp(-2 pi i n k / 16).
// n = 4 n1 + n2, k = k1 + 4 k2:  X[k1 + 4 k2] = sum_n2 W16^(n2 k1) (sum_n1 x[4 n1 + n2] W4^(n1 k1)) W4^(n2 k2)
template <typename V>
ZK_HD void dft16(cpxv<V> (&x)[16]) {
  constexpr float C1 = 0.92387953251128673848f, S1 = 0.38268343236508978178f, R2 = 0.70710678118654752440f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4(x[n2], x[4 + n2], x[8 + n2], x[12 + n2]);  // x[4 k1 + n2] = T[k1][n2]
  // twiddle T[k1][n2] *= W16^(n2 k1)
  x[5] = cmulc(x[5], C1, -S1);     // W^1
  x[6] = cmulc(x[6], R2, -R2);     // W^2
  x[7] = cmulc(x[7], S1, -C1);     // W^3
  x[9] = cmulc(x[9], R2, -R2);     // W^2
  x[10] = {x[10].im, vsub(vbc<V>(0.f), x[10].re)};  // W^4 = -i
  x[11] = cmulc(x[11], -R2, -R2);  // W^6
  x[13] = cmulc(x[13], S1, -C1);   // W^3
  x[14] = cmulc(x[14], -R2, -R2);  // W^6
  x[15] = cmulc(x[15], -C1, S1);   // W^9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4(x[4 * k1 + 0], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);  // -> X[k1 + 4 k2] at x[4 k1 + k2]
  // reorder to natural order: X[k1 + 4 k2] currently at x[4 k1 + k2]  (a 4x4 transpose)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i + 1; j < 4; ++j) {
      const cpxv<V> t = x[4 * i + j];
      x[4 * i + j] = x[4 * j + i];
      x[4 * j + i] = t;
    }
}

// ---- stage 0/1: lane n2 (0..15) owns z[n1] = (y[32 n1 + 2 n2], y[32 n1 + 2 n2 + 1]), n1 < 13 (zero beyond 400).
// Samples and window are addressed by PAIR index i = 16 n1 + n2 (even(i) = x[2 i], odd(i) = x[2 i + 1]): the kernel
// keeps them de-interleaved in shared memory so that the 16 lanes of a group read 16 consecutive words.
// Returns the lane's partial sample sum (for the DC mean) and keeps the raw samples in x[][].
template <typename V, typename Loader>
ZK_HD V lane_load(const Loader& ld, int n2, V (&x0)[13], V (&x1)[13], V (&xp)[13]) {
  V s = vbc<V>(0.f);
#pragma unroll
  for (int n1 = 0; n1 < 13; ++n1) {
    const int i = 16 * n1 + n2;
    if (2 * i < FRAME) {
      x0[n1] = ld.even(i);
      x1[n1] = ld.odd(i);
      xp[n1] = i > 0 ? ld.odd(i - 1) : ld.even(0);  // replicate padding of the pre-emphasis
    } else {
      x0[n1] = vbc<V>(0.f);
      x1[n1] = vbc<V>(0.f);
      xp[n1] = vbc<V>(0.f);
    }
    s = vadd(s, vadd(x0[n1], x1[n1]));
  }
  return s;
}

// DC removal, pre-emphasis, window, first DFT-16 over n1, twiddle by W256^(n2 k1); result a[k1] = A[k1][n2].
// win.even(i) / win.odd(i) and tw[k1] (= W256^(n2 k1) of this lane) are given in the value type.
template <typename V, typename Window>
ZK_HD void lane_stage1(const V (&x0)[13], const V (&x1)[13], const V (&xp)[13], V mean, float preemph, const Window& win,
                       int n2, const cpxv<V>* tw, int tw_stride, cpxv<V> (&z)[16]) {
  const V npre = vbc<V>(-preemph);
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    z[n1] = {vbc<V>(0.f), vbc<V>(0.f)};
    if (n1 < 13) {
      const int i = 16 * n1 + n2;
      if (2 * i < FRAME) {
        const V cp = vsub(xp[n1], mean), c0 = vsub(x0[n1], mean), c1 = vsub(x1[n1], mean);
        const V y0 = vfma(npre, cp, c0), y1 = vfma(npre, c0, c1);
        z[n1] = {vmul(y0, win.even(i)), vmul(y1, win.odd(i))};
      }
    }
  }
  dft16(z);
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) z[k1] = cmulv(z[k1], tw[k1 * tw_stride].re, tw[k1 * tw_stride].im);
}

// ---- real-FFT split + power for one pair (k, NZ - k): a = Z[k], b = Z[NZ - k], w = W512^k.
//   E = a + conj b, O = a - conj b, T = -i w O;  X[k] = (E + T) / 2,  X[NZ-k] = conj(E - T) / 2.
// Returns 4 |X|^2 (the factor 1/4 lives in the mel weights).
template <typename V>
ZK_HD void split_power(cpxv<V> a, cpxv<V> b, V wr, V wi, V& p_k, V& p_nk) {
  const cpxv<V> e = {vadd(a.re, b.re), vsub(a.im, b.im)};
  const cpxv<V> o = {vsub(a.re, b.re), vadd(a.im, b.im)};
  // -i w = (wi, -wr):  T = (wi o.re + wr o.im, wi o.im - wr o.re)
  const cpxv<V> t = {vfma(wi, o.re, vmul(wr, o.im)), vsub(vmul(wi, o.im), vmul(wr, o.re))};
  const V xr = vadd(e.re, t.re), xi = vadd(e.im, t.im);
  const V yr = vsub(e.re, t.re), yi = vsub(e.im, t.im);
  p_k = vfma(xr, xr, vmul(xi, xi));
  p_nk = vfma(yr, yr, vmul(yi, yi));
}

// ---- host: constant tables shared by the plan and the CPU emulation harness --------------------------------
struct HostTables {
  float tw[16 * 16 * 2];       // [k1][lane n2] = W256^(n2 k1) as (re, im)
  float w512[8 * 16 * 2];      // [j][lane L]  = W512^(L + 16 j) as (re, im)
  int seg_start[NMEL];         // first FFT bin of segment r (bins whose lower filter is r)
  int seg_len[NMEL];
  int glen[SEG_GROUPS];        // longest segment of group i (segments 16 i .. 16 i + 15)
  int goff[SEG_GROUPS + 1];    // prefix sum of glen
  float seg_w[SEG_TAPS_MAX * 16 * 2];  // [goff[i] + t][lane L] = (lo, hi) weight of bin seg_start[L + 16 i] + t, x 1/4
};
// h_mel: dense [NMEL][NZ] bank.  Returns 0, or a (1-based) FFT bin whose non-zeros are not two adjacent filters.
inline int build_host_tables(const float* h_mel, HostTables& t) {
  const double PI = 3.14159265358979323846;
  for (int k1 = 0; k1 < 16; ++k1)
    for (int n2 = 0; n2 < 16; ++n2) {
      const double a = -2.0 * PI * (double)(n2 * k1) / 256.0;
      t.tw[(k1 * 16 + n2) * 2] = (float)cos(a);
      t.tw[(k1 * 16 + n2) * 2 + 1] = (float)sin(a);
    }
  for (int j = 0; j < 8; ++j)
    for (int L = 0; L < 16; ++L) {
      const double a = -2.0 * PI * (double)(L + 16 * j) / 512.0;
      t.w512[(j * 16 + L) * 2] = (float)cos(a);
      t.w512[(j * 16 + L) * 2 + 1] = (float)sin(a);
    }
  int row_of[NZ];
  for (int r = 0; r < NMEL; ++r) {
    t.seg_start[r] = 0;
    t.seg_len[r] = 0;
  }
  int prev = -1;
  for (int k = 0; k < NZ; ++k) {
    int lo = -1, nnz = 0;
    for (int r = 0; r < NMEL; ++r)
      if (h_mel[r * NZ + k] != 0.f) {
        if (lo < 0) lo = r;
        ++nnz;
      }
    row_of[k] = lo;
    if (nnz == 0) continue;
    if (nnz > 2 || (nnz == 2 && h_mel[(lo + 1) * NZ + k] == 0.f) || lo < prev) return k + 1;
    if (t.seg_len[lo] == 0) t.seg_start[lo] = k;
    if (t.seg_start[lo] + t.seg_len[lo] != k) return k + 1;  // a segment is a run of consecutive bins
    ++t.seg_len[lo];
    prev = lo;
  }
  t.goff[0] = 0;
  for (int i = 0; i < SEG_GROUPS; ++i) {
    t.glen[i] = 0;
    for (int L = 0; L < 16; ++L)
      if (t.seg_len[L + 16 * i] > t.glen[i]) t.glen[i] = t.seg_len[L + 16 * i];
    t.goff[i + 1] = t.goff[i] + t.glen[i];
  }
  if (t.goff[SEG_GROUPS] > SEG_TAPS_MAX) return NZ + 1;
  for (int i = 0; i < SEG_TAPS_MAX * 16 * 2; ++i) t.seg_w[i] = 0.f;
  for (int i = 0; i < SEG_GROUPS; ++i)
    for (int L = 0; L < 16; ++L) {
      const int r = L + 16 * i;
      for (int tt = 0; tt < t.seg_len[r]; ++tt) {
        const int k = t.seg_start[r] + tt;
        float* w = &t.seg_w[((t.goff[i] + tt) * 16 + L) * 2];
        w[0] = 0.25f * h_mel[r * NZ + k];
        w[1] = (r + 1 < NMEL) ? 0.25f * h_mel[(r + 1) * NZ + k] : 0.f;
      }
    }
  (void)row_of;
  return 0;
}

}  // namespace fb
}  // namespace zk
