// Per-lane arithmetic of one Kaldi fbank frame, written as __host__ __device__ code so the exact same
// index math runs inside the CUDA kernel (16 lanes of a warp per frame) and inside the CPU emulation harness
// (tests/native/fbank_emulate.cu, 16 "lanes" in a loop) that pins it against the oracle without a GPU.
//
// Frame pipeline (TA:compliance/kaldi.py:177-211, 616-633):
//   400 samples -> DC removal -> pre-emphasis (replicate pad) -> window -> zero-pad to 512
//   -> 512-point real FFT computed as a 256-point complex FFT of z[n] = y[2n] + i y[2n+1]
//      (256 = 16 x 16: in-lane DFT-16, twiddle, transpose through a 16x17 buffer, in-lane DFT-16)
//   -> split into the real spectrum, power |X[k]|^2 for k < 256 -> sparse mel (<= MELW taps per bin) -> log.
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

constexpr int FRAME = 400, SHIFT = 160, NFFT = 512, NZ = 256, NMEL = 128, MELW = 16;
constexpr int TPAD = 17;                 // row pitch (complex) of the 16x16 transpose buffer
constexpr int ZBUF = 16 * TPAD * 2;      // floats per frame scratch (>= 2*256 for the Z exchange)
constexpr int PBUF = NZ + MELW;          // power spectrum + zero tail so padded mel taps read zeros

struct cpx {
  float re, im;
};
ZK_HD cpx cmul(cpx a, cpx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }

// forward 4-point DFT (W4 = -i), in place on (a,b,c,d) -> (X0,X1,X2,X3)
ZK_HD void dft4(cpx& a, cpx& b, cpx& c, cpx& d) {
  const cpx s0 = {a.re + c.re, a.im + c.im}, d0 = {a.re - c.re, a.im - c.im};
  const cpx s1 = {b.re + d.re, b.im + d.im}, d1 = {b.re - d.re, b.im - d.im};
  a = {s0.re + s1.re, s0.im + s1.im};
  c = {s0.re - s1.re, s0.im - s1.im};
  b = {d0.re + d1.im, d0.im - d1.re};  // d0 - i*d1
  d = {d0.re - d1.im, d0.im + d1.re};  // d0 + i*d1
}

// forward 16-point DFT, natural order in and out: X[k] = sum_n x[n] exp(-2 pi i n k / 16).
// n = 4 n1 + n2, k = k1 + 4 k2:  X[k1 + 4 k2] = sum_n2 W16^(n2 k1) (sum_n1 x[4 n1 + n2] W4^(n1 k1)) W4^(n2 k2)
ZK_HD void dft16(cpx (&x)[16]) {
  constexpr float C1 = 0.92387953251128673848f, S1 = 0.38268343236508978178f, R2 = 0.70710678118654752440f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4(x[n2], x[4 + n2], x[8 + n2], x[12 + n2]);  // x[4 k1 + n2] = T[k1][n2]
  // twiddle T[k1][n2] *= W16^(n2 k1)
  x[5] = cmul(x[5], cpx{C1, -S1});    // W^1
  x[6] = cmul(x[6], cpx{R2, -R2});    // W^2
  x[7] = cmul(x[7], cpx{S1, -C1});    // W^3
  x[9] = cmul(x[9], cpx{R2, -R2});    // W^2
  x[10] = cpx{x[10].im, -x[10].re};   // W^4 = -i
  x[11] = cmul(x[11], cpx{-R2, -R2}); // W^6
  x[13] = cmul(x[13], cpx{S1, -C1});  // W^3
  x[14] = cmul(x[14], cpx{-R2, -R2}); // W^6
  x[15] = cmul(x[15], cpx{-C1, S1});  // W^9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4(x[4 * k1 + 0], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);  // -> X[k1 + 4 k2] at x[4 k1 + k2]
  // reorder to natural order: X[k1 + 4 k2] currently at x[4 k1 + k2]  (a 4x4 transpose)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i + 1; j < 4; ++j) {
      const cpx t = x[4 * i + j];
      x[4 * i + j] = x[4 * j + i];
      x[4 * j + i] = t;
    }
}

// ---- phase 1: lane n2 (0..15) loads its samples, returns the partial sum for the DC mean.
// xs = the frame's 400 samples (shared memory on the device). x[n1][0/1] = xs[32 n1 + 2 n2 + 0/1], n1 < 13.
ZK_HD float lane_load(const float* xs, int n2, float (&x)[13][2]) {
  float s = 0.f;
#pragma unroll
  for (int n1 = 0; n1 < 13; ++n1) {
    const int m = 32 * n1 + 2 * n2;
    if (m < FRAME) {
      x[n1][0] = xs[m];
      x[n1][1] = xs[m + 1];
    } else {
      x[n1][0] = 0.f;
      x[n1][1] = 0.f;
    }
    s += x[n1][0] + x[n1][1];
  }
  return s;
}

#ifdef __CUDA_ARCH__
#define ZK_MUL(a, b) __fmul_rn(a, b)
#define ZK_SUB(a, b) __fsub_rn(a, b)
#else
#define ZK_MUL(a, b) ((a) * (b))
#define ZK_SUB(a, b) ((a) - (b))
#endif

// ---- phase 2: DC removal, pre-emphasis, window, first DFT-16 over n1, twiddle by W256^(n2 k1), and the
// transposing store A[k1][n2] -> tbuf[(k1*TPAD + n2)*2].  tw[k1] = W256^(n2 k1) for this lane.
ZK_HD void lane_stage1(const float* xs, const float* win, int n2, const float (&x)[13][2], float mean, float preemph,
                       const cpx (&tw)[16], float* tbuf) {
  cpx z[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    z[n1] = {0.f, 0.f};
    if (n1 < 13) {
      const int m = 32 * n1 + 2 * n2;
      if (m < FRAME) {
        const float prev = ZK_SUB(xs[m > 0 ? m - 1 : 0], mean);
        const float c0 = ZK_SUB(x[n1][0], mean), c1 = ZK_SUB(x[n1][1], mean);
        const float y0 = ZK_SUB(c0, ZK_MUL(preemph, prev));
        const float y1 = ZK_SUB(c1, ZK_MUL(preemph, c0));
        z[n1] = {ZK_MUL(y0, win[m]), ZK_MUL(y1, win[m + 1])};
      }
    }
  }
  dft16(z);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    const cpx a = (k1 == 0) ? z[0] : cmul(z[k1], tw[k1]);
    tbuf[(k1 * TPAD + n2) * 2 + 0] = a.re;
    tbuf[(k1 * TPAD + n2) * 2 + 1] = a.im;
  }
}

// ---- phase 3: lane k1 reads A[k1][n2] for all n2, second DFT-16 -> Z[k1 + 16 k2] in z[k2].
ZK_HD void lane_stage2(const float* tbuf, int k1, cpx (&z)[16]) {
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) z[n2] = {tbuf[(k1 * TPAD + n2) * 2 + 0], tbuf[(k1 * TPAD + n2) * 2 + 1]};
  dft16(z);
}

// ---- phase 4: publish Z in natural order (index k = L + 16 k2) for the partner exchange.
ZK_HD void lane_store_z(const cpx (&z)[16], int L, float* zbuf) {
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    zbuf[(L + 16 * k2) * 2 + 0] = z[k2].re;
    zbuf[(L + 16 * k2) * 2 + 1] = z[k2].im;
  }
}

// ---- phase 5: real-FFT split + power.  Lane L handles k = L + 16 j (j < 8) together with 256 - k;
// lane 0 additionally handles k = 128.  w512[k] = exp(-2 pi i k / 512) for k < 128.
//   E = (Z[k] + conj Z[N-k]) / 2,  O = (Z[k] - conj Z[N-k]) / (2i),  X[k] = E + W512^k O,  X[N-k] = conj(E - W512^k O)
ZK_HD void lane_power(const cpx (&z)[16], const float* zbuf, const float* w512, int L, float* pbuf) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = L + 16 * j;
    const int kp = (NZ - k) & (NZ - 1);
    const cpx a = z[j];
    const cpx b = {zbuf[kp * 2], zbuf[kp * 2 + 1]};
    const cpx e = {0.5f * (a.re + b.re), 0.5f * (a.im - b.im)};
    const cpx o = {0.5f * (a.im + b.im), -0.5f * (a.re - b.re)};
    const cpx w = {w512[2 * k], w512[2 * k + 1]};
    const cpx wo = cmul(w, o);
    const float xr = e.re + wo.re, xi = e.im + wo.im;
    const float yr = e.re - wo.re, yi = e.im - wo.im;
    pbuf[k] = xr * xr + xi * xi;
    if (k != 0) pbuf[NZ - k] = yr * yr + yi * yi;
  }
  if (L == 0) pbuf[128] = z[8].re * z[8].re + z[8].im * z[8].im;
}

// ---- phase 6: sparse mel + log for mel bins r = L + 16 i.  mel_start[r], mel_w[t*NMEL + r], group_len[i].
ZK_HD void lane_mel(const float* pbuf, const int* mel_start, const float* mel_w, const int* group_len, int L,
                    float log_floor, float (&out)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = L + 16 * i;
    const int s = mel_start[r];
    const int len = group_len[i];
    float e = 0.f;
    for (int t = 0; t < len; ++t) e = fmaf(pbuf[s + t], mel_w[t * NMEL + r], e);
    out[i] = logf(fmaxf(e, log_floor));
  }
}

// ---- host: constant tables shared by the plan and the CPU emulation harness --------------------------------
struct HostTables {
  float tw[16 * 16 * 2];     // [lane n2][k1] = W256^(n2 k1)
  float w512[256];           // [k < 128] = W512^k
  float melw[MELW * NMEL];   // [tap][mel bin]
  int start[NMEL];
  int glen[8];
};
// h_mel: dense [NMEL][NZ] bank.  Returns 0, or the (1-based) index of a filter wider than MELW.
inline int build_host_tables(const float* h_mel, HostTables& t) {
  const double PI = 3.14159265358979323846;
  for (int n2 = 0; n2 < 16; ++n2)
    for (int k1 = 0; k1 < 16; ++k1) {
      const double a = -2.0 * PI * (double)(n2 * k1) / 256.0;
      t.tw[(n2 * 16 + k1) * 2] = (float)cos(a);
      t.tw[(n2 * 16 + k1) * 2 + 1] = (float)sin(a);
    }
  for (int k = 0; k < 128; ++k) {
    const double a = -2.0 * PI * (double)k / 512.0;
    t.w512[2 * k] = (float)cos(a);
    t.w512[2 * k + 1] = (float)sin(a);
  }
  for (int i = 0; i < MELW * NMEL; ++i) t.melw[i] = 0.f;
  for (int i = 0; i < 8; ++i) t.glen[i] = 0;
  for (int r = 0; r < NMEL; ++r) {
    t.start[r] = 0;
    int lo = -1, hi = -1;
    for (int k = 0; k < NZ; ++k)
      if (h_mel[r * NZ + k] != 0.f) {
        if (lo < 0) lo = k;
        hi = k;
      }
    if (lo < 0) continue;  // empty filter: always log(floor)
    const int len = hi - lo + 1;
    if (len > MELW) return r + 1;
    t.start[r] = lo;
    for (int k = 0; k < len; ++k) t.melw[k * NMEL + r] = h_mel[r * NZ + lo + k];
    if (len > t.glen[r / 16]) t.glen[r / 16] = len;
  }
  return 0;
}

}  // namespace fb
}  // namespace zk
