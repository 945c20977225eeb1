// CPU emulation of the fbank CUDA kernel's per-lane arithmetic (TEST INFRASTRUCTURE).
// Includes the very header the kernel is built from (zk_fbank_math.cuh), instantiated for V = float (one frame per
// 16-lane group instead of the kernel's two packed in f32x2), and runs the 16 "lanes" in a loop with plain arrays
// standing in for shared memory, so the index math of the FFT decomposition, the real-FFT split and the segment form
// of the mel bank can be pinned against the oracle in the GPU-less container.
#include <stdlib.h>
#include <string.h>

#include "zk_fbank_math.cuh"

using namespace zk::fb;
typedef cpxv<float> cpx;
struct ScalarLoad {  // pair-indexed view of a contiguous array (the kernel keeps even / odd samples apart)
  const float* xs;
  float even(int i) const { return xs[2 * i]; }
  float odd(int i) const { return xs[2 * i + 1]; }
};

extern "C" int zk_emu_fbank(const float* wave, long n, const float* window, const float* mel_dense, float preemph,
                            float log_floor, float* out, long m) {
  HostTables* t = new HostTables();
  if (build_host_tables(mel_dense, *t)) {
    delete t;
    return -2;
  }
  if (m > (n < FRAME ? 0 : 1 + (n - FRAME) / SHIFT)) {
    delete t;
    return -1;
  }
  static cpx T[16][TPITCH];
  static cpx Z[16][16];
  float P[NZ + 1];
  for (long f = 0; f < m; ++f) {
    const float* xs = wave + f * SHIFT;
    static float x0[16][13], x1[16][13], xp[16][13];
    float s = 0.f;
    for (int L = 0; L < 16; ++L) s += lane_load<float>(ScalarLoad{xs}, L, x0[L], x1[L], xp[L]);
    const float mean = s / (float)FRAME;
    for (int L = 0; L < 16; ++L) {
      cpx tw[16], z[16];
      for (int k1 = 0; k1 < 16; ++k1) tw[k1] = {t->tw[(k1 * 16 + L) * 2], t->tw[(k1 * 16 + L) * 2 + 1]};
      lane_stage1<float>(x0[L], x1[L], xp[L], mean, preemph, ScalarLoad{window}, L, tw, 1, z);
      for (int k1 = 0; k1 < 16; ++k1) T[k1][L] = z[k1];
    }
    for (int k1 = 0; k1 < 16; ++k1) {
      cpx z[16];
      for (int n2 = 0; n2 < 16; ++n2) z[n2] = T[k1][n2];
      dft16(z);
      for (int k2 = 0; k2 < 16; ++k2) Z[k1][k2] = z[k2];  // Z[k1 + 16 k2]
    }
    for (int L = 0; L < 16; ++L) {
      for (int j = 0; j < 8; ++j) {
        const cpx a = Z[L][j];
        const cpx b = (L == 0) ? Z[0][(16 - j) & 15] : Z[16 - L][15 - j];
        float pk, pnk;
        split_power<float>(a, b, t->w512[(j * 16 + L) * 2], t->w512[(j * 16 + L) * 2 + 1], pk, pnk);
        P[L + 16 * j] = pk;
        if (L + 16 * j != 0) P[NZ - L - 16 * j] = pnk;
      }
    }
    P[128] = 4.0f * (Z[0][8].re * Z[0][8].re + Z[0][8].im * Z[0][8].im);
    float lo[NMEL], hi[NMEL];
    for (int i = 0; i < SEG_GROUPS; ++i)
      for (int L = 0; L < 16; ++L) {
        const int r = L + 16 * i;
        float al = 0.f, ah = 0.f;
        for (int tt = 0; tt < t->glen[i]; ++tt) {
          int k = t->seg_start[r] + tt;
          if (k > NZ - 1) k = NZ - 1;
          const float* w = &t->seg_w[((t->goff[i] + tt) * 16 + L) * 2];
          al = vfma(P[k], w[0], al);
          ah = vfma(P[k], w[1], ah);
        }
        lo[r] = al;
        hi[r] = ah;
      }
    for (int r = 0; r < NMEL; ++r) {
      const float e = lo[r] + (r > 0 ? hi[r - 1] : 0.f);
      out[f * NMEL + r] = logf(e > log_floor ? e : log_floor);
    }
  }
  delete t;
  return 0;
}
