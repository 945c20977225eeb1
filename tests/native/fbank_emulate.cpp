// CPU emulation of the fbank CUDA kernel's per-lane arithmetic (TEST INFRASTRUCTURE).
// Includes the very header the kernel is built from (zk_fbank_math.cuh) and runs its 16 "lanes" in a loop
// with plain arrays standing in for shared memory, so the index math of the FFT decomposition, the
// real-FFT split and the sparse mel bank can be pinned against the oracle in the GPU-less container.
#include <stdlib.h>
#include <string.h>

#include "zk_fbank_math.cuh"

using namespace zk::fb;

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
  float tbuf[ZBUF], pbuf[PBUF];
  for (int i = 0; i < PBUF; ++i) pbuf[i] = 0.f;
  for (long f = 0; f < m; ++f) {
    const float* xs = wave + f * SHIFT;
    float x[16][13][2];
    float s = 0.f;
    for (int L = 0; L < 16; ++L) s += lane_load(xs, L, x[L]);
    const float mean = s / (float)FRAME;
    for (int L = 0; L < 16; ++L) {
      cpx tw[16];
      for (int k1 = 0; k1 < 16; ++k1) tw[k1] = {t->tw[(L * 16 + k1) * 2], t->tw[(L * 16 + k1) * 2 + 1]};
      lane_stage1(xs, window, L, x[L], mean, preemph, tw, tbuf);
    }
    cpx z[16][16];
    for (int L = 0; L < 16; ++L) lane_stage2(tbuf, L, z[L]);
    for (int L = 0; L < 16; ++L) lane_store_z(z[L], L, tbuf);
    for (int L = 0; L < 16; ++L) lane_power(z[L], tbuf, t->w512, L, pbuf);
    for (int L = 0; L < 16; ++L) {
      float o[8];
      lane_mel(pbuf, t->start, t->melw, t->glen, L, log_floor, o);
      for (int i = 0; i < 8; ++i) out[f * NMEL + L + 16 * i] = o[i];
    }
  }
  delete t;
  return 0;
}
