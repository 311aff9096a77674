// Kaldi-compatible log-mel filterbank + global CMVN, one CTA per 25 ms frame.
// Arithmetic follows torchaudio.compliance.kaldi.fbank with the arguments the reference passes
// (fairseq/fairseq/data/audio/audio_utils.py:241-247): x*2^15, per-frame DC removal,
// pre-emphasis 0.97 (replicate pad), Povey window, zero-pad 400 -> 512, |rfft|^2, 80 mel bins, log(max(., eps)),
// then (x - mean) / std (agent/speech_to_speech.streamspeech.agent.py:89-98).
// The 400 samples of a frame are staged in shared memory once (1.6 KB in, 320 B out per frame): HBM-bound.
#include "common.cuh"
#include "kernels.h"

namespace ss {
namespace {

constexpr int FRAME = 400, SHIFT = 160, NFFT = 512, NBIN = 257, NMEL = 80;

__global__ void __launch_bounds__(256) fbank_kernel(const float* __restrict__ samples, int64_t n_samples, int f0,
                                                    const float* __restrict__ mel_bank, const float* __restrict__ window,
                                                    const float* __restrict__ cmvn_mean, const float* __restrict__ cmvn_std,
                                                    float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float re[NFFT], im[NFFT];
  __shared__ float twc[NFFT / 2], tws[NFFT / 2];
  __shared__ float frame[FRAME];
  __shared__ float red[32];
  const int tid = threadIdx.x;
  const int f = f0 + blockIdx.x;
  const int64_t base = (int64_t)f * SHIFT;
  float s = 0.f;
  for (int j = tid; j < FRAME; j += 256) {
    float v = (base + j < n_samples) ? samples[base + j] * 32768.0f : 0.f;
    frame[j] = v;
    s += v;
  }
  {
    float sn, cs;
    sincospif(-(float)tid / 256.0f, &sn, &cs);  // exp(-2*pi*i*tid/512)
    twc[tid] = cs;
    tws[tid] = sn;
  }
  float mean = block_sum(s, red) / (float)FRAME;
  // bit-reversed load of the windowed, pre-emphasised frame (zero padded to 512)
  for (int j = tid; j < NFFT; j += 256) {
    float v = 0.f;
    if (j < FRAME) {
      float x0 = frame[j] - mean;
      float xm = frame[j > 0 ? j - 1 : 0] - mean;
      v = (x0 - 0.97f * xm) * window[j];
    }
    int r = __brev((unsigned)j) >> (32 - 9);
    re[r] = v;
    im[r] = 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int stage = 0; stage < 9; ++stage) {
    int half = 1 << stage;
    int grp = tid >> stage, pos = tid & (half - 1);
    int i = (grp << (stage + 1)) + pos, j = i + half;
    int tw = pos << (8 - stage);
    float wr = twc[tw], wi = tws[tw];
    float xr = re[j], xi = im[j];
    float tr = wr * xr - wi * xi, ti = wr * xi + wi * xr;
    float ur = re[i], ui = im[i];
    re[i] = ur + tr;
    im[i] = ui + ti;
    re[j] = ur - tr;
    im[j] = ui - ti;
    __syncthreads();
  }
  // power spectrum: spectrum.abs().pow(2)
  for (int k = tid; k < NBIN; k += 256) {
    float a = sqrtf(re[k] * re[k] + im[k] * im[k]);
    frame[k] = a * a;  // reuse (257 <= 400)
  }
  __syncthreads();
  if (tid < NMEL) {
    const float* wrow = mel_bank + tid * NBIN;
    float acc = 0.f;
    for (int k = 0; k < NBIN; ++k) acc = fmaf(frame[k], wrow[k], acc);
    float v = logf(fmaxf(acc, 1.1920928955078125e-07f));
    if (cmvn_mean) v = (v - cmvn_mean[tid]) / cmvn_std[tid];
    out[(int64_t)blockIdx.x * NMEL + tid] = v;
  }
}

}  // namespace

void fbank_cmvn(const float* samples, int64_t n_samples, int f0, int nf, const float* mel_bank, const float* window,
                const float* cmvn_mean, const float* /*unused*/, const float* cmvn_std, float* out, cudaStream_t st) {
  ++g_launches;
  if (nf <= 0) return;
  launch_pdl(fbank_kernel, dim3(nf), dim3(256), 0, st, samples, n_samples, f0, mel_bank, window, cmvn_mean, cmvn_std, out);
}

}  // namespace ss
