// Multi-stream (batched streaming) kernels, see kernels_multistream.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ss {

// One stream of a batched step (device array, one entry per batch element).
struct MsStream {
  int slot;          // pool slot: every per-stream buffer is base + slot * stride
  int a0;            // first active (not yet final) encoder row = rows final before this step
  int T;             // encoder rows after this step (keys 0 .. T-1)
  int F;             // fbank frames after this step
  int f_lo;          // first fbank frame the subsampler window of this step reads
  int frame0;        // fbank: first new frame
  int n_new_frames;  // fbank: number of new frames
  int pad_;
  int64_t out_off;   // CTC collapse: int64 offset of this stream's packed output
};

void ms_gather_rows(const float* src_base, int64_t slot_stride, const MsStream* S, int n, int which /*0: rows from f_lo (limit F), 1: rows from a0 (limit T)*/,
                    int rows, int C, float* dst, cudaStream_t st);
void ms_scatter_rows(const float* s0, const float* s1, const float* s2, int lds, float* d0, float* d1, float* d2, int64_t slot_stride, const MsStream* S,
                     int n, int nA, int C, cudaStream_t st);
void ms_relpos_attention(const float* q, int ldq, const float* kc, const float* vc, int64_t slot_stride, int D, const float* pos, int Tpos,
                         const float* bias_u, const float* bias_v, float* out, int ldo, const MsStream* S, int n, int nA, int H, int chunk,
                         int Tmax, cudaStream_t st);
void ms_depthwise(const float* gc, int64_t slot_stride, const float* w, const float* scale, const float* shift, float* y, int ldy, const MsStream* S, int n,
                  int nA, int C, int k, int chunk, cudaStream_t st);
void ms_fbank(const float* audio_base, int64_t audio_stride, float* feat_base, int64_t feat_stride, const MsStream* S, int n, int max_new_frames,
              const float* melT /*[257][80]*/, const float* window, const float* cmvn_mean, const float* cmvn_std, cudaStream_t st);
void ms_ctc_argmax(const float* logits, int ld, int V, const int* masked, int n_masked, int64_t* am_base, int64_t am_stride, const MsStream* S, int n,
                   int nA, int heads, cudaStream_t st);
void ms_ctc_collapse(const int64_t* am_base, int64_t am_stride, const MsStream* S, int n, int heads, int blank, int pad, int64_t* out, cudaStream_t st);

}  // namespace ss
