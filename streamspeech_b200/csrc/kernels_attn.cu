// Attention kernels (fp32, head_dim 64).  One CTA per (query tile, head, batch element); the score
// tile lives in shared memory, softmax uses warp-shuffle reductions.
//   relpos_attention: RelPositionMultiHeadedAttention.forward with the rel_shift folded into the index
//                     bd[i][j] = (q_i + v) . P[i - j]   (uni_unity/modules/espnet_multihead_attention.py:133-209)
//   mha_attention   : fairseq MultiheadAttention slow path (ctc_unity/modules/multihead_attention.py:555-784)
#include "common.cuh"
#include "kernels.h"

namespace ss {
namespace {

constexpr int QT = 16;    // queries per CTA
constexpr int ATT_NT = 128;
constexpr int HD = 64;    // head dim

__device__ __forceinline__ float dot64(const float* __restrict__ a_smem, const float* __restrict__ b_gmem) {
  float acc = 0.f;
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    float4 b = *reinterpret_cast<const float4*>(b_gmem + d);
    acc = fmaf(a_smem[d + 0], b.x, acc);
    acc = fmaf(a_smem[d + 1], b.y, acc);
    acc = fmaf(a_smem[d + 2], b.z, acc);
    acc = fmaf(a_smem[d + 3], b.w, acc);
  }
  return acc;
}

// softmax over S[q][0..n) for every query row of the tile; rows are handled one warp at a time
__device__ __forceinline__ void softmax_rows(float* S, int ld, int nq, const int* nvis) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = warp; q < nq; q += ATT_NT / 32) {
    float* row = S + q * ld;
    int n = nvis[q];
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) {
      float e = expf(row[j] - mx);
      row[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    for (int j = lane; j < n; j += 32) row[j] = row[j] / sum;
  }
}

// out[q][d] = sum_j S[q][j] * V[j][d]; thread -> (q = tid/8, 8 dims starting at (tid%8)*8)
__device__ __forceinline__ void pv_store(const float* S, int ld, int nq, const int* nvis, const float* __restrict__ vbase,
                                         int64_t ldv, float* __restrict__ obase, int64_t ldo, int i0) {
  int q = threadIdx.x >> 3, d0 = (threadIdx.x & 7) * 8;
  if (q >= nq) return;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  int n = nvis[q];
  const float* row = S + q * ld;
  for (int j = 0; j < n; ++j) {
    float p = row[j];
    const float* vp = vbase + (int64_t)j * ldv + d0;
    float4 a = *reinterpret_cast<const float4*>(vp);
    float4 b = *reinterpret_cast<const float4*>(vp + 4);
    acc[0] = fmaf(p, a.x, acc[0]);
    acc[1] = fmaf(p, a.y, acc[1]);
    acc[2] = fmaf(p, a.z, acc[2]);
    acc[3] = fmaf(p, a.w, acc[3]);
    acc[4] = fmaf(p, b.x, acc[4]);
    acc[5] = fmaf(p, b.y, acc[5]);
    acc[6] = fmaf(p, b.z, acc[6]);
    acc[7] = fmaf(p, b.w, acc[7]);
  }
  float* op = obase + (int64_t)(i0 + q) * ldo + d0;
  *reinterpret_cast<float4*>(op) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  *reinterpret_cast<float4*>(op + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

__global__ void __launch_bounds__(ATT_NT) relpos_attention_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k,
                                                                  int ldk, const float* __restrict__ v, int ldv,
                                                                  const float* __restrict__ pos, int Tpos,
                                                                  const float* __restrict__ bias_u, const float* __restrict__ bias_v,
                                                                  float* __restrict__ out, int nQ, int q_offset, int T, int H, int D,
                                                                  int chunk, const int* __restrict__ lengths, int ldS) {
  extern __shared__ __align__(16) float smem[];
  float* Qu = smem;                 // [QT][64]
  float* Qv = Qu + QT * HD;         // [QT][64]
  float* S = Qv + QT * HD;          // [QT][ldS]
  __shared__ int nvis[QT];
  const int b = blockIdx.z, h = blockIdx.y, r0 = blockIdx.x * QT;  // r0: first query row of the tile (buffer-relative)
  const int nq = min(QT, nQ - r0);
  const int len = lengths ? min(lengths[b], T) : T;
  const float* qb = q + ((int64_t)b * nQ) * ldq + h * HD;
  const float* kb = k + ((int64_t)b * T) * ldk + h * HD;
  const float* vb = v + ((int64_t)b * T) * ldv + h * HD;
  for (int e = threadIdx.x; e < QT * HD; e += ATT_NT) {
    int qq = e / HD, d = e % HD;
    float val = (qq < nq) ? qb[(int64_t)(r0 + qq) * ldq + d] : 0.f;
    Qu[e] = val + bias_u[h * HD + d];
    Qv[e] = val + bias_v[h * HD + d];
  }
  if (threadIdx.x < QT) {
    int i = q_offset + r0 + threadIdx.x;                           // absolute query position
    int lim = chunk > 0 ? min((i / chunk + 1) * chunk, T) : T;   // chunk mask (s2t_conformer.py:195-213)
    nvis[threadIdx.x] = max(1, min(lim, len));                   // key padding mask (forward_attention)
  }
  __syncthreads();
  const int kmax = nvis[nq - 1];  // limits are non-decreasing in i
  const float* pb = pos + h * HD;
  for (int e = threadIdx.x; e < nq * kmax; e += ATT_NT) {
    int qq = e / kmax, j = e - qq * kmax;
    if (j >= nvis[qq]) continue;
    int i = q_offset + r0 + qq;
    float ac = dot64(Qu + qq * HD, kb + (int64_t)j * ldk);
    float bd = dot64(Qv + qq * HD, pb + (int64_t)(i - j + Tpos - 1) * D);
    S[qq * ldS + j] = (ac + bd) * 0.125f;  // / sqrt(d_k), d_k = 64
  }
  __syncthreads();
  softmax_rows(S, ldS, nq, nvis);
  __syncthreads();
  pv_store(S, ldS, nq, nvis, vb, ldv, out + ((int64_t)b * nQ) * D + h * HD, D, r0);
}

__global__ void __launch_bounds__(ATT_NT) mha_attention_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k,
                                                               int ldk, const float* __restrict__ v, int ldv,
                                                               float* __restrict__ out, int ldo, int Tq, int Tk, float scale,
                                                               int causal, int causal_offset,
                                                               const int* __restrict__ kv_len, int ldS) {
  extern __shared__ __align__(16) float smem[];
  float* Q = smem;            // [QT][64]
  float* S = Q + QT * HD;     // [QT][ldS]
  __shared__ int nvis[QT];
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * QT;
  const int nq = min(QT, Tq - i0);
  const int len = kv_len ? min(kv_len[b], Tk) : Tk;
  const float* qb = q + ((int64_t)b * Tq) * ldq + h * HD;
  const float* kb = k + ((int64_t)b * Tk) * ldk + h * HD;
  const float* vb = v + ((int64_t)b * Tk) * ldv + h * HD;
  for (int e = threadIdx.x; e < QT * HD; e += ATT_NT) {
    int qq = e / HD, d = e % HD;
    Q[e] = (qq < nq) ? qb[(int64_t)(i0 + qq) * ldq + d] * scale : 0.f;  // q *= scaling (multihead_attention.py:573)
  }
  if (threadIdx.x < QT) {
    int i = i0 + threadIdx.x;
    int lim = causal ? min(i + causal_offset + 1, Tk) : Tk;
    nvis[threadIdx.x] = max(1, min(lim, len));
  }
  __syncthreads();
  int kmax = 0;
  for (int qq = 0; qq < nq; ++qq) kmax = max(kmax, nvis[qq]);
  for (int e = threadIdx.x; e < nq * kmax; e += ATT_NT) {
    int qq = e / kmax, j = e - qq * kmax;
    if (j >= nvis[qq]) continue;
    S[qq * ldS + j] = dot64(Q + qq * HD, kb + (int64_t)j * ldk);
  }
  __syncthreads();
  softmax_rows(S, ldS, nq, nvis);
  __syncthreads();
  pv_store(S, ldS, nq, nvis, vb, ldv, out + ((int64_t)b * Tq) * ldo + h * HD, ldo, i0);
}

// ---- one CTA per (query row, head, batch): used when there are too few query tiles to fill the GPU (streaming steps)
__device__ __forceinline__ float block_max128(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
}
__device__ __forceinline__ float block_sum128(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return (red[0] + red[1]) + (red[2] + red[3]);
}

// softmax over S[0..n) by the whole CTA, then out[d] = sum_j S[j] * V[j][d]
__device__ __forceinline__ void row_softmax_pv(float* S, int n, const float* __restrict__ vb, int64_t ldv, float* __restrict__ op,
                                               float* red, float* part) {
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n; j += ATT_NT) mx = fmaxf(mx, S[j]);
  mx = block_max128(mx, red);
  float sum = 0.f;
  for (int j = threadIdx.x; j < n; j += ATT_NT) {
    float e = expf(S[j] - mx);
    S[j] = e;
    sum += e;
  }
  sum = block_sum128(sum, red);
  __syncthreads();
  const int d = threadIdx.x & 63, half = threadIdx.x >> 6;
  float acc = 0.f;
  for (int j = half; j < n; j += 2) acc = fmaf(S[j] / sum, vb[(int64_t)j * ldv + d], acc);
  if (half == 1) part[d] = acc;
  __syncthreads();
  if (half == 0) op[d] = acc + part[d];
}

__global__ void __launch_bounds__(ATT_NT) relpos_attention_row_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k,
                                                                      int ldk, const float* __restrict__ v, int ldv,
                                                                      const float* __restrict__ pos, int Tpos,
                                                                      const float* __restrict__ bias_u, const float* __restrict__ bias_v,
                                                                      float* __restrict__ out, int nQ, int q_offset, int T, int D,
                                                                      int chunk, const int* __restrict__ lengths) {
  extern __shared__ __align__(16) float smem[];
  float* S = smem;  // [T]
  __shared__ __align__(16) float qu[HD], qv[HD], part[HD];
  __shared__ float red[4];
  const int r = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int i = q_offset + r;
  const int len = lengths ? min(lengths[b], T) : T;
  const int lim = chunk > 0 ? min((i / chunk + 1) * chunk, T) : T;
  const int n = max(1, min(lim, len));
  const float* qb = q + ((int64_t)b * nQ + r) * ldq + h * HD;
  const float* kb = k + ((int64_t)b * T) * ldk + h * HD;
  const float* vb = v + ((int64_t)b * T) * ldv + h * HD;
  if (threadIdx.x < HD) {
    float val = qb[threadIdx.x];
    qu[threadIdx.x] = val + bias_u[h * HD + threadIdx.x];
    qv[threadIdx.x] = val + bias_v[h * HD + threadIdx.x];
  }
  __syncthreads();
  const float* pb = pos + h * HD;
  for (int j = threadIdx.x; j < n; j += ATT_NT) {
    float ac = dot64(qu, kb + (int64_t)j * ldk);
    float bd = dot64(qv, pb + (int64_t)(i - j + Tpos - 1) * D);
    S[j] = (ac + bd) * 0.125f;
  }
  __syncthreads();
  row_softmax_pv(S, n, vb, ldv, out + ((int64_t)b * nQ + r) * D + h * HD, red, part);
}

__global__ void __launch_bounds__(ATT_NT) mha_attention_row_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                                                                   const float* __restrict__ v, int ldv, float* __restrict__ out, int ldo,
                                                                   int Tq, int Tk, float scale, int causal, int causal_offset,
                                                                   const int* __restrict__ kv_len) {
  extern __shared__ __align__(16) float smem[];
  float* S = smem;
  __shared__ __align__(16) float qs[HD], part[HD];
  __shared__ float red[4];
  const int r = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int len = kv_len ? min(kv_len[b], Tk) : Tk;
  const int lim = causal ? min(r + causal_offset + 1, Tk) : Tk;
  const int n = max(1, min(lim, len));
  const float* qb = q + ((int64_t)b * Tq + r) * ldq + h * HD;
  const float* kb = k + ((int64_t)b * Tk) * ldk + h * HD;
  const float* vb = v + ((int64_t)b * Tk) * ldv + h * HD;
  if (threadIdx.x < HD) qs[threadIdx.x] = qb[threadIdx.x] * scale;
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += ATT_NT) S[j] = dot64(qs, kb + (int64_t)j * ldk);
  __syncthreads();
  row_softmax_pv(S, n, vb, ldv, out + ((int64_t)b * Tq + r) * ldo + h * HD, red, part);
}

}  // namespace

void relpos_attention(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, const float* pos, int Tpos,
                      const float* bias_u, const float* bias_v, float* out, int B, int nQ, int q_offset, int T, int H, int D,
                      int chunk, const int* lengths_dev, cudaStream_t st) {
  ++g_launches;
  if (B <= 0 || T <= 0 || nQ <= 0) return;
  if ((long)((nQ + QT - 1) / QT) * H * B < 96 && (size_t)T * sizeof(float) <= 160 * 1024) {  // too few tiles: one CTA per query row
    size_t smem_row = (size_t)((T + 3) & ~3) * sizeof(float);
    static size_t configured_row = 0;
    if (smem_row > 48 * 1024 && smem_row > configured_row) {
      cudaFuncSetAttribute(relpos_attention_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      configured_row = 160 * 1024;
    }
    relpos_attention_row_kernel<<<dim3(nQ, H, B), ATT_NT, smem_row, st>>>(q, ldq, k, ldk, v, ldv, pos, Tpos, bias_u, bias_v, out, nQ, q_offset,
                                                                          T, D, chunk, lengths_dev);
    return;
  }
  int ldS = (T + 3) & ~3;
  size_t smem = (size_t)(2 * QT * HD + QT * ldS) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(relpos_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  dim3 grid((nQ + QT - 1) / QT, H, B);
  relpos_attention_kernel<<<grid, ATT_NT, smem, st>>>(q, ldq, k, ldk, v, ldv, pos, Tpos, bias_u, bias_v, out, nQ, q_offset, T, H, D,
                                                      chunk, lengths_dev, ldS);
}

void mha_attention(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out, int ldo, int B,
                   int Tq, int Tk, int H, float scale, int causal, int causal_offset, const int* kv_len_dev,
                   cudaStream_t st) {
  ++g_launches;
  if (B <= 0 || Tq <= 0 || Tk <= 0) return;
  if ((long)((Tq + QT - 1) / QT) * H * B < 96 && (size_t)Tk * sizeof(float) <= 160 * 1024) {
    size_t smem_row = (size_t)((Tk + 3) & ~3) * sizeof(float);
    static size_t configured_row = 0;
    if (smem_row > 48 * 1024 && smem_row > configured_row) {
      cudaFuncSetAttribute(mha_attention_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      configured_row = 160 * 1024;
    }
    mha_attention_row_kernel<<<dim3(Tq, H, B), ATT_NT, smem_row, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, Tq, Tk, scale, causal, causal_offset,
                                                                       kv_len_dev);
    return;
  }
  int ldS = (Tk + 3) & ~3;
  size_t smem = (size_t)(QT * HD + QT * ldS) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(mha_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  dim3 grid((Tq + QT - 1) / QT, H, B);
  mha_attention_kernel<<<grid, ATT_NT, smem, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, Tq, Tk, scale, causal, causal_offset,
                                                   kv_len_dev, ldS);
}

}  // namespace ss
