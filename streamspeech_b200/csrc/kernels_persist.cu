// Persistent cooperative kernel for one streaming step of the chunk-Conformer encoder stack.
//
// At batch 1 a 320 ms step touches 16 active rows and ~11 MB of fp32 weights per layer; run as separate kernels it is
// bounded by kernel boundaries (137 dependent launches x ~8 us, profiles/r1_*), not by HBM or FLOPs.  This kernel keeps
// one CTA per SM resident for ALL layers and replaces the kernel boundaries by grid-wide barriers:
//   per layer:  [LN+W1+SiLU] | [W2 + 0.5 res] | [LN + QKV -> q, K-cache, V-cache] | [rel-pos attention] | [out + res] |
//               [LN + PW1 + GLU -> conv cache] | [depthwise k31 + BN + SiLU] | [PW2 + res] | [LN+W1+SiLU] | [W2 + 0.5 res] | [LN]
// The GEMM phases use the skinny-GEMM scheme (warp per 4 columns x K-slice, 128-bit streaming loads of W, fixed-order
// reduction); attention is one CTA per (query row, head).  Same arithmetic as the multi-kernel path
// (ss_encoder_stream_step), which remains the fallback for shapes this kernel does not cover (nA > 16, D != 256, ...).
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"
#include "kernels_persist.h"

namespace cg = cooperative_groups;

namespace ss {
namespace {

constexpr int PW = 8;          // warps per CTA
constexpr int PT = PW * 32;    // threads per CTA
constexpr int PMR = 16;        // max active rows
constexpr int PHD = 64;        // head dim

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ float4 ldw(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

struct GemmEpi {
  const float* bias = nullptr;
  int act = ACT_NONE;       // ignored when GLU
  float alpha = 1.f;
  float* out = nullptr;     // column block 0
  int ldo = 0;
  bool residual = false;    // out += (in place)
  int split_n = 0;          // > 0: route column blocks of this width to out / out2 / out3
  float* out2 = nullptr;
  float* out3 = nullptr;
  int ldo2 = 0, ldo3 = 0;
};

struct Smem {
  float part[PW][4][PMR];
  float ln_mean[PMR], ln_rstd[PMR];
  float S[1024];
  float qa[PHD], qb[PHD];
  float pv[PW][PHD];
  float red[PW];
  unsigned long long* fine;  // profile mode, CTA 0, layer 1: clock64 stamps inside the phases
  int nfine;
};

__device__ __forceinline__ void fine_stamp(Smem& sm) {
  if (threadIdx.x == 0 && sm.fine != nullptr && sm.nfine < 200) sm.fine[sm.nfine++] = (unsigned long long)clock64();
}

// Activations are written by other CTAs between barriers: they are read with plain (coherent) loads, never through
// __restrict__ / ld.global.nc; only weights take the non-coherent streaming path.
// out[M][N'] = epi( LN?(A)[M][K] @ W[N][K]^T ), M <= 16.  All CTAs of the grid take part.
template <int CPT, bool GLU>
__device__ void phase_gemm(Smem& sm, const float* A, int lda, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                           const float* __restrict__ W, int M, int N, int K, int KS, const GemmEpi& ep) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool fuse_ln = ln_g != nullptr;
  fine_stamp(sm);
  if (fuse_ln) {
    // K == 256 here (checked by encoder_layers_persistent_supported): warp w owns rows w and w + 8, 8 values per lane each
    float v[2][8];
    float s[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = warp + h * PW;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[h][i] = m < M ? A[(int64_t)m * lda + lane + (i << 5)] : 0.f;
        s[h] += v[h][i];
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = warp + h * PW;
      const float mean = warp_sum(s[h]) / (float)K;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float d = v[h][i] - mean;
        q = fmaf(d, d, q);
      }
      const float var = warp_sum(q) / (float)K;
      if (lane == 0 && m < M) {
        sm.ln_mean[m] = mean;
        sm.ln_rstd[m] = 1.0f / sqrtf(var + 1e-5f);
      }
    }
    __syncthreads();
  }
  fine_stamp(sm);
  const int tpc = PW / KS;
  const int slice = warp % KS, tslot = warp / KS;
  const int ntasks = N / CPT;
  const int kslice = K / KS;
  const int k_lo = slice * kslice, k_hi = k_lo + kslice;
  for (int tbase = blockIdx.x * tpc; tbase < ntasks; tbase += gridDim.x * tpc) {
    const int task = tbase + tslot;
    const bool active = task < ntasks;
    const int n0 = task * CPT;
    float acc[CPT][PMR];
#pragma unroll
    for (int c = 0; c < CPT; ++c)
#pragma unroll
      for (int r = 0; r < PMR; ++r) acc[c][r] = 0.f;
    if (active) {
      const float* w0 = W + (int64_t)n0 * K;
#pragma unroll 2
      for (int k = k_lo + lane * 4; k < k_hi; k += 128) {
        float4 wv[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) wv[c] = ldw(w0 + (int64_t)c * K + k);
        float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fuse_ln) {
          g4 = *reinterpret_cast<const float4*>(ln_g + k);
          b4 = *reinterpret_cast<const float4*>(ln_b + k);
        }
#pragma unroll
        for (int r = 0; r < PMR; ++r) {
          if (r < M) {
            float4 x = *reinterpret_cast<const float4*>(A + (int64_t)r * lda + k);
            if (fuse_ln) {
              const float mu = sm.ln_mean[r], rs = sm.ln_rstd[r];
              x.x = (x.x - mu) * rs * g4.x + b4.x;
              x.y = (x.y - mu) * rs * g4.y + b4.y;
              x.z = (x.z - mu) * rs * g4.z + b4.z;
              x.w = (x.w - mu) * rs * g4.w + b4.w;
            }
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              acc[c][r] = fmaf(x.x, wv[c].x, acc[c][r]);
              acc[c][r] = fmaf(x.y, wv[c].y, acc[c][r]);
              acc[c][r] = fmaf(x.z, wv[c].z, acc[c][r]);
              acc[c][r] = fmaf(x.w, wv[c].w, acc[c][r]);
            }
          }
        }
      }
    }
    fine_stamp(sm);
    // 16 row sums per column across the warp: recursive halving (16 shuffles per column instead of 16 x 5); even lane
    // 2m ends up with the sum of row m.  Fixed order -> deterministic.
    float mine[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
      float w8[8], w4[4], w2[2];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float send = b4 ? acc[c][i] : acc[c][i + 8];
        float keep = b4 ? acc[c][i + 8] : acc[c][i];
        w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float send = b3 ? w8[i] : w8[i + 4];
        float keep = b3 ? w8[i + 4] : w8[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float send = b2 ? w4[i] : w4[i + 2];
        float keep = b2 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      float send = b1 ? w2[0] : w2[1];
      float keep = b1 ? w2[1] : w2[0];
      float w1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      mine[c] = w1 + __shfl_xor_sync(0xffffffffu, w1, 1);
    }
    const int m = lane >> 1;
    const bool owner = (lane & 1) == 0;
    fine_stamp(sm);
    if (KS > 1) {
      if (owner) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) sm.part[warp][c][m] = mine[c];
      }
      __syncthreads();
      if (slice == 0 && owner) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          float t = sm.part[warp][c][m];
          for (int s = 1; s < KS; ++s) t += sm.part[warp + s][c][m];
          mine[c] = t;
        }
      }
    }
    if (active && slice == 0 && owner && m < M) {
#pragma unroll
      for (int c = 0; c < (GLU ? 1 : CPT); ++c) {
        float y;
        int oc;
        if (GLU) {
          float av = mine[0] + (ep.bias ? ep.bias[n0] : 0.f);
          float gv = mine[CPT - 1] + (ep.bias ? ep.bias[n0 + 1] : 0.f);
          y = ep.alpha * (av * (1.0f / (1.0f + expf(-gv))));
          oc = n0 >> 1;
        } else {
          float v = mine[c] + (ep.bias ? ep.bias[n0 + c] : 0.f);
          if (ep.act == ACT_SILU) v = v / (1.0f + expf(-v));
          y = ep.alpha * v;
          oc = n0 + c;
        }
        float* obase = ep.out;
        int ld = ep.ldo;
        if (ep.split_n > 0) {
          int p = oc / ep.split_n;
          oc -= p * ep.split_n;
          if (p == 1) { obase = ep.out2; ld = ep.ldo2; }
          else if (p == 2) { obase = ep.out3; ld = ep.ldo3; }
        }
        int64_t o = (int64_t)m * ld + oc;
        if (ep.residual) y += obase[o];
        obase[o] = y;
      }
    }
    if (KS > 1) __syncthreads();
    fine_stamp(sm);
  }
}

// rel-pos attention for the active rows: one CTA per (row, head) task
__device__ void phase_attention(Smem& sm, const float* q, const float* kc, const float* vc,
                                const float* __restrict__ pos, int Tpos, const float* __restrict__ bias_u, const float* __restrict__ bias_v,
                                float* out, int nA, int a0, int T, int D, int H, int chunk) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int task = blockIdx.x; task < nA * H; task += gridDim.x) {
    const int r = task / H, h = task - r * H;
    const int i = a0 + r;
    const int lim = chunk > 0 ? min((i / chunk + 1) * chunk, T) : T;
    const int n = max(1, lim);
    __syncthreads();  // smem reuse across tasks
    if (tid < PHD) {
      float val = q[(int64_t)r * D + h * PHD + tid];
      sm.qa[tid] = val + bias_u[h * PHD + tid];
      sm.qb[tid] = val + bias_v[h * PHD + tid];
    }
    __syncthreads();
    const float* kb = kc + h * PHD;
    const float* vb = vc + h * PHD;
    const float* pb = pos + h * PHD;
    float mx = -INFINITY;
    for (int j = tid; j < n; j += PT) {
      const float* kr = kb + (int64_t)j * D;
      const float* pr = pb + (int64_t)(i - j + Tpos - 1) * D;
      float ac = 0.f, bd = 0.f;
#pragma unroll
      for (int d = 0; d < PHD; d += 4) {
        float4 kk = *reinterpret_cast<const float4*>(kr + d);
        float4 pp = *reinterpret_cast<const float4*>(pr + d);
        ac = fmaf(sm.qa[d], kk.x, ac); ac = fmaf(sm.qa[d + 1], kk.y, ac); ac = fmaf(sm.qa[d + 2], kk.z, ac); ac = fmaf(sm.qa[d + 3], kk.w, ac);
        bd = fmaf(sm.qb[d], pp.x, bd); bd = fmaf(sm.qb[d + 1], pp.y, bd); bd = fmaf(sm.qb[d + 2], pp.z, bd); bd = fmaf(sm.qb[d + 3], pp.w, bd);
      }
      float s = (ac + bd) * 0.125f;
      sm.S[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if (lane == 0) sm.red[w] = mx;
    __syncthreads();
    mx = sm.red[0];
#pragma unroll
    for (int x = 1; x < PW; ++x) mx = fmaxf(mx, sm.red[x]);
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < n; j += PT) {
      float e = expf(sm.S[j] - mx);
      sm.S[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) sm.red[w] = sum;
    __syncthreads();
    sum = sm.red[0];
#pragma unroll
    for (int x = 1; x < PW; ++x) sum += sm.red[x];
    float a0_ = 0.f, a1_ = 0.f;
#pragma unroll 4
    for (int j = w; j < n; j += PW) {
      float p = sm.S[j];
      float2 vv = *reinterpret_cast<const float2*>(vb + (int64_t)j * D + 2 * lane);
      a0_ = fmaf(p, vv.x, a0_);
      a1_ = fmaf(p, vv.y, a1_);
    }
    sm.pv[w][2 * lane] = a0_;
    sm.pv[w][2 * lane + 1] = a1_;
    __syncthreads();
    if (tid < PHD) {
      float t = 0.f;
#pragma unroll
      for (int x = 0; x < PW; ++x) t += sm.pv[x][tid];
      out[(int64_t)r * D + h * PHD + tid] = t / sum;
    }
  }
}

__device__ void phase_depthwise(const float* gc, const float* __restrict__ w, const float* __restrict__ scale,
                                const float* __restrict__ shift, float* dw, int nA, int a0, int T, int D, int k, int chunk) {
  const int half = (k - 1) >> 1;
  for (int r = blockIdx.x; r < nA; r += gridDim.x) {
    const int t = a0 + r;
    const int lim = chunk > 0 ? min(T, (t / chunk + 1) * chunk) : T;
    for (int c = threadIdx.x; c < D; c += PT) {
      float acc = 0.f;
      for (int j0 = 0; j0 < k; j0 += 8) {
        float xv[8], wv[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          int j = j0 + jj, p = t - half + j;
          bool ok = j < k && p >= 0 && p < lim;
          xv[jj] = ok ? gc[(int64_t)p * D + c] : 0.f;
          wv[jj] = ok ? w[j * D + c] : 0.f;
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) acc = fmaf(wv[jj], xv[jj], acc);
      }
      float v = acc * scale[c] + shift[c];
      dw[(int64_t)r * D + c] = v / (1.0f + expf(-v));
    }
  }
}

__device__ void phase_layer_norm(float* x, const float* __restrict__ g, const float* __restrict__ b, int M, int D) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * PW + (threadIdx.x >> 5);
  if (gw >= M) return;
  float* xr = x + (int64_t)gw * D;
  float v[8];  // D == 256
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = xr[lane + (i << 5)];
    s += v[i];
  }
  float mean = warp_sum(s) / (float)D;
  float ss_ = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float d = v[i] - mean;
    ss_ = fmaf(d, d, ss_);
  }
  float rstd = 1.0f / sqrtf(warp_sum(ss_) / (float)D + 1e-5f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int c = lane + (i << 5);
    xr[c] = (v[i] - mean) * rstd * g[c] + b[c];
  }
}

__global__ void __launch_bounds__(PT, 1) encoder_layers_persistent_kernel(const PersistLayer* __restrict__ layers, int n_layers, float* x,
                                                                          float* hid, float* qb, float* att, float* dw, float* kc_all,
                                                                          float* vc_all, float* gc_all, int nA, int a0, int T, int D, int FFN,
                                                                          int H, int Tpos, int chunk, int conv_chunk, int dw_k, unsigned long long* ts) {
  cg::grid_group grid = cg::this_grid();
  __shared__ Smem sm;
  int nts = 0;
  const bool stamp = ts != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (ts != nullptr) {  // profile mode: two back-to-back barriers first (pure barrier cost)
    if (stamp) ts[nts++] = globaltimer_ns();
    grid.sync();
    if (stamp) ts[nts++] = globaltimer_ns();
    grid.sync();
    if (stamp) ts[nts++] = globaltimer_ns();
  }
  if (threadIdx.x == 0) { sm.fine = nullptr; sm.nfine = 0; }
  for (int li = 0; li < n_layers; ++li) {
    if (threadIdx.x == 0) sm.fine = (ts != nullptr && blockIdx.x == 0 && li == 1) ? ts + 256 : nullptr;
    const PersistLayer L = layers[li];
    float* kc = kc_all + (size_t)li * Tpos * D;
    float* vc = vc_all + (size_t)li * Tpos * D;
    float* gc = gc_all + (size_t)li * Tpos * D;
    GemmEpi e;
    // x = x + 0.5 * W2(SiLU(W1 LN(x)))
    e = GemmEpi(); e.bias = L.ffn1_b1; e.act = ACT_SILU; e.out = hid; e.ldo = FFN;
    phase_gemm<4, false>(sm, x, D, L.ffn1_g, L.ffn1_b, L.ffn1_w1, nA, FFN, D, 2, e);
    grid.sync();
    if (stamp) ts[nts++] = globaltimer_ns();
    fine_stamp(sm);
    e = GemmEpi(); e.bias = L.ffn1_b2; e.alpha = 0.5f; e.out = x; e.ldo = D; e.residual = true;
    phase_gemm<2, false>(sm, hid, FFN, nullptr, nullptr, L.ffn1_w2, nA, D, FFN, 8, e);
    grid.sync();
    if (stamp) ts[nts++] = globaltimer_ns();
    // q -> qb, k / v -> cache rows a0..
    e = GemmEpi(); e.bias = L.bqkv; e.out = qb; e.ldo = D; e.split_n = D; e.out2 = kc + (size_t)a0 * D; e.ldo2 = D; e.out3 = vc + (size_t)a0 * D; e.ldo3 = D;
    phase_gemm<2, false>(sm, x, D, L.attn_g, L.attn_b, L.wqkv, nA, 3 * D, D, 2, e);
    grid.sync();
    if (stamp) ts[nts++] = globaltimer_ns();
    fine_stamp(sm);
    phase_attention(sm, qb, kc, vc, L.pos_proj, Tpos, L.pos_u, L.pos_v, att, nA, a0, T, D, H, chunk);
    grid.sync();
    if (stamp) ts[nts++] = globaltimer_ns();
    fine_stamp(sm);
    e = GemmEpi(); e.bias = L.bo; e.out = x; e.ldo = D; e.residual = true;
    phase_gemm<1, false>(sm, att, D, nullptr, nullptr, L.wo, nA, D, D, 2, e);
    grid.sync();
    if (stamp) ts[nts++] = globaltimer_ns();
    // conv module
    e = GemmEpi(); e.bias = L.pw1_b; e.out = gc + (size_t)a0 * D; e.ldo = D;
    phase_gemm<2, true>(sm, x, D, L.conv_g, L.conv_b, L.pw1, nA, 2 * D, D, 2, e);
    grid.sync();
    if (stamp) ts[nts++] = globaltimer_ns();
    fine_stamp(sm);
    phase_depthwise(gc, L.dw_w, L.bn_scale, L.bn_shift, dw, nA, a0, T, D, dw_k, conv_chunk);
    grid.sync();
    if (stamp) ts[nts++] = globaltimer_ns();
    fine_stamp(sm);
    e = GemmEpi(); e.bias = L.pw2_b; e.out = x; e.ldo = D; e.residual = true;
    phase_gemm<1, false>(sm, dw, D, nullptr, nullptr, L.pw2, nA, D, D, 2, e);
    grid.sync();
    if (stamp) ts[nts++] = globaltimer_ns();
    fine_stamp(sm);
    e = GemmEpi(); e.bias = L.ffn2_b1; e.act = ACT_SILU; e.out = hid; e.ldo = FFN;
    phase_gemm<4, false>(sm, x, D, L.ffn2_g, L.ffn2_b, L.ffn2_w1, nA, FFN, D, 2, e);
    grid.sync();
    if (stamp) ts[nts++] = globaltimer_ns();
    fine_stamp(sm);
    e = GemmEpi(); e.bias = L.ffn2_b2; e.alpha = 0.5f; e.out = x; e.ldo = D; e.residual = true;
    phase_gemm<2, false>(sm, hid, FFN, nullptr, nullptr, L.ffn2_w2, nA, D, FFN, 8, e);
    grid.sync();
    if (stamp) ts[nts++] = globaltimer_ns();
    fine_stamp(sm);
    phase_layer_norm(x, L.fin_g, L.fin_b, nA, D);
    grid.sync();
    if (stamp) ts[nts++] = globaltimer_ns();
  }
}

}  // namespace

bool encoder_layers_persistent_supported(int nA, int D, int FFN, int H, int T, int dw_k) {
  return nA >= 1 && nA <= PMR && D == 256 && (FFN % 1024) == 0 && H * PHD == D && T <= 1024 && dw_k <= 64;
}

int encoder_layers_persistent(const PersistLayer* layers_dev, int n_layers, float* x, float* hid, float* qb, float* att, float* dw, float* kc,
                              float* vc, float* gc, int nA, int a0, int T, int D, int FFN, int H, int Tpos, int chunk, int conv_chunk, int dw_k,
                              unsigned long long* ts, cudaStream_t st) {
  ++g_launches;
  static int grid = 0;
  if (grid == 0) {
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, encoder_layers_persistent_kernel, PT, 0);
    if (occ < 1) return -1;
    grid = sms;
  }
  void* args[] = {(void*)&layers_dev, (void*)&n_layers, (void*)&x, (void*)&hid, (void*)&qb, (void*)&att, (void*)&dw, (void*)&kc, (void*)&vc,
                  (void*)&gc, (void*)&nA, (void*)&a0, (void*)&T, (void*)&D, (void*)&FFN, (void*)&H, (void*)&Tpos, (void*)&chunk,
                  (void*)&conv_chunk, (void*)&dw_k, (void*)&ts};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)encoder_layers_persistent_kernel, dim3(grid), dim3(PT), args, 0, st);
  return e == cudaSuccess ? 0 : -2;
}

}  // namespace ss
