// Skinny GEMM for the streaming regime (M <= 64 rows: cached encoder step M = 16, MT decode step M = 1):
//   out[m][n] = epilogue( sum_k A[m][k] * W[n][k] ),  A row-major [M][K], W row-major [N][K].
// These GEMMs are weight-streaming problems (2 MB of W for 16 rows), so the kernel is organised around memory-level
// parallelism instead of tiles:
//   * one warp per (output column, K-slice); lanes stride K with 128-bit loads, so a whole slice of the W row is in
//     flight at once; W is read exactly once (streaming, L1 no-allocate), A (<= 128 KB) is re-read through L1/L2;
//   * the KS K-slices of a column live in one CTA and are combined through shared memory in a fixed order
//     (deterministic results), then the epilogue (bias / ReLU / SiLU / GLU / residual) runs one row per lane.
// HBM/L2-bound by construction: algorithmic bytes = N*K*4 (weights) + M*K*4 + M*N*4.
#include "common.cuh"
#include "kernels.h"

namespace ss {
namespace {

constexpr int SK_WARPS = 8;
constexpr int SK_MR = 16;  // rows per pass (accumulators per lane)

__device__ __forceinline__ float sk_act(float x, int act) {
  switch (act) {
    case ACT_RELU: return x > 0.f ? x : 0.f;
    case ACT_SILU: return x / (1.0f + expf(-x));
    case ACT_TANH: return tanhf(x);
    default: return x;
  }
}

__device__ __forceinline__ float4 ld_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// KS = K-slices per column (warps cooperating on one column), CPT = columns per task (2 for GLU pairs)
template <int MR, int KS, int CPT>
__global__ void __launch_bounds__(SK_WARPS * 32) skinny_gemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                                                                    int M, int N, int K, Epilogue ep) {
  constexpr int TASKS_PER_CTA = SK_WARPS / KS;
  __shared__ float part[SK_WARPS][CPT][MR];
  __shared__ float ln_mean[64], ln_rstd[64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool fuse_ln = ep.ln_gamma != nullptr;
  // Programmatic dependent launch: let the next kernel of the stream start its own prologue now, and fetch this
  // warp's first slice of W (weights do not depend on the previous kernel) BEFORE waiting for the previous kernel's
  // results -- the weight-fetch latency of GEMM i+1 hides behind the execution of GEMM i.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  constexpr int PF = 2;  // prefetched 128-bit W loads per column and lane
  float4 wpre[CPT][PF];
  {
    const int slice_ = warp % KS, tslot_ = warp / KS;
    const int kslice_ = K / KS;
    const int task_ = blockIdx.x * (SK_WARPS / KS) + tslot_;
#pragma unroll
    for (int c = 0; c < CPT; ++c)
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        int k = slice_ * kslice_ + lane * 4 + i * 128;
        wpre[c][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (task_ < N / CPT && k < (slice_ + 1) * kslice_) wpre[c][i] = ld_stream(W + ((int64_t)task_ * CPT + c) * K + k);
      }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (fuse_ln) {
    // row statistics with the same operation order as layer_norm_kernel (two-pass, lane-strided, shuffle tree)
    for (int m = warp; m < M; m += SK_WARPS) {
      const float* xr = A + (int64_t)m * lda;
      float s = 0.f;
      for (int c = lane; c < K; c += 32) s += xr[c];
      float mean = warp_sum(s) / (float)K;
      float ss_ = 0.f;
      for (int c = lane; c < K; c += 32) {
        float d = xr[c] - mean;
        ss_ = fmaf(d, d, ss_);
      }
      float var = warp_sum(ss_) / (float)K;
      if (lane == 0) {
        ln_mean[m] = mean;
        ln_rstd[m] = 1.0f / sqrtf(var + 1e-5f);
      }
    }
    __syncthreads();
  }
  const int slice = warp % KS, tslot = warp / KS;
  const int ntasks = N / CPT;
  const int kslice = K / KS;  // multiple of 128 (checked by the host)
  const int k_lo = slice * kslice, k_hi = k_lo + kslice;
  for (int tbase = blockIdx.x * TASKS_PER_CTA; tbase < ntasks; tbase += gridDim.x * TASKS_PER_CTA) {
    const int task = tbase + tslot;
    const bool active = task < ntasks;
    const int n0 = task * CPT;
    for (int mb = 0; mb < M; mb += MR) {
      float acc[CPT][MR];
#pragma unroll
      for (int c = 0; c < CPT; ++c)
#pragma unroll
        for (int r = 0; r < MR; ++r) acc[c][r] = 0.f;
      if (active) {
        const float* w0 = W + (int64_t)n0 * K;
        const bool first = (tbase == blockIdx.x * TASKS_PER_CTA) && mb == 0;
        auto step = [&](int k, const float4* wv) {
          float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (fuse_ln) {
            g4 = *reinterpret_cast<const float4*>(ep.ln_gamma + k);
            b4 = *reinterpret_cast<const float4*>(ep.ln_beta + k);
          }
#pragma unroll
          for (int r = 0; r < MR; ++r) {
            if (mb + r < M) {
              float4 x = *reinterpret_cast<const float4*>(A + (int64_t)(mb + r) * lda + k);
              if (fuse_ln) {
                const float mu = ln_mean[mb + r], rs = ln_rstd[mb + r];
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
        };
        int k = k_lo + lane * 4;
#pragma unroll
        for (int i = 0; i < PF; ++i, k += 128) {
          if (k < k_hi) {
            float4 wv[CPT];
#pragma unroll
            for (int c = 0; c < CPT; ++c) wv[c] = first ? wpre[c][i] : ld_stream(w0 + (int64_t)c * K + k);
            step(k, wv);
          }
        }
#pragma unroll 2
        for (; k < k_hi; k += 128) {
          float4 wv[CPT];
#pragma unroll
          for (int c = 0; c < CPT; ++c) wv[c] = ld_stream(w0 + (int64_t)c * K + k);
          step(k, wv);
        }
      }
      // lane r ends up holding the warp total of row r
      float mine[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        mine[c] = 0.f;
#pragma unroll
        for (int r = 0; r < MR; ++r) {
          float t = warp_sum(acc[c][r]);
          if (lane == r) mine[c] = t;
        }
      }
      if (KS > 1) {
        if (lane < MR) {
#pragma unroll
          for (int c = 0; c < CPT; ++c) part[warp][c][lane] = mine[c];
        }
        __syncthreads();
        if (slice == 0 && lane < MR) {
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            float t = part[warp][c][lane];
#pragma unroll
            for (int s = 1; s < KS; ++s) t += part[warp + s][c][lane];
            mine[c] = t;
          }
        }
      }
      const int m = mb + lane;
      if (active && slice == 0 && lane < MR && m < M) {
        float y;
        int oc;
        if (CPT == 2) {
          float av = mine[0] + (ep.bias ? ep.bias[n0] : 0.f);
          float gv = mine[CPT - 1] + (ep.bias ? ep.bias[n0 + 1] : 0.f);
          y = ep.alpha * (av * (1.0f / (1.0f + expf(-gv))));
          oc = n0 >> 1;
        } else {
          y = ep.alpha * sk_act(mine[0] + (ep.bias ? ep.bias[n0] : 0.f), ep.act);
          oc = n0;
        }
        float* obase = ep.out;
        int ld = ep.ldo;
        if (ep.split_n > 0) {  // fused Q|K|V: route the column block to its own buffer
          int p = oc / ep.split_n;
          oc -= p * ep.split_n;
          if (p == 1) { obase = ep.out2; ld = ep.ldo2; }
          else if (p == 2) { obase = ep.out3; ld = ep.ldo3; }
        }
        int64_t o = (int64_t)m * ld + oc;
        if (ep.residual) y += ep.res_scale * ep.residual[o];
        if (ep.accumulate) y += obase[o];
        obase[o] = y;
      }
      if (KS > 1) __syncthreads();  // `part` is reused by the next pass / task
    }
  }
}

template <int MR, int KS>
void launch_skinny(const float* A, int lda, const float* W, int M, int N, int K, const Epilogue& ep, cudaStream_t st) {
  const int cpt = ep.glu ? 2 : 1;
  const int ntasks = N / cpt;
  const int per_cta = SK_WARPS / KS;
  int grid = (ntasks + per_cta - 1) / per_cta;
  // with a fused LayerNorm every CTA recomputes the row statistics: keep the grid at two CTAs per SM and loop over tasks
  const int cap = ep.ln_gamma ? 148 * 2 : 148 * 8;
  if (grid > cap) grid = cap;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(SK_WARPS * 32);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (ep.glu)
    cudaLaunchKernelEx(&cfg, skinny_gemm_kernel<MR, KS, 2>, A, lda, W, M, N, K, ep);
  else
    cudaLaunchKernelEx(&cfg, skinny_gemm_kernel<MR, KS, 1>, A, lda, W, M, N, K, ep);
}

}  // namespace

bool skinny_gemm_supported(int M, int N, int K, const Epilogue& ep) {
  if (M < 1 || M > 64 || (K & 127) != 0 || ep.out_L > 0) return false;
  if (ep.ln_gamma && K > 1024) return false;
  if (ep.glu && (N & 1)) return false;
  return true;
}

void skinny_gemm(const float* A, int lda, const float* W, int M, int N, int K, const Epilogue& ep, cudaStream_t st) {
  ++g_launches;
  // K-slices per column: enough warp-tasks to cover 148 SMs x 16 warps, every slice a multiple of 128 columns
  const int ntasks = ep.glu ? N / 2 : N;
  int ks = 1;
  while (ks < 8 && (long)ntasks * ks < 2400 && K % (ks * 2 * 128) == 0) ks *= 2;
  if (M == 1) {
    switch (ks) {
      case 1: launch_skinny<1, 1>(A, lda, W, M, N, K, ep, st); break;
      case 2: launch_skinny<1, 2>(A, lda, W, M, N, K, ep, st); break;
      case 4: launch_skinny<1, 4>(A, lda, W, M, N, K, ep, st); break;
      default: launch_skinny<1, 8>(A, lda, W, M, N, K, ep, st); break;
    }
  } else if (M <= 8) {
    switch (ks) {
      case 1: launch_skinny<8, 1>(A, lda, W, M, N, K, ep, st); break;
      case 2: launch_skinny<8, 2>(A, lda, W, M, N, K, ep, st); break;
      case 4: launch_skinny<8, 4>(A, lda, W, M, N, K, ep, st); break;
      default: launch_skinny<8, 8>(A, lda, W, M, N, K, ep, st); break;
    }
  } else {
    switch (ks) {
      case 1: launch_skinny<SK_MR, 1>(A, lda, W, M, N, K, ep, st); break;
      case 2: launch_skinny<SK_MR, 2>(A, lda, W, M, N, K, ep, st); break;
      case 4: launch_skinny<SK_MR, 4>(A, lda, W, M, N, K, ep, st); break;
      default: launch_skinny<SK_MR, 8>(A, lda, W, M, N, K, ep, st); break;
    }
  }
}

}  // namespace ss
