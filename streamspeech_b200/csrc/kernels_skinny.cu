// Skinny GEMM for the streaming regime (M <= 64 rows: cached encoder step M = 16, MT decode step M = 1):
//   out[m][n] = epilogue( sum_k A[m][k] * W[n][k] ),  A row-major [M][K], W row-major [N][K].
// These GEMMs are weight-streaming problems (2 MB of W for 16 rows), so the kernel is organised around memory-level
// parallelism instead of tiles: one warp per output column, lanes stride K with 128-bit loads (a whole W row is in
// flight at once), A staged once per CTA in shared memory, M accumulators per lane, shuffle reduction, fused epilogue
// (bias / ReLU / SiLU / GLU / residual).  Deterministic summation order.  HBM/L2-bound by construction.
#include "common.cuh"
#include "kernels.h"

namespace ss {
namespace {

constexpr int SK_WARPS = 8;

__device__ __forceinline__ float sk_act(float x, int act) {
  switch (act) {
    case ACT_RELU: return x > 0.f ? x : 0.f;
    case ACT_SILU: return x / (1.0f + expf(-x));
    case ACT_TANH: return tanhf(x);
    default: return x;
  }
}

// MR = rows handled per pass (accumulators per lane)
template <int MR>
__global__ void __launch_bounds__(SK_WARPS * 32) skinny_gemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                                                                    int M, int N, int K, Epilogue ep) {
  extern __shared__ __align__(16) float As[];  // [M][K]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // stage A (coalesced 128-bit loads)
  const int kv = K >> 2;
  for (int i = tid; i < M * kv; i += SK_WARPS * 32) {
    int m = i / kv, k4 = i - m * kv;
    reinterpret_cast<float4*>(As)[i] = *reinterpret_cast<const float4*>(A + (int64_t)m * lda + (k4 << 2));
  }
  __syncthreads();
  const int cols_per_task = ep.glu ? 2 : 1;
  const int ntasks = N / cols_per_task;
  for (int task = blockIdx.x * SK_WARPS + warp; task < ntasks; task += gridDim.x * SK_WARPS) {
    const int n0 = task * cols_per_task;
    for (int mb = 0; mb < M; mb += MR) {
      float acc0[MR], acc1[MR];
#pragma unroll
      for (int r = 0; r < MR; ++r) acc0[r] = acc1[r] = 0.f;
      const float* w0 = W + (int64_t)n0 * K;
      const float* w1 = w0 + K;
#pragma unroll 4
      for (int k = lane * 4; k < K; k += 128) {
        float4 a0 = *reinterpret_cast<const float4*>(w0 + k);
        float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cols_per_task == 2) a1 = *reinterpret_cast<const float4*>(w1 + k);
#pragma unroll
        for (int r = 0; r < MR; ++r) {
          if (mb + r < M) {
            float4 x = *reinterpret_cast<const float4*>(As + (mb + r) * K + k);
            acc0[r] = fmaf(x.x, a0.x, acc0[r]);
            acc0[r] = fmaf(x.y, a0.y, acc0[r]);
            acc0[r] = fmaf(x.z, a0.z, acc0[r]);
            acc0[r] = fmaf(x.w, a0.w, acc0[r]);
            if (cols_per_task == 2) {
              acc1[r] = fmaf(x.x, a1.x, acc1[r]);
              acc1[r] = fmaf(x.y, a1.y, acc1[r]);
              acc1[r] = fmaf(x.z, a1.z, acc1[r]);
              acc1[r] = fmaf(x.w, a1.w, acc1[r]);
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < MR; ++r) {
        acc0[r] = warp_sum(acc0[r]);
        if (cols_per_task == 2) acc1[r] = warp_sum(acc1[r]);
      }
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < MR; ++r) {
          int m = mb + r;
          if (m >= M) break;
          float y;
          int oc;
          if (ep.glu) {
            float av = acc0[r] + (ep.bias ? ep.bias[n0] : 0.f);
            float gv = acc1[r] + (ep.bias ? ep.bias[n0 + 1] : 0.f);
            y = ep.alpha * (av * (1.0f / (1.0f + expf(-gv))));
            oc = n0 >> 1;
          } else {
            y = ep.alpha * sk_act(acc0[r] + (ep.bias ? ep.bias[n0] : 0.f), ep.act);
            oc = n0;
          }
          int64_t o = (int64_t)m * ep.ldo + oc;
          if (ep.residual) y += ep.res_scale * ep.residual[o];
          if (ep.accumulate) y += ep.out[o];
          ep.out[o] = y;
        }
      }
    }
  }
}

}  // namespace

bool skinny_gemm_supported(int M, int N, int K, const Epilogue& ep) {
  if (M < 1 || M > 64 || (K & 127) != 0 || ep.out_L > 0) return false;
  if (ep.glu && (N & 1)) return false;
  return (size_t)M * K * sizeof(float) <= 200 * 1024;
}

void skinny_gemm(const float* A, int lda, const float* W, int M, int N, int K, const Epilogue& ep, cudaStream_t st) {
  ++g_launches;
  const size_t smem = (size_t)M * K * sizeof(float);
  const int ntasks = ep.glu ? N / 2 : N;
  int grid = (ntasks + SK_WARPS - 1) / SK_WARPS;
  const int max_grid = smem > 100 * 1024 ? 148 : 148 * 4;
  if (grid > max_grid) grid = max_grid;
  static size_t configured[3] = {0, 0, 0};
  auto launch = [&](auto kernel, int slot) {
    if (smem > 48 * 1024 && smem > configured[slot]) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      configured[slot] = 200 * 1024;
    }
    kernel<<<grid, SK_WARPS * 32, smem, st>>>(A, lda, W, M, N, K, ep);
  };
  if (M == 1) launch(skinny_gemm_kernel<1>, 0);
  else if (M <= 8) launch(skinny_gemm_kernel<8>, 1);
  else launch(skinny_gemm_kernel<16>, 2);
}

}  // namespace ss
