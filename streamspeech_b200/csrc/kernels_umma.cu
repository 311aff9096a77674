// tcgen05 (5th-gen tensor core) GEMM with fp32-grade accuracy through bf16 operand splitting.
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T ),  A = fp32 matrix or implicit-im2col view of a channels-last activation
//   (same ConvA / Epilogue contract as gemm_conv), W = fp32 row-major [N][K].
//
// Every fp32 operand x is split into bf16 pieces x = x0 + x1 (+ x2) while it is staged into shared memory, and the
// product is accumulated in TMEM (fp32) as a sum of bf16 MMAs:
//   NSPLIT = 3 :  a0*w0 + a0*w1 + a1*w0                        (~2^-16 relative: vocoder, 1e-3 waveform tolerance)
//   NSPLIT = 6 :  + a0*w2 + a2*w0 + a1*w1                      (~2^-23 relative: layers that feed an arg-max)
// so the tensor pipe does 3-6x the nominal work, which is irrelevant next to the operand staging cost.
//
// Structure (one CTA = 128 threads = one 128 x BN output tile, cta_group::1):
//   * all threads gather A / W tiles (BK = 32 columns) from global memory, split, and store them in the canonical
//     no-swizzle K-major UMMA layout: 8x8 bf16 core matrices (128 B), SBO = 128 B between 8-row groups, LBO between
//     8-column chunks (cute::UMMA::make_umma_desc<Major::K>, LayoutType::INTERLEAVE);
//   * fence.proxy.async + __syncthreads, then ONE thread issues tcgen05.mma (M = 128, N = BN, K = 16) per k-step and
//     piece pair and commits to an mbarrier; the accumulator lives in TMEM (BN fp32 columns x 128 lanes);
//   * epilogue: each warp reads its 32 TMEM lanes with tcgen05.ld.32x32b, applies bias / activation / GLU / residual
//     and writes fp32 rows.
// Two smem stages are used so that the gather of tile k+1 overlaps the MMAs of tile k.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace ss {
namespace {

constexpr int UM_BM = 128;
constexpr int UM_BK = 32;
constexpr int UM_NT = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float um_act(float x, int act) {
  switch (act) {
    case ACT_RELU: return x > 0.f ? x : 0.f;
    case ACT_SILU: return x / (1.0f + expf(-x));
    case ACT_TANH: return tanhf(x);
    default: return x;
  }
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);            // start address  [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset [16,30): next 8-column chunk along K
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;  // stride byte offset  [32,46): next 8-row group along M/N
  d |= (uint64_t)1 << 46;                            // version = 1 (Blackwell)
  return d;                                          // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(int n) {
  uint32_t d = 0;
  d |= 1u << 4;                    // c_format = F32
  d |= 1u << 7;                    // a_format = BF16
  d |= 1u << 10;                   // b_format = BF16
  d |= (uint32_t)(n >> 3) << 17;   // n_dim
  d |= (uint32_t)(UM_BM >> 4) << 24;  // m_dim
  return d;
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

// split 4 consecutive fp32 values into NP bf16 pieces and store each piece's 4 values (8 bytes) at byte offset `off`
// of its own tile
template <int NP>
__device__ __forceinline__ void split_store(float4 v, unsigned char* const* piece_base, uint32_t off) {
  float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    __nv_bfloat16 b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      b[i] = __float2bfloat16_rn(r[i]);
      r[i] -= __bfloat162float(b[i]);
    }
    uint2 packed;
    packed.x = (uint32_t)__bfloat16_as_ushort(b[0]) | ((uint32_t)__bfloat16_as_ushort(b[1]) << 16);
    packed.y = (uint32_t)__bfloat16_as_ushort(b[2]) | ((uint32_t)__bfloat16_as_ushort(b[3]) << 16);
    *reinterpret_cast<uint2*>(piece_base[p] + off) = packed;
  }
}

__device__ __forceinline__ float4 load_a4_conv(const ConvA& a, int m, int kk, int M, int K) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (m >= M || kk >= K) return v;
  int b = m / a.L_rows;
  int t = m - b * a.L_rows + a.t_offset;
  int tap = kk / a.C_in;
  int ci = kk - tap * a.C_in;
  int center = t * a.stride;
  int pos = center + tap * a.dil - a.pad_left;
  if (pos < 0 || pos >= a.L_in) return v;
  if (a.chunk > 0 && pos >= (center / a.chunk + 1) * a.chunk) return v;
  int xr = a.x_rows > 0 ? a.x_rows : a.L_in;
  int pr = pos - a.x_row0;
  if (pr < 0 || pr >= xr) return v;
  v = *reinterpret_cast<const float4*>(a.x + ((int64_t)b * xr + pr) * a.ldx + ci);
  if (a.pre_lrelu != 1.0f) {
    v.x = v.x > 0.f ? v.x : v.x * a.pre_lrelu;
    v.y = v.y > 0.f ? v.y : v.y * a.pre_lrelu;
    v.z = v.z > 0.f ? v.z : v.z * a.pre_lrelu;
    v.w = v.w > 0.f ? v.w : v.w * a.pre_lrelu;
  }
  return v;
}

// NP = bf16 pieces per operand (2 -> 3 MMAs, 3 -> 6 MMAs)
// gridDim.z > 1: split-K, slice z covers k-tiles [z*tiles_per_split, ...) and stores raw partial sums to ws[z][M][N]
// (reduced in a fixed order by splitk_epilogue, which also bounds the length of each tensor-core accumulation chain).
template <int BN, int NP>
__global__ void __launch_bounds__(UM_NT) umma_gemm_kernel(ConvA a, const float* __restrict__ W, int M, int N, int K, Epilogue ep,
                                                          int tiles_per_split, float* __restrict__ ws) {
  constexpr int A_TILE = UM_BM * UM_BK * 2;  // bytes of one bf16 piece of the A tile
  constexpr int B_TILE = BN * UM_BK * 2;
  constexpr int STAGE = NP * (A_TILE + B_TILE);
  constexpr uint32_t A_LBO = (UM_BM / 8) * 128, B_LBO = (BN / 8) * 128, SBO = 128;
  constexpr int TM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * UM_BM, n0 = blockIdx.x * BN;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar[0])), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar[1])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = make_idesc(BN);

  const int nk_all = (K + UM_BK - 1) / UM_BK;
  const int kt0 = blockIdx.z * tiles_per_split;
  const int nk = min(nk_all, kt0 + tiles_per_split) - kt0;
  uint32_t phase[2] = {0, 0};
  // per-thread staging assignment is the same for every k-tile: rows r = (tid>>3) + 32*i, quad q = tid&7
  constexpr int A_IT = UM_BM * (UM_BK / 4) / UM_NT;  // 4
  constexpr int B_IT = (BN * (UM_BK / 4) + UM_NT - 1) / UM_NT;
  const int q = tid & 7;
  // register double buffering: the global loads of tile kt+1 are issued before tile kt is split / stored, so the
  // load latency overlaps the bf16 conversion and the (asynchronous) MMAs of the previous tiles
  float4 va[A_IT], vb[B_IT];
  auto issue_loads = [&](int kt_abs) {
    const int k0 = kt_abs * UM_BK;
#pragma unroll
    for (int i = 0; i < A_IT; ++i) va[i] = load_a4_conv(a, m0 + (tid >> 3) + 32 * i, k0 + q * 4, M, K);
#pragma unroll
    for (int i = 0; i < B_IT; ++i) {
      int r = (tid >> 3) + 32 * i;
      int n = n0 + r, kk = k0 + q * 4;
      vb[i] = (r < BN && n < N && kk < K) ? *reinterpret_cast<const float4*>(W + (int64_t)n * K + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (nk > 0) issue_loads(kt0);
  for (int kt = 0; kt < nk; ++kt) {
    const int st = kt & 1;
    unsigned char* sbase = smem + st * STAGE;
    unsigned char* a_piece[NP];
    unsigned char* b_piece[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      a_piece[p] = sbase + p * A_TILE;
      b_piece[p] = sbase + NP * A_TILE + p * B_TILE;
    }
    // current tile: registers -> local copies, then start fetching the next tile
    float4 ca[A_IT], cb[B_IT];
#pragma unroll
    for (int i = 0; i < A_IT; ++i) ca[i] = va[i];
#pragma unroll
    for (int i = 0; i < B_IT; ++i) cb[i] = vb[i];
    if (kt + 1 < nk) issue_loads(kt0 + kt + 1);
    if (kt >= 2) {  // the MMAs that read this stage two tiles ago must have completed
      mbar_wait(smem_u32(&mbar[st]), phase[st]);
      phase[st] ^= 1;
    }
#pragma unroll
    for (int i = 0; i < A_IT; ++i) {
      int r = (tid >> 3) + 32 * i;
      uint32_t off = (uint32_t)(((q >> 1) * (UM_BM / 8) + (r >> 3)) * 128 + (r & 7) * 16 + (q & 1) * 8);
      split_store<NP>(ca[i], a_piece, off);
    }
#pragma unroll
    for (int i = 0; i < B_IT; ++i) {
      int r = (tid >> 3) + 32 * i;
      if (r < BN) {
        uint32_t off = (uint32_t)(((q >> 1) * (BN / 8) + (r >> 3)) * 128 + (r & 7) * 16 + (q & 1) * 8);
        split_store<NP>(cb[i], b_piece, off);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < UM_BK / 16; ++ks) {
        uint64_t ad[NP], bd[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          ad[p] = make_desc(smem_u32(a_piece[p]) + ks * 2 * A_LBO, A_LBO, SBO);
          bd[p] = make_desc(smem_u32(b_piece[p]) + ks * 2 * B_LBO, B_LBO, SBO);
        }
        const uint32_t first = (kt == 0 && ks == 0) ? 0u : 1u;
        mma_bf16(tmem_d, ad[0], bd[0], idesc, first);
        mma_bf16(tmem_d, ad[0], bd[1], idesc, 1u);
        mma_bf16(tmem_d, ad[1], bd[0], idesc, 1u);
        if (NP == 3) {
          mma_bf16(tmem_d, ad[1], bd[1], idesc, 1u);
          mma_bf16(tmem_d, ad[0], bd[NP - 1], idesc, 1u);
          mma_bf16(tmem_d, ad[NP - 1], bd[0], idesc, 1u);
        }
      }
      // arrives on the stage's mbarrier when every MMA issued so far has finished reading smem / writing TMEM
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[st])) : "memory");
    }
  }
  // wait for the last commits (both stages) before reading the accumulator
  for (int s = 0; s < 2; ++s) {
    int uses = (nk + 1 - s) / 2;  // tiles that used stage s
    if (uses > 0) mbar_wait(smem_u32(&mbar[s]), phase[s]);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: warp w reads TMEM lane quadrant (w & 3) = output rows m0 + 32*(w&3) + lane; the two warpgroups
  // split the columns
  const int quad = warp & 3;
  const int m = m0 + quad * 32 + lane;
  constexpr int COLS_PER_GROUP = BN >= 32 ? BN / 2 : BN;
  const int c_begin = (BN >= 32) ? (warp >> 2) * COLS_PER_GROUP : 0;
  const bool epi_active = (BN >= 32) || warp < 4;
  int64_t orow = m;
  if (ep.out_L > 0 && m < M) {
    int b = m / a.L_rows;
    int t = m - b * a.L_rows;
    orow = (int64_t)b * ep.out_L + (int64_t)t * ep.out_row_stride + ep.out_row_offset;
  }
#pragma unroll 1
  for (int c0 = c_begin; epi_active && c0 < c_begin + COLS_PER_GROUP; c0 += 16) {
    uint32_t r[16];
    uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (ws != nullptr) {  // split-K: raw partial sums
      if (m < M) {
        float* wz = ws + ((int64_t)blockIdx.z * M + m) * N;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          int n = n0 + c0 + j;
          if (n < N) wz[n] = __uint_as_float(r[j]);
        }
      }
      continue;
    }
    if (m < M) {
      float* orow_p = ep.out + orow * ep.ldo;
      const float* rrow_p = ep.residual ? ep.residual + orow * ep.ldo : nullptr;
      if (ep.glu) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          int n = n0 + c0 + j;
          if (n >= N) continue;
          float av = __uint_as_float(r[j]) + (ep.bias ? ep.bias[n] : 0.f);
          float gv = __uint_as_float(r[j + 1]) + (ep.bias ? ep.bias[n + 1] : 0.f);
          float y = ep.alpha * (av * (1.0f / (1.0f + expf(-gv))));
          int oc = n >> 1;
          if (rrow_p) y += ep.res_scale * rrow_p[oc];
          if (ep.accumulate) y += orow_p[oc];
          orow_p[oc] = y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          int n = n0 + c0 + j;
          if (n >= N) continue;
          float y = ep.alpha * um_act(__uint_as_float(r[j]) + (ep.bias ? ep.bias[n] : 0.f), ep.act);
          if (rrow_p) y += ep.res_scale * rrow_p[n];
          if (ep.accumulate) y += orow_p[n];
          orow_p[n] = y;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TM_COLS) : "memory");
  }
}

template <int BN, int NP>
void launch_umma(const ConvA& a, const float* W, int M, int N, int K, const Epilogue& ep, cudaStream_t st) {
  constexpr int STAGE = NP * (UM_BM * UM_BK * 2 + BN * UM_BK * 2);
  const size_t smem = 2 * STAGE + 1024;
  if (first_time_on_device((const void*)umma_gemm_kernel<BN, NP>))
    cudaFuncSetAttribute(umma_gemm_kernel<BN, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((N + BN - 1) / BN, (M + UM_BM - 1) / UM_BM);
  const long ctas = (long)grid.x * grid.y;
  const int nk = (K + UM_BK - 1) / UM_BK;
  int splits = 1;
  float* ws = nullptr;
  // the register-staged operand path keeps only ~32 KB of loads in flight per CTA: aim for 3 resident CTAs per SM
  if (ctas < 444 && nk >= 8) {
    splits = (int)std::min<long>((444 + ctas - 1) / ctas, nk / 4);
    ws = splitk_workspace((size_t)splits * M * N * sizeof(float));
    if (!ws) splits = 1;
  }
  int tiles = nk;
  if (splits > 1) {
    tiles = (nk + splits - 1) / splits;
    splits = (nk + tiles - 1) / tiles;
    grid.z = splits;
  } else {
    ws = nullptr;
  }
  umma_gemm_kernel<BN, NP><<<grid, UM_NT, smem, st>>>(a, W, M, N, K, ep, tiles, ws);
  if (splits > 1) splitk_epilogue(ws, splits, M, N, a.L_rows, ep, st);
}

}  // namespace

bool umma_gemm_supported(const ConvA& a, int N, const Epilogue& ep) {
  const int K = a.ksize * a.C_in;
  if (N < 16 || (N & 15) != 0 || (K & 3) != 0 || (a.C_in & 3) != 0) return false;
  if (a.lengths != nullptr || ep.ln_gamma != nullptr || ep.split_n > 0) return false;
  return true;
}

void umma_gemm_conv(const ConvA& a, const float* W, int N, const Epilogue& ep, int pieces, cudaStream_t st) {
  ++g_launches;
  const int M = a.B * a.L_rows;
  const int K = a.ksize * a.C_in;
  if (M <= 0) return;
  const int bn = (N % 128 == 0) ? 128 : (N % 64 == 0) ? 64 : (N % 32 == 0) ? 32 : 16;
  if (pieces >= 3) {
    switch (bn) {
      case 128: launch_umma<128, 3>(a, W, M, N, K, ep, st); break;
      case 64: launch_umma<64, 3>(a, W, M, N, K, ep, st); break;
      case 32: launch_umma<32, 3>(a, W, M, N, K, ep, st); break;
      default: launch_umma<16, 3>(a, W, M, N, K, ep, st); break;
    }
  } else {
    switch (bn) {
      case 128: launch_umma<128, 2>(a, W, M, N, K, ep, st); break;
      case 64: launch_umma<64, 2>(a, W, M, N, K, ep, st); break;
      case 32: launch_umma<32, 2>(a, W, M, N, K, ep, st); break;
      default: launch_umma<16, 2>(a, W, M, N, K, ep, st); break;
    }
  }
}

}  // namespace ss
