// tcgen05 implicit-GEMM convolution / linear kernel (warp-specialised, weights pre-packed).
//
//   out[t][n] = epilogue( sum_{tap j, channel ci}  pre(x[t + j*dil - pad_left][ci]) * W[n][j*C_in + ci] )
//
// for stride-1 convolutions over one channels-last sequence (B = 1; a plain linear layer is ksize = 1).  fp32 operands are
// split into NP bf16 pieces (x = x0 + x1 [+ x2]) and the product is accumulated in TMEM as 3 (NP = 2, ~2^-16 relative) or
// 6 (NP = 3, ~fp32) bf16 tcgen05 MMAs.  Structure:
//
//   * weights are split and laid out ONCE (umma2_pack_kernel, cached per weight matrix) as ready-made shared-memory tiles
//     [n-tile][channel chunk][tap][piece][BN x CK bf16, canonical no-swizzle K-major core matrices]; a producer thread
//     streams them with cp.async.bulk (the TMA engine's 1-D bulk copy) into a ring of SB stages, completion on mbarriers;
//   * the activation rows of an output tile are converted ONCE per channel chunk, halo included: rows
//     [m0 - pad_left, m0 + 128 + (k-1)*dil - pad_left) x CK channels.  In the no-swizzle layout a plane of 8 channels is a
//     linear array of rows at a 16-byte pitch, so tap j is the same staged data with the descriptor's start address moved
//     by j*dil rows: a k-tap convolution costs k MMAs per staged chunk instead of k im2col conversions;
//   * roles: warps 0-7 convert (global fp32 -> registers, one chunk ahead -> bf16 pieces in smem), warp 8 issues the MMAs,
//     warp 9 runs the weight ring; all hand-offs are mbarriers, so conversion of chunk c+1, the weight copies and the MMAs
//     of chunk c overlap.  Warps 0-7 read the accumulator back with tcgen05.ld and run the fused epilogue.
//   * split over (chunk, tap) units on gridDim.z when the tile grid alone cannot fill the GPU; partial sums are reduced
//     in a fixed order by splitk_epilogue_kernel (kernels_gemm.cu).
#include <cuda_bf16.h>

#include <algorithm>
#include <map>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace ss {

struct Umma2Cache {
  std::map<std::tuple<const float*, int, int, int, int, int, int>, unsigned char*> packed;
  std::vector<void*> allocs;
};

namespace {

constexpr int U2_BM = 128;
constexpr int U2_CONVERTERS = 256;  // warps 0-7
constexpr int U2_THREADS = 320;     // + MMA warp + weight-ring warp
constexpr int U2_MAX_UNITS = 6;     // float4 units per converter thread and chunk: rows * CK/4 <= 6 * 256
constexpr int U2_MAX_SB = 4;
constexpr int U2_MAX_HALO = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float u2_act(float x, int act) {
  switch (act) {
    case ACT_RELU: return x > 0.f ? x : 0.f;
    case ACT_SILU: return x / (1.0f + expf(-x));
    case ACT_TANH: return tanhf(x);
    default: return x;
  }
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, version 1)
__device__ __forceinline__ uint64_t u2_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  // next 8-column plane along K
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;  // next 8-row group along M / N
  d |= (uint64_t)1 << 46;
  return d;
}

// kind::f16 instruction descriptor: D = F32, A = B = BF16, K-major, M = 128, N = n
__device__ __forceinline__ uint32_t u2_idesc(int n) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(U2_BM >> 4) << 24;
  return d;
}

__device__ __forceinline__ void u2_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void u2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void u2_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

__device__ __forceinline__ void u2_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

struct U2Params {
  ConvA a;
  Epilogue ep;
  const unsigned char* wp;  // packed weight tiles
  int M, N;
  int CK;                   // channels per chunk (16 or 32)
  int ksize, dil;
  int rs;                   // staged rows per chunk = 128 + (ksize - 1) * dil
  int rs_pad;               // row capacity of a staged 8-channel plane (plane = rs_pad * 16 bytes)
  int units_total;          // (C_in / CK) * ksize
  int units_per_split;
  int SB;                   // weight ring stages
  float* ws;                // split partial sums [gridDim.z][M][N] or null
  unsigned* tile_ctr;       // ticket per output tile: the last split CTA of a tile reduces and runs the epilogue (null: separate kernel)
  unsigned long long* dbg;  // optional: %globaltimer stamps of CTA (0,0,0) at the hand-off points (tools/umma2_check.py)
};

__device__ __forceinline__ void u2_stamp(const U2Params& p, int slot) {
  if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.dbg[slot] = t;
  }
}

template <int NP>
__device__ __forceinline__ void u2_split_store(float4 v, unsigned char* stage, uint32_t piece_bytes, uint32_t off) {
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 h23 = __floats2bfloat162_rn(v.z, v.w);
    uint2 packed;
    packed.x = *reinterpret_cast<uint32_t*>(&h01);
    packed.y = *reinterpret_cast<uint32_t*>(&h23);
    *reinterpret_cast<uint2*>(stage + p * piece_bytes + off) = packed;
    if (p + 1 < NP) {
      v.x -= __low2float(h01);
      v.y -= __high2float(h01);
      v.z -= __low2float(h23);
      v.w -= __high2float(h23);
    }
  }
}

// Split reduction inside the kernel: the CTA that takes the last ticket of an output tile sums the gridDim.z partial tiles in slice
// order (deterministic, same order and arithmetic as splitk_epilogue_kernel) and applies the epilogue.
__device__ __forceinline__ void u2_reduce_tile(const U2Params& p, int m0, int n0, int bn) {
  const Epilogue& ep = p.ep;
  const int M = p.M, N = p.N, splits = gridDim.z;
  const int rows = min(U2_BM, M - m0);
  const int cols = min(bn, N - n0);
  const int ctile = ep.glu ? cols / 2 : cols;
  for (int idx = threadIdx.x; idx < rows * ctile; idx += U2_THREADS) {
    const int r = idx / ctile, c = idx - r * ctile;
    const int m = m0 + r;
    int64_t orow = m;
    if (ep.out_L > 0) orow = (int64_t)m * ep.out_row_stride + ep.out_row_offset;  // B == 1
    float y;
    int oc;
    if (ep.glu) {
      const int n = n0 + 2 * c;
      oc = n >> 1;
      float av_ = 0.f, gv = 0.f;
      for (int z = 0; z < splits; ++z) {
        const float* q = p.ws + ((int64_t)z * M + m) * N + n;
        av_ += __ldcg(q);
        gv += __ldcg(q + 1);
      }
      if (ep.bias) {
        av_ += ep.bias[n];
        gv += ep.bias[n + 1];
      }
      y = ep.alpha * (av_ * (1.0f / (1.0f + expf(-gv))));
    } else {
      oc = n0 + c;
      float acc = 0.f;
      for (int z = 0; z < splits; ++z) acc += __ldcg(p.ws + ((int64_t)z * M + m) * N + oc);
      if (ep.bias) acc += ep.bias[oc];
      y = ep.alpha * u2_act(acc, ep.act);
    }
    float* op = ep.out + orow * ep.ldo + oc;
    if (ep.residual) y += ep.res_scale * ep.residual[orow * ep.ldo + oc];
    if (ep.accumulate) y += *op;
    *op = y;
  }
}

template <int BN, int NP, int UPT, int SETS>
__global__ void __launch_bounds__(U2_THREADS) umma2_kernel(const __grid_constant__ U2Params p) {
  constexpr int TM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[5 + 2 * U2_MAX_SB];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * U2_BM, nt = blockIdx.x, n0 = nt * BN;
  const int CK = p.CK, k = p.ksize, SB = p.SB;
  const uint32_t a_plane = (uint32_t)p.rs_pad * 16u;
  const uint32_t a_piece = (uint32_t)(CK >> 3) * a_plane;
  const uint32_t a_stage = NP * a_piece;
  const uint32_t b_piece = (uint32_t)BN * CK * 2u;
  const uint32_t b_unit = NP * b_piece;
  unsigned char* a_smem = smem + ((128u - (smem_u32(smem) & 127u)) & 127u);  // 128-byte aligned whatever the static layout is
  unsigned char* b_smem = a_smem + ((2 * a_stage + 127u) & ~127u);
  const uint32_t a_full = smem_u32(&bars[0]), a_empty = smem_u32(&bars[2]), b_full = smem_u32(&bars[4]);
  const uint32_t b_empty = smem_u32(&bars[4 + U2_MAX_SB]), acc_full = smem_u32(&bars[4 + 2 * U2_MAX_SB]);

  const int u0 = blockIdx.z * p.units_per_split;
  const int u1 = min(p.units_total, u0 + p.units_per_split);
  const int n_units = u1 - u0;
  const int c_first = u0 / k, c_last = (u1 - 1) / k;
  const int n_local = c_last - c_first + 1;

  if (tid == 0) u2_stamp(p, 0);
  // ---- converter set-up (every warp computes the addresses; only warps 0-7 load)
  const int upr_shift = (CK == 32) ? 3 : 2;  // float4 units per row = CK / 4
  const int upr = 1 << upr_shift;
  const int n_a = p.rs * upr;
  int64_t goff[UPT];
  uint32_t soff[UPT];
#pragma unroll
  for (int i = 0; i < UPT; ++i) {
    const int idx = tid + i * U2_CONVERTERS;
    const int r = idx >> upr_shift, q = idx & (upr - 1);
    const int pos = m0 + r - p.a.pad_left;
    const bool inb = idx < n_a && pos >= 0 && pos < p.a.L_in;
    goff[i] = inb ? (int64_t)pos * p.a.ldx + q * 4 : (int64_t)-1;
    soff[i] = (uint32_t)(q >> 1) * a_plane + (uint32_t)r * 16u + (uint32_t)(q & 1) * 8u;
  }
  const float slope = p.a.pre_lrelu;
  // register ring of SETS chunk loads: while chunk cl is converted, the loads of chunks cl+1 .. cl+SETS-1 are in flight
  // (UPT = 6, SETS = 2 for convolutions with a halo; UPT = 4, SETS = 3 for linears, whose chunks carry few MMAs)
  float4 pf[SETS][UPT];
  auto issue = [&](int c, float4* dst) {
    const float* xc = p.a.x + (int64_t)c * CK;
#pragma unroll
    for (int i = 0; i < UPT; ++i)
      dst[i] = goff[i] >= 0 ? __ldg(reinterpret_cast<const float4*>(xc + goff[i])) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  if (warp < 8) {  // the first chunk loads do not depend on the barriers: they fly while TMEM / mbarriers are set up
#pragma unroll
    for (int d = 0; d < SETS - 1; ++d)
      if (d < n_local) issue(c_first + d, pf[d]);
  }

  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a_full + 8 * i), "r"(U2_CONVERTERS / 32) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a_empty + 8 * i), "r"(1) : "memory");
    }
    for (int i = 0; i < U2_MAX_SB; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b_full + 8 * i), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b_empty + 8 * i), "r"(1) : "memory");
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(acc_full), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  if (tid == 0) u2_stamp(p, 1);

  if (warp < 8) {
    // ---------------- converters: activation rows -> bf16 pieces, one channel chunk per stage
    for (int clb = 0; clb < n_local; clb += SETS) {
#pragma unroll
      for (int d = 0; d < SETS; ++d) {
        const int cl = clb + d;
        if (cl < n_local) {
          const int sa = cl & 1, ua = cl >> 1;
          if (cl + SETS - 1 < n_local) issue(c_first + cl + SETS - 1, pf[(d + SETS - 1) % SETS]);
          if (ua >= 1) u2_wait(a_empty + 8 * sa, (uint32_t)((ua - 1) & 1));  // the MMAs of chunk cl-2 have read this stage
          unsigned char* stage = a_smem + (size_t)sa * a_stage;
#pragma unroll
          for (int i = 0; i < UPT; ++i) {
            if (tid + i * U2_CONVERTERS < n_a) {
              float4 v = pf[d][i];
              if (slope != 1.0f) {
                v.x = v.x > 0.f ? v.x : v.x * slope;
                v.y = v.y > 0.f ? v.y : v.y * slope;
                v.z = v.z > 0.f ? v.z : v.z * slope;
                v.w = v.w > 0.f ? v.w : v.w * slope;
              }
              u2_split_store<NP>(v, stage, a_piece, soff[i]);
            }
          }
          // every lane publishes its own generic-proxy writes to the async proxy, then one lane per warp arrives (256
          // arrivals on one mbarrier serialise; 8 do not)
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) u2_arrive(a_full + 8 * sa);
          if (tid == 0) u2_stamp(p, cl == 0 ? 2 : 6);
        }
      }
    }
  } else if (warp == 8) {
    // ---------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = u2_idesc(BN);
      const uint32_t b_plane = (uint32_t)BN * 16u;
      uint32_t uc = 0, accumulate = 0;
      for (int cl = 0; cl < n_local; ++cl) {
        const int sa = cl & 1, ua = cl >> 1;
        u2_wait(a_full + 8 * sa, (uint32_t)(ua & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (cl == 0) u2_stamp(p, 3);
        const int c = c_first + cl;
        const int j_lo = (cl == 0) ? u0 - c * k : 0;
        const int j_hi = (cl == n_local - 1) ? (u1 - 1) - c * k : k - 1;
        const uint32_t a_base = smem_u32(a_smem + (size_t)sa * a_stage);
        for (int j = j_lo; j <= j_hi; ++j, ++uc) {
          const uint32_t sb = uc % (uint32_t)SB, ub = uc / (uint32_t)SB;
          u2_wait(b_full + 8 * sb, ub & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (uc == 0) u2_stamp(p, 4);
          const uint32_t a_tap = a_base + (uint32_t)(j * p.dil) * 16u;
          const uint32_t b_base = smem_u32(b_smem + (size_t)sb * b_unit);
          for (int ks = 0; ks < (CK >> 4); ++ks) {
            uint64_t ad[NP], bd[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              ad[q] = u2_desc(a_tap + q * a_piece + ks * 2 * a_plane, a_plane, 128);
              bd[q] = u2_desc(b_base + q * b_piece + ks * 2 * b_plane, b_plane, 128);
            }
            u2_mma(tmem_d, ad[0], bd[0], idesc, accumulate);
            accumulate = 1;
            u2_mma(tmem_d, ad[0], bd[1], idesc, 1u);
            u2_mma(tmem_d, ad[1], bd[0], idesc, 1u);
            if (NP == 3) {
              u2_mma(tmem_d, ad[1], bd[1], idesc, 1u);
              u2_mma(tmem_d, ad[0], bd[NP - 1], idesc, 1u);
              u2_mma(tmem_d, ad[NP - 1], bd[0], idesc, 1u);
            }
          }
          u2_commit(b_empty + 8 * sb);  // weight stage free once these MMAs have read it
        }
        u2_commit(a_empty + 8 * sa);
      }
      u2_commit(acc_full);
      u2_stamp(p, 5);
    }
  } else {
    // ---------------- weight ring: one bulk copy per (chunk, tap) unit, all pieces
    if (lane == 0) {
      const unsigned char* src = p.wp + ((size_t)nt * p.units_total + u0) * b_unit;
      for (uint32_t uc = 0; uc < (uint32_t)n_units; ++uc) {
        const uint32_t sb = uc % (uint32_t)SB, ub = uc / (uint32_t)SB;
        if (ub >= 1) u2_wait(b_empty + 8 * sb, (ub - 1) & 1u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b_full + 8 * sb), "r"(b_unit) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(b_smem + (size_t)sb * b_unit)), "l"(src + (size_t)uc * b_unit), "r"(b_unit), "r"(b_full + 8 * sb)
                     : "memory");
      }
    }
  }

  // ---------------- epilogue (warps 0-7): warp w owns TMEM lanes 32*(w&3).. (= 32 output rows); with BN >= 64 the two
  // warpgroups split the columns.  tcgen05.ld hands every lane ONE row (registers = columns); written out like that, each
  // store instruction would touch 32 rows x 4 bytes = 32 sectors (measured: 10.7 us of a 15 us CTA).  So each 32 x CW
  // block goes through a padded shared-memory tile (the operand stages are free once acc_full has completed) and is
  // written row by row with lanes = consecutive columns: full 128-byte segments, bias / residual reads coalesced too.
  if (warp < 8) {
    u2_wait(acc_full, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) u2_stamp(p, 7);
    const Epilogue& ep = p.ep;
    const int M = p.M, N = p.N;
    const int quad = warp & 3;
    constexpr int GROUPS = BN >= 64 ? 2 : 1;
    constexpr int COLS_PER_GROUP = BN / GROUPS;
    constexpr int CW = COLS_PER_GROUP >= 32 ? 32 : 16;  // columns per tcgen05.ld
    constexpr int RPI = 32 / CW;                         // rows written per store instruction
    const int c_begin = (warp >> 2) * COLS_PER_GROUP;
    const bool epi_active = warp < 4 * GROUPS;
    float* tbuf = reinterpret_cast<float*>(a_smem) + warp * (32 * 33);
    const int cc = lane % CW, rsub = lane / CW;
#pragma unroll 1
    for (int c0 = c_begin; epi_active && c0 < c_begin + COLS_PER_GROUP; c0 += CW) {
      uint32_t r[CW];
      const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
      if (CW == 32) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16 % CW]), "=r"(r[17 % CW]),
              "=r"(r[18 % CW]), "=r"(r[19 % CW]), "=r"(r[20 % CW]), "=r"(r[21 % CW]), "=r"(r[22 % CW]), "=r"(r[23 % CW]), "=r"(r[24 % CW]),
              "=r"(r[25 % CW]), "=r"(r[26 % CW]), "=r"(r[27 % CW]), "=r"(r[28 % CW]), "=r"(r[29 % CW]), "=r"(r[30 % CW]), "=r"(r[31 % CW])
            : "r"(taddr)
            : "memory");
      } else {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr)
            : "memory");
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (tid == 0 && c0 == c_begin) u2_stamp(p, 9);
#pragma unroll
      for (int j = 0; j < CW; ++j) tbuf[lane * 33 + j] = __uint_as_float(r[j]);
      __syncwarp();
      if (tid == 0 && c0 == c_begin) u2_stamp(p, 10);
      const int n = n0 + c0 + cc;
      const bool n_ok = n < N;
      const int m_first = m0 + quad * 32 + rsub;  // row of iteration 0; iteration `it` handles row m_first + it * RPI
      // The rows are independent: 8 shared-memory reads are issued back to back, then their 8 stores (one row per
      // iteration was a ~160-cycle dependent chain with two warps per scheduler: 2.7 us per 32 x 32 block).  The mode
      // (split partial sums / GLU / plain) is decided outside the loops.
      if (p.ws != nullptr) {  // split: raw partial sums
        float* wp = p.ws + ((int64_t)blockIdx.z * M + m_first) * N + n;
#pragma unroll 1
        for (int it0 = 0; it0 < CW; it0 += 8) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = tbuf[((it0 + u) * RPI + rsub) * 33 + cc];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (n_ok && m_first + (it0 + u) * RPI < M) wp[(int64_t)(it0 + u) * RPI * N] = v[u];
        }
      } else {
        const float bias_n = (ep.bias != nullptr && n_ok) ? ep.bias[n] : 0.f;
        const int64_t row_step = (ep.out_L > 0 ? (int64_t)ep.out_row_stride : (int64_t)1) * RPI * ep.ldo;  // B == 1
        const int64_t o_first = (ep.out_L > 0 ? (int64_t)m_first * ep.out_row_stride + ep.out_row_offset : (int64_t)m_first) * ep.ldo;
        if (ep.glu) {  // columns are interleaved (a, gate) pairs: the gate sits in the next lane
#pragma unroll 1
          for (int it0 = 0; it0 < CW; it0 += 8) {
            float v[8], res[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = tbuf[((it0 + u) * RPI + rsub) * 33 + cc] + bias_n;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float gate = __shfl_down_sync(0xffffffffu, v[u], 1);
              const bool ok = n_ok && (n & 1) == 0 && m_first + (it0 + u) * RPI < M;
              const int64_t o = o_first + (int64_t)(it0 + u) * row_step + (n >> 1);
              res[u] = 0.f;
              if (ok && ep.residual) res[u] = ep.res_scale * ep.residual[o];
              if (ok && ep.accumulate) res[u] += ep.out[o];
              v[u] = ep.alpha * (v[u] * (1.0f / (1.0f + expf(-gate))));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const bool ok = n_ok && (n & 1) == 0 && m_first + (it0 + u) * RPI < M;
              if (ok) ep.out[o_first + (int64_t)(it0 + u) * row_step + (n >> 1)] = v[u] + res[u];
            }
          }
        } else {
          const int act = ep.act;
#pragma unroll 1
          for (int it0 = 0; it0 < CW; it0 += 8) {
            float v[8], res[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = tbuf[((it0 + u) * RPI + rsub) * 33 + cc] + bias_n;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const bool ok = n_ok && m_first + (it0 + u) * RPI < M;
              const int64_t o = o_first + (int64_t)(it0 + u) * row_step + n;
              res[u] = 0.f;
              if (ok && ep.residual) res[u] = ep.res_scale * ep.residual[o];
              if (ok && ep.accumulate) res[u] += ep.out[o];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const bool ok = n_ok && m_first + (it0 + u) * RPI < M;
              if (ok) ep.out[o_first + (int64_t)(it0 + u) * row_step + n] = ep.alpha * u2_act(v[u], act) + res[u];
            }
          }
        }
      }
      if (tid == 0 && c0 == c_begin) u2_stamp(p, 11);
      __syncwarp();  // the tile is rewritten by the next column block
    }
  }
  if (tid == 0) u2_stamp(p, 8);
  if (p.tile_ctr != nullptr) __threadfence();  // this CTA's partial sums are visible device-wide before its ticket
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TM_COLS) : "memory");
  }
  if (p.tile_ctr != nullptr) {
    __shared__ unsigned s_last;
    unsigned* ctr = p.tile_ctr + (blockIdx.y * gridDim.x + blockIdx.x);
    if (tid == 0) s_last = (atomicAdd(ctr, 1u) == gridDim.z - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
      __threadfence();
      if (tid == 0) *ctr = 0u;  // at rest for the next launch that uses this slot
      u2_reduce_tile(p, m0, n0, BN);
    }
  }
}

// W[N][ksize*C_in] fp32 (tap-major columns) -> [n-tile][chunk][tap][piece][plane (8 channels)][row (BN)][8 bf16]
template <int NP>
__global__ void umma2_pack_kernel(const float* __restrict__ W, int N, int C_in, int ksize, int BN, int CK, int n_tiles,
                                  unsigned char* __restrict__ out) {
  const int planes = CK >> 3;
  const int n_chunks = C_in / CK;
  const int64_t total = (int64_t)n_tiles * n_chunks * ksize * planes * BN;  // one thread per (tile row, plane) = 8 channels
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int r = (int)(idx % BN);
  int64_t t = idx / BN;
  int pl = (int)(t % planes); t /= planes;
  int j = (int)(t % ksize); t /= ksize;
  int c = (int)(t % n_chunks);
  int nt = (int)(t / n_chunks);
  const int n = nt * BN + r;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = n < N ? W[(int64_t)n * ksize * C_in + (int64_t)j * C_in + c * CK + pl * 8 + e] : 0.f;
  const size_t piece_bytes = (size_t)BN * CK * 2;
  unsigned char* unit = out + (((size_t)nt * n_chunks + c) * ksize + j) * NP * piece_bytes;
#pragma unroll
  for (int pc = 0; pc < NP; ++pc) {
    uint32_t w4[4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[e], v[e + 1]);
      w4[e >> 1] = *reinterpret_cast<uint32_t*>(&h);
      v[e] -= __low2float(h);
      v[e + 1] -= __high2float(h);
    }
    *reinterpret_cast<uint4*>(unit + pc * piece_bytes + (size_t)pl * BN * 16 + (size_t)r * 16) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
  }
}

}  // namespace
int g_umma2_split_below = 60;   // tile grids smaller than this are split over (chunk, tap) units (ss_set_option umma2_split_below;
                                // measured: 60 beats 148 by 0.9 ms per utterance, the reduce launch costs more than the idle SMs)
int g_umma2_min_units = 4;      // ... into slices of at least this many units
int g_umma2_fused_reduce = 0;   // 1: the last-ticket CTA of a tile reduces the split partial sums inside the kernel.  Measured on B200
                                // (round 2, run 5): bit-identical but SLOWER -- one CTA reads splits x tile bytes alone (up to 22 x 64 KB) where
                                // the separate splitk_epilogue_kernel spreads the same reads over the whole GPU: vocoder 10.6 -> 31 ms per
                                // utterance.  Kept as an option (a reduction distributed over the split CTAs needs them co-resident).
unsigned long long* g_umma2_dbg = nullptr;  // device buffer of 16 stamps when the debug option is on
namespace {

int u2_bn(int N) { return N >= 128 ? 128 : N > 32 ? 64 : N > 16 ? 32 : 16; }

template <int BN, int NP, int UPT, int SETS>
void u2_launch_v(const U2Params& p, dim3 grid, size_t smem, cudaStream_t st) {
  if (first_time_on_device((const void*)umma2_kernel<BN, NP, UPT, SETS>))
    cudaFuncSetAttribute(umma2_kernel<BN, NP, UPT, SETS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  prefer_shared_once((const void*)umma2_kernel<BN, NP, UPT, SETS>);
  umma2_kernel<BN, NP, UPT, SETS><<<grid, U2_THREADS, smem, st>>>(p);
}

template <int BN, int NP>
void u2_launch(const U2Params& p, dim3 grid, size_t smem, cudaStream_t st) {
  const int n_a = p.rs * (p.CK / 4);  // float4 units per staged chunk
  if (n_a <= 4 * U2_CONVERTERS) u2_launch_v<BN, NP, 4, 3>(p, grid, smem, st);
  else u2_launch_v<BN, NP, U2_MAX_UNITS, 2>(p, grid, smem, st);
}

}  // namespace

Umma2Cache* umma2_cache_create() { return new Umma2Cache(); }

void umma2_cache_clear(Umma2Cache* c) {
  if (!c) return;
  cudaDeviceSynchronize();
  for (void* p : c->allocs) cudaFree(p);
  c->allocs.clear();
  c->packed.clear();
}

void umma2_cache_destroy(Umma2Cache* c) {
  if (!c) return;
  umma2_cache_clear(c);
  delete c;
}

bool umma2_supported(const ConvA& a, int N, const Epilogue& ep) {
  if (a.B != 1 || a.stride != 1 || a.chunk != 0 || a.lengths != nullptr) return false;
  if (a.t_offset != 0 || a.x_row0 != 0 || a.x_rows != 0) return false;
  if ((a.C_in & 15) != 0 || (a.ldx & 3) != 0 || a.C_in > a.ldx) return false;
  if ((reinterpret_cast<uintptr_t>(a.x) & 15) != 0) return false;
  if ((a.ksize - 1) * a.dil > U2_MAX_HALO || a.ksize < 1 || a.dil < 1) return false;
  if (N < 16 || (ep.glu && (N & 1))) return false;
  if (ep.ln_gamma != nullptr || ep.split_n > 0) return false;
  if (ep.out_L > 0 && a.L_rows <= 0) return false;
  return true;
}

void umma2_conv(Umma2Cache* cache, const ConvA& a, const float* W, int N, const Epilogue& ep, int pieces, cudaStream_t st) {
  ++g_launches;
  const int M = a.L_rows;
  if (M <= 0) return;
  const int NP = pieces >= 3 ? 3 : 2;
  const int BN = u2_bn(N);
  const int CK = (a.C_in % 32 == 0) ? 32 : 16;
  const int n_chunks = a.C_in / CK;
  const int n_tiles = (N + BN - 1) / BN;
  const int k = a.ksize;
  // ---- packed weights (once per weight matrix and tiling)
  auto key = std::make_tuple(W, N, a.C_in, k, BN, NP, CK);
  unsigned char* wp = nullptr;
  auto it = cache->packed.find(key);
  if (it == cache->packed.end()) {
    const size_t bytes = (size_t)n_tiles * n_chunks * k * NP * BN * CK * 2;
    if (cudaMalloc((void**)&wp, bytes) != cudaSuccess) {
      cudaGetLastError();
      --g_launches;
      gemm_conv(a, W, N, ep, st);  // out of memory for the packed copy: exact fp32 path
      return;
    }
    cache->allocs.push_back(wp);
    cache->packed[key] = wp;
    const int64_t total = (int64_t)n_tiles * n_chunks * k * (CK >> 3) * BN;
    const int blocks = (int)((total + 255) / 256);
    if (NP == 3)
      umma2_pack_kernel<3><<<blocks, 256, 0, st>>>(W, N, a.C_in, k, BN, CK, n_tiles, wp);
    else
      umma2_pack_kernel<2><<<blocks, 256, 0, st>>>(W, N, a.C_in, k, BN, CK, n_tiles, wp);
  } else {
    wp = it->second;
  }
  U2Params p;
  p.a = a;
  p.ep = ep;
  p.wp = wp;
  p.M = M;
  p.N = N;
  p.CK = CK;
  p.ksize = k;
  p.dil = a.dil;
  p.rs = U2_BM + (k - 1) * a.dil;
  p.rs_pad = ((p.rs + 7) & ~7) + 4;  // planes 64 bytes out of phase: halves the bank conflicts of the 8-byte piece stores
  p.units_total = n_chunks * k;
  // ---- split over (chunk, tap) units when the tile grid cannot fill the GPU
  const int m_tiles = (M + U2_BM - 1) / U2_BM;
  const long base = (long)m_tiles * n_tiles;
  const int min_units = std::max(1, g_umma2_min_units);
  int splits = 1;
  if (base < g_umma2_split_below && p.units_total >= 2 * min_units) {
    splits = (int)std::min<long>((148 + base - 1) / base, p.units_total / min_units);
    while (splits > 1 && (size_t)splits * M * N * sizeof(float) > (32u << 20)) --splits;
  }
  int ups = (p.units_total + splits - 1) / splits;
  splits = (p.units_total + ups - 1) / ups;
  float* ws = nullptr;
  if (splits > 1) {
    ws = splitk_workspace((size_t)splits * M * N * sizeof(float));
    if (!ws) {
      splits = 1;
      ups = p.units_total;
    }
  }
  p.units_per_split = ups;
  p.ws = ws;
  p.tile_ctr = (splits > 1 && g_umma2_fused_reduce) ? splitk_counters(n_tiles * m_tiles) : nullptr;
  p.dbg = g_umma2_dbg;
  const size_t a_bytes = (size_t)2 * NP * (CK >> 3) * p.rs_pad * 16;
  const size_t b_unit = (size_t)NP * BN * CK * 2;
  p.SB = (((a_bytes + 127) & ~(size_t)127) + 4 * b_unit <= 110 * 1024) ? 4 : 3;
  const size_t smem = std::max<size_t>(((a_bytes + 127) & ~(size_t)127) + p.SB * b_unit, 8 * 32 * 33 * sizeof(float)) + 256;  // >= the epilogue's transpose tiles
  dim3 grid(n_tiles, m_tiles, splits);
  if (NP == 3) {
    switch (BN) {
      case 128: u2_launch<128, 3>(p, grid, smem, st); break;
      case 64: u2_launch<64, 3>(p, grid, smem, st); break;
      case 32: u2_launch<32, 3>(p, grid, smem, st); break;
      default: u2_launch<16, 3>(p, grid, smem, st); break;
    }
  } else {
    switch (BN) {
      case 128: u2_launch<128, 2>(p, grid, smem, st); break;
      case 64: u2_launch<64, 2>(p, grid, smem, st); break;
      case 32: u2_launch<32, 2>(p, grid, smem, st); break;
      default: u2_launch<16, 2>(p, grid, smem, st); break;
    }
  }
  if (splits > 1 && p.tile_ctr == nullptr) splitk_epilogue(ws, splits, M, N, a.L_rows, ep, st);
}

}  // namespace ss
