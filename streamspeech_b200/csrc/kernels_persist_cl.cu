// Cluster version of the streaming encoder step (option persistent_encoder_cluster): the <= 16 active rows are split over 4
// thread-block clusters of 16 CTAs (4 rows per cluster), and inside a cluster the activations NEVER leave the chip: every CTA keeps
// the cluster's residual rows in shared memory, GEMM outputs are exchanged through distributed shared memory (stores into the peers'
// shared memory + barrier.cluster, ~0.3 us) instead of global memory + a 148-CTA grid barrier (~3 us per phase in
// kernels_persist.cu, profiles/r2_persist_phases_*).  Only two steps per layer involve the other clusters -- the K / V rows and the
// conv-module GLU rows of the step are published to the per-layer caches in global memory -- and keep a grid barrier (64 CTAs).
//
//   per layer and cluster (CTA rank c of 16):
//     FFN  : LN (local) -> W1 rows [128c, 128c+128) -> SiLU -> rank-128 update with W2^T rows [128c, ..) -> partial [4][256]
//            -> reduce-scatter over DSMEM (rank d sums the 16 partials of columns [16d, 16d+16) in rank order) -> bias, 0.5, residual
//            -> all-gather of the new residual columns
//     MHA  : LN -> q | k | v columns [16c, 16c+16) -> k, v to the global cache, q to the CTA of its (row, head) -> GRID BARRIER ->
//            rel-pos attention of (row c % 4, head c / 4) -> all-gather -> Wo columns [16c, ..) + residual -> all-gather
//     conv : LN -> PW1 GLU channels [16c, ..) -> GLU rows to the global conv cache -> GRID BARRIER -> depthwise k31 + BN + SiLU on those
//            channels -> all-gather -> PW2 columns + residual -> all-gather
//     FFN, final LayerNorm (local: every CTA holds complete rows)
//   weights: each (layer, rank) owns one contiguous blob in consumption order (21 chunks of <= 32 KB, packed once); thread 0 streams
//   it with cp.async.bulk (TMA) into a 6-slot ring of shared memory, 5 chunks ahead of the compute and independent of every barrier.
//   Each cluster reads all weights (4 x 123 MB through L2, 1 x from HBM).
// fp32 CUDA-core arithmetic as in kernels_persist.cu (different summation order: ~1e-6).
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"
#include "kernels_persist.h"

namespace cg = cooperative_groups;

namespace ss {
namespace {

constexpr int CT = 256;
constexpr int CWP = CT / 32;
constexpr int CS = 16;               // cluster size
constexpr int CR = 4;                // rows per cluster
constexpr int NCL = 4;               // clusters
constexpr int CD = 256;              // model dim
constexpr int CHD = 64;
constexpr int NSLOT = 6;
constexpr int SLOT_FLOATS = 32 * CD; // 32 KB
constexpr int CHUNKS_PER_LAYER = 21;

// rows (of 256 floats) of chunk j of a layer blob: W1 4 x 32 | W2T 4 x 32 | q+k 32 | v 16 | wo 16 | pw1 32 | pw2 16 | W1 4 x 32 | W2T 4 x 32
__host__ __device__ constexpr int chunk_rows(int j) { return (j == 9 || j == 10 || j == 12) ? 16 : 32; }
__host__ __device__ constexpr int chunk_row0(int j) {
  int r = 0;
  for (int i = 0; i < j; ++i) r += chunk_rows(i);
  return r;
}
constexpr int BLOB_ROWS = chunk_row0(CHUNKS_PER_LAYER);  // 624

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
  }
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target) {
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned v, spins = 0;
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while ((int)(v - target) < 0 && ++spins < (1u << 20));
    if ((int)(v - target) < 0) atomicExch(ctr + SS_BAR_ERR_WORD, 1u);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}

struct ClSmem {
  float xs[CR][CD];          // residual rows of the cluster (replicated in every CTA)
  float As[CR][CD];          // staged GEMM input: LayerNorm output, gathered attention rows or gathered depthwise rows
  float hs[CR][128];         // this CTA's 128 hidden units
  float red[CS][CR][16];     // reduce-scatter receive buffer: red[src][row][col of this CTA's 16-column slice]
  float qh[CHD];             // query of this CTA's (row, head), gathered from the 4 column owners
  float S[1024];             // attention scores
  float qa[CHD], qb[CHD];
  float pv[CWP][CHD];
  float redw[CWP];
  float outc[CR][48];        // raw outputs of a column-split GEMM (q | k | v, or 32 GLU inputs, or 16 columns)
  unsigned long long full[NSLOT];
};

// As[r][:] = LN(xs[r][:]) * g + b (warp r; rows are complete in every CTA); ends with a CTA barrier
__device__ __forceinline__ void stage_ln(ClSmem& sm, const float* __restrict__ g, const float* __restrict__ b) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < CR) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = sm.xs[warp][lane + (i << 5)];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    const float mean = warp_sum(s) / (float)CD;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = v[i] - mean;
      q = fmaf(d, d, q);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)CD + 1e-5f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + (i << 5);
      sm.As[warp][c] = (v[i] - mean) * rstd * g[c] + b[c];
    }
  }
  __syncthreads();
}

// out(j, r, value) = sum_k As[r][k] * w[j][k] for the `rows` (<= 32) weight rows of a ring chunk (row-major, 256 floats each).
// Warp w takes weight rows w, w + 8, w + 16, w + 24; the 16 (weight row, activation row) sums of a warp are reduced with a halving
// tree (16 shuffles); even lane 2v ends with the value v = slot * 4 + r.
template <typename F>
__device__ __forceinline__ void chunk_gemm(const ClSmem& sm, const float* wchunk, int rows, F&& out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 x[CR][2];
#pragma unroll
  for (int r = 0; r < CR; ++r) {
    x[r][0] = *reinterpret_cast<const float4*>(&sm.As[r][lane * 4]);
    x[r][1] = *reinterpret_cast<const float4*>(&sm.As[r][128 + lane * 4]);
  }
  float acc[16];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int j = warp + s * CWP;
    float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
    if (j < rows) {
      w0 = *reinterpret_cast<const float4*>(wchunk + j * CD + lane * 4);
      w1 = *reinterpret_cast<const float4*>(wchunk + j * CD + 128 + lane * 4);
    }
#pragma unroll
    for (int r = 0; r < CR; ++r) {
      float a = x[r][0].x * w0.x;
      a = fmaf(x[r][0].y, w0.y, a);
      a = fmaf(x[r][0].z, w0.z, a);
      a = fmaf(x[r][0].w, w0.w, a);
      a = fmaf(x[r][1].x, w1.x, a);
      a = fmaf(x[r][1].y, w1.y, a);
      a = fmaf(x[r][1].z, w1.z, a);
      a = fmaf(x[r][1].w, w1.w, a);
      acc[s * 4 + r] = a;
    }
  }
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
  float w8[8], w4[4], w2[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = b4 ? acc[i] : acc[i + 8];
    const float keep = b4 ? acc[i + 8] : acc[i];
    w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b3 ? w8[i] : w8[i + 4];
    const float keep = b3 ? w8[i + 4] : w8[i];
    w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b2 ? w4[i] : w4[i + 2];
    const float keep = b2 ? w4[i + 2] : w4[i];
    w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = b1 ? w2[0] : w2[1];
  const float keep = b1 ? w2[1] : w2[0];
  float v1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
  if ((lane & 1) == 0) {
    const int v = lane >> 1;  // = slot * 4 + r
    const int j = warp + (v >> 2) * CWP;
    if (j < rows) out(j, v & 3, v1);
  }
}

struct ClParams {
  const PersistLayer* layers;  // LayerNorm parameters, biases, pos tables, depthwise weights (device pointers as in kernels_persist.cu)
  const float* blobs;          // [n_layers][CS][BLOB_ROWS][256]
  int n_layers;
  float* x;                    // [nA][256] active rows (in: encoder.linear output, out: layer-stack output)
  float *kc, *vc, *gc;         // per-layer caches [n_layers][Tpos][256]
  int nA, a0, T, Tpos, chunk, conv_chunk, dw_k;
  unsigned* bar_ctr;
  unsigned bar_target;
  unsigned long long* ts;      // profiling (option persistent_profile): ts[0] = number of stamps, then (id, ns) pairs of CTA 0, layer 1
};

__global__ void __launch_bounds__(CT, 1) encoder_layers_cluster_kernel(ClParams P) {
  extern __shared__ __align__(128) unsigned char dyn[];
  ClSmem& sm = *reinterpret_cast<ClSmem*>(dyn);
  float* ring = reinterpret_cast<float*>(dyn + ((sizeof(ClSmem) + 127) & ~(size_t)127));
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = (int)cluster.block_rank();           // rank in the cluster
  const int g = blockIdx.x / CS;                     // cluster index = row group
  const int r_lo = g * CR;
  const int nr = max(0, min(CR, P.nA - r_lo));       // valid rows of this cluster
  unsigned bar_target = P.bar_target;
  const int total_chunks = P.n_layers * CHUNKS_PER_LAYER;

  // ---- weight ring: chunk qi of this rank's stream goes to slot qi % NSLOT (thread 0 issues)
  auto issue = [&](int qi) {
    if (qi >= total_chunks) return;
    const int slot = qi % NSLOT;
    const int li = qi / CHUNKS_PER_LAYER, j = qi - li * CHUNKS_PER_LAYER;
    const uint32_t bytes = (uint32_t)chunk_rows(j) * CD * 4u;
    const float* src = P.blobs + ((size_t)(li * CS + c) * BLOB_ROWS + chunk_row0(j)) * CD;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the slot are ordered before the async write
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm.full[slot])), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + (size_t)slot * SLOT_FLOATS)),
                 "l"(src), "r"(bytes), "r"(smem_u32(&sm.full[slot]))
                 : "memory");
  };
  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < NSLOT; ++i) issue(i);
  }
  // residual rows of the cluster (zero rows where the cluster has fewer than 4)
  for (int i = tid; i < CR * CD; i += CT) {
    const int r = i / CD;
    sm.xs[r][i - r * CD] = r < nr ? P.x[(int64_t)(r_lo + r) * CD + (i - r * CD)] : 0.f;
  }
  __syncthreads();
  cluster.sync();  // every CTA of the cluster runs and has initialised its shared memory before any DSMEM traffic

  int q = 0;  // chunk counter
  auto acquire = [&]() -> const float* {
    mbar_wait(smem_u32(&sm.full[q % NSLOT]), (uint32_t)((q / NSLOT) & 1));
    return ring + (size_t)(q % NSLOT) * SLOT_FLOATS;
  };
  auto release = [&]() {  // every thread has finished reading the slot -> refill it with the chunk NSLOT ahead
    __syncthreads();
    if (tid == 0) issue(q + NSLOT);
    ++q;
  };
  int nts = 0;
  bool stamping = false;
  auto stamp = [&](int id) {
    if (stamping && tid == 0) {
      P.ts[1 + 2 * nts] = (unsigned long long)id;
      P.ts[2 + 2 * nts] = globaltimer_ns();
      ++nts;
    }
  };
  auto peer = [&](float* p, int rank) -> float* { return cluster.map_shared_rank(p, rank); };

  // ---- FFN block: xs += 0.5 * (W2 silu(W1 LN(xs) + b1) + b2); chunks [base, base + 8) of the layer
  auto ffn = [&](const float* __restrict__ g_, const float* __restrict__ b_, const float* __restrict__ b1, const float* __restrict__ b2) {
    stage_ln(sm, g_, b_);
    for (int part = 0; part < 4; ++part) {
      const float* w = acquire();
      chunk_gemm(sm, w, 32, [&](int j, int r, float v) {
        const float y = v + b1[c * 128 + part * 32 + j];
        sm.hs[r][part * 32 + j] = y / (1.0f + expf(-y));
      });
      release();  // (its CTA barrier also publishes hs)
    }
    stamp(1);
    float acc[CR] = {0.f, 0.f, 0.f, 0.f};  // thread n = tid: output column n of the rank-128 update
    for (int part = 0; part < 4; ++part) {
      const float* w = acquire();
#pragma unroll 8
      for (int u = 0; u < 32; ++u) {
        const float wv = w[u * CD + tid];
#pragma unroll
        for (int r = 0; r < CR; ++r) acc[r] = fmaf(sm.hs[r][part * 32 + u], wv, acc[r]);
      }
      release();
    }
    stamp(2);
    // reduce-scatter: column n belongs to rank n / 16
#pragma unroll
    for (int r = 0; r < CR; ++r) peer(&sm.red[c][r][tid & 15], tid >> 4)[0] = acc[r];
    cluster.sync();
    stamp(3);
    if (tid < CR * 16) {
      const int r = tid >> 4, j = tid & 15;
      float t = sm.red[0][r][j];
#pragma unroll
      for (int s = 1; s < CS; ++s) t += sm.red[s][r][j];
      const float y = sm.xs[r][c * 16 + j] + 0.5f * (t + b2[c * 16 + j]);
#pragma unroll
      for (int d = 0; d < CS; ++d) peer(&sm.xs[r][c * 16 + j], d)[0] = y;
    }
    cluster.sync();
    stamp(4);
  };
  // xs[:, 16c..16c+16) += outc[:, 0..16) + bias, all-gathered (column-split GEMM epilogue)
  auto residual_gather = [&](const float* __restrict__ bias) {
    __syncthreads();
    if (tid < CR * 16) {
      const int r = tid >> 4, j = tid & 15;
      const float y = sm.xs[r][c * 16 + j] + (sm.outc[r][j] + (bias ? bias[c * 16 + j] : 0.f));
#pragma unroll
      for (int d = 0; d < CS; ++d) peer(&sm.xs[r][c * 16 + j], d)[0] = y;
    }
    cluster.sync();
  };

  for (int li = 0; li < P.n_layers; ++li) {
    const PersistLayer L = P.layers[li];
    stamping = P.ts != nullptr && blockIdx.x == 0 && li == 1;
    stamp(0);
    float* kc = P.kc + (size_t)li * P.Tpos * CD;
    float* vc = P.vc + (size_t)li * P.Tpos * CD;
    float* gc = P.gc + (size_t)li * P.Tpos * CD;
    ffn(L.ffn1_g, L.ffn1_b, L.ffn1_b1, L.ffn1_b2);
    // ================= attention block =================
    stage_ln(sm, L.attn_g, L.attn_b);
    {
      const float* w = acquire();  // q rows 0..15, k rows 16..31
      chunk_gemm(sm, w, 32, [&](int j, int r, float v) { sm.outc[r][j] = v + L.bqkv[(j >> 4) * CD + c * 16 + (j & 15)]; });
      release();
      w = acquire();               // v rows
      chunk_gemm(sm, w, 16, [&](int j, int r, float v) { sm.outc[r][32 + j] = v + L.bqkv[2 * CD + c * 16 + j]; });
      release();
    }
    if (tid < CR * 16) {
      const int r = tid >> 4, j = tid & 15;
      if (r < nr) {
        kc[(int64_t)(P.a0 + r_lo + r) * CD + c * 16 + j] = sm.outc[r][16 + j];
        vc[(int64_t)(P.a0 + r_lo + r) * CD + c * 16 + j] = sm.outc[r][32 + j];
      }
      // q of (row r, head c / 4) goes to the CTA that owns that task: rank 4 * (c / 4) + r
      peer(&sm.qh[(c & 3) * 16 + j], (c & ~3) + r)[0] = sm.outc[r][j];
    }
    stamp(10);
    cluster.sync();                          // q gathered
    stamp(11);
    grid_barrier(P.bar_ctr, bar_target);     // K / V rows of all clusters are in the cache
    stamp(12);
    {
      // rel-pos attention of (row c % 4, head c / 4) over keys 0 .. lim-1 (phase_attention of kernels_persist.cu, one task per CTA)
      const int r = c & 3, h = c >> 2;
      if (r < nr) {
        const int i = P.a0 + r_lo + r;
        const int lim = P.chunk > 0 ? min((i / P.chunk + 1) * P.chunk, P.T) : P.T;
        const int n = max(1, lim);
        if (tid < CHD) {
          const float val = sm.qh[tid];
          sm.qa[tid] = val + L.pos_u[h * CHD + tid];
          sm.qb[tid] = val + L.pos_v[h * CHD + tid];
        }
        __syncthreads();
        const float* kb = kc + h * CHD;
        const float* vb = vc + h * CHD;
        const float* pb = L.pos_proj + h * CHD;
        float mx = -INFINITY;
        for (int j = tid; j < n; j += CT) {
          const float* kr = kb + (int64_t)j * CD;
          const float* pr = pb + (int64_t)(i - j + P.Tpos - 1) * CD;
          float4 kk[CHD / 4], pp[CHD / 4];
#pragma unroll
          for (int d = 0; d < CHD / 4; ++d) {
            kk[d] = *reinterpret_cast<const float4*>(kr + 4 * d);
            pp[d] = *reinterpret_cast<const float4*>(pr + 4 * d);
          }
          float ac = 0.f, bd = 0.f;
#pragma unroll
          for (int d = 0; d < CHD / 4; ++d) {
            ac = fmaf(sm.qa[4 * d], kk[d].x, ac); ac = fmaf(sm.qa[4 * d + 1], kk[d].y, ac);
            ac = fmaf(sm.qa[4 * d + 2], kk[d].z, ac); ac = fmaf(sm.qa[4 * d + 3], kk[d].w, ac);
            bd = fmaf(sm.qb[4 * d], pp[d].x, bd); bd = fmaf(sm.qb[4 * d + 1], pp[d].y, bd);
            bd = fmaf(sm.qb[4 * d + 2], pp[d].z, bd); bd = fmaf(sm.qb[4 * d + 3], pp[d].w, bd);
          }
          const float s = (ac + bd) * 0.125f;
          sm.S[j] = s;
          mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        if (lane == 0) sm.redw[warp] = mx;
        __syncthreads();
        mx = sm.redw[0];
#pragma unroll
        for (int x = 1; x < CWP; ++x) mx = fmaxf(mx, sm.redw[x]);
        __syncthreads();
        float sum = 0.f;
        for (int j = tid; j < n; j += CT) {
          const float e = expf(sm.S[j] - mx);
          sm.S[j] = e;
          sum += e;
        }
        sum = warp_sum(sum);
        if (lane == 0) sm.redw[warp] = sum;
        __syncthreads();
        sum = sm.redw[0];
#pragma unroll
        for (int x = 1; x < CWP; ++x) sum += sm.redw[x];
        float a0_ = 0.f, a1_ = 0.f;
        for (int j0 = warp; j0 < n; j0 += CWP * 16) {
          float2 vv[16];
          float p[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int j = j0 + u * CWP;
            const bool ok = j < n;
            vv[u] = ok ? *reinterpret_cast<const float2*>(vb + (int64_t)j * CD + 2 * lane) : make_float2(0.f, 0.f);
            p[u] = ok ? sm.S[j] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            a0_ = fmaf(p[u], vv[u].x, a0_);
            a1_ = fmaf(p[u], vv[u].y, a1_);
          }
        }
        sm.pv[warp][2 * lane] = a0_;
        sm.pv[warp][2 * lane + 1] = a1_;
        __syncthreads();
        if (tid < CHD) {
          float t = 0.f;
#pragma unroll
          for (int x = 0; x < CWP; ++x) t += sm.pv[x][tid];
          t /= sum;
#pragma unroll
          for (int d = 0; d < CS; ++d) peer(&sm.As[r][h * CHD + tid], d)[0] = t;  // all-gather of the attention rows
        }
      } else if (tid < CHD) {
#pragma unroll
        for (int d = 0; d < CS; ++d) peer(&sm.As[r][h * CHD + tid], d)[0] = 0.f;
      }
    }
    stamp(13);
    cluster.sync();
    stamp(14);
    {
      const float* w = acquire();  // Wo rows [16c, 16c+16)
      chunk_gemm(sm, w, 16, [&](int j, int r, float v) { sm.outc[r][j] = v; });
      release();
    }
    stamp(15);
    residual_gather(L.bo);
    stamp(16);
    // ================= conv module =================
    stage_ln(sm, L.conv_g, L.conv_b);
    {
      const float* w = acquire();  // PW1: interleaved (value, gate) rows of channels [16c, 16c+16)
      chunk_gemm(sm, w, 32, [&](int j, int r, float v) { sm.outc[r][j] = v + (L.pw1_b ? L.pw1_b[c * 32 + j] : 0.f); });
      release();
    }
    if (tid < CR * 16) {
      const int r = tid >> 4, ch = tid & 15;
      if (r < nr) {
        const float a = sm.outc[r][2 * ch], gate = sm.outc[r][2 * ch + 1];
        gc[(int64_t)(P.a0 + r_lo + r) * CD + c * 16 + ch] = a * (1.0f / (1.0f + expf(-gate)));
      }
    }
    stamp(20);
    grid_barrier(P.bar_ctr, bar_target);  // GLU rows of all clusters are in the conv cache
    stamp(21);
    if (tid < CR * 16) {
      const int r = tid >> 4, ch = tid & 15, oc = c * 16 + ch;
      float y = 0.f;
      if (r < nr) {
        const int t = P.a0 + r_lo + r, half = (P.dw_k - 1) >> 1;
        const int lim = P.conv_chunk > 0 ? min(P.T, (t / P.conv_chunk + 1) * P.conv_chunk) : P.T;
        float a = 0.f;
        for (int j = 0; j < P.dw_k; ++j) {
          const int p = t - half + j;
          if (p >= 0 && p < lim) a = fmaf(L.dw_w[j * CD + oc], gc[(int64_t)p * CD + oc], a);
        }
        const float v = a * L.bn_scale[oc] + L.bn_shift[oc];
        y = v / (1.0f + expf(-v));
      }
#pragma unroll
      for (int d = 0; d < CS; ++d) peer(&sm.As[r][oc], d)[0] = y;  // all-gather of the depthwise rows
    }
    stamp(22);
    cluster.sync();
    stamp(23);
    {
      const float* w = acquire();  // PW2 rows [16c, 16c+16)
      chunk_gemm(sm, w, 16, [&](int j, int r, float v) { sm.outc[r][j] = v; });
      release();
    }
    stamp(24);
    residual_gather(L.pw2_b);
    stamp(25);
    ffn(L.ffn2_g, L.ffn2_b, L.ffn2_b1, L.ffn2_b2);
    // final LayerNorm of the layer, in place (every CTA holds the complete rows: no exchange)
    if (warp < CR) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = sm.xs[warp][lane + (i << 5)];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[i];
      const float mean = warp_sum(s) / (float)CD;
      float qv = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[i] - mean;
        qv = fmaf(d, d, qv);
      }
      const float rstd = 1.0f / sqrtf(warp_sum(qv) / (float)CD + 1e-5f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int cc = lane + (i << 5);
        sm.xs[warp][cc] = (v[i] - mean) * rstd * L.fin_g[cc] + L.fin_b[cc];
      }
    }
    __syncthreads();
    stamp(30);
    if (stamping && tid == 0) P.ts[0] = (unsigned long long)nts;
  }
  if (c == 0) {
    for (int i = tid; i < nr * CD; i += CT) P.x[(int64_t)r_lo * CD + i] = sm.xs[i / CD][i % CD];
  }
  cluster.sync();  // no CTA exits while a peer may still address its shared memory
}

// blob row `row` of (layer li, rank c): see chunk table above.  One thread per float4.
__global__ void cluster_pack_kernel(const PersistLayer* __restrict__ layers, int n_layers, int FFN, float* __restrict__ blobs) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // float4 index
  const int64_t total = (int64_t)n_layers * CS * BLOB_ROWS * (CD / 4);
  if (idx >= total) return;
  const int k4 = (int)(idx % (CD / 4));
  int64_t t = idx / (CD / 4);
  const int row = (int)(t % BLOB_ROWS);
  t /= BLOB_ROWS;
  const int c = (int)(t % CS), li = (int)(t / CS);
  const PersistLayer& L = layers[li];
  const float* src;
  if (row < 128) src = L.ffn1_w1 + (int64_t)(c * 128 + row) * CD;
  else if (row < 256) src = L.ffn1_w2t + (int64_t)(c * 128 + row - 128) * CD;
  else if (row < 272) src = L.wqkv + (int64_t)(c * 16 + row - 256) * CD;
  else if (row < 288) src = L.wqkv + (int64_t)(CD + c * 16 + row - 272) * CD;
  else if (row < 304) src = L.wqkv + (int64_t)(2 * CD + c * 16 + row - 288) * CD;
  else if (row < 320) src = L.wo + (int64_t)(c * 16 + row - 304) * CD;
  else if (row < 352) src = L.pw1 + (int64_t)(c * 32 + row - 320) * CD;
  else if (row < 368) src = L.pw2 + (int64_t)(c * 16 + row - 352) * CD;
  else if (row < 496) src = L.ffn2_w1 + (int64_t)(c * 128 + row - 368) * CD;
  else src = L.ffn2_w2t + (int64_t)(c * 128 + row - 496) * CD;
  (void)FFN;
  reinterpret_cast<float4*>(blobs)[idx] = reinterpret_cast<const float4*>(src)[k4];
}

}  // namespace

size_t encoder_layers_cluster_blob_floats(int n_layers) { return (size_t)n_layers * CS * BLOB_ROWS * CD; }

void encoder_layers_cluster_pack(const PersistLayer* layers_dev, int n_layers, int FFN, float* blobs_dev, cudaStream_t st) {
  const int64_t total = (int64_t)n_layers * CS * BLOB_ROWS * (CD / 4);
  cluster_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(layers_dev, n_layers, FFN, blobs_dev);
}

bool encoder_layers_cluster_supported(int nA, int D, int FFN, int H, int T, int dw_k) {
  return nA >= 1 && nA <= NCL * CR && D == CD && FFN == 2048 && H * CHD == D && T <= 1024 && (dw_k & 1) == 1 && dw_k <= 31;
}

int encoder_layers_cluster(const PersistLayer* layers_dev, const float* blobs_dev, int n_layers, float* x, float* kc, float* vc, float* gc, int nA,
                           int a0, int T, int Tpos, int chunk, int conv_chunk, int dw_k, unsigned* bar_ctr, unsigned* bar_target_host,
                           unsigned long long* ts_or_null, cudaStream_t st) {
  ++g_launches;
  const size_t smem = ((sizeof(ClSmem) + 127) & ~(size_t)127) + (size_t)NSLOT * SLOT_FLOATS * sizeof(float);
  if (first_time_on_device((const void*)encoder_layers_cluster_kernel)) {
    if (cudaFuncSetAttribute(encoder_layers_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(encoder_layers_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return -1;
  }
  ClParams P;
  P.layers = layers_dev; P.blobs = blobs_dev; P.n_layers = n_layers; P.x = x; P.kc = kc; P.vc = vc; P.gc = gc;
  P.nA = nA; P.a0 = a0; P.T = T; P.Tpos = Tpos; P.chunk = chunk; P.conv_chunk = conv_chunk; P.dw_k = dw_k;
  P.bar_ctr = bar_ctr; P.bar_target = *bar_target_host; P.ts = ts_or_null;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(NCL * CS);
  cfg.blockDim = dim3(CT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;  // all 4 clusters co-resident (they meet in a grid barrier)
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, encoder_layers_cluster_kernel, &cfg) != cudaSuccess || max_clusters < NCL) {
    cudaGetLastError();
    return -1;  // the 64 CTAs could not be co-resident: their grid barrier would time out
  }
  if (cudaLaunchKernelEx(&cfg, encoder_layers_cluster_kernel, P) != cudaSuccess) return -2;
  *bar_target_host += (unsigned)(NCL * CS) * (unsigned)(2 * n_layers);
  return 0;
}

}  // namespace ss
