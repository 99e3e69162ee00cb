// (b2) TAGConv(32 -> cout, K hops) [+ dropout + ReLU] [+ residual], forward and recompute-based backward.
// Replaces PyG TAGConv.forward + gcn_norm (recomputed per layer in the reference) and the inline
// Dropout/ReLU of networks.py:268-269, 330-331, plus what autograd derives for them.
//
// Design (small-graph / tile path; a tile = a few WHOLE graphs, so hops never leave the CTA):
//   * persistent CTAs loop over tiles; the tile's node features sit in shared memory as [rows][32] fp32,
//     one warp lane per hidden feature, so hop gathers are conflict-free and HBM rows are 128-byte lines;
//   * A_hat^k x is built hop by hop in shared memory walking the CSR-by-destination in PyG scatter
//     order: segmented sums without atomics;
//   * the (K+1) 32x32 transforms run on CUDA cores with the weight operand resident in registers
//     (96 registers per lane for K=2) and the activation operand delivered by warp-broadcast LDS.128,
//     i.e. 1 shared-memory wavefront per 4 FFMA warp-instructions;
//   * forward epilogue fuses bias, dropout (counter-based Philox or an injected mask), ReLU, residual and
//     emits one 32-bit word per node with the sign pattern of the output - all the backward needs;
//   * backward recomputes A_hat x, A_hat^2 x from the saved layer input and is warp-specialised: half of
//     the warps produce grad_x by Horner's rule (A_hat is symmetric on the doubled graph),
//     the other half accumulate grad_W / grad_b in registers across all tiles of the CTA; per-CTA
//     partial sums are written once and reduced in fixed order by dss2_reduce_partials (deterministic).
#include "common.cuh"

namespace {

constexpr int FWD_THREADS = 256;
constexpr int FWD_WARPS = FWD_THREADS / 32;
constexpr int BWD_THREADS = 512;
constexpr int BWD_ROLE_WARPS = BWD_THREADS / 64;   // warps per role
constexpr int MAXK = 3;

struct TagFwdArgs {
  dss2_graph_t g;
  const float* x;
  const float* w;
  const float* bias;
  int cout;
  int act;
  float scale;          // 1/(1-p)
  uint32_t keep_thr;    // keep iff rnd < keep_thr
  int drop_mode;        // 0 none, 1 philox, 2 mask
  const uint64_t* rng;
  uint32_t layer_uid;
  const uint8_t* mask;
  const float* res;
  int64_t res_stride;
  float* y;
  uint32_t* bits;
};

struct TagBwdArgs {
  dss2_graph_t g;
  const float* x;
  const float* w;
  int cout;
  int act;
  float scale;
  const uint32_t* bits;
  const float* gy;
  float* gx;
  float* partials;
  int64_t partial_stride;
  int64_t bias_offset;
};

__host__ __device__ inline int round4(int v) { return (v + 3) & ~3; }

// Tile topology in shared memory: local rowptr and one packed (source row offset, gcn weight) pair per CSR entry.
struct TileTopo {
  int* rowptr;
  int2* cw;     // .x = local source node * 32 (float offset of its feature row), .y = weight bits
};

__device__ __forceinline__ void load_topo(const dss2_graph_t& g, const TileRange& r, TileTopo s, int tid, int nthreads) {
  const int nT = r.n1 - r.n0, nZ = r.z1 - r.z0;
  for (int i = tid; i <= nT; i += nthreads) s.rowptr[i] = g.rowptr[r.n0 + i] - r.z0;
  for (int i = tid; i < nZ; i += nthreads) s.cw[i] = make_int2((g.col[r.z0 + i] - r.n0) * HID, __float_as_int(g.w[r.z0 + i]));
}

// acc[i] += sum_{e in row r0+i} w_e * src[col_e][lane] for R consecutive rows handled by one warp.
// Degrees are tiny (<= 3 on the feeders): the first 4 entries of each row are fully unrolled and predicated on the
// warp-uniform degree, so the R rows' loads interleave and no divergence bookkeeping is emitted; longer rows fall through
// to a plain loop.  Entries are visited in CSR order = PyG scatter order.
template <int R>
__device__ __forceinline__ void hop_rows(const TileTopo& s, const float* __restrict__ src, int r0, int nT, int lane, float (&acc)[R]) {
  int beg[R], deg[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const bool valid = r0 + i < nT;
    beg[i] = valid ? s.rowptr[r0 + i] : 0;
    deg[i] = valid ? s.rowptr[r0 + i + 1] - beg[i] : 0;
  }
#pragma unroll
  for (int d = 0; d < 4; ++d) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
      if (d < deg[i]) {
        const int2 cw = s.cw[beg[i] + d];
        acc[i] = fmaf(__int_as_float(cw.y), src[cw.x + lane], acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    for (int z = beg[i] + 4; z < beg[i] + deg[i]; ++z) {
      const int2 cw = s.cw[z];
      acc[i] = fmaf(__int_as_float(cw.y), src[cw.x + lane], acc[i]);
    }
  }
}

// contiguous rows global -> shared, 16 bytes per thread per step
__device__ __forceinline__ void load_rows32(float* dst, const float* __restrict__ src, int nrows, int tid, int nthreads) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = tid; i < nrows * (HID / 4); i += nthreads) d4[i] = ldg_stream4(s4 + i);
}

// -------------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(FWD_THREADS, 2) k_tag_fwd(TagFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const dss2_graph_t& g = a.g;
  const int TR = round4(g.max_tile_nodes);
  float* X = smem;                                   // [K+1][TR][32]
  TileTopo topo;
  topo.cw = reinterpret_cast<int2*>(X + (K + 1) * TR * HID);
  topo.rowptr = reinterpret_cast<int*>(topo.cw + g.max_tile_nnz + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cout = a.cout;

  // weight operand: lane c keeps row c of every W_k (96 registers for K = 2)
  float W[K + 1][HID];
#pragma unroll
  for (int k = 0; k <= K; ++k) {
#pragma unroll
    for (int j4 = 0; j4 < HID / 4; ++j4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane < cout) v = *reinterpret_cast<const float4*>(a.w + ((size_t)k * cout + lane) * HID + 4 * j4);
      W[k][4 * j4 + 0] = v.x;
      W[k][4 * j4 + 1] = v.y;
      W[k][4 * j4 + 2] = v.z;
      W[k][4 * j4 + 3] = v.w;
    }
  }
  const float bias = lane < cout ? a.bias[lane] : 0.0f;
  uint2 key = make_uint2(0u, 0u);
  uint32_t step_lo = 0;
  if (a.drop_mode == 1) {
    uint64_t seed = a.rng[0], step = a.rng[1];
    key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (a.layer_uid * 0x9E3779B9u) ^ (uint32_t)(step >> 32));
    step_lo = (uint32_t)step;
  }

  for (int t = blockIdx.x; t < g.num_tiles; t += gridDim.x) {
    const TileRange r = tile_range(g, t);
    const int nT = r.n1 - r.n0;
    load_rows32(X, a.x + (size_t)r.n0 * HID, nT, tid, FWD_THREADS);
    load_topo(g, r, topo, tid, FWD_THREADS);
    __syncthreads();
    // hops 1..K-1 need every row of the previous hop -> block barrier; the last hop is fused below
    const int nblk = (nT + 3) >> 2;
#pragma unroll
    for (int k = 1; k < K; ++k) {
      for (int blk = warp; blk < nblk; blk += FWD_WARPS) {
        float h[4] = {0.f, 0.f, 0.f, 0.f};
        hop_rows<4>(topo, X + (k - 1) * TR * HID, blk * 4, nT, lane, h);
#pragma unroll
        for (int i = 0; i < 4; ++i) X[(k * TR + blk * 4 + i) * HID + lane] = h[i];
      }
      __syncthreads();
    }
    for (int blk = warp; blk < nblk; blk += FWD_WARPS) {
      const int r0 = blk * 4;
      {
        float h[4] = {0.f, 0.f, 0.f, 0.f};
        hop_rows<4>(topo, X + (K - 1) * TR * HID, r0, nT, lane, h);
#pragma unroll
        for (int i = 0; i < 4; ++i) X[(K * TR + r0 + i) * HID + lane] = h[i];
        __syncwarp();
      }
      // 4 rows x (K+1)*32 inputs: the 4 accumulators are independent, FFMAs are issued row-interleaved
      float acc[4] = {bias, bias, bias, bias};
#pragma unroll
      for (int k = 0; k <= K; ++k) {
        const float* Xk = X + (k * TR + r0) * HID;
#pragma unroll
        for (int j4 = 0; j4 < HID / 4; ++j4) {
          float4 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(Xk + i * HID + 4 * j4);   // warp-broadcast
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fmaf(v[i].x, W[k][4 * j4 + 0], acc[i]);
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fmaf(v[i].y, W[k][4 * j4 + 1], acc[i]);
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fmaf(v[i].z, W[k][4 * j4 + 2], acc[i]);
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fmaf(v[i].w, W[k][4 * j4 + 3], acc[i]);
        }
      }
      // epilogue: dropout -> ReLU -> sign word -> residual -> store (a row is one 128-byte line)
      uint32_t rnd[4] = {0u, 0u, 0u, 0u};
      if (a.act && a.drop_mode == 1) {
        uint4 q = philox4x32_10(make_uint4((uint32_t)t, (uint32_t)blk, (uint32_t)lane, step_lo), key);
        rnd[0] = q.x;
        rnd[1] = q.y;
        rnd[2] = q.z;
        rnd[3] = q.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = r0 + i;
        if (row >= nT) break;   // warp-uniform
        const size_t n = (size_t)r.n0 + row;
        float v = acc[i];
        if (a.act) {
          bool keep = true;
          if (a.drop_mode == 1) keep = rnd[i] < a.keep_thr;
          else if (a.drop_mode == 2) keep = a.mask[n * HID + lane] != 0;
          if (a.drop_mode != 0) v = keep ? v * a.scale : 0.0f;
          v = fmaxf(v, 0.0f);
          const uint32_t word = __ballot_sync(0xffffffffu, v > 0.0f);
          if (lane == 0 && a.bits) a.bits[n] = word;
        }
        if (lane < cout) {
          if (a.res) v += a.res[n * a.res_stride + lane];
          a.y[n * cout + lane] = v;
        }
      }
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------------------
// backward
// -------------------------------------------------------------------------------------------------
template <int K, int CP>   // CP: cout padded to {2, 8, 32}; padded columns carry zeros
__global__ void __launch_bounds__(BWD_THREADS, 1) k_tag_bwd(TagBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const dss2_graph_t& g = a.g;
  const int TR = round4(g.max_tile_nodes);
  float* X = smem;                          // [K+1][TR][32]  x, A x, A^2 x   (W role recomputes hops)
  float* G = X + (K + 1) * TR * HID;        // [TR][32]       grad wrt pre-activation output, zero padded
  float* H = G + TR * HID;                  // [2][TR][32]    Horner ping-pong (X role)
  TileTopo topo;
  topo.cw = reinterpret_cast<int2*>(H + 2 * TR * HID);
  topo.rowptr = reinterpret_cast<int*>(topo.cw + g.max_tile_nnz + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool x_role = warp < BWD_ROLE_WARPS;
  const int rw = x_role ? warp : warp - BWD_ROLE_WARPS;   // warp index inside the role
  const int cout = a.cout;

  // X role: lane j keeps column j of every W_k (W_k[c][j], c < CP).  W role: lane j accumulates
  // grad_W_k[c][j] for all c, plus grad_b[c = lane].
  float R[K + 1][CP];
#pragma unroll
  for (int k = 0; k <= K; ++k)
#pragma unroll
    for (int c = 0; c < CP; ++c) R[k][c] = (x_role && c < cout) ? a.w[((size_t)k * cout + c) * HID + lane] : 0.0f;
  float gb = 0.0f;

  for (int t = blockIdx.x; t < g.num_tiles; t += gridDim.x) {
    const TileRange r = tile_range(g, t);
    const int nT = r.n1 - r.n0;
    load_rows32(X, a.x + (size_t)r.n0 * HID, nT, tid, BWD_THREADS);
    load_topo(g, r, topo, tid, BWD_THREADS);
    // grad tile: g_out = grad_y * [y > 0] / (1-p)   (ReLU'(0) = 0; dropped units have y = 0)
    for (int i = tid; i < nT * HID; i += BWD_THREADS) {
      const int row = i >> 5, c = i & 31;
      float v = 0.0f;
      if (c < cout) {
        v = a.gy[((size_t)r.n0 + row) * cout + c];
        if (a.act) v = ((a.bits[(size_t)r.n0 + row] >> c) & 1u) ? v * a.scale : 0.0f;
      }
      G[i] = v;
    }
    __syncthreads();

    const int nblk = (nT + 3) >> 2;
    if (x_role) {
      // grad_x = G W_0 + A (G W_1 + A (G W_2)): Horner over hops, all rows of a stage before the next; 4 rows per warp pass
#pragma unroll
      for (int k = K; k >= 0; --k) {
        float* out = H + ((K - k) & 1) * TR * HID;
        const float* prev = H + ((K - k + 1) & 1) * TR * HID;
        for (int blk = rw; blk < nblk; blk += BWD_ROLE_WARPS) {
          const int r0 = blk * 4;
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          if (k < K) hop_rows<4>(topo, prev, r0, nT, lane, acc);
          const float* grow = G + r0 * HID;
          if (CP >= 4) {
#pragma unroll
            for (int c4 = 0; c4 < CP / 4; ++c4) {
              float4 v[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(grow + i * HID + 4 * c4);
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] = fmaf(v[i].x, R[k][4 * c4 + 0], acc[i]);
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] = fmaf(v[i].y, R[k][4 * c4 + 1], acc[i]);
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] = fmaf(v[i].z, R[k][4 * c4 + 2], acc[i]);
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] = fmaf(v[i].w, R[k][4 * c4 + 3], acc[i]);
            }
          } else {
#pragma unroll
            for (int c = 0; c < CP; ++c)
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] = fmaf(grow[i * HID + c], R[k][c], acc[i]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (r0 + i < nT) {
              if (k > 0) out[(r0 + i) * HID + lane] = acc[i];
              else a.gx[((size_t)r.n0 + r0 + i) * HID + lane] = acc[i];
            }
          }
        }
        if (k > 0) named_bar_sync(1, BWD_THREADS / 2);
      }
    } else {
      // recompute A^k x, then rank-1 updates of grad_W held in registers
#pragma unroll
      for (int k = 1; k <= K; ++k) {
        for (int blk = rw; blk < nblk; blk += BWD_ROLE_WARPS) {
          float h[4] = {0.f, 0.f, 0.f, 0.f};
          hop_rows<4>(topo, X + (k - 1) * TR * HID, blk * 4, nT, lane, h);
#pragma unroll
          for (int i = 0; i < 4; ++i) X[(k * TR + blk * 4 + i) * HID + lane] = h[i];
        }
        named_bar_sync(2, BWD_THREADS / 2);
      }
      for (int row = rw; row < nT; row += BWD_ROLE_WARPS) {
        float xk[K + 1];
#pragma unroll
        for (int k = 0; k <= K; ++k) xk[k] = X[(k * TR + row) * HID + lane];
        const float* grow = G + row * HID;
        gb += grow[lane];
        if (CP >= 4) {
#pragma unroll
          for (int c4 = 0; c4 < CP / 4; ++c4) {
            const float4 v = *reinterpret_cast<const float4*>(grow + 4 * c4);
#pragma unroll
            for (int k = 0; k <= K; ++k) {
              R[k][4 * c4 + 0] = fmaf(v.x, xk[k], R[k][4 * c4 + 0]);
              R[k][4 * c4 + 1] = fmaf(v.y, xk[k], R[k][4 * c4 + 1]);
              R[k][4 * c4 + 2] = fmaf(v.z, xk[k], R[k][4 * c4 + 2]);
              R[k][4 * c4 + 3] = fmaf(v.w, xk[k], R[k][4 * c4 + 3]);
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < CP; ++c)
#pragma unroll
            for (int k = 0; k <= K; ++k) R[k][c] = fmaf(grow[c], xk[k], R[k][c]);
        }
      }
    }
    __syncthreads();
  }

  // cross-warp reduction of the W role's accumulators through shared memory, one partial per CTA
  float* red = smem;   // [ROLE_WARPS][(K+1)*CP + 1][32]
  constexpr int PER = (K + 1) * CP + 1;
  if (!x_role) {
    float* mine = red + (size_t)rw * PER * HID;
#pragma unroll
    for (int k = 0; k <= K; ++k)
#pragma unroll
      for (int c = 0; c < CP; ++c) mine[(k * CP + c) * HID + lane] = R[k][c];
    mine[(K + 1) * CP * HID + lane] = gb;
  }
  __syncthreads();
  float* part = a.partials + (size_t)blockIdx.x * a.partial_stride;
  const int nW = (K + 1) * cout * HID;
  for (int i = tid; i < nW + cout; i += BWD_THREADS) {
    int src;
    if (i < nW) {
      const int k = i / (cout * HID), rem = i - k * cout * HID;   // rem = c*32 + j
      src = (k * CP) * HID + rem;
    } else {
      src = (K + 1) * CP * HID + (i - nW);
    }
    float s = 0.0f;
#pragma unroll
    for (int w8 = 0; w8 < BWD_ROLE_WARPS; ++w8) s += red[(size_t)w8 * PER * HID + src];
    if (i < nW) part[i] = s;
    else part[a.bias_offset + (i - nW)] = s;
  }
}

// -------------------------------------------------------------------------------------------------
// large-graph path (a graph does not fit a tile): hops are separate SpMM launches over the global CSR, the transform streams rows.
//   forward : L1 = A x, L2 = A L1 (scratch) ; y = sum_k L_k W_k^T + b [+ dropout, ReLU, sign word, residual]
//   backward: G0 = grad_y * [y>0]/(1-p), G1 = A G0, G2 = A G1 ; grad_x = sum_k G_k W_k (same transform kernel, transposed weights) ;
//             grad_W_k = G_k^T x, grad_b via the row-streaming tcgen05 kernel k_tag_gw (tag_tc2.cu) - it never needed tiles
// -------------------------------------------------------------------------------------------------
constexpr int GEN_THREADS = 256;

// one hop Y = A_hat X over the whole batch: 8 lanes own one row (a float4 of features each), so a warp advances 4 rows per instruction;
// entries are walked in CSR order = PyG scatter order
__global__ void __launch_bounds__(GEN_THREADS) k_hop_generic(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                             const float* __restrict__ w, const float* __restrict__ X, float* __restrict__ Y,
                                                             int64_t Nt) {
  const int sub = threadIdx.x & 7;
  const int64_t grp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3, ngrp = ((int64_t)gridDim.x * blockDim.x) >> 3;
  const float4* X4 = reinterpret_cast<const float4*>(X);
  float4* Y4 = reinterpret_cast<float4*>(Y);
  for (int64_t n = grp; n < Nt; n += ngrp) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int z1 = rowptr[n + 1];
    for (int z = rowptr[n]; z < z1; ++z) {
      const float wz = w[z];
      const float4 v = X4[(size_t)col[z] * (HID / 4) + sub];
      acc.x = fmaf(wz, v.x, acc.x);
      acc.y = fmaf(wz, v.y, acc.y);
      acc.z = fmaf(wz, v.z, acc.z);
      acc.w = fmaf(wz, v.w, acc.w);
    }
    Y4[(size_t)n * (HID / 4) + sub] = acc;
  }
}

__global__ void __launch_bounds__(GEN_THREADS) k_mask_grad(const float* __restrict__ gy, const uint32_t* __restrict__ bits, int cout,
                                                           float scale, float* __restrict__ G0, int64_t Nt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Nt * HID; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i >> 5;
    const int c = (int)(i & 31);
    float v = c < cout ? gy[n * cout + c] : 0.0f;
    if (bits) v = ((bits[n] >> c) & 1u) ? v * scale : 0.0f;
    G0[i] = v;
  }
}

// WT[k][j][c] = W[k][c][j] for c < cout, 0 beyond: the transform kernel then yields grad_x from the hop levels of the gradient
__global__ void k_transpose_w(const float* __restrict__ W, int cout, int K, float* __restrict__ WT) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (K + 1) * HID * HID; i += gridDim.x * blockDim.x) {
    const int k = i / (HID * HID), j = (i / HID) % HID, c = i % HID;
    WT[i] = c < cout ? W[((size_t)k * cout + c) * HID + j] : 0.0f;
  }
}

struct DenseArgs {
  const float* L[MAXK + 1];   // hop levels [Nt,32]
  const float* w;             // [K+1][cout][32]
  const float* bias;          // may be NULL
  int cout, act, drop_mode;
  float scale;
  uint32_t keep_thr;
  const uint64_t* rng;
  uint32_t layer_uid;
  const uint8_t* mask;
  const float* res;
  int64_t res_stride;
  float* y;
  uint32_t* bits;
  int64_t Nt;
};

template <int K>
__global__ void __launch_bounds__(GEN_THREADS, 2) k_dense_tag(DenseArgs a) {
  __shared__ __align__(16) float stage[GEN_THREADS / 32][K + 1][4][HID];   // per warp: 4 rows of every level
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int cout = a.cout;
  float W[K + 1][HID];
#pragma unroll
  for (int k = 0; k <= K; ++k)
#pragma unroll
    for (int j4 = 0; j4 < HID / 4; ++j4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane < cout) v = *reinterpret_cast<const float4*>(a.w + ((size_t)k * cout + lane) * HID + 4 * j4);
      W[k][4 * j4 + 0] = v.x;
      W[k][4 * j4 + 1] = v.y;
      W[k][4 * j4 + 2] = v.z;
      W[k][4 * j4 + 3] = v.w;
    }
  const float bias = (a.bias && lane < cout) ? a.bias[lane] : 0.0f;
  uint2 key = make_uint2(0u, 0u);
  uint32_t step_lo = 0;
  if (a.drop_mode == 1) {
    const uint64_t seed = a.rng[0], step = a.rng[1];
    key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (a.layer_uid * 0x9E3779B9u) ^ (uint32_t)(step >> 32));
    step_lo = (uint32_t)step;
  }
  const int64_t nblk = (a.Nt + 3) >> 2;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t blk = warp; blk < nblk; blk += nwarps) {
    const int64_t r0 = blk * 4;
#pragma unroll
    for (int k = 0; k <= K; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) stage[wib][k][i][lane] = (r0 + i < a.Nt) ? a.L[k][(size_t)(r0 + i) * HID + lane] : 0.0f;
    __syncwarp();
    float acc[4] = {bias, bias, bias, bias};
#pragma unroll
    for (int k = 0; k <= K; ++k)
#pragma unroll
      for (int j4 = 0; j4 < HID / 4; ++j4) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(&stage[wib][k][i][4 * j4]);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = fmaf(v[i].x, W[k][4 * j4 + 0], acc[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = fmaf(v[i].y, W[k][4 * j4 + 1], acc[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = fmaf(v[i].z, W[k][4 * j4 + 2], acc[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = fmaf(v[i].w, W[k][4 * j4 + 3], acc[i]);
      }
    uint32_t rnd[4] = {0u, 0u, 0u, 0u};
    if (a.act && a.drop_mode == 1) {
      const uint4 q = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)lane, step_lo), key);
      rnd[0] = q.x;
      rnd[1] = q.y;
      rnd[2] = q.z;
      rnd[3] = q.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t n = r0 + i;
      if (n >= a.Nt) break;
      float v = acc[i];
      if (a.act) {
        bool keep = true;
        if (a.drop_mode == 1) keep = rnd[i] < a.keep_thr;
        else if (a.drop_mode == 2) keep = a.mask[n * HID + lane] != 0;
        if (a.drop_mode != 0) v = keep ? v * a.scale : 0.0f;
        v = fmaxf(v, 0.0f);
        const uint32_t word = __ballot_sync(0xffffffffu, v > 0.0f);
        if (lane == 0 && a.bits) a.bits[n] = word;
      }
      if (lane < cout) {
        if (a.res) v += a.res[n * a.res_stride + lane];
        a.y[n * cout + lane] = v;
      }
    }
    __syncwarp();
  }
}

inline int gen_grid(int64_t work_items, int per_block) {
  return (int)max((int64_t)1, min((int64_t)dss2_sm_count() * 8, (work_items + per_block - 1) / per_block));
}

template <int K>
int launch_dense(const DenseArgs& a, cudaStream_t s) {
  k_dense_tag<K><<<gen_grid((a.Nt + 3) / 4, GEN_THREADS / 32), GEN_THREADS, 0, s>>>(a);
  DSS2_LAUNCH_CHECK();
  return 0;
}
int launch_dense_k(const DenseArgs& a, int K, cudaStream_t s) {
  switch (K) {
    case 1: return launch_dense<1>(a, s);
    case 2: return launch_dense<2>(a, s);
    default: return launch_dense<3>(a, s);
  }
}
int launch_hop(const dss2_graph_t* g, const float* X, float* Y, cudaStream_t s) {
  k_hop_generic<<<gen_grid(g->num_nodes, GEN_THREADS / 8), GEN_THREADS, 0, s>>>(g->rowptr, g->col, g->w, X, Y, g->num_nodes);
  DSS2_LAUNCH_CHECK();
  return 0;
}

// grad[i] = sum over the per-CTA partial rows, in a fixed order: 8 row groups per column (row r in group r % 8) summed in parallel,
// then the 8 group sums in group order.  148 dependent loads per thread made the one-group version latency bound (~20 us per launch
// whatever the column count); the per-sub-net launches of the two-stream backward are short.
constexpr int RP_COLS = 32, RP_GROUPS = 8;
__global__ void __launch_bounds__(RP_COLS* RP_GROUPS) k_reduce_partials(const float* __restrict__ partials, int64_t stride, int num, int64_t count,
                                                                        float* grad, int accumulate) {
  __shared__ float red[RP_GROUPS][RP_COLS];
  const int cx = threadIdx.x & (RP_COLS - 1), gy = threadIdx.x / RP_COLS;
  for (int64_t base = (int64_t)blockIdx.x * RP_COLS; base < count; base += (int64_t)gridDim.x * RP_COLS) {
    const int64_t i = base + cx;
    float s = 0.0f;
    if (i < count) {
      int c = gy;
      for (; c + 3 * RP_GROUPS < num; c += 4 * RP_GROUPS) {
        const float v0 = partials[(size_t)c * stride + i], v1 = partials[(size_t)(c + RP_GROUPS) * stride + i];
        const float v2 = partials[(size_t)(c + 2 * RP_GROUPS) * stride + i], v3 = partials[(size_t)(c + 3 * RP_GROUPS) * stride + i];
        s += v0;
        s += v1;
        s += v2;
        s += v3;
      }
      for (; c < num; c += RP_GROUPS) s += partials[(size_t)c * stride + i];
    }
    red[gy][cx] = s;
    __syncthreads();
    if (gy == 0 && i < count) {
      float t = red[0][cx];
#pragma unroll
      for (int q = 1; q < RP_GROUPS; ++q) t += red[q][cx];
      grad[i] = accumulate ? grad[i] + t : t;
    }
    __syncthreads();
  }
}

size_t topo_bytes(const dss2_graph_t* g) {
  const int TR = round4(g->max_tile_nodes);
  return (size_t)(g->max_tile_nnz + 2) * 8 + (size_t)(TR + 8) * 4;
}
size_t fwd_smem(const dss2_graph_t* g, int K) { return (size_t)(K + 1) * round4(g->max_tile_nodes) * HID * 4 + topo_bytes(g); }
size_t bwd_smem(const dss2_graph_t* g, int K, int CP) {
  size_t tile = (size_t)(K + 4) * round4(g->max_tile_nodes) * HID * 4 + topo_bytes(g);
  size_t red = (size_t)BWD_ROLE_WARPS * ((K + 1) * CP + 1) * HID * 4;
  return tile > red ? tile : red;
}

template <typename Kern>
int set_smem(Kern kern, size_t bytes) {
  if (bytes > 48 * 1024) DSS2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <int K>
int launch_fwd(const TagFwdArgs& a, int grid, size_t smem, cudaStream_t s) {
  if (set_smem(k_tag_fwd<K>, smem)) return -2;
  k_tag_fwd<K><<<grid, FWD_THREADS, smem, s>>>(a);
  DSS2_LAUNCH_CHECK();
  return 0;
}
template <int K, int CP>
int launch_bwd(const TagBwdArgs& a, int grid, size_t smem, cudaStream_t s) {
  if (set_smem(k_tag_bwd<K, CP>, smem)) return -2;
  k_tag_bwd<K, CP><<<grid, BWD_THREADS, smem, s>>>(a);
  DSS2_LAUNCH_CHECK();
  return 0;
}
template <int K>
int launch_bwd_cp(const TagBwdArgs& a, int cp, int grid, size_t smem, cudaStream_t s) {
  if (cp == 2) return launch_bwd<K, 2>(a, grid, smem, s);
  if (cp == 8) return launch_bwd<K, 8>(a, grid, smem, s);
  return launch_bwd<K, 32>(a, grid, smem, s);
}
inline int pad_cout(int cout) { return cout <= 2 ? 2 : (cout <= 8 ? 8 : 32); }

}  // namespace

extern "C" int dss2_num_partials(void) { return dss2_sm_count(); }

// DSS2_DENSE_TC=0 keeps the large-graph transform on the CUDA cores (k_dense_tag): measurement / bisecting aid
static bool dense_tc_enabled() {
  const char* e = getenv("DSS2_DENSE_TC");
  return !(e && e[0] == '0');
}

extern "C" int dss2_tag_fwd(const dss2_graph_t* g, const float* x, const float* w, const float* bias, int cout, int K, int act,
                            float p_drop, int drop_mode, const uint64_t* rng_state, uint32_t layer_uid, const uint8_t* mask,
                            const float* res, int64_t res_stride, float* y, uint32_t* act_bits, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSS2_CHECK_ARG(g && x && w && bias && y, "dss2_tag_fwd: null argument");
  DSS2_CHECK_ARG(cout >= 1 && cout <= HID, "dss2_tag_fwd: cout %d outside 1..%d", cout, HID);
  DSS2_CHECK_ARG(K >= 1 && K <= MAXK, "dss2_tag_fwd: K %d outside 1..%d", K, MAXK);
  DSS2_CHECK_ARG(p_drop >= 0.0f && p_drop < 1.0f, "dss2_tag_fwd: dropout p %f outside [0,1)", p_drop);
  DSS2_CHECK_ARG(!(act && drop_mode == 1) || rng_state, "dss2_tag_fwd: philox dropout needs rng_state");
  DSS2_CHECK_ARG(!(act && drop_mode == 2) || mask, "dss2_tag_fwd: mask dropout needs a mask");
  if (g->num_nodes == 0) return 0;
  if (g->num_tiles == 0) {   // large-graph path
    DSS2_NEED_SCRATCH(g, "dss2_tag_fwd");
    const int64_t Nt = g->num_nodes;
    float* lv = g->scratch;
    DenseArgs d = {};
    d.L[0] = x;
    for (int k = 1; k <= K; ++k) {
      if (launch_hop(g, d.L[k - 1], lv + (size_t)(k - 1) * Nt * HID, stream)) return -3;
      d.L[k] = lv + (size_t)(k - 1) * Nt * HID;
    }
    if (dense_tc_enabled() && K <= 2 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0)   // transform on the tensor cores (same kernel as the tiled path, hop-free)
      return dss2_tc2_dense_fwd(g, x, lv, w, bias, cout, K, act, p_drop, drop_mode, rng_state, layer_uid, mask, res, res_stride, y, act_bits,
                                stream);
    d.w = w;
    d.bias = bias;
    d.cout = cout;
    d.act = act;
    if (p_drop == 0.0f) drop_mode = 0;
    d.drop_mode = act ? drop_mode : 0;
    d.scale = 1.0f / (float)(1.0 - (double)p_drop);
    double thr_ = (1.0 - (double)p_drop) * 4294967296.0;
    d.keep_thr = thr_ >= 4294967295.0 ? 0xffffffffu : (uint32_t)thr_;
    d.rng = rng_state;
    d.layer_uid = layer_uid;
    d.mask = mask;
    d.res = res;
    d.res_stride = res_stride;
    d.y = y;
    d.bits = act_bits;
    d.Nt = Nt;
    return launch_dense_k(d, K, stream);
  }
  TagFwdArgs a;
  a.g = *g;
  a.x = x;
  a.w = w;
  a.bias = bias;
  a.cout = cout;
  a.act = act;
  if (p_drop == 0.0f) drop_mode = 0;
  a.drop_mode = act ? drop_mode : 0;
  a.scale = 1.0f / (float)(1.0 - (double)p_drop);
  double thr = (1.0 - (double)p_drop) * 4294967296.0;
  a.keep_thr = thr >= 4294967295.0 ? 0xffffffffu : (uint32_t)thr;
  a.rng = rng_state;
  a.layer_uid = layer_uid;
  a.mask = mask;
  a.res = res;
  a.res_stride = res_stride;
  a.y = y;
  a.bits = act_bits;
  size_t smem = fwd_smem(g, K);
  DSS2_CHECK_ARG(smem <= 227 * 1024, "dss2_tag_fwd: tile needs %zu bytes of shared memory", smem);
  int grid = max(1, min(g->num_tiles, 2 * dss2_sm_count()));
  switch (K) {
    case 1: return launch_fwd<1>(a, grid, smem, stream);
    case 2: return launch_fwd<2>(a, grid, smem, stream);
    default: return launch_fwd<3>(a, grid, smem, stream);
  }
}

extern "C" int dss2_tag_bwd(const dss2_graph_t* g, const float* x, const float* w, int cout, int K, int act, float p_drop,
                            const uint32_t* act_bits, const float* grad_y, float* grad_x, float* partials, int64_t partial_stride,
                            int64_t bias_offset, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSS2_CHECK_ARG(g && x && w && grad_y && grad_x && partials, "dss2_tag_bwd: null argument");
  DSS2_CHECK_ARG(cout >= 1 && cout <= HID, "dss2_tag_bwd: cout %d outside 1..%d", cout, HID);
  DSS2_CHECK_ARG(K >= 1 && K <= MAXK, "dss2_tag_bwd: K %d outside 1..%d", K, MAXK);
  DSS2_CHECK_ARG(!act || act_bits, "dss2_tag_bwd: activation layers need act_bits from the forward");
  DSS2_CHECK_ARG(partial_stride >= (int64_t)(K + 1) * cout * HID + cout, "dss2_tag_bwd: partial_stride too small");
  DSS2_CHECK_ARG(bias_offset >= (int64_t)(K + 1) * cout * HID || bias_offset <= -(int64_t)cout, "dss2_tag_bwd: bias_offset overlaps grad_W");
  if (g->num_nodes == 0) return 0;
  if (g->num_tiles == 0) {   // large-graph path: hops on the masked output gradient, then two GEMMs
    DSS2_NEED_SCRATCH(g, "dss2_tag_bwd");
    DSS2_CHECK_ARG(K <= 2, "dss2_tag_bwd: the large-graph backward supports K <= 2 (weight-gradient kernel)");
    const int64_t Nt = g->num_nodes;
    float* lvl = g->scratch;                         // [K][Nt,32] = A g, A^2 g  (layout k_tag_gw expects)
    float* G0 = g->scratch + (size_t)K * Nt * HID;   // masked, zero-padded gradient
    float* WT = G0 + (size_t)Nt * HID;               // transposed weights
    const float scale = 1.0f / (float)(1.0 - (double)p_drop);
    k_mask_grad<<<gen_grid(Nt * HID, GEN_THREADS), GEN_THREADS, 0, stream>>>(grad_y, act ? act_bits : nullptr, cout, scale, G0, Nt);
    DSS2_LAUNCH_CHECK();
    DenseArgs d = {};
    d.L[0] = G0;
    for (int k = 1; k <= K; ++k) {
      if (launch_hop(g, d.L[k - 1], lvl + (size_t)(k - 1) * Nt * HID, stream)) return -3;
      d.L[k] = lvl + (size_t)(k - 1) * Nt * HID;
    }
    if (dense_tc_enabled() && ((uintptr_t)grad_x & 15) == 0 && (cout != HID || ((uintptr_t)grad_y & 15) == 0)) {   // tensor cores, hop-free (see dss2_tag_fwd)
      if (dss2_tc2_dense_bgx(g, grad_y, act ? act_bits : nullptr, p_drop, lvl, w, cout, K, grad_x, stream)) return -3;
    } else {
      k_transpose_w<<<12, 256, 0, stream>>>(w, cout, K, WT);
      DSS2_LAUNCH_CHECK();
      d.w = WT;
      d.cout = HID;
      d.y = grad_x;
      d.Nt = Nt;
      if (launch_dense_k(d, K, stream)) return -3;
    }
    // the hop levels above are plain rows: say so in the format word the weight-gradient kernels read behind them
    DSS2_CUDA(cudaMemsetAsync(lvl + (size_t)K * Nt * HID, 0, sizeof(uint32_t), stream));
    // weight gradients: the tcgen05 streaming GEMM when its alignment contract holds (it measures 2x faster), else the exact FMA pass
    const bool aligned = (((uintptr_t)x | (uintptr_t)grad_y | (uintptr_t)act_bits | (uintptr_t)partials) & 15) == 0;
    if (dense_tc_enabled() && aligned)
      return dss2_tag_bwd_tc2_gw(Nt, x, cout, K, act, p_drop, act_bits, grad_y, partials, partial_stride, bias_offset, lvl,
                                 dss2_tag_bwd_tc2_workspace_bytes(Nt, K), stream_);
    return dss2_tag_gw_ffma(Nt, x, cout, K, act, p_drop, act_bits, grad_y, partials, partial_stride, bias_offset, lvl,
                            dss2_tag_bwd_tc2_workspace_bytes(Nt, K), stream_);
  }
  TagBwdArgs a;
  a.g = *g;
  a.x = x;
  a.w = w;
  a.cout = cout;
  a.act = act;
  a.scale = 1.0f / (float)(1.0 - (double)p_drop);
  a.bits = act_bits;
  a.gy = grad_y;
  a.gx = grad_x;
  a.partials = partials;
  a.partial_stride = partial_stride;
  a.bias_offset = bias_offset;
  const int cp = pad_cout(cout);
  size_t smem = bwd_smem(g, K, cp);
  DSS2_CHECK_ARG(smem <= 227 * 1024, "dss2_tag_bwd: tile needs %zu bytes of shared memory", smem);
  // exactly dss2_num_partials() CTAs so that every partial row is written (idle CTAs write zeros)
  int grid = dss2_sm_count();
  switch (K) {
    case 1: return launch_bwd_cp<1>(a, cp, grid, smem, stream);
    case 2: return launch_bwd_cp<2>(a, cp, grid, smem, stream);
    default: return launch_bwd_cp<3>(a, cp, grid, smem, stream);
  }
}

extern "C" int dss2_reduce_partials(const float* partials, int64_t partial_stride, int num_partials, int64_t count, float* grad,
                                    int accumulate, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSS2_CHECK_ARG(partials && grad && num_partials >= 1 && count >= 0, "dss2_reduce_partials: bad argument");
  if (count == 0) return 0;
  int grid = (int)max((int64_t)1, min((int64_t)148 * 8, (count + RP_COLS - 1) / RP_COLS));
  k_reduce_partials<<<grid, RP_COLS * RP_GROUPS, 0, stream>>>(partials, partial_stride, num_partials, count, grad, accumulate);
  DSS2_LAUNCH_CHECK();
  return 0;
}
