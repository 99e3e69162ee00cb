// (b1) EdgeAggregation forward / backward.  Replaces networks.py:159-209 (PyG propagate: gather x_i, x_j,
// concat with edge_attr, 22->32->32 MLP per doubled edge, scatter-add by target) and its autograd.
//
// The reference materialises a [2Et, 22] concat and a [2Et, 32] message tensor.  Here a tile of whole
// graphs lives in shared memory and the first Linear is split by operand,
//     W1 [x_dst | x_src | a_e] + b1 = (W1a x_dst + b1) + W1b x_src + W1c a_e = P[dst] + Q[src] + W1c a_e,
// so the node part is computed once per node instead of once per edge, and the second Linear is
// applied once per node to the sum of hidden activations (it is linear):
//     out[n] = W2 * sum_e relu(P[n] + Q[src_e] + W1c a_e) + deg(n) * b2.
// One warp owns a destination node (lane = hidden unit) and walks its CSR row in PyG scatter order:
// segmented sum with no atomics.  Reversed edges use a_e with columns 0 and 2 negated
// (networks.py:252) straight from the forward edge's row: the doubled attribute tensor never exists.
// The backward re-evaluates each in-edge and its twin (every edge has one in the doubled graph), which
// yields all contributions to grad_x[n] inside the warp that owns n.
#include "common.cuh"

namespace {

constexpr int EA_THREADS = 256;
constexpr int EA_WARPS = EA_THREADS / 32;
constexpr int FP = 8;   // padded node / edge feature count

struct EaArgs {
  dss2_graph_t g;
  const float* x;
  int64_t xs;
  int fn;
  const float* ea;
  int64_t eas;
  int fe;
  const float* w1;
  const float* b1;
  const float* w2;
  const float* b2;
  float* out;            // fwd
  const float* gout;     // bwd
  const float* skip;
  int64_t skip_stride;
  float* gx;
  float* partials;
  int64_t partial_stride;
};

__host__ __device__ inline int round4(int v) { return (v + 3) & ~3; }

struct EaSmem {
  float* xs;     // [TR][8]
  float* at;     // [ER][8]   edge attributes of the tile's one-way edges
  float* P;      // [TR][32]
  float* Q;      // [TR][32]
  int* rowptr;   // [TR+1]
  int* col;      // [Z]
  uint32_t* eid; // [Z]  local one-way edge id | reversed << 31
};

__device__ __forceinline__ EaSmem carve(float* smem, int TR, int ER, int Z, int extra_bufs, float** extra) {
  EaSmem s;
  s.P = smem;
  s.Q = s.P + TR * HID;
  float* p = s.Q + TR * HID;
  for (int i = 0; i < extra_bufs; ++i) {
    extra[i] = p;
    p += TR * HID;
  }
  s.xs = p;
  s.at = s.xs + TR * FP;
  s.rowptr = reinterpret_cast<int*>(s.at + ER * FP);
  s.col = s.rowptr + TR + 4;
  s.eid = reinterpret_cast<uint32_t*>(s.col + Z + 4);
  return s;
}

__device__ __forceinline__ void load_tile(const EaArgs& a, const TileRange& r, const EaSmem& s, int tid, int nthr) {
  const dss2_graph_t& g = a.g;
  const int nT = r.n1 - r.n0, nE = (int)(r.e1 - r.e0), nZ = r.z1 - r.z0;
  for (int i = tid; i < nT * FP; i += nthr) {
    const int row = i >> 3, c = i & 7;
    s.xs[i] = c < a.fn ? a.x[((size_t)r.n0 + row) * a.xs + c] : 0.0f;
  }
  for (int i = tid; i < nE * FP; i += nthr) {
    const int row = i >> 3, c = i & 7;
    s.at[i] = c < a.fe ? a.ea[((size_t)r.e0 + row) * a.eas + c] : 0.0f;
  }
  for (int i = tid; i <= nT; i += nthr) s.rowptr[i] = g.rowptr[r.n0 + i] - r.z0;
  for (int i = tid; i < nZ; i += nthr) {
    s.col[i] = g.col[r.z0 + i] - r.n0;
    const uint32_t id = g.eid[r.z0 + i];
    s.eid[i] = (uint32_t)((int64_t)(id & 0x7fffffffu) - r.e0) | (id & 0x80000000u);
  }
}

// lane h keeps row h of W1 split into its three operand blocks (zero padded to 8 columns each)
struct W1Row {
  float a[FP], b[FP], c[FP], bias;
};
__device__ __forceinline__ W1Row load_w1(const EaArgs& a, int lane) {
  W1Row w;
  const int ld = 2 * a.fn + a.fe;
  const float* row = a.w1 + (size_t)lane * ld;
#pragma unroll
  for (int i = 0; i < FP; ++i) {
    w.a[i] = i < a.fn ? row[i] : 0.0f;
    w.b[i] = i < a.fn ? row[a.fn + i] : 0.0f;
    w.c[i] = i < a.fe ? row[2 * a.fn + i] : 0.0f;
  }
  w.bias = a.b1[lane];
  return w;
}

__device__ __forceinline__ void node_pq(const W1Row& w, const float* xrow, float& p, float& q) {
  const float4 x0 = *reinterpret_cast<const float4*>(xrow), x1 = *reinterpret_cast<const float4*>(xrow + 4);
  const float xv[FP] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
  p = w.bias;
  q = 0.0f;
#pragma unroll
  for (int i = 0; i < FP; ++i) {
    p = fmaf(w.a[i], xv[i], p);
    q = fmaf(w.b[i], xv[i], q);
  }
}

// W1c a_e for the (possibly reversed) edge; `sgn` = -1 flips columns 0 and 2 (networks.py:252)
__device__ __forceinline__ float edge_term(const W1Row& w, const float* arow, float sgn, float* av) {
  const float4 a0 = *reinterpret_cast<const float4*>(arow), a1 = *reinterpret_cast<const float4*>(arow + 4);
  av[0] = sgn * a0.x;
  av[1] = a0.y;
  av[2] = sgn * a0.z;
  av[3] = a0.w;
  av[4] = a1.x;
  av[5] = a1.y;
  av[6] = a1.z;
  av[7] = a1.w;
  float t = 0.0f;
#pragma unroll
  for (int i = 0; i < FP; ++i) t = fmaf(w.c[i], av[i], t);
  return t;
}

// Sum each of 8 per-lane values over the warp with 9 shuffles instead of 40: halve the value set while halving the lane set
// (xor 16, 8, 4), then two plain butterfly steps.  Lane l ends up with the total of v[l >> 2].
__device__ __forceinline__ float reduce8_transposed(const float (&v)[FP], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float u[4], t[2];
#pragma unroll
  for (int j = 0; j < 4; ++j) u[j] = (b4 ? v[4 + j] : v[j]) + __shfl_xor_sync(0xffffffffu, b4 ? v[j] : v[4 + j], 16);
#pragma unroll
  for (int j = 0; j < 2; ++j) t[j] = (b3 ? u[2 + j] : u[j]) + __shfl_xor_sync(0xffffffffu, b3 ? u[j] : u[2 + j], 8);
  float s = (b2 ? t[1] : t[0]) + __shfl_xor_sync(0xffffffffu, b2 ? t[0] : t[1], 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  return s;
}

// out[lane] = sum_h M[lane][h] * vec[h], vec broadcast from shared memory, M row in registers
__device__ __forceinline__ float matvec32(const float (&Mrow)[HID], const float* vec) {
  float acc = 0.0f;
#pragma unroll
  for (int j4 = 0; j4 < HID / 4; ++j4) {
    const float4 v = *reinterpret_cast<const float4*>(vec + 4 * j4);
    acc = fmaf(Mrow[4 * j4 + 0], v.x, acc);
    acc = fmaf(Mrow[4 * j4 + 1], v.y, acc);
    acc = fmaf(Mrow[4 * j4 + 2], v.z, acc);
    acc = fmaf(Mrow[4 * j4 + 3], v.w, acc);
  }
  return acc;
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_edgeagg_fwd(EaArgs a) {
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) float smem[];
  const dss2_graph_t& g = a.g;
  const int TR = round4(g.max_tile_nodes), ER = round4(g.max_tile_edges);
  EaSmem s = carve(smem, TR, ER, g.max_tile_nnz, 0, nullptr);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const W1Row w1 = load_w1(a, lane);
  float W2row[HID];   // lane o keeps row o of W2
#pragma unroll
  for (int h = 0; h < HID; ++h) W2row[h] = a.w2[lane * HID + h];
  const float b2 = a.b2[lane];

  for (int t = blockIdx.x; t < g.num_tiles; t += gridDim.x) {
    const TileRange r = tile_range(g, t);
    const int nT = r.n1 - r.n0;
    load_tile(a, r, s, tid, NT);
    __syncthreads();
    for (int row = warp; row < nT; row += NW) {
      float p, q;
      node_pq(w1, s.xs + row * FP, p, q);
      s.P[row * HID + lane] = p;
      s.Q[row * HID + lane] = q;
    }
    __syncthreads();
    for (int row = warp; row < nT; row += NW) {
      const int beg = s.rowptr[row], end = s.rowptr[row + 1];
      const float p = s.P[row * HID + lane];
      float S = 0.0f;
      for (int z = beg; z < end; ++z) {
        const uint32_t id = s.eid[z];
        float av[FP];
        const float pre = p + s.Q[s.col[z] * HID + lane] + edge_term(w1, s.at + (id & 0x7fffffffu) * FP, (id >> 31) ? -1.0f : 1.0f, av);
        S += fmaxf(pre, 0.0f);
      }
      __syncwarp();
      s.P[row * HID + lane] = S;     // P[row] is only ever read by this warp: reuse it as the broadcast buffer
      __syncwarp();
      const float o = fmaf((float)(end - beg), b2, matvec32(W2row, s.P + row * HID));
      a.out[((size_t)r.n0 + row) * HID + lane] = o;
    }
    __syncthreads();
  }
}

// partial layout: w1 [32][ld], b1 [32], w2 [32][32], b2 [32]
template <int NT>
__global__ void __launch_bounds__(NT, 1) k_edgeagg_bwd(EaArgs a) {
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) float smem[];
  const dss2_graph_t& g = a.g;
  const int TR = round4(g.max_tile_nodes), ER = round4(g.max_tile_edges);
  float* extra[2];
  EaSmem s = carve(smem, TR, ER, g.max_tile_nnz, 2, extra);
  float* GO = extra[0];   // grad_out tile
  float* GS = extra[1];   // grad wrt S = sum of hidden activations
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fn = a.fn, fe = a.fe, ld = 2 * fn + fe;
  const W1Row w1 = load_w1(a, lane);

  // accumulators (lane = hidden unit h, except gb2 where lane = output unit o)
  float gW1a[FP], gW1b[FP], gW1c[FP], gb1 = 0.0f, gb2 = 0.0f, gW2[HID];
#pragma unroll
  for (int i = 0; i < FP; ++i) gW1a[i] = gW1b[i] = gW1c[i] = 0.0f;
#pragma unroll
  for (int o = 0; o < HID; ++o) gW2[o] = 0.0f;

  for (int t = blockIdx.x; t < g.num_tiles; t += gridDim.x) {
    const TileRange r = tile_range(g, t);
    const int nT = r.n1 - r.n0;
    load_tile(a, r, s, tid, NT);
    {
      const float4* src = reinterpret_cast<const float4*>(a.gout + (size_t)r.n0 * HID);
      float4* dst = reinterpret_cast<float4*>(GO);
      for (int i = tid; i < nT * (HID / 4); i += NT) dst[i] = ldg_stream4(src + i);
    }
    __syncthreads();
    {
      float W2col[HID];   // lane h keeps column h of W2: W2[o][h]; only live in this phase (reloaded per tile from L1)
#pragma unroll
      for (int o = 0; o < HID; ++o) W2col[o] = __ldg(a.w2 + o * HID + lane);
    for (int row = warp; row < nT; row += NW) {
      float p, q;
      node_pq(w1, s.xs + row * FP, p, q);
      s.P[row * HID + lane] = p;
      s.Q[row * HID + lane] = q;
      // grad_S[h] = sum_o W2[o][h] * grad_out[o]
      GS[row * HID + lane] = matvec32(W2col, GO + row * HID);
      gb2 = fmaf((float)(s.rowptr[row + 1] - s.rowptr[row]), GO[row * HID + lane], gb2);
    }
    }
    __syncthreads();
    for (int row = warp; row < nT; row += NW) {
      const int beg = s.rowptr[row], end = s.rowptr[row + 1];
      const float p = s.P[row * HID + lane], q = s.Q[row * HID + lane], gs = GS[row * HID + lane];
      float S = 0.0f, gP = 0.0f, gQ = 0.0f;
      for (int z = beg; z < end; ++z) {
        const uint32_t id = s.eid[z];
        const int c = s.col[z];
        const float sgn = (id >> 31) ? -1.0f : 1.0f;
        float av[FP];
        const float et = edge_term(w1, s.at + (id & 0x7fffffffu) * FP, sgn, av);
        // in-edge (c -> row)
        const float pre_in = p + s.Q[c * HID + lane] + et;
        S += fmaxf(pre_in, 0.0f);
        const float gp = pre_in > 0.0f ? gs : 0.0f;
        gP += gp;
#pragma unroll
        for (int i = 0; i < FP; ++i) gW1c[i] = fmaf(gp, av[i], gW1c[i]);
        // its twin (row -> c): same attributes with columns 0 and 2 flipped once more
        float et_tw = 0.0f;
#pragma unroll
        for (int i = 0; i < FP; ++i) et_tw = fmaf(w1.c[i], (i == 0 || i == 2) ? -av[i] : av[i], et_tw);
        const float pre_tw = s.P[c * HID + lane] + q + et_tw;
        gQ += pre_tw > 0.0f ? GS[c * HID + lane] : 0.0f;
      }
      // parameter gradients
      gb1 += gP;
      {
        const float* xrow = s.xs + row * FP;
#pragma unroll
        for (int i = 0; i < FP; ++i) {
          gW1a[i] = fmaf(gP, xrow[i], gW1a[i]);
          gW1b[i] = fmaf(gQ, xrow[i], gW1b[i]);
        }
        const float* go = GO + row * HID;
#pragma unroll
        for (int o4 = 0; o4 < HID / 4; ++o4) {
          const float4 v = *reinterpret_cast<const float4*>(go + 4 * o4);
          gW2[4 * o4 + 0] = fmaf(v.x, S, gW2[4 * o4 + 0]);
          gW2[4 * o4 + 1] = fmaf(v.y, S, gW2[4 * o4 + 1]);
          gW2[4 * o4 + 2] = fmaf(v.z, S, gW2[4 * o4 + 2]);
          gW2[4 * o4 + 3] = fmaf(v.w, S, gW2[4 * o4 + 3]);
        }
      }
      // grad_x[row][i] = sum_h W1a[h][i] gP[h] + W1b[h][i] gQ[h]  (+ residual-path gradient)
      if (a.gx) {
        float v[FP];
#pragma unroll
        for (int i = 0; i < FP; ++i) v[i] = fmaf(w1.a[i], gP, w1.b[i] * gQ);
        float mine = reduce8_transposed(v, lane);   // lane l: total of feature l >> 2
        const int i = lane >> 2;
        if ((lane & 3) == 0 && i < fn) {
          const size_t n = (size_t)r.n0 + row;
          if (a.skip) mine += a.skip[n * a.skip_stride + i];
          a.gx[n * fn + i] = mine;
        }
      }
    }
    __syncthreads();
  }

  // per-CTA partial: reduce the 8 warps through shared memory
  constexpr int PER = 3 * FP + 2 + HID;   // gW1a, gW1b, gW1c, gb1, gb2, gW2[32]
  float* red = smem;                      // [NW][PER][32]
  {
    float* mine = red + (size_t)warp * PER * HID;
#pragma unroll
    for (int i = 0; i < FP; ++i) {
      mine[(i)*HID + lane] = gW1a[i];
      mine[(FP + i) * HID + lane] = gW1b[i];
      mine[(2 * FP + i) * HID + lane] = gW1c[i];
    }
    mine[(3 * FP) * HID + lane] = gb1;
    mine[(3 * FP + 1) * HID + lane] = gb2;
#pragma unroll
    for (int o = 0; o < HID; ++o) mine[(3 * FP + 2 + o) * HID + lane] = gW2[o];
  }
  __syncthreads();
  float* part = a.partials + (size_t)blockIdx.x * a.partial_stride;
  const int n_w1 = HID * ld, off_b1 = n_w1, off_w2 = n_w1 + HID, off_b2 = off_w2 + HID * HID, total = off_b2 + HID;
  for (int i = tid; i < total; i += NT) {
    int slot, ln;
    if (i < n_w1) {
      const int h = i / ld, c = i - h * ld;
      slot = c < fn ? c : (c < 2 * fn ? FP + (c - fn) : 2 * FP + (c - 2 * fn));
      ln = h;
    } else if (i < off_w2) {
      slot = 3 * FP;
      ln = i - off_b1;
    } else if (i < off_b2) {
      const int o = (i - off_w2) / HID, h = (i - off_w2) - o * HID;
      slot = 3 * FP + 2 + o;
      ln = h;
    } else {
      slot = 3 * FP + 1;
      ln = i - off_b2;
    }
    float sum = 0.0f;
#pragma unroll
    for (int w8 = 0; w8 < NW; ++w8) sum += red[((size_t)w8 * PER + slot) * HID + ln];
    part[i] = sum;
  }
}

// -------------------------------------------------------------------------------------------------
// large-graph path: same arithmetic, P / Q / grad_S live in global scratch ([3][Nt,32]) and every phase is its own launch
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_x8(const EaArgs& a, int64_t n, float (&xv)[FP]) {
#pragma unroll
  for (int i = 0; i < FP; ++i) xv[i] = i < a.fn ? a.x[n * a.xs + i] : 0.0f;
}
__device__ __forceinline__ void load_a8(const EaArgs& a, uint32_t id, float (&av)[FP]) {
  const int64_t e = id & 0x7fffffffu;
  const float sgn = (id >> 31) ? -1.0f : 1.0f;
#pragma unroll
  for (int i = 0; i < FP; ++i) av[i] = i < a.fe ? a.ea[e * a.eas + i] : 0.0f;
  av[0] *= sgn;
  av[2] *= sgn;
}
__device__ __forceinline__ float dot8(const float (&w)[FP], const float (&v)[FP]) {
  float t = 0.0f;
#pragma unroll
  for (int i = 0; i < FP; ++i) t = fmaf(w[i], v[i], t);
  return t;
}
// out[lane] = sum_h M[h] * shfl(vec, h): matrix row/column in registers, vector spread over the lanes
__device__ __forceinline__ float matvec_shfl(const float (&M)[HID], float vec) {
  float acc = 0.0f;
#pragma unroll
  for (int h = 0; h < HID; ++h) acc = fmaf(M[h], __shfl_sync(0xffffffffu, vec, h), acc);
  return acc;
}

// phase 1: P = W1a x + b1, Q = W1b x  [, grad_S = W2^T grad_out]
__global__ void __launch_bounds__(EA_THREADS) k_ea_nodes_g(EaArgs a, float* P, float* Q, float* GS) {
  const dss2_graph_t& g = a.g;
  const int lane = threadIdx.x & 31;
  const W1Row w1 = load_w1(a, lane);
  float W2col[HID];
  if (GS) {
#pragma unroll
    for (int o = 0; o < HID; ++o) W2col[o] = a.w2[o * HID + lane];
  }
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t n = warp; n < g.num_nodes; n += nwarps) {
    float xv[FP];
    load_x8(a, n, xv);
    P[n * HID + lane] = w1.bias + dot8(w1.a, xv);
    Q[n * HID + lane] = dot8(w1.b, xv);
    if (GS) GS[n * HID + lane] = matvec_shfl(W2col, a.gout[n * HID + lane]);
  }
}

// phase 2 forward: out[n] = W2 * sum_e relu(P[n] + Q[src] + W1c a_e) + deg * b2
__global__ void __launch_bounds__(EA_THREADS) k_ea_fwd_g(EaArgs a, const float* __restrict__ P, const float* __restrict__ Q) {
  const dss2_graph_t& g = a.g;
  const int lane = threadIdx.x & 31;
  const W1Row w1 = load_w1(a, lane);
  float W2row[HID];
#pragma unroll
  for (int h = 0; h < HID; ++h) W2row[h] = a.w2[lane * HID + h];
  const float b2 = a.b2[lane];
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t n = warp; n < g.num_nodes; n += nwarps) {
    const int beg = g.rowptr[n], end = g.rowptr[n + 1];
    const float p = P[n * HID + lane];
    float S = 0.0f;
    for (int z = beg; z < end; ++z) {
      float av[FP];
      load_a8(a, g.eid[z], av);
      S += fmaxf(p + Q[(size_t)g.col[z] * HID + lane] + dot8(w1.c, av), 0.0f);
    }
    a.out[n * HID + lane] = fmaf((float)(end - beg), b2, matvec_shfl(W2row, S));
  }
}

// phase 2 backward: identical to the tile kernel's edge loop, with global gathers
template <int NT>
__global__ void __launch_bounds__(NT, 1) k_ea_bwd_g(EaArgs a, const float* __restrict__ P, const float* __restrict__ Q,
                                                            const float* __restrict__ GS) {
  extern __shared__ __align__(16) float smem[];
  const dss2_graph_t& g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, warp_in_block = tid >> 5;
  const int fn = a.fn, fe = a.fe, ld = 2 * fn + fe;
  const W1Row w1 = load_w1(a, lane);
  float gW1a[FP], gW1b[FP], gW1c[FP], gb1 = 0.0f, gb2 = 0.0f, gW2[HID];
#pragma unroll
  for (int i = 0; i < FP; ++i) gW1a[i] = gW1b[i] = gW1c[i] = 0.0f;
#pragma unroll
  for (int o = 0; o < HID; ++o) gW2[o] = 0.0f;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + tid) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t n = warp; n < g.num_nodes; n += nwarps) {
    const int beg = g.rowptr[n], end = g.rowptr[n + 1];
    const float p = P[n * HID + lane], q = Q[n * HID + lane], gs = GS[n * HID + lane], go = a.gout[n * HID + lane];
    float S = 0.0f, gP = 0.0f, gQ = 0.0f;
    for (int z = beg; z < end; ++z) {
      const int64_t c = g.col[z];
      float av[FP];
      load_a8(a, g.eid[z], av);
      const float pre_in = p + Q[c * HID + lane] + dot8(w1.c, av);
      S += fmaxf(pre_in, 0.0f);
      const float gp = pre_in > 0.0f ? gs : 0.0f;
      gP += gp;
#pragma unroll
      for (int i = 0; i < FP; ++i) gW1c[i] = fmaf(gp, av[i], gW1c[i]);
      av[0] = -av[0];   // twin edge (n -> c): columns 0 and 2 flipped once more
      av[2] = -av[2];
      const float pre_tw = P[c * HID + lane] + q + dot8(w1.c, av);
      gQ += pre_tw > 0.0f ? GS[c * HID + lane] : 0.0f;
    }
    gb1 += gP;
    gb2 = fmaf((float)(end - beg), go, gb2);
    float xv[FP];
    load_x8(a, n, xv);
#pragma unroll
    for (int i = 0; i < FP; ++i) {
      gW1a[i] = fmaf(gP, xv[i], gW1a[i]);
      gW1b[i] = fmaf(gQ, xv[i], gW1b[i]);
    }
#pragma unroll
    for (int o = 0; o < HID; ++o) gW2[o] = fmaf(__shfl_sync(0xffffffffu, go, o), S, gW2[o]);
    if (a.gx) {
      float v[FP];
#pragma unroll
      for (int i = 0; i < FP; ++i) v[i] = fmaf(w1.a[i], gP, w1.b[i] * gQ);
      float mine = reduce8_transposed(v, lane);
      const int i = lane >> 2;
      if ((lane & 3) == 0 && i < fn) {
        if (a.skip) mine += a.skip[n * a.skip_stride + i];
        a.gx[n * fn + i] = mine;
      }
    }
  }
  // per-CTA partial (same layout and reduction as k_edgeagg_bwd)
  constexpr int PER = 3 * FP + 2 + HID;
  float* red = smem;
  {
    float* mine = red + (size_t)warp_in_block * PER * HID;
#pragma unroll
    for (int i = 0; i < FP; ++i) {
      mine[(i)*HID + lane] = gW1a[i];
      mine[(FP + i) * HID + lane] = gW1b[i];
      mine[(2 * FP + i) * HID + lane] = gW1c[i];
    }
    mine[(3 * FP) * HID + lane] = gb1;
    mine[(3 * FP + 1) * HID + lane] = gb2;
#pragma unroll
    for (int o = 0; o < HID; ++o) mine[(3 * FP + 2 + o) * HID + lane] = gW2[o];
  }
  __syncthreads();
  float* part = a.partials + (size_t)blockIdx.x * a.partial_stride;
  const int n_w1 = HID * ld, off_b1 = n_w1, off_w2 = n_w1 + HID, off_b2 = off_w2 + HID * HID, total = off_b2 + HID;
  for (int i = tid; i < total; i += NT) {
    int slot, ln;
    if (i < n_w1) {
      const int h = i / ld, c = i - h * ld;
      slot = c < fn ? c : (c < 2 * fn ? FP + (c - fn) : 2 * FP + (c - 2 * fn));
      ln = h;
    } else if (i < off_w2) {
      slot = 3 * FP;
      ln = i - off_b1;
    } else if (i < off_b2) {
      const int o = (i - off_w2) / HID, h = (i - off_w2) - o * HID;
      slot = 3 * FP + 2 + o;
      ln = h;
    } else {
      slot = 3 * FP + 1;
      ln = i - off_b2;
    }
    float sum = 0.0f;
#pragma unroll
    for (int w8 = 0; w8 < NT / 32; ++w8) sum += red[((size_t)w8 * PER + slot) * HID + ln];
    part[i] = sum;
  }
}

inline int ea_gen_grid(int64_t rows) { return (int)max((int64_t)1, min((int64_t)dss2_sm_count() * 8, (rows + EA_WARPS - 1) / EA_WARPS)); }

size_t ea_smem(const dss2_graph_t* g, int bufs) {
  const int TR = round4(g->max_tile_nodes), ER = round4(g->max_tile_edges);
  size_t tile = (size_t)bufs * TR * HID * 4 + (size_t)TR * FP * 4 + (size_t)ER * FP * 4 + (size_t)(TR + 4) * 4 +
                (size_t)(g->max_tile_nnz + 4) * 8 + 64;
  return tile;
}

// threads per CTA of the tile kernels; the environment override exists for measurements
constexpr int EA_FWD_THREADS = 256, EA_BWD_THREADS = 384;
int ea_threads(const char* env, int dflt) {
  const char* v = getenv(env);
  const int n = v ? atoi(v) : dflt;
  return (n == 256 || n == 384 || n == 512) ? n : dflt;
}

// DSS2_EA_IMPL=warp keeps the round-1 warp-per-row kernels of this file (measured alternative); default = edgeagg_row.cu
bool ea_use_row(const dss2_graph_t* g, int64_t x_stride, int64_t ea_stride, int fe, int bwd) {
  const char* v = getenv("DSS2_EA_IMPL");
  if (v && v[0] == 'w') return false;
  return dss2_ea_row_fits(g, x_stride, ea_stride, fe, bwd) != 0;
}

int check_common(const char* who, const dss2_graph_t* g, const float* x, int fn, const float* ea, int fe, const float* w1,
                 const float* b1, const float* w2, const float* b2) {
  DSS2_CHECK_ARG(g && x && ea && w1 && b1 && w2 && b2, "%s: null argument", who);
  DSS2_CHECK_ARG(fn >= 1 && fn <= FP && fe >= 1 && fe <= FP, "%s: feature counts (%d node, %d edge) outside 1..%d", who, fn, fe, FP);
  return 0;
}

}  // namespace

extern "C" int dss2_edgeagg_fwd(const dss2_graph_t* g, const float* x, int64_t x_stride, int fn, const float* edge_attr,
                                int64_t ea_stride, int fe, const float* w1, const float* b1, const float* w2, const float* b2,
                                float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (check_common("dss2_edgeagg_fwd", g, x, fn, edge_attr, fe, w1, b1, w2, b2)) return -1;
  DSS2_CHECK_ARG(out, "dss2_edgeagg_fwd: null output");
  EaArgs a = {};
  a.g = *g;
  a.x = x;
  a.xs = x_stride;
  a.fn = fn;
  a.ea = edge_attr;
  a.eas = ea_stride;
  a.fe = fe;
  a.w1 = w1;
  a.b1 = b1;
  a.w2 = w2;
  a.b2 = b2;
  a.out = out;
  if (g->num_nodes == 0) return 0;
  if (g->num_tiles == 0) {   // large-graph path
    DSS2_NEED_SCRATCH(g, "dss2_edgeagg_fwd");
    float* P = g->scratch;
    float* Q = P + (size_t)g->num_nodes * HID;
    k_ea_nodes_g<<<ea_gen_grid(g->num_nodes), EA_THREADS, 0, stream>>>(a, P, Q, nullptr);
    DSS2_LAUNCH_CHECK();
    k_ea_fwd_g<<<ea_gen_grid(g->num_nodes), EA_THREADS, 0, stream>>>(a, P, Q);
    DSS2_LAUNCH_CHECK();
    return 0;
  }
  if (ea_use_row(g, x_stride, ea_stride, fe, 0)) {   // weights through the scratch constant-memory slot, then the thread-per-row kernel
    const int slot = dss2_ea_row_scratch_slot();
    if (dss2_ea_row_upload(slot, 1, &w1, &b1, &w2, &b2, fn, fe, stream)) return -1;
    return dss2_ea_row_fwd(g, x, x_stride, fn, edge_attr, ea_stride, fe, slot, out, stream);
  }
  size_t smem = ea_smem(g, 2);
  DSS2_CHECK_ARG(smem <= 113 * 1024, "dss2_edgeagg_fwd: tile needs %zu bytes of shared memory", smem);
  int grid = max(1, min(g->num_tiles, 2 * dss2_sm_count()));
#define DSS2_EA_FWD(NT)                                                                                                          \
  {                                                                                                                              \
    if (smem > 48 * 1024) DSS2_CUDA(cudaFuncSetAttribute(k_edgeagg_fwd<NT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_edgeagg_fwd<NT, 2><<<grid, NT, smem, stream>>>(a);                                                                         \
  }
  const int nt = ea_threads("DSS2_EA_FWD_THREADS", EA_FWD_THREADS);
  if (nt == 256) DSS2_EA_FWD(256) else if (nt == 384) DSS2_EA_FWD(384) else DSS2_EA_FWD(512)
#undef DSS2_EA_FWD
  DSS2_LAUNCH_CHECK();
  return 0;
}

extern "C" int dss2_edgeagg_bwd(const dss2_graph_t* g, const float* x, int64_t x_stride, int fn, const float* edge_attr,
                                int64_t ea_stride, int fe, const float* w1, const float* b1, const float* w2, const float* b2,
                                const float* grad_out, const float* skip_grad, int64_t skip_stride, float* grad_x, float* partials,
                                int64_t partial_stride, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (check_common("dss2_edgeagg_bwd", g, x, fn, edge_attr, fe, w1, b1, w2, b2)) return -1;
  DSS2_CHECK_ARG(grad_out && partials, "dss2_edgeagg_bwd: null argument");
  DSS2_CHECK_ARG(partial_stride >= (int64_t)HID * (2 * fn + fe) + HID + HID * HID + HID, "dss2_edgeagg_bwd: partial_stride too small");
  EaArgs a = {};
  a.g = *g;
  a.x = x;
  a.xs = x_stride;
  a.fn = fn;
  a.ea = edge_attr;
  a.eas = ea_stride;
  a.fe = fe;
  a.w1 = w1;
  a.b1 = b1;
  a.w2 = w2;
  a.b2 = b2;
  a.gout = grad_out;
  a.skip = skip_grad;
  a.skip_stride = skip_stride;
  a.gx = grad_x;
  a.partials = partials;
  a.partial_stride = partial_stride;
  DSS2_CHECK_ARG(g->undirected == 1, "dss2_edgeagg_bwd: needs the one-way edge list of the reference's data (reversed twins are derived, "
                 "not looked up); edge lists that already hold both directions are forward-only");
  if (g->num_tiles == 0) {   // large-graph path
    DSS2_NEED_SCRATCH(g, "dss2_edgeagg_bwd");
    const int64_t Nt = g->num_nodes;
    float* P = g->scratch;
    float* Q = P + (size_t)Nt * HID;
    float* GS = Q + (size_t)Nt * HID;
    k_ea_nodes_g<<<ea_gen_grid(Nt), EA_THREADS, 0, stream>>>(a, P, Q, GS);
    DSS2_LAUNCH_CHECK();
    constexpr int NTG = 512;   // 16 warps per SM: the row loop is latency bound (global gathers), one CTA per SM because of the per-CTA partial
    const size_t red_bytes = (size_t)(NTG / 32) * (3 * FP + 2 + HID) * HID * 4;
    DSS2_CUDA(cudaFuncSetAttribute(k_ea_bwd_g<NTG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)red_bytes));
    k_ea_bwd_g<NTG><<<dss2_sm_count(), NTG, red_bytes, stream>>>(a, P, Q, GS);
    DSS2_LAUNCH_CHECK();
    return 0;
  }
  if (ea_use_row(g, x_stride, ea_stride, fe, 1)) {
    const int slot = dss2_ea_row_scratch_slot();
    if (dss2_ea_row_upload(slot, 1, &w1, &b1, &w2, &b2, fn, fe, stream)) return -1;
    return dss2_ea_row_bwd(g, x, x_stride, fn, edge_attr, ea_stride, fe, slot, grad_out, skip_grad, skip_stride, grad_x, partials, partial_stride,
                           stream);
  }
  const int nt = ea_threads("DSS2_EA_BWD_THREADS", EA_BWD_THREADS);
  size_t smem = ea_smem(g, 4);
  size_t red = (size_t)(nt / 32) * (3 * FP + 2 + HID) * HID * 4;
  if (red > smem) smem = red;
  DSS2_CHECK_ARG(smem <= 227 * 1024, "dss2_edgeagg_bwd: tile needs %zu bytes of shared memory", smem);
#define DSS2_EA_BWD(NT)                                                                                                       \
  {                                                                                                                           \
    if (smem > 48 * 1024) DSS2_CUDA(cudaFuncSetAttribute(k_edgeagg_bwd<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_edgeagg_bwd<NT><<<dss2_sm_count(), NT, smem, stream>>>(a);                                                              \
  }
  if (nt == 256) DSS2_EA_BWD(256) else if (nt == 384) DSS2_EA_BWD(384) else DSS2_EA_BWD(512)
#undef DSS2_EA_BWD
  DSS2_LAUNCH_CHECK();
  return 0;
}

// Prepared-weights variant (throughput tier): dss2_edgeagg_upload moves the weights of up to 7 EdgeAggregation modules into
// constant-memory slots once per step; the _slot entry points then run the thread-per-row kernels without touching the weight tensors.
extern "C" int dss2_edgeagg_slots_ok(const dss2_graph_t* g, int64_t x_stride, int64_t ea_stride, int fe) {
  if (!g) return 0;
  return ea_use_row(g, x_stride, ea_stride, fe, 0) && ea_use_row(g, x_stride, ea_stride, fe, 1) ? dss2_ea_row_scratch_slot() : 0;
}

extern "C" int dss2_edgeagg_upload(int slot0, int n, const float* const* w1, const float* const* b1, const float* const* w2,
                                   const float* const* b2, int fn, int fe, void* stream_) {
  DSS2_CHECK_ARG(w1 && b1 && w2 && b2, "dss2_edgeagg_upload: null argument");
  DSS2_CHECK_ARG(slot0 >= 0 && n >= 1 && slot0 + n <= dss2_ea_row_scratch_slot(), "dss2_edgeagg_upload: slots %d..%d outside 0..%d", slot0,
                 slot0 + n - 1, dss2_ea_row_scratch_slot() - 1);
  return dss2_ea_row_upload(slot0, n, w1, b1, w2, b2, fn, fe, (cudaStream_t)stream_);
}

extern "C" int dss2_edgeagg_fwd_slot(const dss2_graph_t* g, const float* x, int64_t x_stride, int fn, const float* edge_attr, int64_t ea_stride,
                                     int fe, int slot, float* out, void* stream_) {
  DSS2_CHECK_ARG(g && x && edge_attr && out, "dss2_edgeagg_fwd_slot: null argument");
  DSS2_CHECK_ARG(slot >= 0 && slot < dss2_ea_row_scratch_slot(), "dss2_edgeagg_fwd_slot: slot %d", slot);
  DSS2_CHECK_ARG(fn >= 1 && fn <= FP && fe >= 1, "dss2_edgeagg_fwd_slot: feature counts");
  if (g->num_nodes == 0) return 0;
  DSS2_CHECK_ARG(dss2_ea_row_fits(g, x_stride, ea_stride, fe, 0), "dss2_edgeagg_fwd_slot: batch does not fit the thread-per-row kernels "
                 "(check dss2_edgeagg_slots_ok and use dss2_edgeagg_fwd)");
  return dss2_ea_row_fwd(g, x, x_stride, fn, edge_attr, ea_stride, fe, slot, out, (cudaStream_t)stream_);
}

extern "C" int dss2_edgeagg_bwd_slot(const dss2_graph_t* g, const float* x, int64_t x_stride, int fn, const float* edge_attr, int64_t ea_stride,
                                     int fe, int slot, const float* grad_out, const float* skip_grad, int64_t skip_stride, float* grad_x,
                                     float* partials, int64_t partial_stride, void* stream_) {
  DSS2_CHECK_ARG(g && x && edge_attr && grad_out && partials, "dss2_edgeagg_bwd_slot: null argument");
  DSS2_CHECK_ARG(slot >= 0 && slot < dss2_ea_row_scratch_slot(), "dss2_edgeagg_bwd_slot: slot %d", slot);
  DSS2_CHECK_ARG(fn >= 1 && fn <= FP && fe >= 1, "dss2_edgeagg_bwd_slot: feature counts");
  DSS2_CHECK_ARG(partial_stride >= (int64_t)HID * (2 * fn + fe) + HID + HID * HID + HID, "dss2_edgeagg_bwd_slot: partial_stride too small");
  DSS2_CHECK_ARG(g->undirected == 1, "dss2_edgeagg_bwd_slot: needs the one-way edge list of the reference's data");
  DSS2_CHECK_ARG(dss2_ea_row_fits(g, x_stride, ea_stride, fe, 1), "dss2_edgeagg_bwd_slot: batch does not fit the thread-per-row kernels "
                 "(check dss2_edgeagg_slots_ok and use dss2_edgeagg_bwd)");
  return dss2_ea_row_bwd(g, x, x_stride, fn, edge_attr, ea_stride, fe, slot, grad_out, skip_grad, skip_stride, grad_x, partials, partial_stride,
                         (cudaStream_t)stream_);
}
