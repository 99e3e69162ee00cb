// (b1) EdgeAggregation, thread-per-row kernels (round 2).  Same operator and the same algebra as edgeagg.cu (networks.py:159-209 and its
// autograd: first Linear split by operand, second Linear applied once per bus to the sum of hidden activations, CSR walk in PyG
// scatter order, reversed attributes by sign flip, no atomics), laid out for the FMA pipe instead of for one warp per bus:
//
//   * thread = bus row (forward) or (bus row, half of the 32 hidden units) (backward); a row's 16-32 accumulators live in registers and
//     every dense product is a chain of packed `fma.rn.f32x2` (FFMA2): a scalar of the row times a PAIR of adjacent weights.  The
//     weight pairs come from constant memory through the uniform datapath (one LDCU.128 feeds two FFMA2): measured in isolation
//     (tools/exp/ffma2_bench.cu) that form runs at 125 of the SM's 128 fp32 FMA lanes per clock, the scalar FFMA form at 100.
//     Constant addresses must be IMMEDIATES for that (slot and half are template parameters): with a run-time slot index ptxas emits
//     one indexed LDCU.64 per FFMA2 and the same product runs 4.9x slower.
//   * the weights reach constant memory through `dss2_edgeagg_upload`: a small kernel writes the input-major layouts into a staging
//     struct and ONE device-to-device copy node moves the slots into constant memory - capturable, so a replayed CUDA graph picks up
//     the weights of the step.  The adjoint products (W2^T g, W1^T g) use the same layouts with vector x vector FFMA2 (even / odd
//     partial sums), so no transposed copy is needed: 7.4 KB per EdgeAggregation, 8 slots.
//   * the tile's inputs (x rows, edge attributes, CSR row pointers / columns / edge ids) arrive by 1-D bulk copies of the raw, 16-byte
//     aligned byte ranges one tile ahead (2-stage ring, one elected thread, mbarrier) - no thread issues a global load for them; the
//     backward's upstream gradient rows come through a SWIZZLE_128B tensor map.
//   * neighbour rows (Q, and in the backward P and grad_S) are gathered from 16-byte-chunk-swizzled shared-memory rows (chunk c of row r
//     at c ^ (r & 7)): a quarter warp of consecutive rows reads or writes 8 distinct bank groups.
//   * backward phase C (weight gradients = reductions over rows) is role-split over the 16 warps (two groups for the 32x32 weight, one for
//     the node blocks of the first Linear, one for its edge block), each role with <= 17 accumulators, so the reduction state no longer
//     competes with phase B for registers; ReLU gates of the in-edges travel from phase B to phase C as one bit per hidden unit.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int FP = 8;          // padded node feature count
constexpr int FE = 6;          // edge features of the row kernels (fe <= 6; wider edge attributes keep the warp-per-row kernels)
constexpr int EA_SLOTS = 8;    // constant-memory slots: 0..6 for prepared weights (dss2_edgeagg_upload), 7 = scratch of the pointer API
constexpr int SCRATCH_SLOT = EA_SLOTS - 1;

struct __align__(16) EaConst {
  float w1t[3 * FP][HID];   // first Linear, input-major: rows 0-7 x_dst block, 8-15 x_src block, 16-23 edge block (zero rows beyond fn / fe)
  float b1[HID];
  float w2t[HID][HID];      // second Linear, input-major: [h][o]
  float b2[HID];
};

__constant__ EaConst c_ea[EA_SLOTS];
__device__ EaConst g_ea_stage[EA_SLOTS];

struct EaUpload {
  const float* w1[EA_SLOTS];
  const float* b1[EA_SLOTS];
  const float* w2[EA_SLOTS];
  const float* b2[EA_SLOTS];
  int slot0, fn, fe;
};

__global__ void __launch_bounds__(256) k_ea_prep(EaUpload u) {
  const int l = blockIdx.x;
  EaConst& s = g_ea_stage[u.slot0 + l];
  const float *w1 = u.w1[l], *w2 = u.w2[l];
  const int fn = u.fn, fe = u.fe, ld = 2 * fn + fe;
  for (int i = threadIdx.x; i < 3 * FP * HID; i += blockDim.x) {
    const int row = i >> 5, h = i & 31, blk = row >> 3, c = row & 7;
    const int lim = blk == 2 ? fe : fn;
    const int col = blk == 0 ? c : (blk == 1 ? fn + c : 2 * fn + c);
    s.w1t[row][h] = c < lim ? w1[h * ld + col] : 0.0f;
  }
  for (int i = threadIdx.x; i < HID * HID; i += blockDim.x) s.w2t[i & 31][i >> 5] = w2[i];
  if (threadIdx.x < HID) {
    s.b1[threadIdx.x] = u.b1[l][threadIdx.x];
    s.b2[threadIdx.x] = u.b2[l][threadIdx.x];
  }
}

struct EaRowArgs {
  dss2_graph_t g;
  const float* x;
  int xs;
  int fn;
  const float* ea;
  int eas;
  int fe;
  float* out;            // fwd
  const float* skip;     // bwd
  int64_t skip_stride;
  float* gx;
  float* partials;
  int64_t partial_stride;
};

// ---- packed fp32 helpers -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long& u64(float2& v) { return reinterpret_cast<unsigned long long&>(v); }
__device__ __forceinline__ void fma2(float2& d, float2 w, float s) {   // d += w * (s, s)
  float2 b = make_float2(s, s);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(u64(d)) : "l"(u64(w)), "l"(u64(b)));
}
__device__ __forceinline__ void fma2v(float2& d, float2 w, float2 v) {   // d += w * v, lane-wise
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(u64(d)) : "l"(u64(w)), "l"(u64(v)));
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(u64(d)) : "l"(u64(a)), "l"(u64(b)));
  return d;
}
__device__ __forceinline__ float2 cpair(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ int swz(int row, int chunk) { return row * HID + ((chunk ^ (row & 7)) << 2); }   // float offset of a 16-byte chunk
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void sts4(float* p, float2 a, float2 b) { *reinterpret_cast<float4*>(p) = make_float4(a.x, a.y, b.x, b.y); }

__host__ __device__ inline int round8(int v) { return (v + 7) & ~7; }
__host__ __device__ inline uint32_t round16u(uint32_t v) { return (v + 15u) & ~15u; }

// ---- input ring: one stage holds the raw byte ranges of a tile -------------------------------------------------------------------
struct StageLayout {
  uint32_t xraw, eraw, rowptr, col, eid, info, bytes;
};
__host__ __device__ inline StageLayout stage_layout(int TR, int ER, int Z, int xs, int eas) {
  StageLayout L;
  uint32_t o = 0;
  L.xraw = o;
  o += round16u((uint32_t)(TR * xs + 4) * 4u);
  L.eraw = o;
  o += round16u((uint32_t)(ER * eas + 4) * 4u);
  L.rowptr = o;
  o += round16u((uint32_t)(TR + 1 + 4) * 4u);
  L.col = o;
  o += round16u((uint32_t)(Z + 4) * 4u);
  L.eid = o;
  o += round16u((uint32_t)(Z + 4) * 4u);
  L.info = o;
  o += 64;
  L.bytes = (o + 127u) & ~127u;
  return L;
}
struct TileInfo {
  int n0, nT, z0, nZ;
  long long e0;
  int dx, de, dr, dz;   // element offsets of the tile's first x value / attribute / row pointer / CSR entry inside the aligned copies
};

// lands tile `r` in `stage` (elected thread): five bulk copies of 16-byte aligned supersets of the tile's ranges
__device__ __forceinline__ void issue_tile(const EaRowArgs& a, const TileRange& r, char* stage, const StageLayout& L, uint64_t* bar) {
  const dss2_graph_t& g = a.g;
  const int nT = r.n1 - r.n0, nE = (int)(r.e1 - r.e0), nZ = r.z1 - r.z0;
  const uintptr_t xa = (uintptr_t)(a.x + (size_t)r.n0 * a.xs), ea = (uintptr_t)(a.ea + (size_t)r.e0 * a.eas);
  const uintptr_t ra = (uintptr_t)(g.rowptr + r.n0), ca = (uintptr_t)(g.col + r.z0), ia = (uintptr_t)(g.eid + r.z0);
  TileInfo ti;
  ti.n0 = r.n0;
  ti.nT = nT;
  ti.z0 = r.z0;
  ti.nZ = nZ;
  ti.e0 = r.e0;
  ti.dx = (int)(xa & 15) >> 2;
  ti.de = (int)(ea & 15) >> 2;
  ti.dr = (int)(ra & 15) >> 2;
  ti.dz = (int)(ca & 15) >> 2;
  *reinterpret_cast<TileInfo*>(stage + L.info) = ti;
  const uint32_t xb = round16u((uint32_t)(ti.dx + nT * a.xs) * 4u), eb = nE ? round16u((uint32_t)(ti.de + nE * a.eas) * 4u) : 0u;
  const uint32_t rb = round16u((uint32_t)(ti.dr + nT + 1) * 4u), zb = nZ ? round16u((uint32_t)(ti.dz + nZ) * 4u) : 0u;
  mbar_expect_tx(bar, xb + eb + rb + 2u * zb);
  bulk_g2s(stage + L.xraw, reinterpret_cast<const void*>(xa & ~(uintptr_t)15), xb, bar);
  if (eb) bulk_g2s(stage + L.eraw, reinterpret_cast<const void*>(ea & ~(uintptr_t)15), eb, bar);
  bulk_g2s(stage + L.rowptr, reinterpret_cast<const void*>(ra & ~(uintptr_t)15), rb, bar);
  if (zb) {
    bulk_g2s(stage + L.col, reinterpret_cast<const void*>(ca & ~(uintptr_t)15), zb, bar);
    bulk_g2s(stage + L.eid, reinterpret_cast<const void*>(ia & ~(uintptr_t)15), zb, bar);
  }
}

// view of the current stage
struct TileView {
  const float* x;       // first x value of the tile (row stride a.xs)
  const float* ea;      // first attribute of the tile's first one-way edge (row stride a.eas)
  const int* rowptr;    // [nT + 1] global CSR offsets
  const int* col;       // [nZ] global source rows
  const uint32_t* eid;  // [nZ] global one-way edge id | reversed << 31
  int n0, nT, z0;
  long long e0;
};
__device__ __forceinline__ TileView view_stage(const char* stage, const StageLayout& L) {
  const TileInfo ti = *reinterpret_cast<const TileInfo*>(stage + L.info);
  TileView v;
  v.x = reinterpret_cast<const float*>(stage + L.xraw) + ti.dx;
  v.ea = reinterpret_cast<const float*>(stage + L.eraw) + ti.de;
  v.rowptr = reinterpret_cast<const int*>(stage + L.rowptr) + ti.dr;
  v.col = reinterpret_cast<const int*>(stage + L.col) + ti.dz;
  v.eid = reinterpret_cast<const uint32_t*>(stage + L.eid) + ti.dz;
  v.n0 = ti.n0;
  v.nT = ti.nT;
  v.z0 = ti.z0;
  v.e0 = ti.e0;
  return v;
}

__device__ __forceinline__ void load_x(const TileView& v, int row, int xs, int fn, float (&xv)[FP]) {
  const float* p = v.x + row * xs;
  if ((xs & 3) == 0 && (((uintptr_t)p) & 15) == 0) {
    const float4 a = lds4(p), b = lds4(p + 4);
    xv[0] = a.x, xv[1] = a.y, xv[2] = a.z, xv[3] = a.w, xv[4] = b.x, xv[5] = b.y, xv[6] = b.z, xv[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < FP; ++i) xv[i] = p[i];
  }
  if (fn < FP) {
#pragma unroll
    for (int i = 0; i < FP; ++i)
      if (i >= fn) xv[i] = 0.0f;
  }
}
// attributes of a (possibly reversed) doubled edge: columns 0 and 2 change sign (networks.py:252)
__device__ __forceinline__ void load_attr(const TileView& v, uint32_t id, int eas, int fe, float (&av)[FE]) {
  const float* p = v.ea + (int)((long long)(id & 0x7fffffffu) - v.e0) * eas;
#pragma unroll
  for (int i = 0; i < FE; ++i) av[i] = p[i];
  if (fe < FE) {
#pragma unroll
    for (int i = 0; i < FE; ++i)
      if (i >= fe) av[i] = 0.0f;
  }
  if (id >> 31) {
    av[0] = -av[0];
    av[2] = -av[2];
  }
}

// -------------------------------------------------------------------------------------------------
// forward: thread = bus row
// -------------------------------------------------------------------------------------------------
constexpr int FWD_THREADS = 256;

__host__ __device__ inline size_t fwd_smem_bytes(int TR, int ER, int Z, int xs, int eas) {
  return 1024 + (size_t)TR * HID * 4 + 2 * (size_t)stage_layout(TR, ER, Z, xs, eas).bytes + 64;
}

template <int SLOT>
__global__ void __launch_bounds__(FWD_THREADS, 2) k_ea_row_fwd(EaRowArgs a) {
  extern __shared__ __align__(16) char smem_raw[];
  const dss2_graph_t& g = a.g;
  const int TR = round8(g.max_tile_nodes);
  const StageLayout L = stage_layout(TR, round8(g.max_tile_edges), g.max_tile_nnz, a.xs, a.eas);
  char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the pointer in the shared window (LDS / STS)
  float* Qs = reinterpret_cast<float*>(base);
  char* stage0 = base + (size_t)TR * HID * 4;
  uint64_t* bar = reinterpret_cast<uint64_t*>(stage0 + 2 * (size_t)L.bytes);
  const EaConst& C = c_ea[SLOT];
  const int tid = threadIdx.x, ntiles = g.num_tiles, stride = gridDim.x;

  TileRange rn = {};
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
    if ((int)blockIdx.x < ntiles) issue_tile(a, tile_range(g, blockIdx.x), stage0, L, &bar[0]);
    if ((int)blockIdx.x + stride < ntiles) rn = tile_range(g, blockIdx.x + stride);
  }
  __syncthreads();

  int it = 0;
  for (int t = blockIdx.x; t < ntiles; t += stride, ++it) {
    const int s = it & 1;
    if (tid == 0 && t + stride < ntiles) {   // stage s^1 was last read before the barrier that closed the previous iteration
      issue_tile(a, rn, stage0 + (size_t)(s ^ 1) * L.bytes, L, &bar[s ^ 1]);
      if (t + 2 * stride < ntiles) rn = tile_range(g, t + 2 * stride);
    }
    mbar_wait(&bar[s], (uint32_t)(it >> 1) & 1u);
    const TileView v = view_stage(stage0 + (size_t)s * L.bytes, L);
    const int nT = v.nT;
    const bool active = tid < nT;
    float2 P[HID / 2];
    if (active) {
      float xv[FP];
      load_x(v, tid, a.xs, a.fn, xv);
      float2 Q[HID / 2];
#pragma unroll
      for (int j = 0; j < HID / 2; ++j) {
        P[j] = cpair(&C.b1[2 * j]);
        Q[j] = make_float2(0.0f, 0.0f);
      }
#pragma unroll
      for (int i = 0; i < FP; ++i) {
#pragma unroll
        for (int j = 0; j < HID / 2; ++j) {
          fma2(P[j], cpair(&C.w1t[i][2 * j]), xv[i]);
          fma2(Q[j], cpair(&C.w1t[FP + i][2 * j]), xv[i]);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) sts4(Qs + swz(tid, k), Q[2 * k], Q[2 * k + 1]);
    }
    __syncthreads();
    float2 O[HID / 2];
    if (active) {
      float2 S[HID / 2];
#pragma unroll
      for (int j = 0; j < HID / 2; ++j) S[j] = make_float2(0.0f, 0.0f);
      const int beg = v.rowptr[tid] - v.z0, end = v.rowptr[tid + 1] - v.z0;
      for (int z = beg; z < end; ++z) {
        const int c = v.col[z] - v.n0;
        float av[FE];
        load_attr(v, v.eid[z], a.eas, a.fe, av);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 q = lds4(Qs + swz(c, k));
          float2 e0 = make_float2(0.0f, 0.0f), e1 = make_float2(0.0f, 0.0f);
#pragma unroll
          for (int i = 0; i < FE; ++i) {
            fma2(e0, cpair(&C.w1t[2 * FP + i][4 * k]), av[i]);
            fma2(e1, cpair(&C.w1t[2 * FP + i][4 * k + 2]), av[i]);
          }
          // (P + Q) + edge term, rounded exactly like the backward re-evaluates it (same gates, same S)
          const float2 p0 = add2(add2(P[2 * k], make_float2(q.x, q.y)), e0), p1 = add2(add2(P[2 * k + 1], make_float2(q.z, q.w)), e1);
          S[2 * k] = add2(S[2 * k], make_float2(fmaxf(p0.x, 0.0f), fmaxf(p0.y, 0.0f)));
          S[2 * k + 1] = add2(S[2 * k + 1], make_float2(fmaxf(p1.x, 0.0f), fmaxf(p1.y, 0.0f)));
        }
      }
      const float deg = (float)(end - beg);
#pragma unroll
      for (int j = 0; j < HID / 2; ++j) {
        const float2 b = cpair(&C.b2[2 * j]);
        O[j] = make_float2(deg * b.x, deg * b.y);
      }
#pragma unroll
      for (int h = 0; h < HID; ++h) {
        const float sh = (h & 1) ? S[h >> 1].y : S[h >> 1].x;
#pragma unroll
        for (int j = 0; j < HID / 2; ++j) fma2(O[j], cpair(&C.w2t[h][2 * j]), sh);
      }
    }
    __syncthreads();   // every gather from Qs is done: the buffer becomes the output staging tile
    if (active) {
#pragma unroll
      for (int k = 0; k < 8; ++k) sts4(Qs + swz(tid, k), O[2 * k], O[2 * k + 1]);
    }
    __syncthreads();
    {
      float4* dst = reinterpret_cast<float4*>(a.out + (size_t)v.n0 * HID);
      for (int j = tid; j < nT * 8; j += FWD_THREADS) dst[j] = lds4(Qs + swz(j >> 3, j & 7));
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------------------
// backward: thread = (bus row, half of the hidden units); partial layout: w1 [32][ld], b1 [32], w2 [32][32], b2 [32]
// -------------------------------------------------------------------------------------------------
constexpr int NHALF = HID / 2;   // hidden units per thread
constexpr int BWD_THREADS = 512;
constexpr int BWD_WARPS = BWD_THREADS / 32;
constexpr int RED_SLOTS = 18;    // 8 packed accumulators + 1 packed scalar per lane and warp
struct BwdBufs {
  float *Pb, *Qb, *Gb, *Sb, *GO;   // [TR][32] swizzled: P -> grad_P, Q -> grad_Q, grad_S, S, upstream gradient
  float* gxp;                      // [TR][8] grad_x share of the upper half
  float* x8;                       // [TR][8] the tile's node features as aligned 32-byte rows (phase A1 has them in registers; the raw
                                   // stage rows are 44 bytes apart, so phase C's x roles would need 8 scalar loads per row)
  float* avz;                      // [Z][8] one record per CSR entry for phase C's edge role: the 6 attributes as phase B used them
                                   // (sign-flipped for reversed entries), the destination row (tile-local), the ReLU gate word of the
                                   // in-edge (bit h = pre-activation of hidden unit h is positive; one 16-bit half per thread half)
};

__host__ __device__ inline size_t bwd_smem_bytes(int TR, int ER, int Z, int xs, int eas) {
  size_t b = 1024 + 5 * (size_t)TR * HID * 4 + 2 * (size_t)TR * FP * 4 + (size_t)Z * 32 +
             2 * (size_t)stage_layout(TR, ER, Z, xs, eas).bytes + 64;
  const size_t red = 1024 + (size_t)BWD_WARPS * RED_SLOTS * HID * 4;
  return b > red ? b : red;
}

__device__ __forceinline__ void tma_load_rows32(void* smem_dst, const CUtensorMap* map, int row0, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(map), "r"(0), "r"(row0), "r"(smem_u32(bar))
               : "memory");
}

// phase A1: P, Q of the own half row -> shared memory
template <int SLOT, int H0>
__device__ __forceinline__ void bwd_phase_a1(const EaRowArgs& a, const TileView& v, const BwdBufs& b, int row) {
  const EaConst& C = c_ea[SLOT];
  constexpr int K0 = H0 / 4;
  float xv[FP];
  load_x(v, row, a.xs, a.fn, xv);
  if (H0 == 0) {
    sts4(b.x8 + row * FP, make_float2(xv[0], xv[1]), make_float2(xv[2], xv[3]));
    sts4(b.x8 + row * FP + 4, make_float2(xv[4], xv[5]), make_float2(xv[6], xv[7]));
  }
  float2 P[NHALF / 2], Q[NHALF / 2];
#pragma unroll
  for (int j = 0; j < NHALF / 2; ++j) {
    P[j] = cpair(&C.b1[H0 + 2 * j]);
    Q[j] = make_float2(0.0f, 0.0f);
  }
#pragma unroll
  for (int i = 0; i < FP; ++i) {
#pragma unroll
    for (int j = 0; j < NHALF / 2; ++j) {
      fma2(P[j], cpair(&C.w1t[i][H0 + 2 * j]), xv[i]);
      fma2(Q[j], cpair(&C.w1t[FP + i][H0 + 2 * j]), xv[i]);
    }
  }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    sts4(b.Pb + swz(row, K0 + kk), P[2 * kk], P[2 * kk + 1]);
    sts4(b.Qb + swz(row, K0 + kk), Q[2 * kk], Q[2 * kk + 1]);
  }
}

// phase A2: grad_S[h] = sum_o W2[o][h] g[o] for the own 16 hidden units: even / odd o in the two lanes of a packed accumulator
template <int SLOT, int H0>
__device__ __forceinline__ void bwd_phase_a2(const BwdBufs& b, int row) {
  const EaConst& C = c_ea[SLOT];
  constexpr int K0 = H0 / 4;
  float2 GS[NHALF / 2];
  float2 acc[NHALF];
#pragma unroll
  for (int h = 0; h < NHALF; ++h) acc[h] = make_float2(0.0f, 0.0f);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 go = lds4(b.GO + swz(row, k));
#pragma unroll
    for (int h = 0; h < NHALF; ++h) {
      fma2v(acc[h], cpair(&C.w2t[H0 + h][4 * k]), make_float2(go.x, go.y));
      fma2v(acc[h], cpair(&C.w2t[H0 + h][4 * k + 2]), make_float2(go.z, go.w));
    }
  }
#pragma unroll
  for (int j = 0; j < NHALF / 2; ++j) GS[j] = make_float2(acc[2 * j].x + acc[2 * j].y, acc[2 * j + 1].x + acc[2 * j + 1].y);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) sts4(b.Gb + swz(row, K0 + kk), GS[2 * kk], GS[2 * kk + 1]);
}

// phase B: every in-edge and its twin; S, grad_P, grad_Q of the own half row in registers (the row's own P, Q, grad_S chunks come back
// from shared memory per edge: holding them too does not fit 128 registers); then this half's share of grad_x
template <int SLOT, int H0>
__device__ __forceinline__ void bwd_phase_b(const EaRowArgs& a, const TileView& v, const BwdBufs& b, int row, float2 (&S)[NHALF / 2],
                                            float2 (&gP)[NHALF / 2], float2 (&gQ)[NHALF / 2], float (&gxh)[FP], bool want_gx) {
  const EaConst& C = c_ea[SLOT];
  constexpr int K0 = H0 / 4;
  const int beg = v.rowptr[row] - v.z0, end = v.rowptr[row + 1] - v.z0;
  for (int z = beg; z < end; ++z) {
    const int c = v.col[z] - v.n0;
    float av[FE];
    load_attr(v, v.eid[z], a.eas, a.fe, av);
    if (H0 == 0) {   // keep them for phase C (FE = 6: one 16-byte and one 8-byte store)
      static_assert(FE == 6, "avz layout");
      sts4(b.avz + 8 * z, make_float2(av[0], av[1]), make_float2(av[2], av[3]));
      *reinterpret_cast<float2*>(b.avz + 8 * z + 4) = make_float2(av[4], av[5]);
    }
    const float m0 = -2.0f * av[0], m2 = -2.0f * av[2];   // the twin's attributes: columns 0 and 2 flipped once more
    uint32_t gate = 0u;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float2 e0 = make_float2(0.0f, 0.0f), e1 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int i = 0; i < FE; ++i) {
        fma2(e0, cpair(&C.w1t[2 * FP + i][H0 + 4 * kk]), av[i]);
        fma2(e1, cpair(&C.w1t[2 * FP + i][H0 + 4 * kk + 2]), av[i]);
      }
      const float4 qc = lds4(b.Qb + swz(c, K0 + kk));
      const float4 pc = lds4(b.Pb + swz(c, K0 + kk));
      const float4 gc = lds4(b.Gb + swz(c, K0 + kk));
      const float4 qo = lds4(b.Qb + swz(row, K0 + kk)), po = lds4(b.Pb + swz(row, K0 + kk)), go = lds4(b.Gb + swz(row, K0 + kk));
      // in-edge (c -> row)
      const float2 i0 = add2(add2(make_float2(po.x, po.y), make_float2(qc.x, qc.y)), e0);
      const float2 i1 = add2(add2(make_float2(po.z, po.w), make_float2(qc.z, qc.w)), e1);
      S[2 * kk] = add2(S[2 * kk], make_float2(fmaxf(i0.x, 0.0f), fmaxf(i0.y, 0.0f)));
      S[2 * kk + 1] = add2(S[2 * kk + 1], make_float2(fmaxf(i1.x, 0.0f), fmaxf(i1.y, 0.0f)));
      const bool b0 = i0.x > 0.0f, b1 = i0.y > 0.0f, b2 = i1.x > 0.0f, b3 = i1.y > 0.0f;
      gate |= ((b0 ? 1u : 0u) | (b1 ? 2u : 0u) | (b2 ? 4u : 0u) | (b3 ? 8u : 0u)) << (4 * kk);
      gP[2 * kk] = add2(gP[2 * kk], make_float2(b0 ? go.x : 0.0f, b1 ? go.y : 0.0f));
      gP[2 * kk + 1] = add2(gP[2 * kk + 1], make_float2(b2 ? go.z : 0.0f, b3 ? go.w : 0.0f));
      // twin (row -> c)
      fma2(e0, cpair(&C.w1t[2 * FP + 0][H0 + 4 * kk]), m0);
      fma2(e1, cpair(&C.w1t[2 * FP + 0][H0 + 4 * kk + 2]), m0);
      fma2(e0, cpair(&C.w1t[2 * FP + 2][H0 + 4 * kk]), m2);
      fma2(e1, cpair(&C.w1t[2 * FP + 2][H0 + 4 * kk + 2]), m2);
      const float2 t0 = add2(add2(make_float2(pc.x, pc.y), make_float2(qo.x, qo.y)), e0);
      const float2 t1 = add2(add2(make_float2(pc.z, pc.w), make_float2(qo.z, qo.w)), e1);
      gQ[2 * kk] = add2(gQ[2 * kk], make_float2(t0.x > 0.0f ? gc.x : 0.0f, t0.y > 0.0f ? gc.y : 0.0f));
      gQ[2 * kk + 1] = add2(gQ[2 * kk + 1], make_float2(t1.x > 0.0f ? gc.z : 0.0f, t1.y > 0.0f ? gc.w : 0.0f));
    }
    // the entry's record for phase C: [6 attributes | destination row | gate word] = 32 bytes, read back with two vector loads
    reinterpret_cast<unsigned short*>(b.avz + 8 * z + 7)[H0 ? 1 : 0] = (unsigned short)gate;
    if (H0 == 0) reinterpret_cast<uint32_t*>(b.avz)[8 * z + 6] = (uint32_t)row;
  }
  // grad_x share of this half: sum_h W1a[h][i] gP[h] + W1b[h][i] gQ[h], pairs of hidden units in the two lanes
  if (want_gx) {
    float2 acc[FP];
#pragma unroll
    for (int i = 0; i < FP; ++i) acc[i] = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int j = 0; j < NHALF / 2; ++j) {
#pragma unroll
      for (int i = 0; i < FP; ++i) {
        fma2v(acc[i], cpair(&C.w1t[i][H0 + 2 * j]), gP[j]);
        fma2v(acc[i], cpair(&C.w1t[FP + i][H0 + 2 * j]), gQ[j]);
      }
    }
#pragma unroll
    for (int i = 0; i < FP; ++i) gxh[i] = acc[i].x + acc[i].y;
  }
}

template <int SLOT>
__global__ void __launch_bounds__(BWD_THREADS, 1) k_ea_row_bwd(EaRowArgs a, const __grid_constant__ CUtensorMap go_map) {
  extern __shared__ __align__(16) char smem_raw[];
  const dss2_graph_t& g = a.g;
  const int TR = round8(g.max_tile_nodes), Z = g.max_tile_nnz;
  const StageLayout L = stage_layout(TR, round8(g.max_tile_edges), Z, a.xs, a.eas);
  char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the pointer in the shared window (LDS / STS)
  const size_t BUF = (size_t)TR * HID * 4;
  BwdBufs b;
  b.GO = reinterpret_cast<float*>(base);   // 1024-byte aligned: the tensor map's 128-byte swizzle equals swz()
  b.Pb = reinterpret_cast<float*>(base + BUF);
  b.Qb = reinterpret_cast<float*>(base + 2 * BUF);
  b.Gb = reinterpret_cast<float*>(base + 3 * BUF);
  b.Sb = reinterpret_cast<float*>(base + 4 * BUF);
  b.gxp = reinterpret_cast<float*>(base + 5 * BUF);
  b.x8 = b.gxp + (size_t)TR * FP;
  b.avz = b.x8 + (size_t)TR * FP;
  char* stage0 = reinterpret_cast<char*>(b.avz) + (size_t)Z * 32;
  uint64_t* bar = reinterpret_cast<uint64_t*>(stage0 + 2 * (size_t)L.bytes);   // [0], [1]: input stages, [2]: upstream gradient tile
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, ntiles = g.num_tiles, stride = gridDim.x;
  const int row = tid & 255, half = tid >> 8, K0 = half * 4;   // the half is warp-uniform
  const int fn = a.fn, fe = a.fe, ld = 2 * fn + fe;
  const bool want_gx = a.gx != nullptr;
  const uint32_t go_bytes = (uint32_t)BUF;

  // phase C state: 8 packed accumulators + 1 packed scalar, meaning by role.  A warp handles TWO rows (or CSR entries) per
  // instruction: lanes 0-15 one, lanes 16-31 the other, so every role instruction of phase C covers two rows.
  //   warps 0-7  : W2 rows 8q .. 8q+7 (q = warp >> 1), lane = (block of 4 output rows, block of 4 hidden units), row pairs j = warp & 1 mod 2
  //   warps 8-9  : first Linear, x_dst block + b1, lane = pair of hidden units          warps 10-11: x_src block + b2 (lane = pair of outputs)
  //   warps 12-15: first Linear, edge block: gated grad_S of every CSR entry times its attributes, lane = pair of hidden units
  const int rsel = lane >> 4, l16 = lane & 15;
  float2 acc[8], accp = make_float2(0.0f, 0.0f);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(0.0f, 0.0f);

  TileRange rn = {};
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&bar[2], 1);
    fence_mbar_init();
    if ((int)blockIdx.x < ntiles) {
      const TileRange r0 = tile_range(g, blockIdx.x);
      issue_tile(a, r0, stage0, L, &bar[0]);
      mbar_expect_tx(&bar[2], go_bytes);
      tma_load_rows32(b.GO, &go_map, r0.n0, &bar[2]);
    }
    if ((int)blockIdx.x + stride < ntiles) rn = tile_range(g, blockIdx.x + stride);
  }
  __syncthreads();

  int it = 0;
  for (int t = blockIdx.x; t < ntiles; t += stride, ++it) {
    const int s = it & 1;
    int next_n0 = -1;
    if (tid == 0 && t + stride < ntiles) {
      issue_tile(a, rn, stage0 + (size_t)(s ^ 1) * L.bytes, L, &bar[s ^ 1]);
      next_n0 = rn.n0;
      if (t + 2 * stride < ntiles) rn = tile_range(g, t + 2 * stride);
    }
    mbar_wait(&bar[s], (uint32_t)(it >> 1) & 1u);
    const TileView v = view_stage(stage0 + (size_t)s * L.bytes, L);
    const int nT = v.nT;
    const bool active = row < nT;
    if (active) {
      if (half == 0) bwd_phase_a1<SLOT, 0>(a, v, b, row);
      else bwd_phase_a1<SLOT, NHALF>(a, v, b, row);
    }
    mbar_wait(&bar[2], (uint32_t)it & 1u);
    if (active) {
      if (half == 0) bwd_phase_a2<SLOT, 0>(b, row);
      else bwd_phase_a2<SLOT, NHALF>(b, row);
    }
    __syncthreads();
    float2 S[NHALF / 2], gP[NHALF / 2], gQ[NHALF / 2];
    float gxh[FP];
#pragma unroll
    for (int j = 0; j < NHALF / 2; ++j) S[j] = gP[j] = gQ[j] = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < FP; ++i) gxh[i] = 0.0f;
    if (active) {
      if (half == 0) bwd_phase_b<SLOT, 0>(a, v, b, row, S, gP, gQ, gxh, want_gx);
      else bwd_phase_b<SLOT, NHALF>(a, v, b, row, S, gP, gQ, gxh, want_gx);
    }
    __syncthreads();   // every gather from P / Q is done: those buffers take grad_P, grad_Q for phase C
    if (active) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        sts4(b.Pb + swz(row, K0 + kk), gP[2 * kk], gP[2 * kk + 1]);
        sts4(b.Qb + swz(row, K0 + kk), gQ[2 * kk], gQ[2 * kk + 1]);
        sts4(b.Sb + swz(row, K0 + kk), S[2 * kk], S[2 * kk + 1]);
      }
      if (half == 1 && want_gx) {
        *reinterpret_cast<float4*>(b.gxp + row * FP) = make_float4(gxh[0], gxh[1], gxh[2], gxh[3]);
        *reinterpret_cast<float4*>(b.gxp + row * FP + 4) = make_float4(gxh[4], gxh[5], gxh[6], gxh[7]);
      }
    }
    __syncthreads();
    if (active && half == 0 && want_gx) {
      const float4 u0 = lds4(b.gxp + row * FP), u1 = lds4(b.gxp + row * FP + 4);
      float o[FP] = {gxh[0] + u0.x, gxh[1] + u0.y, gxh[2] + u0.z, gxh[3] + u0.w, gxh[4] + u1.x, gxh[5] + u1.y, gxh[6] + u1.z, gxh[7] + u1.w};
      const size_t n = (size_t)v.n0 + row;
      if (a.skip) {
#pragma unroll
        for (int i = 0; i < FP; ++i)
          if (i < fn) o[i] += a.skip[n * a.skip_stride + i];
      }
      if (fn == FP) {
        *reinterpret_cast<float4*>(a.gx + n * FP) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(a.gx + n * FP + 4) = make_float4(o[4], o[5], o[6], o[7]);
      } else {
#pragma unroll
        for (int i = 0; i < FP; ++i)
          if (i < fn) a.gx[n * fn + i] = o[i];
      }
    }
    // ---- phase C: weight gradients (reductions over the tile's rows), roles over warps, two rows per instruction
    if (warp < 8) {
      const int q = warp >> 1, ob = l16 >> 3, hb = l16 & 7;
      for (int rr = 2 * (warp & 1) + rsel; rr < nT; rr += 4) {
        const float4 go = lds4(b.GO + swz(rr, 2 * q + ob)), s4 = lds4(b.Sb + swz(rr, hb));
        const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          fma2(acc[hh], make_float2(go.x, go.y), sv[hh]);
          fma2(acc[4 + hh], make_float2(go.z, go.w), sv[hh]);
        }
      }
    } else if (warp < 12) {
      const bool src_blk = warp >= 10;
      const float* G = src_blk ? b.Qb : b.Pb;
      for (int rr = 2 * (warp & 1) + rsel; rr < nT; rr += 4) {
        const int off = rr * HID + (((l16 >> 1) ^ (rr & 7)) << 2) + ((l16 & 1) << 1);   // elements 2 l16, 2 l16 + 1 of a swizzled row
        const float2 gpair = *reinterpret_cast<const float2*>(G + off);
        const float4 x0 = lds4(b.x8 + rr * FP), x1 = lds4(b.x8 + rr * FP + 4);
        const float xv[FP] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int i = 0; i < FP; ++i) fma2(acc[i], gpair, xv[i]);
        if (src_blk) {   // b2[o] += deg * g[o]
          const float deg = (float)(v.rowptr[rr + 1] - v.rowptr[rr]);
          fma2(accp, *reinterpret_cast<const float2*>(b.GO + off), deg);
        } else {         // b1[h] += grad_P[h]
          accp = add2(accp, gpair);
        }
      }
    } else {
      const int nZ = v.rowptr[nT] - v.z0;
      for (int z = 2 * (warp & 3) + rsel; z < nZ; z += 8) {
        const float4 a0 = lds4(b.avz + 8 * z), a1 = lds4(b.avz + 8 * z + 4);
        const int rr = (int)__float_as_uint(a1.z);
        const uint32_t gate = __float_as_uint(a1.w) >> (2 * l16);
        const float2 gs = *reinterpret_cast<const float2*>(b.Gb + rr * HID + (((l16 >> 1) ^ (rr & 7)) << 2) + ((l16 & 1) << 1));
        const float2 ge = make_float2((gate & 1u) ? gs.x : 0.0f, (gate & 2u) ? gs.y : 0.0f);
        const float av[FE] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y};
#pragma unroll
        for (int i = 0; i < FE; ++i) fma2(acc[i], ge, av[i]);
      }
    }
    __syncthreads();
    if (next_n0 >= 0) {   // the upstream gradient tile is single-buffered: refill it for the next tile now that phase C has read it
      mbar_expect_tx(&bar[2], go_bytes);
      tma_load_rows32(b.GO, &go_map, next_n0, &bar[2]);
    }
  }

  // per-CTA partial: the warps' accumulators through shared memory, every entry summed over its role's warps and both half warps
  __syncthreads();
  float* red = reinterpret_cast<float*>(base);   // [BWD_WARPS][RED_SLOTS][32]
  {
    float* mine = red + (size_t)warp * RED_SLOTS * HID;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mine[(2 * i) * HID + lane] = acc[i].x;
      mine[(2 * i + 1) * HID + lane] = acc[i].y;
    }
    mine[16 * HID + lane] = accp.x;
    mine[17 * HID + lane] = accp.y;
  }
  __syncthreads();
  float* part = a.partials + (size_t)blockIdx.x * a.partial_stride;
  const int n_w1 = HID * ld, off_b1 = n_w1, off_w2 = n_w1 + HID, off_b2 = off_w2 + HID * HID, total = off_b2 + HID;
  for (int i = tid; i < total; i += BWD_THREADS) {
    int w0, nw, slot, ln;   // warps w0 .. w0+nw-1, lanes ln and ln + 16
    if (i < n_w1) {
      const int h = i / ld, c = i - h * ld;
      ln = h >> 1;
      if (c < fn) {
        w0 = 8, nw = 2, slot = 2 * c + (h & 1);
      } else if (c < 2 * fn) {
        w0 = 10, nw = 2, slot = 2 * (c - fn) + (h & 1);
      } else {
        w0 = 12, nw = 4, slot = 2 * (c - 2 * fn) + (h & 1);
      }
    } else if (i < off_w2) {
      const int h = i - off_b1;
      w0 = 8, nw = 2, slot = 16 + (h & 1), ln = h >> 1;
    } else if (i < off_b2) {
      const int o = (i - off_w2) / HID, h = (i - off_w2) - o * HID, oo = o & 3;
      w0 = 2 * (o >> 3), nw = 2;
      ln = (((o & 7) >> 2) << 3) + (h >> 2);
      slot = ((oo >> 1) * 4 + (h & 3)) * 2 + (oo & 1);
    } else {
      const int o = i - off_b2;
      w0 = 10, nw = 2, slot = 16 + (o & 1), ln = o >> 1;
    }
    float sum = 0.0f;
    for (int w = w0; w < w0 + nw; ++w) {
      const float* r = red + ((size_t)w * RED_SLOTS + slot) * HID;
      sum += r[ln];
      sum += r[ln + 16];
    }
    part[i] = sum;
  }
}

// ---- host side ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
// [rows, 32] fp32 row-major, box = [box_rows, 32], 128-byte swizzle, out-of-range rows read as zeros
int make_row_map(CUtensorMap* m, const float* base, int64_t rows, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1;
  const cuuint64_t dims[2] = {32, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0
             : -1;
}

template <int SLOT>
int launch_fwd(const EaRowArgs& a, int grid, size_t smem, cudaStream_t stream) {
  DSS2_CUDA(cudaFuncSetAttribute(k_ea_row_fwd<SLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_ea_row_fwd<SLOT><<<grid, FWD_THREADS, smem, stream>>>(a);
  DSS2_LAUNCH_CHECK();
  return 0;
}
template <int SLOT>
int launch_bwd(const EaRowArgs& a, const CUtensorMap& map, int grid, size_t smem, cudaStream_t stream) {
  DSS2_CUDA(cudaFuncSetAttribute(k_ea_row_bwd<SLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(BWD_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  const int prio = dss2_launch_priority(2);   // on the step's critical path; the weight-gradient passes of the TAG layers fill in behind
  if (prio != 0) {
    attr[0].id = cudaLaunchAttributePriority;
    attr[0].val.priority = prio;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  DSS2_CUDA(cudaLaunchKernelEx(&cfg, k_ea_row_bwd<SLOT>, a, map));
  DSS2_LAUNCH_CHECK();
  return 0;
}

#define DSS2_EA_SLOT_SWITCH(slot, CALL) \
  switch (slot) {                       \
    case 0: return CALL(0);             \
    case 1: return CALL(1);             \
    case 2: return CALL(2);             \
    case 3: return CALL(3);             \
    case 4: return CALL(4);             \
    case 5: return CALL(5);             \
    case 6: return CALL(6);             \
    default: return CALL(7);            \
  }

}  // namespace

// 1 = the batch structure, strides and feature counts fit the thread-per-row kernels (else: the warp-per-row kernels of edgeagg.cu)
int dss2_ea_row_fits(const dss2_graph_t* g, int64_t x_stride, int64_t ea_stride, int fe, int bwd) {
  if (g->num_tiles == 0 || g->max_tile_nodes > 256 || fe > FE || x_stride > 16 || ea_stride > 16) return 0;
  if (bwd && !encode_fn()) return 0;
  const int TR = round8(g->max_tile_nodes), ER = round8(g->max_tile_edges);
  const size_t b = bwd ? bwd_smem_bytes(TR, ER, g->max_tile_nnz, (int)x_stride, (int)ea_stride)
                       : fwd_smem_bytes(TR, ER, g->max_tile_nnz, (int)x_stride, (int)ea_stride);
  return b <= (bwd ? 227u : 113u) * 1024u;
}

// weights of `n` EdgeAggregation modules -> constant-memory slots slot0 .. slot0+n-1 (two capturable nodes: layout kernel, D2D copy)
int dss2_ea_row_upload(int slot0, int n, const float* const* w1, const float* const* b1, const float* const* w2, const float* const* b2, int fn,
                       int fe, cudaStream_t stream) {
  DSS2_CHECK_ARG(n >= 1 && slot0 >= 0 && slot0 + n <= EA_SLOTS, "dss2_edgeagg_upload: slots %d..%d outside 0..%d", slot0, slot0 + n - 1, EA_SLOTS - 1);
  DSS2_CHECK_ARG(fn >= 1 && fn <= FP && fe >= 1 && fe <= FE, "dss2_edgeagg_upload: feature counts (%d node, %d edge) outside 1..%d / 1..%d", fn, fe,
                 FP, FE);
  EaUpload u = {};
  for (int i = 0; i < n; ++i) {
    DSS2_CHECK_ARG(w1[i] && b1[i] && w2[i] && b2[i], "dss2_edgeagg_upload: null weight pointer");
    u.w1[i] = w1[i];
    u.b1[i] = b1[i];
    u.w2[i] = w2[i];
    u.b2[i] = b2[i];
  }
  u.slot0 = slot0;
  u.fn = fn;
  u.fe = fe;
  k_ea_prep<<<n, 256, 0, stream>>>(u);
  DSS2_LAUNCH_CHECK();
  static void* stage_base = nullptr;
  if (!stage_base) DSS2_CUDA(cudaGetSymbolAddress(&stage_base, g_ea_stage));
  DSS2_CUDA(cudaMemcpyToSymbolAsync(c_ea, (const char*)stage_base + (size_t)slot0 * sizeof(EaConst), (size_t)n * sizeof(EaConst),
                                    (size_t)slot0 * sizeof(EaConst), cudaMemcpyDeviceToDevice, stream));
  return 0;
}

int dss2_ea_row_scratch_slot() { return SCRATCH_SLOT; }

int dss2_ea_row_fwd(const dss2_graph_t* g, const float* x, int64_t x_stride, int fn, const float* edge_attr, int64_t ea_stride, int fe, int slot,
                    float* out, cudaStream_t stream) {
  EaRowArgs a = {};
  a.g = *g;
  a.x = x;
  a.xs = (int)x_stride;
  a.fn = fn;
  a.ea = edge_attr;
  a.eas = (int)ea_stride;
  a.fe = fe;
  a.out = out;
  const size_t smem = fwd_smem_bytes(round8(g->max_tile_nodes), round8(g->max_tile_edges), g->max_tile_nnz, a.xs, a.eas);
  const int grid = max(1, min(g->num_tiles, 2 * dss2_sm_count()));
#define CALL(S) launch_fwd<S>(a, grid, smem, stream)
  DSS2_EA_SLOT_SWITCH(slot, CALL)
#undef CALL
}

int dss2_ea_row_bwd(const dss2_graph_t* g, const float* x, int64_t x_stride, int fn, const float* edge_attr, int64_t ea_stride, int fe, int slot,
                    const float* grad_out, const float* skip_grad, int64_t skip_stride, float* grad_x, float* partials, int64_t partial_stride,
                    cudaStream_t stream) {
  EaRowArgs a = {};
  a.g = *g;
  a.x = x;
  a.xs = (int)x_stride;
  a.fn = fn;
  a.ea = edge_attr;
  a.eas = (int)ea_stride;
  a.fe = fe;
  a.skip = skip_grad;
  a.skip_stride = skip_stride;
  a.gx = grad_x;
  a.partials = partials;
  a.partial_stride = partial_stride;
  const int TR = round8(g->max_tile_nodes);
  CUtensorMap map;
  DSS2_CHECK_ARG(((uintptr_t)grad_out & 15) == 0 && make_row_map(&map, grad_out, g->num_nodes, TR) == 0,
                 "dss2_edgeagg_bwd: tensor map for grad_out (needs a 16-byte aligned [Nt,32] array)");
  const size_t smem = bwd_smem_bytes(TR, round8(g->max_tile_edges), g->max_tile_nnz, a.xs, a.eas);
  const int grid = dss2_sm_count();   // = dss2_num_partials(): every partial row is written (idle CTAs write zeros)
#define CALL(S) launch_bwd<S>(a, map, grid, smem, stream)
  DSS2_EA_SLOT_SWITCH(slot, CALL)
#undef CALL
}
