// (b2) TAG layer forward on 5th-generation tensor cores: the (K+1) 32x32 feature transforms run as 3xTF32 tcgen05.mma
// with fp32 accumulators in TMEM; hops, operand splitting and the fused epilogue stay on CUDA cores.
//
// Why: the CUDA-core kernel (tag.cu) is instruction-issue bound - ncu shows 324 warp-instructions per node row of which only
// ~127 are the transform itself (profiles/r1b_tag_fwd_ncu_summary.txt).  Here the transform costs 36 MMA instructions per
// 128-row block issued by ONE thread (18 k MACs each), and the epilogue is thread-per-row straight out of TMEM
// (tcgen05.ld 32x32b: lane = node row, 32 registers = the row's 32 output features), which removes the lane-per-feature
// epilogue, the ballot and the broadcast LDS traffic.
//
// Per tile (<= 256 node rows = whole graphs):
//   1. global -> registers -> split hi/lo -> two swizzled K-major operand tiles per hop level (tc.cuh)
//   2. hops in shared memory on hi+lo (exact fp32), results split again
//   3. fence.proxy.async; one thread issues 2 x 36 MMAs (two 128-row blocks) and commits each block to an mbarrier
//   4. warps 0-3 / 4-7 wait for block 0 / 1, tcgen05.ld their rows, bias + dropout + ReLU + sign word + residual, 128-bit stores
#include "common.cuh"
#include "tc.cuh"

namespace {

constexpr int TC_THREADS = 256;
constexpr int TC_WARPS = TC_THREADS / 32;
constexpr int TRB = 256;                       // rows per operand tile (2 MMA blocks of 128)
constexpr uint32_t A_TILE = TRB * tc::ROW_BYTES;   // 32 KB
constexpr uint32_t B_TILE = 32 * tc::ROW_BYTES;    // 4 KB

struct TagTcArgs {
  dss2_graph_t g;
  const float* x;
  const float* w;
  const float* bias;
  int cout;
  int act;
  float scale;
  uint32_t keep_thr16;   // keep iff 16-bit uniform < keep_thr16
  int drop_mode;
  const uint64_t* rng;
  uint32_t layer_uid;
  const uint8_t* mask;
  const float* res;
  int64_t res_stride;
  float* y;
  uint32_t* bits;
};

struct TopoTc {
  int* rowptr;
  int2* cw;   // .x = row code of the source row (row*128 | (row&7)<<4), .y = gcn weight bits
};

__device__ __forceinline__ void load_topo_tc(const dss2_graph_t& g, const TileRange& r, TopoTc s, int tid, int nthreads) {
  const int nT = r.n1 - r.n0, nZ = r.z1 - r.z0;
  for (int i = tid; i <= nT; i += nthreads) s.rowptr[i] = g.rowptr[r.n0 + i] - r.z0;
  for (int i = tid; i < nZ; i += nthreads)
    s.cw[i] = make_int2((int)tc::row_code((uint32_t)(g.col[r.z0 + i] - r.n0)), __float_as_int(g.w[r.z0 + i]));
}

// acc[i] += sum_e w_e * (hi + lo)[src_e][lane] for R consecutive rows; same visiting order as hop_rows in tag.cu
template <int R>
__device__ __forceinline__ void hop_rows_tc(const TopoTc& s, const char* hi, const char* lo, int r0, int nT, uint32_t lane4, float (&acc)[R]) {
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
        const uint32_t off = (uint32_t)cw.x ^ lane4;
        const float xv = *reinterpret_cast<const float*>(hi + off) + *reinterpret_cast<const float*>(lo + off);
        acc[i] = fmaf(__int_as_float(cw.y), xv, acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    for (int z = beg[i] + 4; z < beg[i] + deg[i]; ++z) {
      const int2 cw = s.cw[z];
      const uint32_t off = (uint32_t)cw.x ^ lane4;
      const float xv = *reinterpret_cast<const float*>(hi + off) + *reinterpret_cast<const float*>(lo + off);
      acc[i] = fmaf(__int_as_float(cw.y), xv, acc[i]);
    }
  }
}

// 32 Bernoulli(keep) decisions for one node row: 4 Philox4x32-10 calls -> 32 sixteen-bit uniforms
__device__ __forceinline__ uint32_t keep_word(uint2 key, uint32_t tile, uint32_t row, uint32_t step_lo, uint32_t thr16) {
  uint32_t word = 0;
#pragma unroll
  for (uint32_t q = 0; q < 4; ++q) {
    const uint4 r = philox4x32_10(make_uint4(tile, row, q, step_lo), key);
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
      word |= ((u[i] & 0xffffu) < thr16 ? 1u : 0u) << (q * 8 + 2 * i);
      word |= ((u[i] >> 16) < thr16 ? 1u : 0u) << (q * 8 + 2 * i + 1);
    }
  }
  return word;
}

inline int round_up4(int v) { return (v + 3) & ~3; }

__device__ __forceinline__ char* align1024(char* p) {
  const uint32_t a = smem_u32(p);
  return p + (((a + 1023u) & ~1023u) - a);
}

// -------------------------------------------------------------------------------------------------
// self test: D[128,32] = A[128,32] * B[32,32]^T through the exact operand / descriptor / TMEM path of the layer kernel
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_tc_selftest(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ char raw[];
  char* base = align1024(raw);
  char *a_hi = base, *a_lo = base + 16384, *b_hi = base + 32768, *b_lo = base + 36864;
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 40960);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(base + 40976);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int idx = tid; idx < 128 * 8; idx += 128) {
    const uint32_t row = idx >> 3, ch = idx & 7;
    tc::split_store4(*reinterpret_cast<const float4*>(A + row * 32 + ch * 4), a_hi, a_lo, row * 128 + ((ch ^ (row & 7)) << 4));
  }
  for (int idx = tid; idx < 32 * 8; idx += 128) {
    const uint32_t row = idx >> 3, ch = idx & 7;
    tc::split_store4(*reinterpret_cast<const float4*>(B + row * 32 + ch * 4), b_hi, b_lo, row * 128 + ((ch ^ (row & 7)) << 4));
  }
  if (warp == 0) tc::tmem_alloc(tslot, 32);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tslot;
  if (tid == 0) {
    tc::issue_block(tmem, smem_u32(a_hi), smem_u32(a_lo), smem_u32(b_hi), smem_u32(b_lo), tc::idesc_tf32(128, 32), true);
    tc::mma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc::fence_after_sync();
  float v[32];
  tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  const int row = warp * 32 + lane;
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4)
    *reinterpret_cast<float4*>(D + row * 32 + 4 * c4) = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 32);
}

// self test 2: MN-major operands (SWIZZLE_128B_BASE32B): D[32*t + j, n] = sum_{r < 64} A[t][r][j] * B[r][n], A = 4 tiles [64][32],
// B = 1 tile [64][32] - the shape of the weight-gradient GEMM (grad_W = G^T X, contraction over node rows).
constexpr int ST_KR = 64;
__global__ void __launch_bounds__(128, 1) k_tc_selftest_mn(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ char raw[];
  char* base = align1024(raw);
  const uint32_t TILE = ST_KR * 128;                      // 8 KB
  char* a_hi = base;                                      // 4 tiles
  char* a_lo = base + 4 * TILE;
  char* b_hi = base + 8 * TILE;
  char* b_lo = base + 9 * TILE;
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 10 * TILE);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(base + 10 * TILE + 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int idx = tid; idx < 4 * ST_KR * 8; idx += 128) {   // A given as [4][KR][32]; idx = 16-byte chunk
    const uint32_t tile = idx / (ST_KR * 8), rem = idx % (ST_KR * 8), row = rem >> 3, ch = rem & 7;
    tc::split_store4(*reinterpret_cast<const float4*>(A + (size_t)idx * 4), a_hi + tile * TILE, a_lo + tile * TILE, tc::swz32_off(row, ch * 4));
  }
  for (int idx = tid; idx < ST_KR * 8; idx += 128) {
    const uint32_t row = idx >> 3, ch = idx & 7;
    tc::split_store4(*reinterpret_cast<const float4*>(B + (size_t)idx * 4), b_hi, b_lo, tc::swz32_off(row, ch * 4));
  }
  if (warp == 0) tc::tmem_alloc(tslot, 32);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tslot;
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_tf32(128, 32, 1, 1);
    for (uint32_t ks = 0; ks < ST_KR / 8; ++ks) {
      const uint32_t o = ks * 1024;
      tc::mma_tf32(tmem, tc::smem_desc_mn32(smem_u32(a_lo) + o, TILE), tc::smem_desc_mn32(smem_u32(b_hi) + o, TILE), idesc, ks ? 1u : 0u);
      tc::mma_tf32(tmem, tc::smem_desc_mn32(smem_u32(a_hi) + o, TILE), tc::smem_desc_mn32(smem_u32(b_lo) + o, TILE), idesc, 1u);
      tc::mma_tf32(tmem, tc::smem_desc_mn32(smem_u32(a_hi) + o, TILE), tc::smem_desc_mn32(smem_u32(b_hi) + o, TILE), idesc, 1u);
    }
    tc::mma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc::fence_after_sync();
  float v[32];
  tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  const int row = warp * 32 + lane;
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4)
    *reinterpret_cast<float4*>(D + row * 32 + 4 * c4) = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 32);
}

// -------------------------------------------------------------------------------------------------
// forward layer
// -------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tag_fwd_tc(TagTcArgs a) {
  extern __shared__ char raw[];
  const dss2_graph_t& g = a.g;
  char* base = align1024(raw);
  char* Bt = base;                                  // [(K+1)][hi,lo] x 4 KB
  char* At = Bt + (K + 1) * 2 * B_TILE;             // [(K+1)][hi,lo] x 32 KB
  char* tail = At + (K + 1) * 2 * A_TILE;
  float* bias_s = reinterpret_cast<float*>(tail);   // [32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 128);   // [2]
  uint32_t* tslot = reinterpret_cast<uint32_t*>(tail + 144);
  TopoTc topo;
  topo.cw = reinterpret_cast<int2*>(tail + 160);
  topo.rowptr = reinterpret_cast<int*>(topo.cw + g.max_tile_nnz + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lane4 = (uint32_t)lane << 2;
  const int cout = a.cout;
  auto a_hi = [&](int k) { return At + (size_t)(2 * k) * A_TILE; };
  auto a_lo = [&](int k) { return At + (size_t)(2 * k + 1) * A_TILE; };

  // ---- one-time setup: TMEM, barriers, weight operand tiles (rows >= cout are zero), bias ----
  if (warp == 0) tc::tmem_alloc(tslot, 64);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  for (int idx = tid; idx < (K + 1) * 32 * 8; idx += TC_THREADS) {
    const int k = idx >> 8;
    const uint32_t row = (idx >> 3) & 31, ch = idx & 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((int)row < cout) v = *reinterpret_cast<const float4*>(a.w + ((size_t)k * cout + row) * HID + ch * 4);
    tc::split_store4(v, Bt + (size_t)(2 * k) * B_TILE, Bt + (size_t)(2 * k + 1) * B_TILE, row * 128 + ((ch ^ (row & 7)) << 4));
  }
  if (tid < 32) bias_s[tid] = tid < cout ? a.bias[tid] : 0.0f;
  uint2 key = make_uint2(0u, 0u);
  uint32_t step_lo = 0;
  if (a.drop_mode == 1) {
    const uint64_t seed = a.rng[0], step = a.rng[1];
    key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (a.layer_uid * 0x9E3779B9u) ^ (uint32_t)(step >> 32));
    step_lo = (uint32_t)step;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tslot;
  const uint32_t idesc = tc::idesc_tf32(128, 32);
  uint32_t phase[2] = {0u, 0u};

  for (int t = blockIdx.x; t < g.num_tiles; t += gridDim.x) {
    const TileRange r = tile_range(g, t);
    const int nT = r.n1 - r.n0;
    const int nmb = nT > 128 ? 2 : 1;
    // ---- 1. X -> hi/lo operand tiles of hop level 0; topology ----
    {
      const float4* src = reinterpret_cast<const float4*>(a.x + (size_t)r.n0 * HID);
      for (int idx = tid; idx < nT * 8; idx += TC_THREADS) {
        const uint32_t row = idx >> 3, ch = idx & 7;
        tc::split_store4(ldg_stream4(src + idx), a_hi(0), a_lo(0), row * 128 + ((ch ^ (row & 7)) << 4));
      }
    }
    load_topo_tc(g, r, topo, tid, TC_THREADS);
    __syncthreads();
    // ---- 2. hops: level k from level k-1 (exact fp32 = hi + lo), split again ----
    const int nblk = (nT + 3) >> 2;
#pragma unroll
    for (int k = 1; k <= K; ++k) {
      for (int blk = warp; blk < nblk; blk += TC_WARPS) {
        float h[4] = {0.f, 0.f, 0.f, 0.f};
        hop_rows_tc<4>(topo, a_hi(k - 1), a_lo(k - 1), blk * 4, nT, lane4, h);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t off = tc::row_code((uint32_t)(blk * 4 + i)) ^ lane4;
          const float hi = tc::tf32_rna(h[i]);
          *reinterpret_cast<float*>(a_hi(k) + off) = hi;
          *reinterpret_cast<float*>(a_lo(k) + off) = tc::tf32_rna(h[i] - hi);
        }
      }
      if (k < K) __syncthreads();
    }
    // ---- 3. hand the operand tiles to the tensor core ----
    fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      for (int mb = 0; mb < nmb; ++mb) {
#pragma unroll
        for (int k = 0; k <= K; ++k)
          tc::issue_block(tmem + mb * 32, smem_u32(a_hi(k)) + mb * 128 * tc::ROW_BYTES, smem_u32(a_lo(k)) + mb * 128 * tc::ROW_BYTES,
                          smem_u32(Bt + (size_t)(2 * k) * B_TILE), smem_u32(Bt + (size_t)(2 * k + 1) * B_TILE), idesc, k == 0);
        tc::mma_commit(&bars[mb]);
      }
    }
    // ---- 4. epilogue: thread = node row ----
    const int mb = warp >> 2;
    if (mb < nmb) {
      mbar_wait(&bars[mb], phase[mb]);
      tc::fence_after_sync();
      float v[32];
      tc::tmem_ld32(tmem + mb * 32 + ((uint32_t)((warp & 3) * 32) << 16), v);
      const int row = mb * 128 + (warp & 3) * 32 + lane;
      if (row < nT) {
        const size_t n = (size_t)r.n0 + row;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 b = *reinterpret_cast<const float4*>(bias_s + 4 * c4);
          v[4 * c4 + 0] += b.x;
          v[4 * c4 + 1] += b.y;
          v[4 * c4 + 2] += b.z;
          v[4 * c4 + 3] += b.w;
        }
        if (a.act) {
          uint32_t keep = 0xffffffffu;
          if (a.drop_mode == 1) {
            keep = keep_word(key, (uint32_t)t, (uint32_t)row, step_lo, a.keep_thr16);
          } else if (a.drop_mode == 2) {
            const uint4* mp = reinterpret_cast<const uint4*>(a.mask + n * HID);
            const uint4 m0 = mp[0], m1 = mp[1];
            const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
            keep = 0u;
#pragma unroll
            for (int c = 0; c < 32; ++c) keep |= (((mw[c >> 2] >> ((c & 3) * 8)) & 0xffu) != 0u ? 1u : 0u) << c;
          }
          uint32_t word = 0u;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float xv = v[c];
            if (a.drop_mode != 0) xv = ((keep >> c) & 1u) ? xv * a.scale : 0.0f;
            xv = fmaxf(xv, 0.0f);
            word |= (xv > 0.0f ? 1u : 0u) << c;
            v[c] = xv;
          }
          if (a.bits) a.bits[n] = word;
        }
        if (a.res) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c < cout) v[c] += a.res[n * a.res_stride + c];
        }
        if (cout == 32) {
          float4* dst = reinterpret_cast<float4*>(a.y + n * 32);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) dst[c4] = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
        } else if (cout == 8) {
          float4* dst = reinterpret_cast<float4*>(a.y + n * 8);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        } else if (cout == 2) {
          *reinterpret_cast<float2*>(a.y + n * 2) = make_float2(v[0], v[1]);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c < cout) a.y[n * cout + c] = v[c];
        }
      }
      phase[mb] ^= 1u;
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
  }
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

size_t tc_smem(const dss2_graph_t* g, int K) {
  return 1024 + (size_t)(K + 1) * 2 * (B_TILE + A_TILE) + 160 + (size_t)(g->max_tile_nnz + 2) * 8 + (size_t)(round_up4(g->max_tile_nodes) + 8) * 4;
}

}  // namespace

extern "C" int dss2_tc_selftest(const float* A, const float* B, float* D, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSS2_CHECK_ARG(A && B && D, "dss2_tc_selftest: null argument");
  const int smem = 1024 + 40960 + 64;
  DSS2_CUDA(cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_tc_selftest<<<1, 128, smem, stream>>>(A, B, D);
  DSS2_LAUNCH_CHECK();
  return 0;
}

extern "C" int dss2_tc_selftest_mn(const float* A, const float* B, float* D, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSS2_CHECK_ARG(A && B && D, "dss2_tc_selftest_mn: null argument");
  const int smem = 1024 + 10 * ST_KR * 128 + 64;
  DSS2_CUDA(cudaFuncSetAttribute(k_tc_selftest_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_tc_selftest_mn<<<1, 128, smem, stream>>>(A, B, D);
  DSS2_LAUNCH_CHECK();
  return 0;
}

// 1 if the tensor-core forward can serve this (graph, K), else 0 (the caller then uses the CUDA-core kernel)
extern "C" int dss2_tag_fwd_tc_supported(const dss2_graph_t* g, int K) {
  if (!g || g->num_tiles <= 0 || K < 1 || K > 2) return 0;
  return tc_smem(g, K) <= 227 * 1024 ? 1 : 0;
}

extern "C" int dss2_tag_fwd_tc(const dss2_graph_t* g, const float* x, const float* w, const float* bias, int cout, int K, int act,
                               float p_drop, int drop_mode, const uint64_t* rng_state, uint32_t layer_uid, const uint8_t* mask,
                               const float* res, int64_t res_stride, float* y, uint32_t* act_bits, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSS2_CHECK_ARG(g && x && w && bias && y, "dss2_tag_fwd_tc: null argument");
  DSS2_CHECK_ARG(cout >= 1 && cout <= HID, "dss2_tag_fwd_tc: cout %d outside 1..%d", cout, HID);
  DSS2_CHECK_ARG(dss2_tag_fwd_tc_supported(g, K), "dss2_tag_fwd_tc: unsupported (needs a tiled graph, K in 1..2, tile fits shared memory)");
  DSS2_CHECK_ARG(p_drop >= 0.0f && p_drop < 1.0f, "dss2_tag_fwd_tc: dropout p %f outside [0,1)", p_drop);
  DSS2_CHECK_ARG(!(act && drop_mode == 1) || rng_state, "dss2_tag_fwd_tc: philox dropout needs rng_state");
  DSS2_CHECK_ARG(!(act && drop_mode == 2) || mask, "dss2_tag_fwd_tc: mask dropout needs a mask");
  if (g->num_nodes == 0) return 0;
  TagTcArgs a;
  a.g = *g;
  a.x = x;
  a.w = w;
  a.bias = bias;
  a.cout = cout;
  a.act = act;
  if (p_drop == 0.0f) drop_mode = 0;
  a.drop_mode = act ? drop_mode : 0;
  a.scale = 1.0f / (float)(1.0 - (double)p_drop);
  double thr = (1.0 - (double)p_drop) * 65536.0 + 0.5;
  a.keep_thr16 = thr >= 65536.0 ? 65536u : (uint32_t)thr;
  a.rng = rng_state;
  a.layer_uid = layer_uid;
  a.mask = mask;
  a.res = res;
  a.res_stride = res_stride;
  a.y = y;
  a.bits = act_bits;
  const size_t smem = tc_smem(g, K);
  const int grid = max(1, min(g->num_tiles, dss2_sm_count()));
  if (K == 1) {
    DSS2_CUDA(cudaFuncSetAttribute(k_tag_fwd_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tag_fwd_tc<1><<<grid, TC_THREADS, smem, stream>>>(a);
  } else {
    DSS2_CUDA(cudaFuncSetAttribute(k_tag_fwd_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tag_fwd_tc<2><<<grid, TC_THREADS, smem, stream>>>(a);
  }
  DSS2_LAUNCH_CHECK();
  return 0;
}
