// Shared host/device helpers for libdss2_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dss2_b200.h"

#define HID DSS2_HID
#define WARP 32

// ---------------------------------------------------------------------------------------------
// host: error reporting + launch accounting
// ---------------------------------------------------------------------------------------------
void dss2_set_error(const char* fmt, ...);
void dss2_count_launch(int n);
int dss2_sm_count();
// launch priority of the kernels on the step's critical path when a second stream competes for the SMs (the two-stream backward:
// the weight-gradient passes are fillers).  level 1: per-tile chained launches, level 2: + every TMA-fed TAG launch and the
// EdgeAggregation backward.  0 = no attribute.  DSS2_CHAIN_PRIO selects the highest level that gets the attribute (default 1).
int dss2_launch_priority(int level);

#define DSS2_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      dss2_set_error(__VA_ARGS__);           \
      return -1;                             \
    }                                        \
  } while (0)

#define DSS2_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t _e = (call);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      dss2_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));       \
      return -2;                                                                                 \
    }                                                                                            \
  } while (0)

#define DSS2_LAUNCH_CHECK()                                                                      \
  do {                                                                                           \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (_e != cudaSuccess) {                                                                     \
      dss2_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e));   \
      return -3;                                                                                 \
    }                                                                                            \
    dss2_count_launch(1);                                                                        \
  } while (0)

// tcgen05 transform of a layer on precomputed hop levels (tag_tc2.cu), used by the large-graph path in tag.cu
int dss2_tc2_dense_fwd(const dss2_graph_t* g, const float* x, const float* lvl, const float* w, const float* bias, int cout, int K, int act,
                       float p_drop, int drop_mode, const uint64_t* rng_state, uint32_t layer_uid, const uint8_t* mask, const float* res,
                       int64_t res_stride, float* y, uint32_t* act_bits, cudaStream_t stream);
int dss2_tc2_dense_bgx(const dss2_graph_t* g, const float* grad_y, const uint32_t* act_bits, float p_drop, const float* lvl, const float* w,
                       int cout, int K, float* grad_x, cudaStream_t stream);

// thread-per-row EdgeAggregation kernels (edgeagg_row.cu), selected by the dss2_edgeagg_* entry points in edgeagg.cu
int dss2_ea_row_fits(const dss2_graph_t* g, int64_t x_stride, int64_t ea_stride, int fe, int bwd);
int dss2_ea_row_scratch_slot();
int dss2_ea_row_upload(int slot0, int n, const float* const* w1, const float* const* b1, const float* const* w2, const float* const* b2, int fn,
                       int fe, cudaStream_t stream);
int dss2_ea_row_fwd(const dss2_graph_t* g, const float* x, int64_t x_stride, int fn, const float* edge_attr, int64_t ea_stride, int fe, int slot,
                    float* out, cudaStream_t stream);
int dss2_ea_row_bwd(const dss2_graph_t* g, const float* x, int64_t x_stride, int fn, const float* edge_attr, int64_t ea_stride, int fe, int slot,
                    const float* grad_out, const float* skip_grad, int64_t skip_stride, float* grad_x, float* partials, int64_t partial_stride,
                    cudaStream_t stream);

// large-graph (num_tiles == 0) variants, implemented next to their tiled counterparts
#define DSS2_NEED_SCRATCH(g, who)                                                                                        \
  DSS2_CHECK_ARG((g)->scratch && (g)->scratch_bytes >= dss2_generic_scratch_bytes((g)->num_nodes),                        \
                 "%s: graph larger than a tile needs g->scratch of dss2_generic_scratch_bytes() bytes", who)

// ---------------------------------------------------------------------------------------------
// device: small utilities
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox4x32-10 (Salmon et al.), counter-based: the dropout mask of (layer, node, feature) is a pure
// function of (seed, step) and never has to be stored or replayed by the backward.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// ---- mbarrier + 1-D bulk async copies (TMA engine, no tensor map needed for contiguous rows) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared, completion signalled on `bar` (bytes % 16 == 0, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global (bulk group)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// streaming (read-once) 128-bit load
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// Device view of the tiling of dss2_graph_t: tile t = graphs [t*G, min((t+1)*G, B)).
struct TileRange {
  int n0, n1;      // node range
  int z0, z1;      // CSR entry range
  long long e0, e1;  // one-way edge range
};
__device__ __forceinline__ TileRange tile_range(const dss2_graph_t& g, int t) {
  TileRange r;
  int g0 = t * g.graphs_per_tile;
  int g1 = min(g0 + g.graphs_per_tile, g.num_graphs);
  r.n0 = (int)g.ptr[g0];
  r.n1 = (int)g.ptr[g1];
  r.z0 = g.rowptr[r.n0];
  r.z1 = g.rowptr[r.n1];
  r.e0 = g.eptr[g0];
  r.e1 = g.eptr[g1];
  return r;
}
#endif  // __CUDACC__
