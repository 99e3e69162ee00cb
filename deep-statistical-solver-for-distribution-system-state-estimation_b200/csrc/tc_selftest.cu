// tcgen05 descriptor self tests: D[128,32] = A[128,32] * B[32,32]^T through the operand layouts, shared-memory / instruction descriptors and
// TMEM read-back that the layer kernels use (K-major SWIZZLE_128B operands: tag_tc2.cu / tag_tc3.cu; MN-major SWIZZLE_128B_BASE32B:
// k_tag_gw).  Run by tests/test_gpu_parity.py (`dss2_tc_selftest`, `dss2_tc_selftest_mn`) so that an encoding mistake shows up in
// isolation rather than as a wrong layer output.
#include "common.cuh"
#include "tc.cuh"

namespace {

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
