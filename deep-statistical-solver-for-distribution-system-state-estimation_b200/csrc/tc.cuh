// tcgen05 / TMEM helpers for the 32x32 feature transforms (sm_100a).
//
// The transforms are evaluated as 3xTF32: x = x_hi + x_lo with x_hi = round-to-nearest TF32(x) and x_lo = x - x_hi (exact in
// fp32, then itself rounded to TF32), and x.w ~= x_hi.w_hi + x_lo.w_hi + x_hi.w_lo with fp32 accumulation in TMEM.  The
// dropped terms are <= 2^-22 relative, i.e. fp32-equivalent accuracy (plain TF32 would be ~1e-3 and break 1e-5 parity).
//
// Operand tiles live in shared memory as [rows][32 fp32] = 128-byte rows, K-major, SWIZZLE_128B: 16-byte chunk c of row r
// is stored at chunk (c ^ (r & 7)); an 8-row group is 1024 bytes (SBO), the tile base is 1024-byte aligned.  One
// tcgen05.mma.kind::tf32 consumes K = 8 elements = 32 bytes, so a 32-wide row is 4 MMAs whose descriptors advance by 32 bytes.
// Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (CUTLASS, vendored in the image).
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace tc {

constexpr uint32_t ROW_BYTES = 128;          // 32 fp32
constexpr uint32_t SBO_BYTES = 1024;         // 8 rows
constexpr uint32_t KSTEP_BYTES = 32;         // 8 tf32

// byte offset of element (row, col j) inside a swizzled [rows][32] tile
__device__ __forceinline__ uint32_t swz_off(uint32_t row, uint32_t j) {
  return row * ROW_BYTES + ((((j >> 2) ^ (row & 7u)) << 4) | ((j & 3u) << 2));
}
// same, as "row code" ^ (4*j): code = row*128 | (row&7)<<4
__device__ __forceinline__ uint32_t row_code(uint32_t row) { return (row << 7) | ((row & 7u) << 4); }

// round-to-nearest (ties away from zero) to TF32 = add half a TF32 ulp to the magnitude and truncate.  cvt.rna.tf32.f32 does the same
// plus NaN/Inf special-casing, which ptxas expands to ~5 instructions on sm_100a (ncu: LOP3/FSETP/SEL made up 40% of the kernel);
// activations and weights here are finite, so the two-instruction integer form is used.
__device__ __forceinline__ float tf32_rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);            // start address, bits [0,14)
  d |= (uint64_t)1u << 16;                                 // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)((SBO_BYTES >> 4) & 0x3FFFu) << 32;       // stride byte offset, bits [32,46)
  d |= (uint64_t)1u << 46;                                 // version = 1
  d |= (uint64_t)2u << 61;                                 // layout type = SWIZZLE_128B
  return d;
}

// MN-major TF32 operands must use the SWIZZLE_128B_BASE32B layout (CUTLASS: "for mn-major tf32 operands, SW128_32B is the only
// available smem layout"): a [k rows][32 fp32] tile with 128-byte rows whose 32-byte chunk p of row r is stored at chunk
// p ^ (r & 3) (Swizzle<2,5,2>); the atom is 4 rows = 512 bytes (SBO).  The contraction index is the tile row: one MMA (K = 8)
// consumes 8 rows = 1024 bytes, so K-steps advance the start address by 1024; M/N extents beyond one 32-element row continue in
// the next tile, `lbo_bytes` further on.
__device__ __forceinline__ uint32_t swz32_off(uint32_t row, uint32_t j) {
  return row * ROW_BYTES + ((((j >> 3) ^ (row & 3u)) << 5) | ((j & 7u) << 2));
}
__device__ __forceinline__ uint64_t smem_desc_mn32(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)1u << 61;   // layout type 1 = SWIZZLE_128B_BASE32B
  return d;
}

// ---- no-swizzle ("interleaved") core-matrix tiles ----
// A tile is a grid of 128-byte core matrices, each 8 rows x 4 fp32 (row i of the core at byte 16*i): element (r, c) of a tile with
// Q column groups lives at (r/8)*(Q*128) + (c/4)*128 + (r%8)*16 + (c%4)*4.  The SAME bytes are a valid K-major operand
// (rows = M/N index, columns = contraction) and a valid MN-major operand (columns = M/N index, rows = contraction): that is what
// lets one set of tiles feed both GEMMs of the backward (grad_x = G W and grad_W = G^T X).
__device__ __forceinline__ uint32_t core_off(uint32_t r, uint32_t c, uint32_t q_groups) {
  return (r >> 3) * (q_groups * 128u) + (c >> 2) * 128u + ((r & 7u) << 4) + ((c & 3u) << 2);
}
// K-major use: LBO = distance between the two core matrices one MMA (K = 8) spans = 128 B; SBO = distance between 8-row groups
__device__ __forceinline__ uint64_t smem_desc_core_k(uint32_t smem_addr, uint32_t row_group_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(128u >> 4) << 16;
  d |= (uint64_t)((row_group_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1u << 46;
  return d;   // layout type 0 = no swizzle
}
// MN-major use: SBO = distance between groups of 4 M/N elements = 128 B; LBO = distance between 8-row contraction blocks
__device__ __forceinline__ uint64_t smem_desc_core_mn(uint32_t smem_addr, uint32_t row_group_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((row_group_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(128u >> 4) << 32;
  d |= (uint64_t)1u << 46;
  return d;
}

// instruction descriptor: D fp32, A/B tf32, M x N; a_mn / b_mn = 1 selects an MN-major operand (0 = K-major)
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn = 0, uint32_t b_mn = 0) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// true in exactly one lane of a converged warp.  ptxas recognises the elect.sync idiom: inside `if (elect_one())` every register is
// uniform by construction, so tcgen05.mma operands move to the uniform datapath with one R2UR each instead of the per-lane
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop it emits under an ordinary `lane == 0` branch (11 instructions per MMA in round 1's SASS).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand in tensor memory (lane = row m, column = k, one fp32 word per element; the tensor core reads the top 19
// bits = TF32 by truncation): no shared-memory read for A at all
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32*(warp%4) + laneid)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// split 4 consecutive values and store hi / lo chunks (16 bytes each) at byte offset `off` of the two tiles
__device__ __forceinline__ void split_store4(const float4 v, char* hi_tile, char* lo_tile, uint32_t off) {
  float4 h, l;
  h.x = tf32_rna(v.x);
  h.y = tf32_rna(v.y);
  h.z = tf32_rna(v.z);
  h.w = tf32_rna(v.w);
  // the residual is rounded to TF32 as well: the tensor core would otherwise TRUNCATE it (biased, errors add coherently in long
  // sums); rounded, the representation error is <= 2^-22 |x| and unbiased
  l.x = tf32_rna(v.x - h.x);
  l.y = tf32_rna(v.y - h.y);
  l.z = tf32_rna(v.z - h.z);
  l.w = tf32_rna(v.w - h.w);
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

// ---- "plain + residual" operand pair ----
// The tensor core reads only the top 19 bits of an fp32 word (TF32 = truncation), so the PLAIN fp32 tile can serve as the high
// operand: hi = trunc_tf32(x) implicitly.  The residual tile holds lo = rna_tf32(x - trunc_tf32(x)) (|lo| < 2^-10 |x|, rounded to 11
// bits: unbiased error <= 2^-21 |x|).  Benefits: hops gather the exact fp32 row from ONE tile, and the split costs 4 instructions.
__device__ __forceinline__ float tf32_residual(float x) { return tf32_rna(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u)); }
__device__ __forceinline__ void plain_store4(const float4 v, char* p_tile, char* l_tile, uint32_t off) {
  *reinterpret_cast<float4*>(p_tile + off) = v;
  *reinterpret_cast<float4*>(l_tile + off) = make_float4(tf32_residual(v.x), tf32_residual(v.y), tf32_residual(v.z), tf32_residual(v.w));
}

// arrive (count 1) on a CTA-local mbarrier
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// advance the start-address field of a descriptor by `bytes` (multiple of 16, no carry out of the 14-bit field for our tile sizes)
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

// 16 columns (half a row) of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// the same load without the wait: issue several, then tmem_wait_ld() once
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16 columns of this thread's TMEM lane <- registers (warp-collective); tmem_wait_st() before the data is handed to another thread
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 4-term product of one [128 x 32] activation tile pair with one [32 x 32] weight tile pair from PRE-BUILT descriptors
__device__ __forceinline__ void issue_block_desc(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc,
                                                 bool first) {
#pragma unroll
  for (uint32_t kk = 0; kk < ROW_BYTES / KSTEP_BYTES; ++kk) {
    const uint32_t o = kk * KSTEP_BYTES;
    mma_tf32(d_tmem, desc_advance(a_lo, o), desc_advance(b_lo, o), idesc, (first && kk == 0) ? 0u : 1u);
    mma_tf32(d_tmem, desc_advance(a_lo, o), desc_advance(b_hi, o), idesc, 1u);
    mma_tf32(d_tmem, desc_advance(a_hi, o), desc_advance(b_lo, o), idesc, 1u);
    mma_tf32(d_tmem, desc_advance(a_hi, o), desc_advance(b_hi, o), idesc, 1u);
  }
}

// Issue the 3xTF32 product of one [128 x 32] activation block pair (hi, lo) with one [32 x 32] weight block pair into
// 32 TMEM columns.  `first` = this is the first product of the accumulation chain.
__device__ __forceinline__ void issue_block(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t idesc,
                                            bool first) {
#pragma unroll
  for (uint32_t kk = 0; kk < ROW_BYTES / KSTEP_BYTES; ++kk) {
    const uint32_t o = kk * KSTEP_BYTES;
    // small terms first; the lo.lo term (2^-22) is kept too - MMA issue is nowhere near the bottleneck and it leaves the operand
    // representation (2^-23 per operand, unbiased) as the only error source
    mma_tf32(d_tmem, smem_desc_sw128(a_lo + o), smem_desc_sw128(b_lo + o), idesc, (first && kk == 0) ? 0u : 1u);
    mma_tf32(d_tmem, smem_desc_sw128(a_lo + o), smem_desc_sw128(b_hi + o), idesc, 1u);
    mma_tf32(d_tmem, smem_desc_sw128(a_hi + o), smem_desc_sw128(b_lo + o), idesc, 1u);
    mma_tf32(d_tmem, smem_desc_sw128(a_hi + o), smem_desc_sw128(b_hi + o), idesc, 1u);
  }
}

}  // namespace tc
