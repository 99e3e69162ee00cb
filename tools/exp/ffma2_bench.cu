// Microbenchmark (not product code): rate of a thread-per-row 32x32 matvec with the weights in constant memory,
// packed fma.rn.f32x2 (FFMA2 + LDCU.128 weight pairs) against scalar FFMA with constant-bank operands.
#include <cuda_runtime.h>
#include <cstdio>
__constant__ float2 cw[32][16];
__constant__ float cs[32][32];
__device__ __forceinline__ void fma2(float2& d, float2 a, float2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(reinterpret_cast<unsigned long long&>(d)) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
}
template <int REP>
__global__ void __launch_bounds__(256) k2(const float* __restrict__ in, float* __restrict__ out, int n) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float s[32];
#pragma unroll
  for (int i = 0; i < 8; ++i) { float4 v = reinterpret_cast<const float4*>(in + (size_t)r * 32)[i]; s[4*i]=v.x; s[4*i+1]=v.y; s[4*i+2]=v.z; s[4*i+3]=v.w; }
#pragma unroll 1
  for (int rep = 0; rep < REP; ++rep) {
    float2 acc[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 32; ++h) {
      float2 sh = make_float2(s[h], s[h]);
#pragma unroll
      for (int o = 0; o < 16; ++o) fma2(acc[o], cw[h][o], sh);
    }
#pragma unroll
    for (int o = 0; o < 16; ++o) { s[2*o] = acc[o].x; s[2*o+1] = acc[o].y; }
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) reinterpret_cast<float4*>(out + (size_t)r * 32)[o] = make_float4(s[4*o], s[4*o+1], s[4*o+2], s[4*o+3]);
}
struct __align__(16) CW { float2 w[32][16]; float pad[64]; };
__constant__ CW cws[2];
// same with a run-time (uniform) slot index: ptxas emits LDCU.64 c[0x3][UR+imm], one per FFMA2
template <int REP>
__global__ void __launch_bounds__(256) k3(const float* __restrict__ in, float* __restrict__ out, int n, int slot) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const CW& C = cws[slot];
  float s[32];
#pragma unroll
  for (int i = 0; i < 8; ++i) { float4 v = reinterpret_cast<const float4*>(in + (size_t)r * 32)[i]; s[4*i]=v.x; s[4*i+1]=v.y; s[4*i+2]=v.z; s[4*i+3]=v.w; }
#pragma unroll 1
  for (int rep = 0; rep < REP; ++rep) {
    float2 acc[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 32; ++h) {
      float2 sh = make_float2(s[h], s[h]);
#pragma unroll
      for (int o = 0; o < 16; ++o) fma2(acc[o], C.w[h][o], sh);
    }
#pragma unroll
    for (int o = 0; o < 16; ++o) { s[2*o] = acc[o].x; s[2*o+1] = acc[o].y; }
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) reinterpret_cast<float4*>(out + (size_t)r * 32)[o] = make_float4(s[4*o], s[4*o+1], s[4*o+2], s[4*o+3]);
}
template <int REP>
__global__ void __launch_bounds__(256) k1(const float* __restrict__ in, float* __restrict__ out, int n) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float s[32];
#pragma unroll
  for (int i = 0; i < 8; ++i) { float4 v = reinterpret_cast<const float4*>(in + (size_t)r * 32)[i]; s[4*i]=v.x; s[4*i+1]=v.y; s[4*i+2]=v.z; s[4*i+3]=v.w; }
#pragma unroll 1
  for (int rep = 0; rep < REP; ++rep) {
    float acc[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) acc[o] = 0.f;
#pragma unroll
    for (int h = 0; h < 32; ++h)
#pragma unroll
      for (int o = 0; o < 32; ++o) acc[o] = fmaf(cs[h][o], s[h], acc[o]);
#pragma unroll
    for (int o = 0; o < 32; ++o) s[o] = acc[o];
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) reinterpret_cast<float4*>(out + (size_t)r * 32)[o] = make_float4(s[4*o], s[4*o+1], s[4*o+2], s[4*o+3]);
}
template <typename F>
float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < 10; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / 10 * 1e3f;
}
int main() {
  const int n = 286720;
  float *in, *out; cudaMalloc(&in, (size_t)n * 128); cudaMalloc(&out, (size_t)n * 128);
  cudaMemset(in, 0, (size_t)n * 128);
  float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 0.001f * i;
  cudaMemcpyToSymbol(cw, h, 4096); cudaMemcpyToSymbol(cs, h, 4096);
  int grid = (n + 255) / 256;
  printf("rows %d; 1024 FMA per row per rep\n", n);
  printf("FFMA2+LDCU rep1 %.1f us  rep4 %.1f us  rep16 %.1f us\n", timeit([&] { k2<1><<<grid, 256>>>(in, out, n); }), timeit([&] { k2<4><<<grid, 256>>>(in, out, n); }), timeit([&] { k2<16><<<grid, 256>>>(in, out, n); }));
  printf("FFMA c[]    rep1 %.1f us  rep4 %.1f us  rep16 %.1f us\n", timeit([&] { k1<1><<<grid, 256>>>(in, out, n); }), timeit([&] { k1<4><<<grid, 256>>>(in, out, n); }), timeit([&] { k1<16><<<grid, 256>>>(in, out, n); }));
  cudaMemcpyToSymbol(cws, h, 4096);
  printf("FFMA2+LDCU.64 (run-time slot) rep1 %.1f us  rep4 %.1f us  rep16 %.1f us\n", timeit([&] { k3<1><<<grid, 256>>>(in, out, n, 0); }), timeit([&] { k3<4><<<grid, 256>>>(in, out, n, 0); }), timeit([&] { k3<16><<<grid, 256>>>(in, out, n, 0); }));
  // per-warp issue rate of the FFMA2 + LDCU.128 mix: one CTA per SM with 1, 2, 4, 8 warps per scheduler
  for (int wps = 1; wps <= 8; wps *= 2) {
    const int threads = 128 * wps, rows = 148 * threads;
    if (threads > 256) break;
    float t1 = timeit([&] { k2<16><<<148, threads>>>(in, out, rows); }), t0 = timeit([&] { k2<1><<<148, threads>>>(in, out, rows); });
    // 768 issue slots (512 FFMA2 + 256 LDCU.128) per warp and rep
    printf("occupancy %d warp(s)/scheduler: %.2f us per rep -> %.3f instr/clk/warp (1965 MHz)\n", wps, (t1 - t0) / 15, 768.0 / ((t1 - t0) / 15 * 1965.0));
  }
  for (int nb = 1; nb <= 4; ++nb) {
    const int rows = 148 * nb * 256;
    float t1 = timeit([&] { k2<16><<<148 * nb, 256>>>(in, out, rows); }), t0 = timeit([&] { k2<1><<<148 * nb, 256>>>(in, out, rows); });
    printf("occupancy %d warps/scheduler (%d CTAs of 256 per SM): %.2f us per rep -> %.3f instr/clk/scheduler\n", 2 * nb, nb, (t1 - t0) / 15, 2 * nb * 768.0 / ((t1 - t0) / 15 * 1965.0));
  }
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
