// (f-3) Dataset builder on the device: the feature engineering of `data_from_pickles` (reference data.py:96-206) for S scenarios of one
// grid.  The reference loops over scenarios in Python, perturbs the pandapower results with `np.random.normal`, builds the inverse-variance
// weights, concatenates with O(n^2) `torch.cat` and z-scores the first 8 / 6 columns over their non-zero entries.  Here:
//   k_ds_rows      one thread per bus row / closed-branch row: the 11 / 13 columns exactly as the reference rounds them (float64 products
//                  and sums, one rounding each - no contraction -, cast to float32 where the reference casts), raw rows written once;
//                  per-column count and sum of the non-zero entries, accumulated in float64 per CTA in a fixed order
//   k_ds_finalize  column means (pass 1) / standard deviations (pass 2) from the per-CTA partials, fixed order
//   k_ds_moment    sum of the squared deviations of the non-zero entries ((t - mean)^2 formed in float32 like the reference's tensor op)
//   k_ds_normalise (t - mean) * mask / std with torch.nan_to_num semantics, parameter columns untouched (data.py:182,190)
// The measurement noise comes in as standard-normal draws (so a caller can replay the reference's np.random stream, or draw on the device).
// Raw columns, zero patterns and labels are bit-identical to the reference; the statistics are float64 sums of the same float32 terms
// (the reference sums them in float32 in torch's reduction order), i.e. equal to a few ulp.
#include <float.h>

#include "common.cuh"

namespace {

constexpr int DS_THREADS = 256;
constexpr int NX = 11, NE = 13, FX = 8, FEA = 6, NCOL = FX + FEA;   // columns that are z-scored: 8 node + 6 edge

struct DsArgs {
  const double* nodes;     // [S,N,7]  vn_kv, bool_slack, bool_zero_inj, vm_pu, va_rad, p_mw, q_mvar
  const double* cedges;    // [S,E,11] closed branches: from, to, G, B, Gs, Bs, closed, phase shift, imax or sn, p_from_mw, q_from_mvar
  const double* zn;        // [S,N,4]  standard-normal draws
  const double* ze;        // [S,E,2]
  const uint8_t* meas_v;   // [N] 1 = voltage magnitude measured at this bus (dss2_run.py:48-50)
  const uint8_t* meas_pf;  // [E] 1 = branch flow measured on this closed branch (dss2_run.py:51-53)
  double p_noise, v_noise, pm_noise, zero_inj_coef;
  int64_t S;
  int N, E;
  float* x;                // [S*N,11]
  float* ea;               // [S*E,13]
  double* partials;        // [grid][2*NCOL]: counts then sums (pass 1) / squared deviations in the first NCOL (pass 2)
  float* stats;            // [2*NCOL]: x_mean[8], e_mean[6], x_std[8], e_std[6]
  double* counts;          // [NCOL]
};

__device__ __forceinline__ float inv_var(float s, float floor_, float cap) {   // data.py:137-138, 161-162
  const float m = fmaxf(fabsf(s), floor_);
  const float c = 1.0f / (m * m);
  return c < cap ? c : 0.0f;
}

// fixed-order block reduction of NV doubles per thread; thread 0 of the CTA ends up with the totals
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[warp * NV + i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = 0.0;
      for (int w = 0; w < DS_THREADS / 32; ++w) s += smem[w * NV + i];
      v[i] = s;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(DS_THREADS) k_ds_rows(DsArgs a) {
  __shared__ double red[(DS_THREADS / 32) * 2 * NCOL];
  double acc[2 * NCOL];
#pragma unroll
  for (int i = 0; i < 2 * NCOL; ++i) acc[i] = 0.0;
  const int64_t nrows = a.S * a.N, erows = a.S * a.E;
  for (int64_t r = blockIdx.x * (int64_t)DS_THREADS + threadIdx.x; r < nrows + erows; r += (int64_t)gridDim.x * DS_THREADS) {
    if (r < nrows) {
      const int n = (int)(r % a.N);
      const double* nd = a.nodes + r * 7;
      const double* z = a.zn + r * 4;
      const double slack = nd[1], zinj = nd[2];
      const double mask[4] = {a.meas_v[n] ? 1.0 : 0.0, 0.0, 1.0, 1.0};                                   // data.py:121-124
      const double slack_noise[4] = {a.v_noise, a.zero_inj_coef, a.p_noise, a.p_noise};                   // data.py:111
      const double node_noise[4] = {a.v_noise, a.v_noise, a.pm_noise, a.pm_noise};                        // data.py:109
      float* out = a.x + r * NX;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double mean = __dmul_rn(nd[3 + c], mask[c]);                                                // data.py:127
        const double coef = __dadd_rn(__dmul_rn(slack_noise[c], slack), __dmul_rn(node_noise[c], __dsub_rn(1.0, slack)));
        double std = __dmul_rn(mean, coef);                                                               // data.py:128
        const float val = (float)__dadd_rn(mean, __dmul_rn(fabs(std), z[c]));                             // data.py:131
        if (c >= 2) std = __dadd_rn(std, __dmul_rn(a.zero_inj_coef, zinj));                               // data.py:133
        if (c == 1) std = __dadd_rn(std, __dmul_rn(slack_noise[1], slack));                               // data.py:135
        const float w = inv_var((float)std, 1e-6f, 1e12f);
        out[2 * c] = val;
        out[2 * c + 1] = w;
        if (val != 0.0f) {
          acc[2 * c] += 1.0;
          acc[NCOL + 2 * c] += (double)val;
        }
        if (w != 0.0f) {
          acc[2 * c + 1] += 1.0;
          acc[NCOL + 2 * c + 1] += (double)w;
        }
      }
      out[8] = (float)nd[0];
      out[9] = (float)nd[1];
      out[10] = (float)nd[2];
    } else {
      const int64_t q = r - nrows;
      const int e = (int)(q % a.E);
      const double* ed = a.cedges + q * 11;
      const double* z = a.ze + q * 2;
      const double m = a.meas_pf[e] ? 1.0 : 0.0;                                                          // data.py:148-151
      float* out = a.ea + q * NE;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const double mean = __dmul_rn(ed[9 + c], m);                                                      // data.py:156
        const double std = __dmul_rn(mean, a.p_noise);                                                    // data.py:157
        const float val = (float)__dadd_rn(mean, __dmul_rn(fabs(std), z[c]));                             // data.py:159
        const float w = inv_var((float)std, 1e-5f, 1e10f);
        out[2 * c] = val;
        out[2 * c + 1] = w;
        if (val != 0.0f) {
          acc[FX + 2 * c] += 1.0;
          acc[NCOL + FX + 2 * c] += (double)val;
        }
        if (w != 0.0f) {
          acc[FX + 2 * c + 1] += 1.0;
          acc[NCOL + FX + 2 * c + 1] += (double)w;
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {                                                                       // G, B as features, data.py:164
        const float v = (float)ed[2 + c];
        out[4 + c] = v;
        if (v != 0.0f) {
          acc[FX + 4 + c] += 1.0;
          acc[NCOL + FX + 4 + c] += (double)v;
        }
      }
#pragma unroll
      for (int c = 0; c < 7; ++c) out[6 + c] = (float)ed[2 + c];                                          // data.py:172
    }
  }
  block_sum<2 * NCOL>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 2 * NCOL; ++i) a.partials[(size_t)blockIdx.x * 2 * NCOL + i] = acc[i];
  }
}

// pass 0: counts and means; pass 1: standard deviations.  torch: nan_to_num(float32 sum / count) (data.py:180-181, 186-187)
__global__ void k_ds_finalize(DsArgs a, int grid, int pass) {
  const int c = threadIdx.x;
  if (c >= NCOL) return;
  if (pass == 0) {
    double cnt = 0.0, sum = 0.0;
    for (int b = 0; b < grid; ++b) {
      cnt += a.partials[(size_t)b * 2 * NCOL + c];
      sum += a.partials[(size_t)b * 2 * NCOL + NCOL + c];
    }
    a.counts[c] = cnt;
    const float mean = (float)sum / (float)cnt;
    a.stats[c] = isnan(mean) ? 0.0f : mean;
  } else {
    double sq = 0.0;
    for (int b = 0; b < grid; ++b) sq += a.partials[(size_t)b * 2 * NCOL + c];
    const float std = sqrtf((float)sq / (float)a.counts[c]);
    a.stats[NCOL + c] = isnan(std) ? 0.0f : std;
  }
}

__global__ void __launch_bounds__(DS_THREADS) k_ds_moment(DsArgs a) {
  __shared__ double red[(DS_THREADS / 32) * NCOL];
  double acc[NCOL];
#pragma unroll
  for (int i = 0; i < NCOL; ++i) acc[i] = 0.0;
  const int64_t nrows = a.S * a.N, erows = a.S * a.E;
  for (int64_t r = blockIdx.x * (int64_t)DS_THREADS + threadIdx.x; r < nrows + erows; r += (int64_t)gridDim.x * DS_THREADS) {
    if (r < nrows) {
      const float* row = a.x + r * NX;
#pragma unroll
      for (int c = 0; c < FX; ++c) {
        const float t = row[c], d = t - a.stats[c];
        if (t != 0.0f) acc[c] += (double)(d * d);                                                         // data.py:181
      }
    } else {
      const float* row = a.ea + (r - nrows) * NE;
#pragma unroll
      for (int c = 0; c < FEA; ++c) {
        const float t = row[c], d = t - a.stats[FX + c];
        if (t != 0.0f) acc[FX + c] += (double)(d * d);
      }
    }
  }
  block_sum<NCOL>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NCOL; ++i) a.partials[(size_t)blockIdx.x * 2 * NCOL + i] = acc[i];
  }
}

__device__ __forceinline__ float nan_to_num(float v) { return isnan(v) ? 0.0f : (isinf(v) ? (v > 0.0f ? FLT_MAX : -FLT_MAX) : v); }

__global__ void __launch_bounds__(DS_THREADS) k_ds_normalise(DsArgs a) {
  const int64_t nrows = a.S * a.N, erows = a.S * a.E;
  for (int64_t r = blockIdx.x * (int64_t)DS_THREADS + threadIdx.x; r < nrows + erows; r += (int64_t)gridDim.x * DS_THREADS) {
    if (r < nrows) {
      float* row = a.x + r * NX;
#pragma unroll
      for (int c = 0; c < FX; ++c) {
        const float t = row[c];
        row[c] = nan_to_num(((t - a.stats[c]) * (t != 0.0f ? 1.0f : 0.0f)) / a.stats[NCOL + c]);           // data.py:182
      }
    } else {
      float* row = a.ea + (r - nrows) * NE;
#pragma unroll
      for (int c = 0; c < FEA; ++c) {
        const float t = row[c];
        row[c] = nan_to_num(((t - a.stats[FX + c]) * (t != 0.0f ? 1.0f : 0.0f)) / a.stats[NCOL + FX + c]);  // data.py:188
      }
    }
  }
}

// Synthetic scenario sampler (toy_network.py:83-126 + loadsampling.py:75-107), two element-wise kernels in the reference's rounding order:
//   k_load_profiles  mu[l*H + h] = wa[l] * (base[l] * prof_a[h]) + wb[l] * (base[l] * prof_b[h])          (toy_network.py:106-107, 111-114)
//                    and the two sampler arguments: (mu (1 - s), mu (1 + s)) for 'uniform', (mu, mu s) for 'normal' (toy_network.py:119-123)
//   k_mc_sample      out[u, i] = A[u] + draw[u, i] * (B[u] - A[u])   (samplermontecarlo, loadsampling.py:78-91)
//                    out[u, i] = A[u] + B[u] * draw[u, i]            (samplermontecarlo_normal = MU + SIG * gauss, loadsampling.py:105)
__global__ void __launch_bounds__(DS_THREADS) k_load_profiles(const double* __restrict__ base, const double* __restrict__ wa,
                                                              const double* __restrict__ wb, const double* __restrict__ prof_a,
                                                              const double* __restrict__ prof_b, int L, int H, int dist, double spread,
                                                              double* __restrict__ A, double* __restrict__ B) {
  const int total = L * H;
  for (int u = blockIdx.x * DS_THREADS + threadIdx.x; u < total; u += gridDim.x * DS_THREADS) {
    const int l = u / H, h = u % H;
    const double mu = __dadd_rn(__dmul_rn(wa[l], __dmul_rn(base[l], prof_a[h])), __dmul_rn(wb[l], __dmul_rn(base[l], prof_b[h])));
    if (dist == 0) {
      A[u] = __dmul_rn(mu, __dsub_rn(1.0, spread));
      B[u] = __dmul_rn(mu, __dadd_rn(1.0, spread));
    } else {
      A[u] = mu;
      B[u] = __dmul_rn(mu, spread);
    }
  }
}

__global__ void __launch_bounds__(DS_THREADS) k_mc_sample(const double* __restrict__ A, const double* __restrict__ B, int64_t U, int iters, int dist,
                                                          const double* __restrict__ draws, double* __restrict__ out) {
  const int64_t total = U * iters;
  for (int64_t i = blockIdx.x * (int64_t)DS_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * DS_THREADS) {
    const int64_t u = i / iters;
    out[i] = dist == 0 ? __dadd_rn(A[u], __dmul_rn(draws[i], __dsub_rn(B[u], A[u]))) : __dadd_rn(A[u], __dmul_rn(B[u], draws[i]));
  }
}

int ds_grid(int64_t rows) { return (int)max((int64_t)1, min((int64_t)dss2_sm_count() * 8, (rows + DS_THREADS - 1) / DS_THREADS)); }

}  // namespace

extern "C" size_t dss2_build_scenarios_workspace_bytes(void) { return ((size_t)dss2_sm_count() * 8 * 2 * NCOL + NCOL) * sizeof(double); }

extern "C" int dss2_build_scenarios(const double* nodes, const double* closed_edges, const double* noise_nodes, const double* noise_edges,
                                    const uint8_t* meas_v_mask, const uint8_t* meas_pflow_mask, const double* noise_param6, int64_t S, int N,
                                    int E, float* x, float* edge_attr, float* stats28, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSS2_CHECK_ARG(nodes && closed_edges && noise_nodes && noise_edges && meas_v_mask && meas_pflow_mask && noise_param6 && x && edge_attr &&
                     stats28 && workspace,
                 "dss2_build_scenarios: null argument");
  DSS2_CHECK_ARG(S >= 1 && N >= 1 && E >= 0, "dss2_build_scenarios: bad sizes");
  DSS2_CHECK_ARG(workspace_bytes >= dss2_build_scenarios_workspace_bytes(), "dss2_build_scenarios: workspace too small");
  DsArgs a = {};
  a.nodes = nodes;
  a.cedges = closed_edges;
  a.zn = noise_nodes;
  a.ze = noise_edges;
  a.meas_v = meas_v_mask;
  a.meas_pf = meas_pflow_mask;
  // NOISE_COLS order of the reference's noise_param pickle: p_noise, v_noise, i_noise, pm_noise, sgen_noise, zero_inj_coef (host values)
  a.p_noise = noise_param6[0];
  a.v_noise = noise_param6[1];
  a.pm_noise = noise_param6[3];
  a.zero_inj_coef = noise_param6[5];
  a.S = S;
  a.N = N;
  a.E = E;
  a.x = x;
  a.ea = edge_attr;
  a.partials = (double*)workspace;
  a.counts = a.partials + (size_t)dss2_sm_count() * 8 * 2 * NCOL;
  a.stats = stats28;
  const int grid = ds_grid(S * (int64_t)(N + E));
  k_ds_rows<<<grid, DS_THREADS, 0, stream>>>(a);
  DSS2_LAUNCH_CHECK();
  k_ds_finalize<<<1, 32, 0, stream>>>(a, grid, 0);
  DSS2_LAUNCH_CHECK();
  k_ds_moment<<<grid, DS_THREADS, 0, stream>>>(a);
  DSS2_LAUNCH_CHECK();
  k_ds_finalize<<<1, 32, 0, stream>>>(a, grid, 1);
  DSS2_LAUNCH_CHECK();
  k_ds_normalise<<<grid, DS_THREADS, 0, stream>>>(a);
  DSS2_LAUNCH_CHECK();
  return 0;
}

extern "C" int dss2_load_profiles(const double* base, const double* weight_a, const double* weight_b, const double* profile_a,
                                  const double* profile_b, int L, int H, int dist, double spread, double* arg_a, double* arg_b, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSS2_CHECK_ARG(base && weight_a && weight_b && profile_a && profile_b && arg_a && arg_b, "dss2_load_profiles: null argument");
  DSS2_CHECK_ARG(L >= 1 && H >= 1 && (dist == 0 || dist == 1), "dss2_load_profiles: bad sizes or distribution");
  k_load_profiles<<<ds_grid((int64_t)L * H), DS_THREADS, 0, stream>>>(base, weight_a, weight_b, profile_a, profile_b, L, H, dist, spread, arg_a, arg_b);
  DSS2_LAUNCH_CHECK();
  return 0;
}

extern "C" int dss2_mc_sample(const double* arg_a, const double* arg_b, int64_t U, int iters, int dist, const double* draws, double* out,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DSS2_CHECK_ARG(arg_a && arg_b && draws && out, "dss2_mc_sample: null argument");
  DSS2_CHECK_ARG(U >= 1 && iters >= 1 && (dist == 0 || dist == 1), "dss2_mc_sample: bad sizes or distribution");
  k_mc_sample<<<ds_grid(U * iters), DS_THREADS, 0, stream>>>(arg_a, arg_b, U, iters, dist, draws, out);
  DSS2_LAUNCH_CHECK();
  return 0;
}
