// TEST HARNESS (not product code): drives the per-bus / per-branch math of csrc/wls_math.cuh on the
// host, sequentially and without tiles, so CPU tests can compare those exact source lines (forward
// and hand-derived adjoint) with the oracle's autograd.  Built by tests/test_wls_host_math.py with
// g++ -ffp-contract=off.  The pass structure mirrors k_wls in csrc/wls.cu.
#include <stdint.h>
#include <vector>

#include "../../deep-statistical-solver-for-distribution-system-state-estimation_b200/csrc/wls_math.cuh"

extern "C" int wls_host_loss(const float* x, const float* ea, const int64_t* ei, const float* out, const float* stats,
                             const float* coefs, int64_t Nt, int64_t Et, float* loss, float* grad_out, float* pflow8) {
  WlsStats st;
  for (int i = 0; i < 28; ++i) ((float*)&st)[i] = stats[i];
  WlsCoefs k{coefs[0], coefs[1], coefs[2], coefs[3]};
  float vmin = 1e30f, vmax = 0.f;
  for (int64_t n = 0; n < Nt; ++n) {
    vmin = fminf(vmin, x[n * 11 + 8]);
    vmax = fmaxf(vmax, x[n * 11 + 8]);
  }
  WlsGrid grid = wls_grid(vmin, vmax);
  std::vector<float> v(Nt), th(Nt), pb(Nt, 0.f), qb(Nt, 0.f);
  std::vector<float> spt(Nt, 0.f), sqt(Nt, 0.f), spf(Nt, 0.f), sqf(Nt, 0.f);
  std::vector<WlsBranch> br(Et);
  std::vector<WlsBranchIn> bin(Et);
  for (int64_t n = 0; n < Nt; ++n) {
    v[n] = out[2 * n] * st.xs[0] + st.xm[0];
    th[n] = out[2 * n + 1] * (1.0f - x[n * 11 + 9]);
  }
  double sje = 0, sth = 0, sload = 0, sjn = 0, sv = 0;
  for (int64_t e = 0; e < Et; ++e) {
    int64_t i = ei[e], j = ei[Et + e];
    const float* row = ea + e * 13;
    WlsBranchIn in{v[i], v[j], th[i], th[j], row[6], row[7], row[8], row[9], row[11], row[12]};
    bin[e] = in;
    wls_branch_forward(in, grid, br[e]);
    if (pflow8) {
      pflow8[e] = br[e].ll; pflow8[Et + e] = br[e].lt; pflow8[2 * Et + e] = br[e].pf; pflow8[3 * Et + e] = br[e].qf;
      pflow8[4 * Et + e] = br[e].pt; pflow8[5 * Et + e] = br[e].qt; pflow8[6 * Et + e] = br[e].i_f; pflow8[7 * Et + e] = br[e].i_t;
    }
    spt[j] += br[e].pt; sqt[j] += br[e].qt; spf[i] += br[e].pf; sqf[i] += br[e].qf;
    float eZ0 = wls_unnorm(row[0], st.es[0], st.em[0]), eR0 = wls_unnorm(row[1], st.es[1], st.em[1]);
    float eZ1 = wls_unnorm(row[2], st.es[2], st.em[2]), eR1 = wls_unnorm(row[3], st.es[3], st.em[3]);
    sje += wls_branch_residual(eZ0, eR0, eZ1, eR1, br[e].pf, br[e].qf, k);
    sth += fmaxf(fabsf(br[e].delta) - 0.5f, 0.f);
    sload += fmaxf(br[e].loading - 1.5f, 0.f);
  }
  std::vector<WlsBus> bus(Nt);
  for (int64_t n = 0; n < Nt; ++n) {
    pb[n] = -spt[n] - spf[n];
    qb[n] = -sqt[n] - sqf[n];
    wls_bus_load(x + n * 11, out[2 * n], out[2 * n + 1], st, bus[n]);
    sjn += wls_bus_residual(bus[n], pb[n], qb[n], k);
    sv += wls_bus_vband(bus[n]);
  }
  double N = (double)Nt, E = (double)Et, lam = k.lam_reg;
  double jv = sv / N, jt = sth / E, jl = sload / E;
  *loss = (float)(sjn / N + sje / E + lam * jv * jv + lam * jt * jt + lam * jl * jl);
  if (!grad_out) return 0;
  float cN = 1.0f / (float)Nt, cE = 1.0f / (float)Et, mv = (float)jv, mth = (float)jt, ml = (float)jl;
  std::vector<float> ap(Nt), aq(Nt), gv(Nt), gth(Nt);
  for (int64_t n = 0; n < Nt; ++n) {
    const WlsBus& b = bus[n];
    ap[n] = -2.0f * k.lam_p * b.R[2] * (b.Z[2] - pb[n]) * cN;
    aq[n] = -2.0f * k.lam_p * b.R[3] * (b.Z[3] - qb[n]) * cN;
    float band = (b.v - 1.1f > 0.f ? 1.f : 0.f) - (0.9f - b.v > 0.f ? 1.f : 0.f);
    gv[n] = -2.0f * k.lam_v * b.R[0] * (b.Z[0] - b.v) * cN + 2.0f * k.lam_reg * mv * cN * band;
    gth[n] = -2.0f * k.lam_v * b.R[1] * (b.Z[1] - b.th) * cN;
  }
  for (int64_t e = 0; e < Et; ++e) {
    int64_t i = ei[e], j = ei[Et + e];
    const float* row = ea + e * 13;
    const WlsBranch& b = br[e];
    float eZ0 = wls_unnorm(row[0], st.es[0], st.em[0]), eR0 = wls_unnorm(row[1], st.es[1], st.em[1]);
    float eZ1 = wls_unnorm(row[2], st.es[2], st.em[2]), eR1 = wls_unnorm(row[3], st.es[3], st.em[3]);
    float dpf = -ap[i] - 2.0f * k.lam_pf * eR0 * (eZ0 - b.pf) * cE;
    float dqf = -aq[i] - 2.0f * k.lam_pf * eR1 * (eZ1 - b.qf) * cE;
    float dpt = -ap[j], dqt = -aq[j];
    float ddelta = (fabsf(b.delta) - 0.5f > 0.f) ? 2.0f * k.lam_reg * mth * cE * (b.delta > 0.f ? 1.f : -1.f) : 0.f;
    float dload = (b.loading - 1.5f > 0.f) ? 2.0f * k.lam_reg * ml * cE : 0.f;
    float dvi, dvj, ddel;
    wls_branch_backward(bin[e], grid, b, dpf, dqf, dpt, dqt, ddelta, dload, dvi, dvj, ddel);
    gv[i] += dvi; gv[j] += dvj; gth[i] += ddel; gth[j] -= ddel;
  }
  for (int64_t n = 0; n < Nt; ++n) {
    grad_out[2 * n] = gv[n] * st.xs[0];
    grad_out[2 * n + 1] = gth[n] * (1.0f - x[n * 11 + 9]);
  }
  return 0;
}

// get_pflow and its adjoint w.r.t. y through wls_branch_forward_delta / wls_pflow_backward (k_pflow, k_pflow_bwd in csrc/wls.cu).
extern "C" int wls_host_pflow(const float* y, const float* node_param, int64_t np_stride, const float* edge_param, const int64_t* ei,
                              int64_t Nt, int64_t Et, int use_shift, const float* gout8, float* out8, float* grad_y) {
  float vmin = 1e30f, vmax = 0.f;
  for (int64_t n = 0; n < Nt; ++n) {
    vmin = fminf(vmin, node_param[n * np_stride]);
    vmax = fmaxf(vmax, node_param[n * np_stride]);
  }
  WlsGrid grid = wls_grid(vmin, vmax);
  for (int64_t n = 0; n < 2 * Nt; ++n) grad_y[n] = 0.f;
  for (int64_t e = 0; e < Et; ++e) {
    int64_t i = ei[e], j = ei[Et + e];
    const float* row = edge_param + e * 7;
    WlsBranchIn in{y[2 * i], y[2 * j], y[2 * i + 1], y[2 * j + 1], row[0], row[1], row[2], row[3], row[5], row[6]};
    WlsBranch b;
    wls_branch_forward_delta(in, grid, b, use_shift ? (in.thi - in.thj) - in.shift : in.thi - in.thj);
    const float o[8] = {b.ll, b.lt, b.pf, b.qf, b.pt, b.qt, b.i_f, b.i_t};
    float go[8];
    for (int q = 0; q < 8; ++q) {
      out8[q * Et + e] = o[q];
      go[q] = gout8[q * Et + e];
    }
    float dvi, dvj, ddel;
    wls_pflow_backward(in, grid, b, go, dvi, dvj, ddel);
    grad_y[2 * i] += dvi; grad_y[2 * j] += dvj; grad_y[2 * i + 1] += ddel; grad_y[2 * j + 1] -= ddel;
  }
  return 0;
}
