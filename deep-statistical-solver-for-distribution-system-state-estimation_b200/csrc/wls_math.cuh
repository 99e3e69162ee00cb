// Per-bus / per-branch arithmetic of the physics-informed WLS loss, forward and adjoint.
//
// Plain inline functions, compiled for the device by wls.cu (with -fmad=false so that every product
// and sum rounds separately, like the reference's eager ops) and for the HOST by the test harness
// tests/host_math/wls_host.cpp (g++ -ffp-contract=off), which lets CPU tests check these exact lines
// against the oracle's autograd without a GPU.  The kernels in wls.cu own everything else
// (tiling, shared memory, reductions).
//
// Follows reference data.py:362-388 (get_pflow) and data.py:397-455 (gsp_wls_edge); the operand
// association of the four flow equations is kept exactly (data.py:370-376) because they cancel
// catastrophically in fp32 (hard part 1).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define WLS_HD __host__ __device__ __forceinline__
#else
#define WLS_HD inline
#endif

struct WlsGrid {      // batch-global constants, data.py:334-338,378
  float v_lv, v_hv;   // min / max of vn_kv over the batch
  float ratio;        // v_hv / v_lv
  float base;         // v_lv ** 2
  float sqrt3;        // fp32 sqrt(3)
};

WLS_HD WlsGrid wls_grid(float v_lv, float v_hv) {
  WlsGrid g;
  g.v_lv = v_lv;
  g.v_hv = v_hv;
  g.ratio = v_hv / v_lv;
  g.base = v_lv * v_lv;
  g.sqrt3 = 1.7320508075688772f;
  return g;
}

struct WlsCoefs {
  float lam_v, lam_p, lam_pf, lam_reg;
};

struct WlsStats {     // x_mean[8], x_std[8], edge_mean[6], edge_std[6]
  float xm[8], xs[8], em[6], es[6];
};

// ---- measurements: un-normalise and mask (data.py:397-409) ----
WLS_HD float wls_unnorm(float z, float sd, float mu) { return z != 0.0f ? (z * sd + mu) : 0.0f; }

struct WlsBranchIn {
  float vi, vj, thi, thj;          // V (pu) and masked theta (rad) at from / to bus
  float G, B, Gs, Bs, shift, rating;  // edge_param columns 0,1,2,3,5,6
};

struct WlsBranch {
  float delta, cs, sn;
  float pf, qf, pt, qt;            // data.py:370-376
  float hf, ht;                    // |P - jQ| at both ends
  float i_f, i_t;                  // data.py:378-383
  float tpos, den;                 // ceil(shift), 1 - tpos*(1-ratio)
  float loading;                   // loading_lines + loading_trafo, data.py:387-388,417
  float ll, lt;
};

// `delta` = Th_i - Th_j - shift as the caller formed it (shift = 0 for phase_shift=True, data.py:362-365)
WLS_HD void wls_branch_forward_delta(const WlsBranchIn& in, const WlsGrid& g, WlsBranch& b, float delta) {
  float vi = in.vi, vj = in.vj;
  b.delta = delta;
  b.cs = cosf(b.delta);
  b.sn = sinf(b.delta);
  float nvv = (-vi) * vj;                          // "- V_i * V_j" binds as (-V_i) * V_j
  float vv = vi * vj;
  float vi2 = vi * vi, vj2 = vj * vj;
  float gsh = in.G + in.Gs / 2.0f;
  float bsh = in.B + in.Bs / 2.0f;
  b.pf = (nvv * (in.G * b.cs + in.B * b.sn) + gsh * vi2) * g.base;
  b.qf = (vv * ((-in.G) * b.sn + in.B * b.cs) - bsh * vi2) * g.base;
  b.pt = (nvv * (in.G * b.cs - in.B * b.sn) + gsh * vj2) * g.base;
  b.qt = (vv * (in.G * b.sn + in.B * b.cs) - bsh * vj2) * g.base;
  b.tpos = ceilf(in.shift);                        // data.py:367 (3 on Oberrhein: reference quirk, kept)
  b.den = 1.0f - (b.tpos * (1.0f - g.ratio));
  b.hf = hypotf(b.pf, -b.qf);
  b.ht = hypotf(b.pt, -b.qt);
  b.i_f = b.hf / ((vi * g.v_lv) * g.sqrt3);
  b.i_f = b.i_f / b.den;
  b.i_t = b.ht / ((vj * g.v_lv) * g.sqrt3);
  b.ll = ((1.0f - b.tpos) * fmaxf(b.i_f, b.i_t)) / in.rating;
  b.lt = (b.tpos * fmaxf(b.i_f * g.v_hv, b.i_t * g.v_lv)) / in.rating;
  b.loading = b.ll + b.lt;
}

WLS_HD void wls_branch_forward(const WlsBranchIn& in, const WlsGrid& g, WlsBranch& b) {
  wls_branch_forward_delta(in, g, b, in.thi - in.thj);   // shift ignored: phase_shift=True, data.py:362-363
}

// Adjoint of wls_branch_forward.  Inputs: adjoints of the four flows as seen by the bus-injection and
// branch-measurement residuals (dpf..dqt), of |delta| penalty (ddelta, already signed) and of `loading`.
// Outputs: adjoints of V_i, V_j and delta.  Conventions match torch autograd: maximum() splits the
// gradient 1/2-1/2 on ties, |z| has gradient 0 at 0, ceil/rating/ratio are constants.
WLS_HD void wls_branch_backward(const WlsBranchIn& in, const WlsGrid& g, const WlsBranch& b, float dpf, float dqf,
                                float dpt, float dqt, float ddelta, float dload, float& dvi, float& dvj, float& ddel) {
  float vi = in.vi, vj = in.vj;
  dvi = 0.0f;
  dvj = 0.0f;
  if (dload != 0.0f) {
    float dm1 = dload * (1.0f - b.tpos) / in.rating;
    float dm2 = dload * b.tpos / in.rating;
    float w1f = b.i_f > b.i_t ? 1.0f : (b.i_f < b.i_t ? 0.0f : 0.5f);
    float a = b.i_f * g.v_hv, c = b.i_t * g.v_lv;
    float w2f = a > c ? 1.0f : (a < c ? 0.0f : 0.5f);
    float di_f = dm1 * w1f + dm2 * w2f * g.v_hv;
    float di_t = dm1 * (1.0f - w1f) + dm2 * (1.0f - w2f) * g.v_lv;
    float dhf = di_f / (((vi * g.v_lv) * g.sqrt3) * b.den);
    float dht = di_t / ((vj * g.v_lv) * g.sqrt3);
    if (b.hf != 0.0f) {
      dpf += dhf * (b.pf / b.hf);
      dqf += dhf * (b.qf / b.hf);
    }
    if (b.ht != 0.0f) {
      dpt += dht * (b.pt / b.ht);
      dqt += dht * (b.qt / b.ht);
    }
    dvi += -di_f * b.i_f / vi;
    dvj += -di_t * b.i_t / vj;
  }
  float c = g.base;
  float A = in.G * b.cs + in.B * b.sn;      // from side
  float Bm = -in.G * b.sn + in.B * b.cs;
  float A2 = in.G * b.cs - in.B * b.sn;     // to side
  float B2 = in.G * b.sn + in.B * b.cs;
  float gsh = in.G + in.Gs / 2.0f;
  float bsh = in.B + in.Bs / 2.0f;
  float vv = vi * vj;
  // P_from
  dvi += dpf * (c * (-vj * A + 2.0f * gsh * vi));
  dvj += dpf * (-c * vi * A);
  ddel = ddelta + dpf * (-c * vv * Bm);
  // Q_from
  dvi += dqf * (c * (vj * Bm - 2.0f * bsh * vi));
  dvj += dqf * (c * vi * Bm);
  ddel += dqf * (-c * vv * A);
  // P_to
  dvi += dpt * (-c * vj * A2);
  dvj += dpt * (c * (-vi * A2 + 2.0f * gsh * vj));
  ddel += dpt * (c * vv * B2);
  // Q_to
  dvi += dqt * (c * vj * B2);
  dvj += dqt * (c * (vi * B2 - 2.0f * bsh * vj));
  ddel += dqt * (c * vv * A2);
}

// Adjoint of get_pflow's eight outputs (data.py:390: loading_lines, loading_trafo, P_from, Q_from, P_to, Q_to, I_from, I_to) w.r.t.
// V_i, V_j and delta: the same chain as wls_branch_backward with the two loading terms and the two currents fed separately.
WLS_HD void wls_pflow_backward(const WlsBranchIn& in, const WlsGrid& g, const WlsBranch& b, const float (&go)[8], float& dvi, float& dvj,
                               float& ddel) {
  float vi = in.vi, vj = in.vj;
  float dpf = go[2], dqf = go[3], dpt = go[4], dqt = go[5];
  dvi = 0.0f;
  dvj = 0.0f;
  float dm1 = go[0] * (1.0f - b.tpos) / in.rating;
  float dm2 = go[1] * b.tpos / in.rating;
  float w1f = b.i_f > b.i_t ? 1.0f : (b.i_f < b.i_t ? 0.0f : 0.5f);
  float a = b.i_f * g.v_hv, c2 = b.i_t * g.v_lv;
  float w2f = a > c2 ? 1.0f : (a < c2 ? 0.0f : 0.5f);
  float di_f = go[6] + dm1 * w1f + dm2 * w2f * g.v_hv;
  float di_t = go[7] + dm1 * (1.0f - w1f) + dm2 * (1.0f - w2f) * g.v_lv;
  if (di_f != 0.0f || di_t != 0.0f) {
    float dhf = di_f / (((vi * g.v_lv) * g.sqrt3) * b.den);
    float dht = di_t / ((vj * g.v_lv) * g.sqrt3);
    if (b.hf != 0.0f) {
      dpf += dhf * (b.pf / b.hf);
      dqf += dhf * (b.qf / b.hf);
    }
    if (b.ht != 0.0f) {
      dpt += dht * (b.pt / b.ht);
      dqt += dht * (b.qt / b.ht);
    }
    dvi += -di_f * b.i_f / vi;
    dvj += -di_t * b.i_t / vj;
  }
  float dv2i, dv2j;
  wls_branch_backward(in, g, b, dpf, dqf, dpt, dqt, 0.0f, 0.0f, dv2i, dv2j, ddel);
  dvi += dv2i;
  dvj += dv2j;
}

// ---- per-bus terms ----
struct WlsBus {
  float Z[4], R[4];   // un-normalised measurements and weights (V, theta, P, Q)
  float v, th;        // model state: V pu, theta rad (slack already masked)
  float slack;
};

// xrow: the 11 columns of x for this bus; out0/out1 the model output
WLS_HD void wls_bus_load(const float* xrow, float out0, float out1, const WlsStats& st, WlsBus& b) {
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int c = 0; c < 4; ++c) {
    b.Z[c] = wls_unnorm(xrow[2 * c], st.xs[2 * c], st.xm[2 * c]);
    b.R[c] = wls_unnorm(xrow[2 * c + 1], st.xs[2 * c + 1], st.xm[2 * c + 1]);
  }
  b.slack = xrow[9];
  b.v = out0 * st.xs[0] + st.xm[0];              // data.py:411
  b.th = out1 * (1.0f - b.slack);                // data.py:412-413
}

// residual term of one bus (data.py:446) and the voltage-band penalty argument (data.py:453)
WLS_HD float wls_bus_residual(const WlsBus& b, float p_bus, float q_bus, const WlsCoefs& k) {
  float d0 = b.Z[0] - b.v, d1 = b.Z[1] - b.th, d2 = b.Z[2] - p_bus, d3 = b.Z[3] - q_bus;
  return (((d0 * d0) * b.R[0]) * k.lam_v + ((d1 * d1) * b.R[1]) * k.lam_v) + ((d2 * d2) * b.R[2]) * k.lam_p +
         ((d3 * d3) * b.R[3]) * k.lam_p;
}
WLS_HD float wls_bus_vband(const WlsBus& b) { return fmaxf(b.v - 1.1f, 0.0f) + fmaxf(0.9f - b.v, 0.0f); }

// residual term of one branch measurement pair (data.py:447)
WLS_HD float wls_branch_residual(float eZ0, float eR0, float eZ1, float eR1, float pf, float qf, const WlsCoefs& k) {
  float d0 = eZ0 - pf, d1 = eZ1 - qf;
  return ((d0 * d0) * eR0) * k.lam_pf + ((d1 * d1) * eR1) * k.lam_pf;
}
