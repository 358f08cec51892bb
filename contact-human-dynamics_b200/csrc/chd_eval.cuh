// Work-item evaluators of the batched NLP (product code, device only).
// One item = one constraint instance (a time node, a spline junction, a spline node) or one cost sample.
// Replaces the per-(node x variable-set) virtual calls of ifopt/TOWR listed in SURVEY.md 8(a) B1-B12:
//   humanoid_dynamic_constraint.cpp:63-143 + humanoid_rigid_body_dynamics.cpp:89-206 (B1,B2)
//   leg_length_constraint.cpp:37-111 (B3), ee_dist_constraint.cpp:27-94 (B4), height_constraint.cpp:23-58 (B5)
//   towr SplineAccConstraint / TerrainConstraint / ForceConstraint (B7-B9)
//   data_cost.cpp:40-96 (B10), vel_smooth_cost.cpp:37-100 (B11)
#pragma once
#include "chd_dev.h"

struct ChdCtx {
  const ChdSeq* h;
  const double* x;           // current point (shared or global memory)
  const double *poly_T, *poly_tend, *node_const, *par, *t_dyn, *t_rom, *t_data, *dur0;
  const int *node_var, *itab, *ent_ptr, *poly_ph;
  int* ent_col;              // Jacobian slot columns (written here once the durations move: dyn)
  const ChdSet* sets;
  int Pmax, F_max, Ph_max;
  int dyn;                   // polynomial boundaries are run-time data: node columns of the located blocks are re-assigned
  int opt_dur;               // stage 3: derivatives with respect to the switch times are wanted
};

__device__ __forceinline__ void chd_make_ctx(const ChdDev& D, int b, const double* x, ChdCtx& c) {
  c.h = D.seq + b;
  c.x = x;
  c.poly_T = D.poly_T + (size_t)b * D.S * D.Pmax;
  c.poly_tend = D.poly_tend + (size_t)b * D.S * D.Pmax;
  c.poly_ph = D.poly_ph + (size_t)b * D.S * D.Pmax;
  c.dur0 = D.dur0 + (size_t)b * D.n_ee_max * D.Ph_max;
  c.ent_col = D.ent_col + (size_t)b * D.slots_max;
  c.Ph_max = D.Ph_max;
  c.dyn = 0;
  c.opt_dur = 0;
  c.node_const = D.node_const + (size_t)b * D.S * (D.Pmax + 1) * 6;
  c.node_var = D.node_var + (size_t)b * D.S * (D.Pmax + 1) * 6;
  c.par = D.par + (size_t)b * D.par_stride;
  c.t_dyn = D.t_dyn + (size_t)b * D.Kd_max;
  c.t_rom = D.t_rom + (size_t)b * D.Kr_max;
  c.t_data = D.t_data + (size_t)b * D.F_max;
  c.itab = D.itab + (size_t)b * D.tab_max;
  c.ent_ptr = D.ent_ptr + (size_t)b * (D.m_max + 1);
  c.sets = D.sets + (size_t)b * D.sets_max;
  c.Pmax = D.Pmax;
  c.F_max = D.F_max;
}

// A spline located at time t: active polynomial, basis weights and the 12 node-value slots
// slot = side*6 + nd*3 + dim  (side: start/end node, nd: 0 position / 1 velocity node value).
struct ChdSpl {
  int poly;
  double tl, T;              // local time and duration of the active polynomial
  ChdBasis B;
  const int* var;
  const double* cst;
};
__device__ __forceinline__ void chd_spl_at(const ChdCtx& c, int s, double t, ChdSpl& o) {
  double tl;
  const int np = c.h->sp_npoly[s];
  o.poly = chd_locate(c.poly_tend + (size_t)s * c.Pmax, np, t, &tl);
  o.tl = tl, o.T = c.poly_T[(size_t)s * c.Pmax + o.poly];
  chd_basis(tl, o.T, o.B);
  o.var = c.node_var + ((size_t)s * (c.Pmax + 1) + o.poly) * 6;
  o.cst = c.node_const + ((size_t)s * (c.Pmax + 1) + o.poly) * 6;
}
__device__ __forceinline__ void chd_spl_poly(const ChdCtx& c, int s, int poly, double tl, ChdSpl& o) {
  o.poly = poly;
  o.tl = tl, o.T = c.poly_T[(size_t)s * c.Pmax + poly];
  chd_basis(tl, o.T, o.B);
  o.var = c.node_var + ((size_t)s * (c.Pmax + 1) + poly) * 6;
  o.cst = c.node_const + ((size_t)s * (c.Pmax + 1) + poly) * 6;
}
__device__ __forceinline__ double chd_nodeval(const ChdCtx& c, const ChdSpl& o, int slot) {
  const int v = o.var[slot];
  return v >= 0 ? c.x[v] : o.cst[slot];
}
// out[dim] = deriv-th time derivative of the spline
__device__ __forceinline__ void chd_spl_val(const ChdCtx& c, const ChdSpl& o, int deriv, double out[3]) {
#pragma unroll
  for (int d = 0; d < 3; ++d)
    out[d] = o.B.w[deriv][0] * chd_nodeval(c, o, d) + o.B.w[deriv][1] * chd_nodeval(c, o, 3 + d) +
             o.B.w[deriv][2] * chd_nodeval(c, o, 6 + d) + o.B.w[deriv][3] * chd_nodeval(c, o, 9 + d);
}
// weight of slot (side, nd) for the deriv-th derivative
__device__ __forceinline__ double chd_slot_w(const ChdSpl& o, int deriv, int slot) { return o.B.w[deriv][(slot / 6) * 2 + ((slot % 6) / 3)]; }

// Derivative of the position of a phase-based spline (foot `ee`, motion or force) with respect to the two switch
// times that bound the active phase k: tau_{k-1} (start) and tau_k (end).  With the phase durations d = D tau this
// is towr's PhaseSpline::GetJacobianOfPosWrtDurations / PhaseDurations::GetJacobianOfPos after the
// change of variables: the dense "every earlier phase" columns cancel and two columns remain,
//   d p / d tau_{k-1} = -v - inner,   d p / d tau_k = inner,   inner = (dp/dT_poly - k_prev v) / n_polys.
// va / vb: variable index of tau_{k-1} / tau_k (the slot of d_{k-1} / d_k in x), -1 when fixed (0 and T).
struct ChdTau {
  int va, vb;
  double da[3], db[3];
};
__device__ __forceinline__ void chd_spl_tau(const ChdCtx& c, int s, int ee, const ChdSpl& o, ChdTau& u) {
  const int info = c.poly_ph[(size_t)s * c.Pmax + o.poly];
  const int k = info & 4095, kprev = (info >> 12) & 255, npol = info >> 20;
  const double t = o.tl, T = o.T, iT = 1.0 / T, t2 = t * t * iT * iT, t3 = t2 * t * iT;   // (t/T)^2, (t/T)^3
  // d(basis weight)/dT at fixed local time
  const double wT[4] = {6.0 * (t2 - t3) * iT, 2.0 * (t2 - t3), 6.0 * (t3 - t2) * iT, t2 - 2.0 * t3};
  double v[3];
  chd_spl_val(c, o, 1, v);
  const double inv = 1.0 / npol;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double dT = wT[0] * chd_nodeval(c, o, d) + wT[1] * chd_nodeval(c, o, 3 + d) + wT[2] * chd_nodeval(c, o, 6 + d) +
                      wT[3] * chd_nodeval(c, o, 9 + d);
    const double inner = inv * (dT - kprev * v[d]);
    u.da[d] = -v[d] - inner;
    u.db[d] = inner;
  }
  const int P = c.h->n_phases[ee];
  u.va = k >= 1 ? c.h->dur_xoff[ee] + k - 1 : -1;
  u.vb = k <= P - 2 ? c.h->dur_xoff[ee] + k : -1;
}
// columns of a located block of a phase-based spline (run-time pattern)
__device__ __forceinline__ void chd_put_cols(const ChdCtx& c, int e0, const ChdSpl& o) {
#pragma unroll
  for (int q = 0; q < 12; ++q) c.ent_col[e0 + q] = o.var[q];
}

__device__ __forceinline__ int chd_frame_index(const ChdSeq* h, double t) {  // humanoid_rigid_body_dynamics.cpp:81-87
  int idx = (int)((t / h->T) * h->F);
  if (idx == h->F) idx -= 1;
  return idx;
}
__device__ __forceinline__ void chd_inertia(const ChdCtx& c, double t, double Ib[9]) {
  const double* I = c.par + 6 * c.F_max + 6 * chd_frame_index(c.h, t);
  Ib[0] = I[0]; Ib[1] = I[3]; Ib[2] = I[4];
  Ib[3] = I[3]; Ib[4] = I[1]; Ib[5] = I[5];
  Ib[6] = I[4]; Ib[7] = I[5]; Ib[8] = I[2];
}
__device__ __forceinline__ const double* chd_hip(const ChdCtx& c, int ee, double t) {  // humanoid.h:45-48
  return c.par + ((ee == 0 || ee == 2) ? 0 : 3 * c.F_max) + 3 * chd_frame_index(c.h, t);
}

// ------------------------------------------------------------------ constraint items ------------
template <bool JAC>
__device__ void chd_item_acc(const ChdCtx& c, const ChdSet& st, int j, double* g, double* Jv) {
  const int s = st.a;
  const double T0 = c.poly_T[(size_t)s * c.Pmax + j], T1 = c.poly_T[(size_t)s * c.Pmax + j + 1];
  ChdBasis A, Bn;
  chd_basis(T0, T0, A);
  chd_basis(0.0, T1, Bn);
  const int* var = c.node_var + ((size_t)s * (c.Pmax + 1) + j) * 6;
  const double* cst = c.node_const + ((size_t)s * (c.Pmax + 1) + j) * 6;
  const double wv[6] = {A.w[2][0], A.w[2][1], A.w[2][2] - Bn.w[2][0], A.w[2][3] - Bn.w[2][1], -Bn.w[2][2], -Bn.w[2][3]};
  for (int d = 0; d < 3; ++d) {
    const int R = st.row0 + 3 * j + d;
    double acc = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int nd = 0; nd < 2; ++nd) {
        const int q = a * 6 + nd * 3 + d;
        const int v = var[q];
        acc += wv[a * 2 + nd] * (v >= 0 ? c.x[v] : cst[q]);
      }
    g[R] = acc;
    if (JAC) {
      double* J = Jv + c.ent_ptr[R];
#pragma unroll
      for (int q = 0; q < 6; ++q) J[q] = wv[q];
    }
  }
}

template <bool JAC>
__device__ void chd_item_terrain(const ChdCtx& c, const ChdSet& st, int it, double* g, double* Jv) {
  const int s = chd_sp_motion(st.a), node = it + 1, R = st.row0 + it;
  const int* var = c.node_var + ((size_t)s * (c.Pmax + 1) + node) * 6;
  const double* cst = c.node_const + ((size_t)s * (c.Pmax + 1) + node) * 6;
  double p[3];
  for (int d = 0; d < 3; ++d) p[d] = var[d] >= 0 ? c.x[var[d]] : cst[d];
  const ChdSeq* h = c.h;
  double z = -h->normal[1] * (p[1] - h->point[1]) - h->normal[0] * (p[0] - h->point[0]);  // ground_plane.cpp:18-26
  z /= h->normal[2];
  z += h->point[2];
  g[R] = p[2] - z;
  if (JAC) {
    double* J = Jv + c.ent_ptr[R];
    J[0] = -h->dhdx, J[1] = -h->dhdy, J[2] = 1.0;
  }
}

template <bool JAC>
__device__ void chd_item_force(const ChdCtx& c, const ChdSet& st, int it, double* g, double* Jv) {
  const ChdSeq* h = c.h;
  const int s = chd_sp_force(h->n_ee, st.a), node = c.itab[st.tab + it];
  const int* var = c.node_var + ((size_t)s * (c.Pmax + 1) + node) * 6;
  const double* cst = c.node_const + ((size_t)s * (c.Pmax + 1) + node) * 6;
  double f[3];
  for (int d = 0; d < 3; ++d) f[d] = var[d] >= 0 ? c.x[var[d]] : cst[d];
  for (int k = 0; k < 5; ++k) {
    double dir[3];
    for (int d = 0; d < 3; ++d) {
      const double tn = (k == 1 || k == 2) ? h->tan1[d] : h->tan2[d];
      dir[d] = k == 0 ? h->nrm[d] : (tn + ((k == 1 || k == 3) ? -h->mu : h->mu) * h->nrm[d]);
    }
    const int R = st.row0 + 5 * it + k;
    g[R] = chd_dot(f, dir);
    if (JAC) {
      double* J = Jv + c.ent_ptr[R];
      J[0] = dir[0], J[1] = dir[1], J[2] = dir[2];
    }
  }
}

template <bool JAC>
__device__ void chd_item_rom(const ChdCtx& c, const ChdSet& st, int k, double* g, double* Jv) {
  const double t = c.t_rom[k];
  const int R = st.row0 + k;
  ChdSpl L, A, E;
  chd_spl_at(c, 0, t, L);
  chd_spl_at(c, 1, t, A);
  chd_spl_at(c, chd_sp_motion(st.a), t, E);
  double cp[3], e[3], pe[3], Rm[9], Rh[3], d[3];
  chd_spl_val(c, L, 0, cp);
  chd_spl_val(c, A, 0, e);
  chd_spl_val(c, E, 0, pe);
  const double* hip = chd_hip(c, st.a, t);
  ChdTrig tr;
  chd_trig(e, tr);
  chd_R(tr, Rm);
  chd_mv(Rm, hip, Rh);
  for (int q = 0; q < 3; ++q) d[q] = pe[q] - (Rh[q] + cp[q]);
  g[R] = 0.5 * chd_dot(d, d);
  if (JAC) {
    double* J = Jv + c.ent_ptr[R];
    double dRh_d[3];  // (dR/de_j h) . d
    for (int j = 0; j < 3; ++j) {
      double Dj[9], v[3];
      chd_dR(tr, j, Dj);
      chd_mv(Dj, hip, v);
      dRh_d[j] = chd_dot(v, d);
    }
#pragma unroll
    for (int q = 0; q < 12; ++q) {
      const int dim = q % 3;
      J[q] = -chd_slot_w(L, 0, q) * d[dim];
      J[12 + q] = -chd_slot_w(A, 0, q) * dRh_d[dim];
      J[24 + q] = chd_slot_w(E, 0, q) * d[dim];
    }
    if (c.dyn) {
      const int e0 = c.ent_ptr[R];
      chd_put_cols(c, e0 + 24, E);
      ChdTau u;
      u.va = u.vb = -1;
      u.da[0] = u.da[1] = u.da[2] = u.db[0] = u.db[1] = u.db[2] = 0.0;
      if (c.opt_dur) chd_spl_tau(c, chd_sp_motion(st.a), st.a, E, u);   // leg_length_constraint.cpp:105
      c.ent_col[e0 + 36] = u.va, c.ent_col[e0 + 37] = u.vb;
      J[36] = chd_dot(d, u.da), J[37] = chd_dot(d, u.db);
    } else {
      J[36] = J[37] = 0.0;
    }
  }
}

template <bool JAC>
__device__ void chd_item_heel(const ChdCtx& c, const ChdSet& st, int k, double* g, double* Jv) {
  const double t = c.t_rom[k];
  const int R = st.row0 + k;
  ChdSpl Ea, Eb;
  chd_spl_at(c, chd_sp_motion(st.a), t, Ea);
  chd_spl_at(c, chd_sp_motion(st.b), t, Eb);
  double pa[3], pb[3], d[3];
  chd_spl_val(c, Ea, 0, pa);
  chd_spl_val(c, Eb, 0, pb);
  for (int q = 0; q < 3; ++q) d[q] = pa[q] - pb[q];
  g[R] = 0.5 * chd_dot(d, d);
  if (JAC) {
    double* J = Jv + c.ent_ptr[R];
#pragma unroll
    for (int q = 0; q < 12; ++q) {
      J[q] = chd_slot_w(Ea, 0, q) * d[q % 3];
      J[12 + q] = -chd_slot_w(Eb, 0, q) * d[q % 3];
    }
    J[24] = J[25] = J[26] = J[27] = 0.0;
    if (c.dyn) {
      const int e0 = c.ent_ptr[R];
      chd_put_cols(c, e0, Ea);
      chd_put_cols(c, e0 + 12, Eb);
      for (int q = 24; q < 28; ++q) c.ent_col[e0 + q] = -1;
      if (c.opt_dur) {   // ee_dist_constraint.cpp:76,88
        ChdTau ua, ub;
        chd_spl_tau(c, chd_sp_motion(st.a), st.a, Ea, ua);
        chd_spl_tau(c, chd_sp_motion(st.b), st.b, Eb, ub);
        c.ent_col[e0 + 24] = ua.va, c.ent_col[e0 + 25] = ua.vb, c.ent_col[e0 + 26] = ub.va, c.ent_col[e0 + 27] = ub.vb;
        J[24] = chd_dot(d, ua.da), J[25] = chd_dot(d, ua.db), J[26] = -chd_dot(d, ub.da), J[27] = -chd_dot(d, ub.db);
      }
    }
  }
}

template <bool JAC>
__device__ void chd_item_height(const ChdCtx& c, const ChdSet& st, int k, double* g, double* Jv) {
  const double t = c.t_dyn[k];
  const int R = st.row0 + k;
  ChdSpl E;
  chd_spl_at(c, chd_sp_motion(st.a), t, E);
  double p[3], q3[3];
  chd_spl_val(c, E, 0, p);
  for (int q = 0; q < 3; ++q) q3[q] = p[q] - c.h->point[q];
  g[R] = chd_dot(c.h->normal, q3);
  if (JAC) {
    double* J = Jv + c.ent_ptr[R];
#pragma unroll
    for (int q = 0; q < 12; ++q) J[q] = chd_slot_w(E, 0, q) * c.h->normal[q % 3];
    J[12] = J[13] = 0.0;
    if (c.dyn) {
      const int e0 = c.ent_ptr[R];
      chd_put_cols(c, e0, E);
      c.ent_col[e0 + 12] = c.ent_col[e0 + 13] = -1;
      if (c.opt_dur) {   // height_constraint.cpp:43
        ChdTau u;
        chd_spl_tau(c, chd_sp_motion(st.a), st.a, E, u);
        c.ent_col[e0 + 12] = u.va, c.ent_col[e0 + 13] = u.vb;
        J[12] = chd_dot(c.h->normal, u.da), J[13] = chd_dot(c.h->normal, u.db);
      }
    }
  }
}

// total_duration_constraint.cpp:60-82: sum of the free durations of foot st.a (= its last switch time)
template <bool JAC>
__device__ void chd_item_tottime(const ChdCtx& c, const ChdSet& st, double* g, double* Jv) {
  const int ee = st.a, P = c.h->n_phases[ee];
  double sum = 0.0;
  for (int k = 0; k < P - 1; ++k) sum += c.x[c.h->dur_xoff[ee] + k];
  g[st.row0] = sum;
  if (JAC) Jv[c.ent_ptr[st.row0]] = 1.0;
}
// PhaseDurations lower bound as a row: d_k = tau_k - tau_{k-1} >= 0
template <bool JAC>
__device__ void chd_item_durpos(const ChdCtx& c, const ChdSet& st, int k, double* g, double* Jv) {
  const int R = st.row0 + k;
  g[R] = c.x[c.h->dur_xoff[st.a] + k];
  if (JAC) {
    double* J = Jv + c.ent_ptr[R];
    J[0] = 1.0, J[1] = -1.0;
  }
}

__device__ __forceinline__ void chd_Iw_apply(const double Rm[9], const double Ib[9], const double v[3], double out[3]) {
  double a[3], b[3];
  chd_mtv(Rm, v, a);
  chd_mv(Ib, a, b);
  chd_mv(Rm, b, out);
}
__device__ __forceinline__ void chd_dIw_apply(const double Rm[9], const double Dj[9], const double Ib[9], const double v[3], double out[3]) {
  double a[3], b[3], c1[3], c2[3];
  chd_mtv(Rm, v, a);
  chd_mv(Ib, a, b);
  chd_mv(Dj, b, c1);   // D_j I_b R^T v
  chd_mtv(Dj, v, a);
  chd_mv(Ib, a, b);
  chd_mv(Rm, b, c2);   // R I_b D_j^T v
  for (int q = 0; q < 3; ++q) out[q] = c1[q] + c2[q];
}

// Dynamics work item, split in parts so that several threads share one time node:
//   part 0: residual (6 rows) + base-linear Jacobian block, part 1: base-angular block, part 2+ee: ee blocks
template <bool JAC>
__device__ void chd_item_dyn(const ChdCtx& c, const ChdSet& st, int k, int part, double* g, double* Jv) {
  const ChdSeq* h = c.h;
  const int n_ee = h->n_ee;
  const double t = c.t_dyn[k];
  const int R0 = st.row0 + 6 * k;
  if (part == 0) {
    ChdSpl L, A;
    chd_spl_at(c, 0, t, L);
    chd_spl_at(c, 1, t, A);
    double cp[3], ca[3], e[3], ed[3], edd[3];
    chd_spl_val(c, L, 0, cp);
    chd_spl_val(c, L, 2, ca);
    chd_spl_val(c, A, 0, e);
    chd_spl_val(c, A, 1, ed);
    chd_spl_val(c, A, 2, edd);
    ChdTrig tr;
    chd_trig(e, tr);
    double Rm[9], M[9], dMy[9], dMz[9], Ib[9];
    chd_R(tr, Rm);
    chd_M(tr, M);
    chd_dM_y(tr, dMy);
    chd_dM_z(tr, dMz);
    chd_inertia(c, t, Ib);
    double om[3], omd[3], tmp[3], Md[9];
    chd_mv(M, ed, om);
    for (int q = 0; q < 9; ++q) Md[q] = dMy[q] * ed[1] + dMz[q] * ed[2];
    chd_mv(Md, ed, omd);
    chd_mv(M, edd, tmp);
    for (int q = 0; q < 3; ++q) omd[q] += tmp[q];
    double Iw_om[3], Iw_omd[3], gyro[3];
    chd_Iw_apply(Rm, Ib, om, Iw_om);
    chd_Iw_apply(Rm, Ib, omd, Iw_omd);
    chd_cross(om, Iw_om, gyro);
    double fsum[3] = {0, 0, 0}, tau[3] = {0, 0, 0};
    for (int ee = 0; ee < n_ee; ++ee) {
      ChdSpl Fs, Es;
      chd_spl_at(c, chd_sp_force(n_ee, ee), t, Fs);
      chd_spl_at(c, chd_sp_motion(ee), t, Es);
      double f[3], p[3], r[3], tq[3];
      chd_spl_val(c, Fs, 0, f);
      chd_spl_val(c, Es, 0, p);
      for (int q = 0; q < 3; ++q) r[q] = cp[q] - p[q], fsum[q] += f[q];
      chd_cross(f, r, tq);
      for (int q = 0; q < 3; ++q) tau[q] += tq[q];
    }
    for (int q = 0; q < 3; ++q) {
      g[R0 + q] = Iw_omd[q] + gyro[q] - tau[q];
      g[R0 + 3 + q] = h->mass * ca[q] - fsum[q] - h->mass * h->grav * h->gvec[q];
    }
    if (JAC) {
      // base-linear block: angular rows -skew(sum f) B_pos, linear rows m B_acc
      const double S[9] = {0, -fsum[2], fsum[1], fsum[2], 0, -fsum[0], -fsum[1], fsum[0], 0};
      for (int r = 0; r < 3; ++r) {
        double* Ja = Jv + c.ent_ptr[R0 + r];
        double* Jl = Jv + c.ent_ptr[R0 + 3 + r];
#pragma unroll
        for (int q = 0; q < 12; ++q) {
          const int dim = q % 3;
          Ja[q] = -S[r * 3 + dim] * chd_slot_w(L, 0, q);
          Jl[q] = dim == r ? h->mass * chd_slot_w(L, 2, q) : 0.0;
        }
      }
    }
    return;
  }
  if (!JAC) return;
  if (part == 1) {
    ChdSpl A;
    chd_spl_at(c, 1, t, A);
    double e[3], ed[3], edd[3];
    chd_spl_val(c, A, 0, e);
    chd_spl_val(c, A, 1, ed);
    chd_spl_val(c, A, 2, edd);
    ChdTrig tr;
    chd_trig(e, tr);
    double Rm[9], M[9], dMy[9], dMz[9], Dyy[9], Dyz[9], Dzz[9], Ib[9], Md[9];
    chd_R(tr, Rm);
    chd_M(tr, M);
    chd_dM_y(tr, dMy);
    chd_dM_z(tr, dMz);
    chd_d2M(tr, Dyy, Dyz, Dzz);
    chd_inertia(c, t, Ib);
    double om[3], omd[3], tmp[3];
    chd_mv(M, ed, om);
    for (int q = 0; q < 9; ++q) Md[q] = dMy[q] * ed[1] + dMz[q] * ed[2];
    chd_mv(Md, ed, omd);
    chd_mv(M, edd, tmp);
    for (int q = 0; q < 3; ++q) omd[q] += tmp[q];
    double Iw_om[3];
    chd_Iw_apply(Rm, Ib, om, Iw_om);
    double Ge[9], Ged[9], Gedd[9];
    for (int j = 0; j < 3; ++j) {
      double Dj[9], dMj[9], dMdj[9];
      chd_dR(tr, j, Dj);
      for (int q = 0; q < 9; ++q) {
        dMj[q] = j == 0 ? 0.0 : (j == 1 ? dMy[q] : dMz[q]);
        dMdj[q] = j == 0 ? 0.0 : (j == 1 ? Dyy[q] * ed[1] + Dyz[q] * ed[2] : Dyz[q] * ed[1] + Dzz[q] * ed[2]);
      }
      double wj[3], wdj[3], t1[3], t2[3], t3[3], t4[3], t5[3], t6[3];
      chd_mv(dMj, ed, wj);
      chd_mv(dMdj, ed, wdj);
      chd_mv(dMj, edd, t1);
      for (int q = 0; q < 3; ++q) wdj[q] += t1[q];
      chd_dIw_apply(Rm, Dj, Ib, omd, t1);   // dIw_j omega_dot
      chd_Iw_apply(Rm, Ib, wdj, t2);        // Iw d(omega_dot)/de_j
      chd_cross(wj, Iw_om, t3);             // d(omega)/de_j x Iw omega
      chd_dIw_apply(Rm, Dj, Ib, om, t4);
      chd_Iw_apply(Rm, Ib, wj, t5);
      for (int q = 0; q < 3; ++q) t4[q] += t5[q];
      chd_cross(om, t4, t6);                // omega x d(Iw omega)/de_j
      for (int q = 0; q < 3; ++q) Ge[q * 3 + j] = t1[q] + t2[q] + t3[q] + t6[q];
      // wrt Euler rates
      double Wd[3], Mj[3] = {M[j], M[3 + j], M[6 + j]};
      chd_mv(dMj, ed, Wd);
      for (int q = 0; q < 3; ++q) Wd[q] += Md[q * 3 + j];
      chd_Iw_apply(Rm, Ib, Wd, t1);
      chd_cross(Mj, Iw_om, t2);
      chd_Iw_apply(Rm, Ib, Mj, t3);
      chd_cross(om, t3, t4);
      for (int q = 0; q < 3; ++q) {
        Ged[q * 3 + j] = t1[q] + t2[q] + t4[q];
        Gedd[q * 3 + j] = t3[q];
      }
    }
    for (int r = 0; r < 3; ++r) {
      double* Ja = Jv + c.ent_ptr[R0 + r] + 12;
      double* Jl = Jv + c.ent_ptr[R0 + 3 + r] + 12;
#pragma unroll
      for (int q = 0; q < 12; ++q) {
        const int dim = q % 3;
        Ja[q] = Ge[r * 3 + dim] * chd_slot_w(A, 0, q) + Ged[r * 3 + dim] * chd_slot_w(A, 1, q) + Gedd[r * 3 + dim] * chd_slot_w(A, 2, q);
        Jl[q] = 0.0;
      }
    }
    return;
  }
  {
    const int ee = part - 2;
    ChdSpl L, Fs, Es;
    chd_spl_at(c, 0, t, L);
    chd_spl_at(c, chd_sp_force(n_ee, ee), t, Fs);
    chd_spl_at(c, chd_sp_motion(ee), t, Es);
    double cp[3], f[3], p[3], r[3];
    chd_spl_val(c, L, 0, cp);
    chd_spl_val(c, Fs, 0, f);
    chd_spl_val(c, Es, 0, p);
    for (int q = 0; q < 3; ++q) r[q] = cp[q] - p[q];
    const double Sf[9] = {0, -f[2], f[1], f[2], 0, -f[0], -f[1], f[0], 0};
    const double Sr[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
    for (int rr = 0; rr < 3; ++rr) {
      double* Ja = Jv + c.ent_ptr[R0 + rr] + 24 + 24 * ee;
      double* Jl = Jv + c.ent_ptr[R0 + 3 + rr] + 24 + 24 * ee;
#pragma unroll
      for (int q = 0; q < 12; ++q) {
        const int dim = q % 3;
        Ja[q] = Sf[rr * 3 + dim] * chd_slot_w(Es, 0, q);          // ee position nodes
        Jl[q] = 0.0;
        Ja[12 + q] = Sr[rr * 3 + dim] * chd_slot_w(Fs, 0, q);     // force nodes
        Jl[12 + q] = dim == rr ? -chd_slot_w(Fs, 0, q) : 0.0;
      }
    }
    // switch-time slots of this foot (humanoid_dynamic_constraint.cpp:112-118): two columns shared by the motion and
    // the force spline (they run on the same PhaseDurations)
    const int toff = 24 + 24 * n_ee + 2 * ee;
    ChdTau up, uf;
    up.va = up.vb = -1;
    for (int q = 0; q < 3; ++q) up.da[q] = up.db[q] = uf.da[q] = uf.db[q] = 0.0;
    if (c.dyn && c.opt_dur) {
      chd_spl_tau(c, chd_sp_motion(ee), ee, Es, up);
      chd_spl_tau(c, chd_sp_force(n_ee, ee), ee, Fs, uf);
    }
    for (int rr = 0; rr < 3; ++rr) {
      double* Ja = Jv + c.ent_ptr[R0 + rr] + toff;
      double* Jl = Jv + c.ent_ptr[R0 + 3 + rr] + toff;
      Ja[0] = Sf[rr * 3] * up.da[0] + Sf[rr * 3 + 1] * up.da[1] + Sf[rr * 3 + 2] * up.da[2] +
              Sr[rr * 3] * uf.da[0] + Sr[rr * 3 + 1] * uf.da[1] + Sr[rr * 3 + 2] * uf.da[2];
      Ja[1] = Sf[rr * 3] * up.db[0] + Sf[rr * 3 + 1] * up.db[1] + Sf[rr * 3 + 2] * up.db[2] +
              Sr[rr * 3] * uf.db[0] + Sr[rr * 3 + 1] * uf.db[1] + Sr[rr * 3 + 2] * uf.db[2];
      Jl[0] = -uf.da[rr];
      Jl[1] = -uf.db[rr];
    }
    if (c.dyn) {
      for (int rr = 0; rr < 6; ++rr) {
        const int e0 = c.ent_ptr[R0 + rr];
        chd_put_cols(c, e0 + 24 + 24 * ee, Es);
        chd_put_cols(c, e0 + 36 + 24 * ee, Fs);
        c.ent_col[e0 + toff] = up.va, c.ent_col[e0 + toff + 1] = up.vb;
      }
    }
  }
}

// ------------------------------------------------------------------ cost items -------------------
// data sample i of spline s (data_cost.cpp:40-96); returns the sample's cost, adds to grad (shared/global atomics)
template <bool GRAD>
__device__ double chd_item_data(const ChdCtx& c, int s, int i, double w, double* grad) {
  const double* data = c.par + (s == 0 ? 12 : (s == 1 ? 15 : 18 + 3 * (s - 2))) * c.F_max + 3 * i;
  ChdSpl P;
  chd_spl_at(c, s, c.t_data[i], P);
  double p[3], diff[3];
  chd_spl_val(c, P, 0, p);
  for (int q = 0; q < 3; ++q) diff[q] = data[q] - p[q];
  if (GRAD) {
#pragma unroll
    for (int q = 0; q < 12; ++q) {
      const int v = P.var[q];
      if (v >= 0) atomicAdd(grad + v, -w * chd_slot_w(P, 0, q) * diff[q % 3]);
    }
    if (c.opt_dur && s >= 2) {   // data_cost.cpp:59-75: durations gradient of the foot targets
      ChdTau u;
      chd_spl_tau(c, s, s - 2, P, u);
      if (u.va >= 0) atomicAdd(grad + u.va, -w * chd_dot(diff, u.da));
      if (u.vb >= 0) atomicAdd(grad + u.vb, -w * chd_dot(diff, u.db));
    }
  }
  return 0.5 * w * chd_dot(diff, diff);
}
// smoothing sample i of spline s, deriv 0 ("velocity smoothing") or 1 ("acceleration smoothing")
// (vel_smooth_cost.cpp:37-100)
template <bool GRAD>
__device__ double chd_item_smooth(const ChdCtx& c, int s, int i, int deriv, double w, double* grad) {
  const double t = c.t_data[i];
  ChdSpl P0, P1;
  chd_spl_at(c, s, t, P0);
  chd_spl_at(c, s, t + c.h->dt, P1);
  double a[3], b[3], diff[3];
  chd_spl_val(c, P0, deriv, a);
  chd_spl_val(c, P1, deriv, b);
  for (int q = 0; q < 3; ++q) diff[q] = b[q] - a[q];
  if (GRAD) {
#pragma unroll
    for (int q = 0; q < 12; ++q) {
      const int v1 = P1.var[q], v0 = P0.var[q];
      if (v1 >= 0) atomicAdd(grad + v1, w * chd_slot_w(P1, deriv, q) * diff[q % 3]);
      if (v0 >= 0) atomicAdd(grad + v0, -w * chd_slot_w(P0, deriv, q) * diff[q % 3]);
    }
    if (c.opt_dur && s >= 2 && deriv == 0) {   // vel_smooth_cost.cpp:56-70 (positions only: :72-79 throws for velocities)
      ChdTau u1, u0;
      chd_spl_tau(c, s, s - 2, P1, u1);
      chd_spl_tau(c, s, s - 2, P0, u0);
      if (u1.va >= 0) atomicAdd(grad + u1.va, w * chd_dot(diff, u1.da));
      if (u1.vb >= 0) atomicAdd(grad + u1.vb, w * chd_dot(diff, u1.db));
      if (u0.va >= 0) atomicAdd(grad + u0.va, -w * chd_dot(diff, u0.da));
      if (u0.vb >= 0) atomicAdd(grad + u0.vb, -w * chd_dot(diff, u0.db));
    }
  }
  return 0.5 * w * chd_dot(diff, diff);
}

// DurationCost (duration_cost.cpp:25-50), term k of foot ee: 1/2 w (d0_k - d_k)^2; gradient in switch-time space
// (tau_k collects +g_k, tau_{k-1} collects -g_k)
template <bool GRAD>
__device__ double chd_item_durcost(const ChdCtx& c, int ee, int k, double w, double* grad) {
  const int v = c.h->dur_xoff[ee] + k;
  const double diff = c.dur0[(size_t)ee * c.Ph_max + k] - c.x[v];
  if (GRAD) {
    atomicAdd(grad + v, -w * diff);
    if (k > 0) atomicAdd(grad + v - 1, w * diff);
  }
  return 0.5 * w * diff * diff;
}

// Spline tables of one sequence from the phase durations held in x (stage 3 onwards; before that the host-built
// tables are exact).  Same arithmetic as towr: last duration = total - sum of the free ones (PhaseDurations),
// polynomial duration = phase duration / polynomials in the phase, end times accumulated sequentially.
// One thread per phase-based spline plus one per foot for the phase end times; the caller synchronises.
__device__ __forceinline__ void chd_tables_from_x(const ChdCtx& c, const double* x, double* polyT, double* polyTend, double* phase_tend) {
  const ChdSeq* h = c.h;
  const int n_ee = h->n_ee;
  for (int w = threadIdx.x; w < 3 * n_ee; w += blockDim.x) {
    const int ee = w % n_ee, kind = w / n_ee;   // 0 motion spline, 1 force spline, 2 phase end times
    const int P = h->n_phases[ee];
    double sum = 0.0, tot = 0.0;
    for (int k = 0; k < P - 1; ++k) sum += x[h->dur_xoff[ee] + k];
    for (int k = 0; k < P; ++k) tot += c.dur0[(size_t)ee * c.Ph_max + k];
    const double last = tot - sum;
    if (kind == 2) {
      if (phase_tend) {
        double t = 0.0;
        for (int k = 0; k < P; ++k) t += k < P - 1 ? x[h->dur_xoff[ee] + k] : last, phase_tend[(size_t)ee * c.Ph_max + k] = t;
      }
      continue;
    }
    const int s = kind == 0 ? chd_sp_motion(ee) : chd_sp_force(n_ee, ee);
    double t = 0.0;
    for (int j = 0; j < h->sp_npoly[s]; ++j) {
      const int info = c.poly_ph[(size_t)s * c.Pmax + j];
      const int ph = info & 4095, npol = info >> 20;
      const double Tj = (ph < P - 1 ? x[h->dur_xoff[ee] + ph] : last) / npol;
      t += Tj;
      polyT[(size_t)s * c.Pmax + j] = Tj;
      polyTend[(size_t)s * c.Pmax + j] = t;
    }
  }
}

// number of work items of a set (dynamics nodes are split in 2 + n_ee parts when the Jacobian is wanted)
__device__ __forceinline__ int chd_set_work(const ChdSet& st, int n_ee, bool jac) {
  return st.type == CHD_SET_DYN ? st.nitems * (jac ? 2 + n_ee : 1) : st.nitems;
}

// Evaluates every active constraint item and cost sample of one sequence with the whole CTA.
// g: m doubles (inactive rows untouched), Jv: slot values, grad: n doubles (must be zeroed; atomics),
// returns the cost in *cost_out (thread 0).  red: shared scratch of CHD_THREADS doubles.
template <bool JAC>
__device__ void chd_eval_all(const ChdCtx& c, const ChdStageDev& sg, double* g, double* Jv, double* grad, double* cost_out,
                             double* red) {
  const ChdSeq* h = c.h;
  const int n_ee = h->n_ee;
  for (int si = 0; si < h->nsets; ++si) {
    const ChdSet st = c.sets[si];
    if (!(sg.set_mask & CHD_MASK(st.type))) continue;
    const int work = chd_set_work(st, n_ee, JAC);
    for (int it = threadIdx.x; it < work; it += blockDim.x) {
      switch (st.type) {
        case CHD_SET_ACC: chd_item_acc<JAC>(c, st, it, g, Jv); break;
        case CHD_SET_TERRAIN: chd_item_terrain<JAC>(c, st, it, g, Jv); break;
        case CHD_SET_ROM: chd_item_rom<JAC>(c, st, it, g, Jv); break;
        case CHD_SET_DYN:
          if (JAC) chd_item_dyn<JAC>(c, st, it / (2 + n_ee), it % (2 + n_ee), g, Jv);
          else chd_item_dyn<JAC>(c, st, it, 0, g, Jv);
          break;
        case CHD_SET_FORCE: chd_item_force<JAC>(c, st, it, g, Jv); break;
        case CHD_SET_HEEL: chd_item_heel<JAC>(c, st, it, g, Jv); break;
        case CHD_SET_HEIGHT: chd_item_height<JAC>(c, st, it, g, Jv); break;
        case CHD_SET_TOTTIME: chd_item_tottime<JAC>(c, st, g, Jv); break;
        case CHD_SET_DURPOS: chd_item_durpos<JAC>(c, st, it, g, Jv); break;
      }
    }
  }
  // costs: splines 0 .. 1+n_ee carry data + smoothing terms
  double acc = 0.0;
  const int nsp = 2 + n_ee, F = h->F, ns = h->n_smooth;
  for (int it = threadIdx.x; it < nsp * F; it += blockDim.x) {
    const int s = it / F, i = it % F, cls = s < 2 ? s : 2;
    if (sg.w_data[cls] != 0.0) acc += chd_item_data<JAC>(c, s, i, sg.w_data[cls], grad);
    if (i < ns) {
      if (sg.w_vel[cls] != 0.0) acc += chd_item_smooth<JAC>(c, s, i, 0, sg.w_vel[cls], grad);
      if (sg.w_acc[cls] != 0.0) acc += chd_item_smooth<JAC>(c, s, i, 1, sg.w_acc[cls], grad);
    }
  }
  if (sg.w_dur != 0.0 && h->n_dur) {
    for (int ee = 0, base = 0; ee < n_ee; base += h->n_phases[ee] - 1, ++ee)
      for (int k = threadIdx.x; k < h->n_phases[ee] - 1; k += blockDim.x) acc += chd_item_durcost<JAC>(c, ee, k, sg.w_dur, grad);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) *cost_out = red[0];
  __syncthreads();
}
