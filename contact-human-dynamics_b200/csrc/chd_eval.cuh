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
  const double *poly_T, *poly_tend, *node_const, *par, *t_dyn, *t_rom, *t_data;
  const int *node_var, *itab, *ent_ptr;
  const ChdSet* sets;
  int Pmax, F_max;
};

__device__ __forceinline__ void chd_make_ctx(const ChdDev& D, int b, const double* x, ChdCtx& c) {
  c.h = D.seq + b;
  c.x = x;
  c.poly_T = D.poly_T + (size_t)b * D.S * D.Pmax;
  c.poly_tend = D.poly_tend + (size_t)b * D.S * D.Pmax;
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
  ChdBasis B;
  const int* var;
  const double* cst;
};
__device__ __forceinline__ void chd_spl_at(const ChdCtx& c, int s, double t, ChdSpl& o) {
  double tl;
  const int np = c.h->sp_npoly[s];
  o.poly = chd_locate(c.poly_tend + (size_t)s * c.Pmax, np, t, &tl);
  chd_basis(tl, c.poly_T[(size_t)s * c.Pmax + o.poly], o.B);
  o.var = c.node_var + ((size_t)s * (c.Pmax + 1) + o.poly) * 6;
  o.cst = c.node_const + ((size_t)s * (c.Pmax + 1) + o.poly) * 6;
}
__device__ __forceinline__ void chd_spl_poly(const ChdCtx& c, int s, int poly, double tl, ChdSpl& o) {
  o.poly = poly;
  chd_basis(tl, c.poly_T[(size_t)s * c.Pmax + poly], o.B);
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
  }
  return 0.5 * w * chd_dot(diff, diff);
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
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) *cost_out = red[0];
  __syncthreads();
}
