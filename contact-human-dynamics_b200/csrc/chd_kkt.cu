// KKT kernels of the batched interior-point solver (product code, sm_100a).
//
//   chd_k_hess_base : once per stage -- Gauss-Newton Hessian of the (quadratic, fixed-duration) cost terms
//                     of data_cost.cpp / vel_smooth_cost.cpp, scaled by the objective scaling, in tile format (Kbase)
//   chd_k_kcopy     : per iteration, side stream -- Kwork <- Kbase for the sequences that continue in their stage
//   chd_k_curv      : per iteration, side stream, after the line search -- y+ Jd^T Jd of the squared-distance rows
//   chd_k_asm       : per iteration, right before chd_k_kkt -- Jacobian dependent matrix entries and the right-hand
//                     side as rhs0 + mu * rhs1 (several CTAs per sequence: the cost is the L2 reductions)
//   chd_k_kkt       : per iteration, one CTA per sequence -- IPOPT error measures + barrier update, (first iteration
//                     of a stage: the whole assembly), tiled band LDL^T with dense border (shared-memory window or,
//                     for wide bands, in place on Kwork; panel, trailing and corner updates on the FP64 tensor
//                     core), triangular solves, step recovery and fraction-to-the-boundary rule
//   chd_k_fp64_peak : DFMA / DMMA throughput probe for the roofline denominators of bench.py
// Replaces IPOPT's per-iteration MA57 factorisation (phys_optim.cpp:573) for the block-banded systems this NLP
// produces.
#include <cuda_runtime.h>

#include "chd_block.cuh"
#include "chd_eval.cuh"
#include "chd_kkt_tiles.cuh"

// Gauss-Newton Hessian of a least-squares cost sample by one warp: H += wgt * J^T J, where the sample residual (3 rows)
// is a signed sum over (up to two) located polynomials of B(deriv) node values.  Every Jacobian column is a 3-vector:
// wgt_q e_dim for a node slot, and -- stage 3, foot splines, positions -- the switch-time columns of chd_spl_tau.
// Entry e = polynomial * 14 + slot (12 node slots, 2 switch-time slots); ws: 28 x 4 doubles of per-warp scratch
// (kkt index or -1, column vector); the 28 x 28 pair loop is spread over the lanes.
__device__ __forceinline__ void chd_hess_sample(const ChdCtx& c, const ChdKT& K, const int* vk, int s, const ChdSpl* P, const double* sign, int np,
                                                int deriv, double wgt, int lane, double* ws) {
  const bool tau = c.opt_dur && s >= 2 && deriv == 0;
  if (lane < 28) {
    const int pa = lane / 14, q = lane % 14;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    int ia = -1;
    if (pa < np) {
      if (q < 12) {
        const int va = P[pa].var[q];
        const double wa = sign[pa] * chd_slot_w(P[pa], deriv, q);
        if (va >= 0 && wa != 0.0) {
          ia = vk[va];
          v0 = q % 3 == 0 ? wa : 0.0, v1 = q % 3 == 1 ? wa : 0.0, v2 = q % 3 == 2 ? wa : 0.0;
        }
      } else if (tau) {
        ChdTau u;
        chd_spl_tau(c, s, s - 2, P[pa], u);
        const int var = q == 12 ? u.va : u.vb;
        const double* dv = q == 12 ? u.da : u.db;
        if (var >= 0) ia = vk[var], v0 = sign[pa] * dv[0], v1 = sign[pa] * dv[1], v2 = sign[pa] * dv[2];
      }
    }
    ws[lane * 4 + 0] = (double)ia, ws[lane * 4 + 1] = v0, ws[lane * 4 + 2] = v1, ws[lane * 4 + 3] = v2;
  }
  __syncwarp();
  const int ne = np * 14;
  for (int idx = lane; idx < ne * ne; idx += 32) {
    const int a = idx / ne, bq = idx - a * ne;
    const int ia = (int)ws[a * 4], ib = (int)ws[bq * 4];
    if (ia < 0 || ib < 0 || ia < ib) continue;
    const double v = ws[a * 4 + 1] * ws[bq * 4 + 1] + ws[a * 4 + 2] * ws[bq * 4 + 2] + ws[a * 4 + 3] * ws[bq * 4 + 3];
    if (v != 0.0) chd_kadd(K, ia, ib, wgt * v);
  }
  __syncwarp();
}

// Cost part of the KKT matrix of sequence b: Gauss-Newton Hessian of the cost terms of data_cost.cpp /
// vel_smooth_cost.cpp / duration_cost.cpp at the current x, scaled by the objective scaling, added into `base` (zeroed
// before).  Constant during a fixed-duration stage (the costs are quadratic in the node values: Kbase, built when the
// stage begins); in stage 3 the basis weights move with the durations and it is rebuilt into Kwork every iteration.
// The samples are dealt to `nparts` CTAs (the reductions into L2 are the cost).
__device__ void chd_hess_build(const ChdDev& D, int b, const ChdStageDev& sg, double* base, int part, int nparts) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const ChdSeq* h = D.seq + b;
  ChdKT K;
  chd_kt_init(D, h, base, K);
  K.ovf = &D.ipm[b].band_ovf;
  const double sf = D.ipm[b].sf;
  const int* vk = D.var_kkt + (size_t)b * D.n_max;
  ChdCtx c;
  chd_make_ctx(D, b, D.x + (size_t)b * D.n_max, c);
  c.dyn = D.ipm[b].dyn;
  c.opt_dur = sg.opt_dur;
  const int n_ee = h->n_ee, nsp = 2 + n_ee, F = h->F, ns = h->n_smooth;
  // one warp per cost sample (the 28 x 28 pair loop of a sample is spread over its lanes)
  __shared__ double s_hws[CHD_THREADS / 32][28 * 4];
  const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  for (int it = part * nwarp + warp; it < nsp * F; it += nparts * nwarp) {
    const int s = it / F, i = it % F, cls = s < 2 ? s : 2;
    ChdSpl P[2];
    double sgn[2] = {1.0, -1.0};
    chd_spl_at(c, s, c.t_data[i], P[1]);
    if (sg.w_data[cls] != 0.0) chd_hess_sample(c, K, vk, s, P + 1, sgn, 1, 0, sf * sg.w_data[cls], lane, s_hws[warp]);
    if (i < ns && (sg.w_vel[cls] != 0.0 || sg.w_acc[cls] != 0.0)) {
      chd_spl_at(c, s, c.t_data[i] + h->dt, P[0]);
      if (sg.w_vel[cls] != 0.0) chd_hess_sample(c, K, vk, s, P, sgn, 2, 0, sf * sg.w_vel[cls], lane, s_hws[warp]);
      if (sg.w_acc[cls] != 0.0) chd_hess_sample(c, K, vk, s, P, sgn, 2, 1, sf * sg.w_acc[cls], lane, s_hws[warp]);
    }
  }
  if (part == 0)
    for (int i = K.Na + tid; i < K.Np; i += nt) K.band[((size_t)(i >> 3) * K.Q) * 64 + (i & 7) * 9] = 1.0;  // identity padding of the band
  if (part == 0 && sg.opt_dur && sg.w_dur != 0.0) {   // duration_cost.cpp: w I in the durations = w D^T D in the switch times
    for (int ee = 0; ee < n_ee; ++ee)
      for (int k = tid; k < h->n_phases[ee] - 1; k += nt) {
        const int i = vk[h->dur_xoff[ee] + k];
        chd_kadd(K, i, i, sf * sg.w_dur);
        if (k > 0) {
          const int j = vk[h->dur_xoff[ee] + k - 1];
          chd_kadd(K, j, j, sf * sg.w_dur);
          chd_kadd(K, i, j, -sf * sg.w_dur);
        }
      }
  }
}
// zero fill of one tile-format matrix, slice `part` of `nparts` (the identity padding of the band is written by
// chd_hess_build, i.e. by a later kernel: the padding entries lie in some other CTA's slice)
__device__ void chd_hess_clear(const ChdDev& D, int b, double* base, int part, int nparts) {
  const size_t cnt2 = D.kstride / 2, per = (cnt2 + nparts - 1) / nparts;
  const size_t lo = (size_t)part * per, hi = lo + per < cnt2 ? lo + per : cnt2;
  double2* dst = reinterpret_cast<double2*>(base);
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = make_double2(0.0, 0.0);
}

// stage begin, grid (G, B): Kbase <- 0, then the cost Hessian, then (one CTA per sequence) the flags + BEGIN -> RUN
__global__ void __launch_bounds__(256) chd_k_hess_zero(ChdDev D) {
  const int b = blockIdx.y;
  if (D.ipm[b].phase != CHD_PH_BEGIN) return;
  chd_hess_clear(D, b, D.Kbase + (size_t)b * D.kstride, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(CHD_THREADS) chd_k_hess_base(ChdDev D) {
  const int b = blockIdx.y;
  if (D.ipm[b].phase != CHD_PH_BEGIN) return;
  chd_hess_build(D, b, D.stages[D.ipm[b].stage], D.Kbase + (size_t)b * D.kstride, blockIdx.x, gridDim.x);
}
// stage 3, every iteration after the line search (side stream, grid (G, B)): the cost Hessian at the accepted iterate
// goes straight into Kwork, which chd_k_kcopy has zeroed for the sequences of that stage
__global__ void __launch_bounds__(CHD_THREADS) chd_k_hess_dur(ChdDev D) {
  const int b = blockIdx.y;
  const ChdIpm& I = D.ipm[b];
  if (I.phase != CHD_PH_RUN || !I.kw_req) return;
  const ChdStageDev sg = D.stages[I.stage];
  if (!sg.opt_dur) return;
  chd_hess_build(D, b, sg, D.Kwork + (size_t)b * D.kstride, blockIdx.x, gridDim.x);
}
// after the cost Hessian is complete: foot-motion node values that no cost sample sees (zero diagonal) are flagged for
// chd_assemble's fixed regularisation.  mode 0 (main stream): sequences at the beginning of a stage (Kbase), which then
// start running; mode 1 (side stream): sequences iterating in stage 3 (Kwork, before the curvature terms are added)
__global__ void __launch_bounds__(128) chd_k_hess_fin(ChdDev D, int mode) {
  const int b = blockIdx.x;
  ChdIpm& I = D.ipm[b];
  const double* base;
  if (mode == 0) {
    if (I.phase != CHD_PH_BEGIN) return;
    base = D.Kbase + (size_t)b * D.kstride;
  } else {
    if (I.phase != CHD_PH_RUN || !I.kw_req || !D.stages[I.stage].opt_dur) return;
    base = D.Kwork + (size_t)b * D.kstride;
  }
  const ChdSeq* h = D.seq + b;
  const int* vk = D.var_kkt + (size_t)b * D.n_max;
  const double* corn = base + (size_t)D.nbc_max * D.Q * 64 + (size_t)D.nbc_max * D.nbt * 64;
  unsigned char* un = D.unobs + (size_t)b * D.n_max;
  const int mot_lo = h->sp_xoff[2], mot_hi = h->sp_xoff[2 + h->n_ee], Na = h->Na;
  for (int i = mot_lo + threadIdx.x; i < mot_hi; i += blockDim.x) {
    const int k = vk[i];
    unsigned char f = 0;
    if (k >= 0) {
      const double hd = k < Na ? base[((size_t)(k >> 3) * D.Q) * 64 + (k & 7) * 9] : corn[(size_t)(k - Na) * 8 * D.nbt + (k - Na)];
      f = hd <= CHD_UNOBS_EPS;
    }
    un[i] = f;
  }
  __syncthreads();
  if (mode == 0 && threadIdx.x == 0) I.phase = CHD_PH_RUN;
}

// y^+ * Jd^T Jd of the squared-distance rows (leg length, toe-heel distance) of sequence b added into K: one warp
// per active row, rows dealt round-robin to `nw` warps of which this is number `wid`.  ws: 192 doubles of per-warp
// scratch, slot -> (kkt index, weight, 3-vector column of Jd).  Called from chd_k_kkt (first iteration of a stage)
// and from chd_k_curv (all later iterations, side stream).
__device__ __forceinline__ void chd_curv_rows(const ChdDev& D, int b, const ChdKT& K, const ChdStageDev& sg, int wid, int nw, int lane,
                                              double* ws) {
  const ChdSeq* h = D.seq + b;
  const size_t ro = (size_t)b * D.m_max, vo = (size_t)b * D.n_max;
  const int* vk = D.var_kkt + vo;
  ChdCtx c;
  chd_make_ctx(D, b, D.x + vo, c);
  c.dyn = D.ipm[b].dyn;
  c.opt_dur = sg.opt_dur;
  for (int si = 0; si < h->nsets; ++si) {
    const ChdSet st = c.sets[si];
    if (!(sg.set_mask & CHD_MASK(st.type))) continue;
    if (st.type != CHD_SET_ROM && st.type != CHD_SET_HEEL) continue;
    const int nslot = st.type == CHD_SET_ROM ? 38 : 28;   // node slots + the switch-time slots (stage 3)
    for (int k = wid; k < st.nitems; k += nw) {
      const int R = st.row0 + k;
      const double yc = D.sc[ro + R] * D.y[ro + R];
      if (!(yc > CHD_CURV_MIN)) continue;
      const double t = c.t_rom[k];
      ChdSpl P0, P1, P2;
      double dRh[3][3];
      if (st.type == CHD_SET_ROM) {
        // d = p_ee - R(e) h - c : blocks lin (-B e_dim), ang (-B dR_dim h), ee (+B e_dim)
        chd_spl_at(c, 0, t, P0);
        chd_spl_at(c, 1, t, P1);
        chd_spl_at(c, chd_sp_motion(st.a), t, P2);
        double e[3];
        chd_spl_val(c, P1, 0, e);
        ChdTrig tr;
        chd_trig(e, tr);
        const double* hip = chd_hip(c, st.a, t);
        for (int j = 0; j < 3; ++j) {
          double Dj[9];
          chd_dR(tr, j, Dj);
          chd_mv(Dj, hip, dRh[j]);
        }
      } else {
        chd_spl_at(c, chd_sp_motion(st.a), t, P0);   // d = p_a - p_b
        chd_spl_at(c, chd_sp_motion(st.b), t, P1);
      }
      ChdTau ua, ub;
      ua.va = ua.vb = ub.va = ub.vb = -1;
      if (c.opt_dur) {
        if (st.type == CHD_SET_ROM) chd_spl_tau(c, chd_sp_motion(st.a), st.a, P2, ua);
        else chd_spl_tau(c, chd_sp_motion(st.a), st.a, P0, ua), chd_spl_tau(c, chd_sp_motion(st.b), st.b, P1, ub);
      }
      const int nnode = st.type == CHD_SET_ROM ? 36 : 24;
      for (int a = lane; a < nslot; a += 32) {
        if (a >= nnode) {   // switch-time columns: d(d)/d(tau) is the 3-vector da / db of the foot spline (second foot: minus)
          const int t4 = a - nnode;
          const ChdTau& u = t4 < 2 ? ua : ub;
          const int var = (t4 & 1) ? u.vb : u.va;
          const double* dv = (t4 & 1) ? u.db : u.da;
          const int ia = var >= 0 ? vk[var] : -1;
          ws[a * 5 + 0] = (double)ia;
          ws[a * 5 + 1] = ia >= 0 ? (t4 < 2 ? 1.0 : -1.0) : 0.0;
          ws[a * 5 + 2] = ia >= 0 ? dv[0] : 0.0, ws[a * 5 + 3] = ia >= 0 ? dv[1] : 0.0, ws[a * 5 + 4] = ia >= 0 ? dv[2] : 0.0;
          continue;
        }
        const int ba = a / 12, qa = a % 12, da = qa % 3;
        const ChdSpl& Pa = ba == 0 ? P0 : (ba == 1 ? P1 : P2);
        const int va = Pa.var[qa];
        const int ia = va >= 0 ? vk[va] : -1;
        double sgn, v3[3] = {da == 0 ? 1.0 : 0.0, da == 1 ? 1.0 : 0.0, da == 2 ? 1.0 : 0.0};
        if (st.type == CHD_SET_ROM) {
          sgn = ba == 2 ? 1.0 : -1.0;
          if (ba == 1) v3[0] = dRh[da][0], v3[1] = dRh[da][1], v3[2] = dRh[da][2];
        } else {
          sgn = ba == 0 ? 1.0 : -1.0;
        }
        ws[a * 5 + 0] = (double)ia;
        ws[a * 5 + 1] = ia >= 0 ? sgn * chd_slot_w(Pa, 0, qa) : 0.0;
        ws[a * 5 + 2] = v3[0], ws[a * 5 + 3] = v3[1], ws[a * 5 + 4] = v3[2];
      }
      __syncwarp();
      for (int idx = lane; idx < nslot * nslot; idx += 32) {
        const int a = idx / nslot, bq = idx - a * nslot;
        const double wa = ws[a * 5 + 1], wb = ws[bq * 5 + 1];
        if (wa == 0.0 || wb == 0.0) continue;
        const int ia = (int)ws[a * 5], ib = (int)ws[bq * 5];
        if (ia < ib) continue;
        const double dotv = ws[a * 5 + 2] * ws[bq * 5 + 2] + ws[a * 5 + 3] * ws[bq * 5 + 3] + ws[a * 5 + 4] * ws[bq * 5 + 4];
        if (dotv != 0.0) chd_kadd(K, ia, ib, yc * wa * wb * dotv);
      }
      __syncwarp();
    }
  }
}

// Assembly of the condensed KKT system of sequence b by threads t0, t0 + tstep, ...: matrix entries (do_mat; global
// reductions into K) and / or right-hand side (rhs_s != nullptr; shared-memory atomics, [0, Np) band unknowns,
// [8*nbc_max, +nb) border unknowns).  Narrow inequality rows (<= 12 slots: terrain, friction pyramid, height) are
// condensed into the primal block (J^T Sigma J); wide ones (leg length) keep their multiplier as an unknown with
// diagonal -1/Sigma, which needs 36 instead of 666 matrix updates per row.
// (A per-column gather of the right-hand side, as in section A, was measured slower than the shared-memory atomics.)
// g0 / g1 (global, chd_k_asm): the right-hand side split as rhs = g0 + mu * g1 -- every term is affine in the barrier
// parameter, which is only decided inside chd_k_kkt -- accumulated with global reductions.
__device__ __forceinline__ void chd_assemble(const ChdDev& D, int b, const ChdKT& K, double delta_w, double mu, double sf, double* rhs_s,
                                             bool do_mat, int t0, int tstep, double* g0 = nullptr, double* g1 = nullptr) {
  const ChdSeq* h = D.seq + b;
  const int n = D.stages[D.ipm[b].stage].opt_dur ? h->n : h->n - h->n_dur;   // the durations are unknowns in stage 3 only
  const int m = h->m, Na = K.Na;
  const size_t ro = (size_t)b * D.m_max, vo = (size_t)b * D.n_max;
  const int* rf = D.rflag + ro;
  const int* vk = D.var_kkt + vo;
  const int* rk = D.row_kkt + ro;
  const int* ep = D.ent_ptr + (size_t)b * (D.m_max + 1);
  const int* ec = D.ent_col + (size_t)b * D.slots_max;
  const double* Jv = D.Jv + (size_t)b * D.slots_max;
  const double* grad = D.grad + vo;
  auto rhs_add = [&](int kk, double v) { if (rhs_s) atomicAdd(rhs_s + (kk < Na ? kk : 8 * D.nbc_max + (kk - Na)), v); };
  auto mat_add = [&](int i, int j, double v) { if (do_mat) chd_kadd(K, i, j, v); };
  auto gidx = [&](int kk) { return kk < Na ? kk : D.Na_max + (kk - Na); };
  auto g_add = [&](int kk, double v0, double v1) {
    if (g0) {
      atomicAdd(g0 + gidx(kk), v0);
      if (v1 != 0.0) atomicAdd(g1 + gidx(kk), v1);
    }
  };
  // foot-motion node values that no cost sample sees (polynomials shorter than a frame): without curvature of their own
  // the Newton step uses them as free slack and they drift by orders of magnitude, which stage 3 (where the sample
  // times sweep over those polynomials) cannot digest.  They get a fixed Levenberg-Marquardt weight on top of the
  // adaptive one -- what the initial scaling of IPOPT's L-BFGS matrix does for the reference.
  const unsigned char* unobs = D.unobs + vo;
  for (int i = t0; i < n; i += tstep) {
    const int k = vk[i];
    if (k < 0) continue;
    if (do_mat && unobs[i]) mat_add(k, k, CHD_DW_UNOBS);
    mat_add(k, k, delta_w);
    rhs_add(k, -sf * grad[i]);
    g_add(k, -sf * grad[i], 0.0);
  }
  for (int r = t0; r < m; r += tstep) {
    const int f = rf[r];
    const int k = rk[r];
    if (!(f & CHD_ROW_ACTIVE)) {
      if (k >= 0) mat_add(k, k, -1.0);   // row of an inactive set: decoupled dummy unknown
      continue;
    }
    const double sc = D.sc[ro + r];
    const int e0 = ep[r], e1 = ep[r + 1];
    if (k >= 0) {
      // explicit row: equality, or wide inequality with its slack eliminated
      double diag = -CHD_DELTA_C, rr, rr0, rr1 = 0.0;
      if (f & CHD_ROW_EQ) {
        rr = -(sc * D.g[ro + r] - D.dL[ro + r]);
        rr0 = rr;
      } else {
        const double s = D.s[ro + r];
        const double gapL = (f & CHD_ROW_HASL) ? s - D.dL[ro + r] : 1.0, gapU = (f & CHD_ROW_HASU) ? D.dU[ro + r] - s : 1.0;
        const double Sig = ((f & CHD_ROW_HASL) ? D.zL[ro + r] / gapL : 0.0) + ((f & CHD_ROW_HASU) ? D.zU[ro + r] / gapU : 0.0);
        const double bvec = ((f & CHD_ROW_HASL) ? mu / gapL : 0.0) - ((f & CHD_ROW_HASU) ? mu / gapU : 0.0);
        diag -= 1.0 / Sig;
        rr = -(sc * D.g[ro + r] - s) + (D.y[ro + r] + bvec) / Sig;
        rr0 = -(sc * D.g[ro + r] - s) + D.y[ro + r] / Sig;
        rr1 = (((f & CHD_ROW_HASL) ? 1.0 / gapL : 0.0) - ((f & CHD_ROW_HASU) ? 1.0 / gapU : 0.0)) / Sig;
      }
      mat_add(k, k, diag);
      rhs_add(k, rr);
      g_add(k, rr0, rr1);
      const double ys = sc * D.y[ro + r];
      for (int e = e0; e < e1; ++e) {
        const int col = ec[e];
        if (col < 0) continue;
        const int kc = vk[col];
        if (kc < 0) continue;
        const double jv = Jv[e];
        if (jv == 0.0) continue;
        mat_add(k, kc, sc * jv);
        rhs_add(kc, -ys * jv);
        g_add(kc, -ys * jv, 0.0);
      }
    } else {
      // condensed narrow inequality row
      const double s = D.s[ro + r];
      const double gapL = (f & CHD_ROW_HASL) ? s - D.dL[ro + r] : 1.0, gapU = (f & CHD_ROW_HASU) ? D.dU[ro + r] - s : 1.0;
      const double Sig = ((f & CHD_ROW_HASL) ? D.zL[ro + r] / gapL : 0.0) + ((f & CHD_ROW_HASU) ? D.zU[ro + r] / gapU : 0.0);
      const double bvec = ((f & CHD_ROW_HASL) ? mu / gapL : 0.0) - ((f & CHD_ROW_HASU) ? mu / gapU : 0.0);
      const double coef = Sig * (sc * D.g[ro + r] - s) - bvec;
      const double coef0 = Sig * (sc * D.g[ro + r] - s);
      const double beta = ((f & CHD_ROW_HASL) ? 1.0 / gapL : 0.0) - ((f & CHD_ROW_HASU) ? 1.0 / gapU : 0.0);
      for (int ea = e0; ea < e1; ++ea) {
        const int ca = ec[ea];
        if (ca < 0) continue;
        const int ka = vk[ca];
        if (ka < 0) continue;
        const double va = sc * Jv[ea];
        if (va == 0.0) continue;
        rhs_add(ka, -va * coef);
        g_add(ka, -va * coef0, va * beta);
        if (do_mat)
        for (int eb = e0; eb < e1; ++eb) {
          const int cb = ec[eb];
          if (cb < 0) continue;
          const int kb = vk[cb];
          if (kb < 0 || ka < kb) continue;
          const double vb = sc * Jv[eb];
          if (vb != 0.0) mat_add(ka, kb, Sig * va * vb);
        }
      }
    }
  }
}

// dynamic shared memory layout of chd_k_kkt (doubles):
//   red[CHD_KKT_THREADS] | vecn[n_max] | xs[Np_max + nbp8] | cc[nbp8*nbp8] | ypan[(Q+nbt)*64] | xpan[(Q+nbt)*64] | xs2[Np_max] | dinv[16]
//   | win[Q(Q+1)/2 * 64] | bwin[Q*nbt*64]            (the last two in global scratch when they do not fit)
template <bool WS>
__device__ __forceinline__ void chd_kkt_body(const ChdDev& D) {
  extern __shared__ double sm[];
  __shared__ int s_fail;
  const int b = blockIdx.x;
  ChdIpm& I = D.ipm[b];
  if (I.phase != CHD_PH_RUN) return;
  const int pre_refreshed = I.kw_req;   // Kwork = Kbase + distance-row curvature already prepared on the side stream
  const ChdStageDev sg = D.stages[I.stage];
  const ChdSeq* h = D.seq + b;
  const int n = h->n, m = h->m, tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  const size_t ro = (size_t)b * D.m_max, vo = (size_t)b * D.n_max;
  const int* rf = D.rflag + ro;
  const int* vk = D.var_kkt + vo;
  const int* rk = D.row_kkt + ro;
  const int* ep = D.ent_ptr + (size_t)b * (D.m_max + 1);
  const int* ec = D.ent_col + (size_t)b * D.slots_max;
  const double* Jv = D.Jv + (size_t)b * D.slots_max;
  const double* grad = D.grad + vo;
  ChdKT K;
  double* kw = D.Kwork + (size_t)b * D.kstride;
  chd_kt_init(D, h, kw, K);
  K.ovf = &I.band_ovf;
  if (!sg.opt_dur) K.q = D.Qfix - 1;   // fixed-duration stages: the static pattern needs fewer band tiles than stage 3 may
  // (below, Q is the number of block rows of the elimination window of this stage, Qs the storage stride of a block column)
  const int n_act = sg.opt_dur ? n : n - h->n_dur;          // the durations (last n_dur entries of x) are unknowns in stage 3 only
  const int Qs = K.Q, Q = sg.opt_dur ? K.Q : D.Qfix, nbt = K.nbt, nbp8 = K.nbp8, nbl = sg.opt_dur ? h->nb : h->nb_fix, nbc = K.nbc;
  // the right-hand side rides along as border row NBR, directly behind the border unknowns of this stage; nbt_s = border
  // tiles in use (the switch-time columns of stage 3 are all zero in the fixed-duration stages: not even looked at)
  const int NBR = nbl, nbt_s = (nbl + 1 + 7) >> 3;
  // WS: everything in shared memory.  !WS (long horizons / very wide bands): only the reduction buffer, the
  // corner and the panel buffers stay in shared memory; the per-unknown vectors and the window live in a global
  // (L2 resident) scratch area  vecn | xs | xs2 | win | bwin.
  const size_t n_even = (size_t)((D.n_max + 1) & ~1), xs_len = (size_t)(8 * D.nbc_max + nbp8);
  double* gs = WS ? nullptr : D.scratch + (size_t)b * D.scratch_stride;
  double* red = sm;
  double* cc = red + CHD_KKT_THREADS;
  double* ypan = cc + nbp8 * nbp8;
  double* xpan = ypan + (Q + nbt) * 64;
  double* dinv = ypan + D.pan_doubles;
  // WS: the per-unknown vectors vecn (step recovery) and xs (right-hand side before, solution after the factorisation)
  // alias the tail of the window region, which is dead whenever they are live (the right-hand side moves into the border
  // storage before the window is loaded; the back-substitution stages its tiles in the front part only)
  const size_t win_region = (size_t)D.win_tiles * 64 + (size_t)Qs * nbt * 64;
  double* win = WS ? dinv + 16 : gs + n_even + xs_len + 8 * (size_t)D.nbc_max;
  double* bwin = win + (size_t)D.win_tiles * 64;
  double* vecn = WS ? win + win_region - n_even : gs;
  double* xs = WS ? vecn - xs_len : vecn + n_even;
  double* xs2 = WS ? ypan : xs + xs_len;      // back-substitution accumulator (WS: aliases the then idle panel buffers)
  const double sf = I.sf;
  double mu = I.mu;
  // per-phase cycle counters (scripts/prof_stage.py): compiled in only with -DCHD_PROFILE (extra barriers + clock reads)
#ifdef CHD_PROFILE
  long long tk0 = clock64();
#define CHD_PROF(slot) do { __syncthreads(); if (tid == 0) { long long t_ = clock64(); I.prof[slot] += (double)(t_ - tk0); tk0 = t_; } } while (0)
#else
#define CHD_PROF(slot) do { } while (0)
#endif

  // ---------------- A. error measures, convergence, barrier update ----------------
  // J^T y is gathered per variable from a column-oriented index of the Jacobian slots (no shared-memory fp64
  // atomics, which are compare-and-swap loops): per-row multipliers first, into the (idle) dy array
  const int* erow = D.ent_row + (size_t)b * D.slots_max;
  const int* cptr = D.col_ptr + (size_t)b * (D.n_max + 1);
  const int* cent = D.col_ent + (size_t)b * D.slots_max;
  double* rowv = D.dy + ro;
  double a_ysum = 0, a_zsum = 0, a_cviol = 0, a_theta = 0, a_rs = 0, a_cmax = -INFINITY, a_cmin = INFINITY, a_violu = 0;
  for (int r = tid; r < m; r += nt) {
    const int f = rf[r];
    if (!(f & CHD_ROW_ACTIVE)) {
      rowv[r] = 0.0;
      continue;
    }
    const double sc = D.sc[ro + r], gval = D.g[ro + r], d = sc * gval, y = D.y[ro + r];
    a_ysum += fabs(y);
    a_violu = fmax(a_violu, fmax(D.row_lo[ro + r] - gval, gval - D.row_hi[ro + r]));
    rowv[r] = sc * y;
    if (f & CHD_ROW_EQ) {
      const double re = d - D.dL[ro + r];
      a_cviol = fmax(a_cviol, fabs(re));
      a_theta += fabs(re);
    } else {
      const double s = D.s[ro + r], ri = d - s, zL = D.zL[ro + r], zU = D.zU[ro + r];
      a_cviol = fmax(a_cviol, fabs(ri));
      a_theta += fabs(ri);
      a_rs = fmax(a_rs, fabs(-y - zL + zU));
      a_zsum += zL + zU;
      if (f & CHD_ROW_HASL) { const double cp = (s - D.dL[ro + r]) * zL; a_cmax = fmax(a_cmax, cp); a_cmin = fmin(a_cmin, cp); }
      if (f & CHD_ROW_HASU) { const double cp = (D.dU[ro + r] - s) * zU; a_cmax = fmax(a_cmax, cp); a_cmin = fmin(a_cmin, cp); }
    }
  }
  __syncthreads();
  double a_dual = 0;
  if (!I.dyn) {
    for (int i = tid; i < n_act; i += nt) {
      double v = sf * grad[i];
      for (int t = cptr[i]; t < cptr[i + 1]; ++t) {
        const int e = cent[t];
        v += rowv[erow[e]] * Jv[e];
      }
      if (vk[i] >= 0) a_dual = fmax(a_dual, fabs(v));
    }
  } else {
    // run-time Jacobian columns (stage 3 onwards): the column index of the layout is stale, J^T y goes through
    // reductions into a per-sequence global vector
    double* jty = D.jty + vo;
    for (int i = tid; i < n; i += nt) jty[i] = 0.0;
    __syncthreads();
    for (int r = tid; r < m; r += nt) {
      const double ys = rowv[r];
      if (ys == 0.0) continue;
      for (int e = ep[r]; e < ep[r + 1]; ++e) {
        const int col = ec[e];
        if (col >= 0 && Jv[e] != 0.0) atomicAdd(jty + col, ys * Jv[e]);
      }
    }
    __syncthreads();
    for (int i = tid; i < n_act; i += nt)
      if (vk[i] >= 0) a_dual = fmax(a_dual, fabs(sf * grad[i] + jty[i]));
  }
  const double ysum = chd_block_sum(a_ysum, red), zsum = chd_block_sum(a_zsum, red);
  const double cviol = chd_block_max(a_cviol, red), theta = chd_block_sum(a_theta, red);
  const double dual_inf = fmax(chd_block_max(a_dual, red), chd_block_max(a_rs, red));
  const double cmax = chd_block_max(a_cmax, red), cmin = chd_block_min(a_cmin, red);
  const double violu = fmax(chd_block_max(a_violu, red), 0.0);
  const int nbnd = I.n_bounds;
  const double s_d = fmax(CHD_S_MAX, (ysum + zsum) / fmax((double)(I.m_act + nbnd), 1.0)) / CHD_S_MAX;
  const double s_c = fmax(CHD_S_MAX, zsum / fmax((double)nbnd, 1.0)) / CHD_S_MAX;
  auto compl_err = [&](double mm) { return nbnd > 0 ? fmax(fabs(cmax - mm), fabs(cmin - mm)) : 0.0; };
  const double E0 = fmax(fmax(dual_inf / s_d, cviol), compl_err(0.0) / s_c);
  const double dual_u = dual_inf / sf, compl_u = compl_err(0.0) / sf;
  bool done = false;
  int new_status = 1;
  if (E0 <= CHD_TOL && violu <= CHD_CONSTR_VIOL_TOL && dual_u <= CHD_DUAL_INF_TOL && compl_u <= CHD_COMPL_INF_TOL) new_status = 0, done = true;
  else if (I.iter >= I.max_iter) new_status = -1, done = true;
  if (!done) {
    const double mu_min = fmin(CHD_TOL, CHD_COMPL_INF_TOL) / (CHD_KAPPA_EPS + 1.0);
    while (true) {
      const double Emu = fmax(fmax(dual_inf / s_d, cviol), compl_err(mu) / s_c);
      if (Emu <= CHD_KAPPA_EPS * mu && mu > mu_min) mu = fmax(mu_min, fmin(CHD_KAPPA_MU * mu, pow(mu, CHD_THETA_MU)));
      else break;
    }
  }
  __syncthreads();
  if (tid == 0) {
    I.f = D.cost[2 * b];
    I.E0 = E0, I.viol_u = violu, I.dual_u = dual_u, I.compl_u = compl_u;
    I.status = new_status;
    s_fail = 0;
    if (done) chd_stage_advance(D, I, new_status, sg.snap_after);
    if (!done) {
      I.mu = mu;
      I.tau = fmax(CHD_TAU_MIN, 1.0 - mu);
      if (I.iter == 0) I.theta_max = 1e4 * fmax(1.0, theta), I.theta_min = 1e-4 * fmax(1.0, theta), I.theta_ref = theta;
      if (mu != I.mu_filter) I.nfilt = 0, I.mu_filter = mu;
      I.theta0 = theta;
    }
  }
  if (done) return;
  const double tau = fmax(CHD_TAU_MIN, 1.0 - mu);
  // feasibility polish: every test but the unscaled constraint violation passes -> this step only restores feasibility
  // (a large Levenberg-Marquardt weight makes it the least-norm Newton correction of the constraints; the adaptive
  // weight itself is left alone)
  const bool polish = E0 <= CHD_TOL && dual_u <= CHD_DUAL_INF_TOL && compl_u <= CHD_COMPL_INF_TOL && violu > CHD_CONSTR_VIOL_TOL &&
                      I.delta_w < CHD_DW_POLISH;
  const double delta_w = polish ? CHD_DW_POLISH : I.delta_w;
  CHD_PROF(0);

  // ---------------- B. assemble the KKT system ----------------
  // Narrow inequality rows (<= 12 slots: terrain, friction pyramid, height) are condensed into the primal block
  // (J^T Sigma J); wide ones (leg length) keep their multiplier as an unknown with diagonal
  // -1/Sigma, which needs 36 instead of 666 matrix updates per row.  The right-hand side is accumulated in
  // shared memory (xs) and written out once.
  if (!pre_refreshed) {   // first iteration of a stage; afterwards chd_k_kcopy refreshes Kwork on the side stream
    const double* base = D.Kbase + (size_t)b * D.kstride;
    const double2* src = reinterpret_cast<const double2*>(base);
    double2* dst = reinterpret_cast<double2*>(kw);
    const size_t cnt2 = D.kstride / 2;
    size_t i = tid;
    for (; i + 3 * (size_t)nt < cnt2; i += 4 * (size_t)nt) {   // four independent 16-byte loads in flight per thread
      const double2 v0 = src[i], v1 = src[i + nt], v2 = src[i + 2 * (size_t)nt], v3 = src[i + 3 * (size_t)nt];
      dst[i] = v0, dst[i + nt] = v1, dst[i + 2 * (size_t)nt] = v2, dst[i + 3 * (size_t)nt] = v3;
    }
    for (; i < cnt2; i += nt) dst[i] = src[i];
  }
  const int Na = K.Na;
  double* rhs_s = xs;   // [0, Np) band unknowns, [8*nbc_max, +nb) border unknowns
  for (int i = tid; i < 8 * D.nbc_max + nbp8; i += nt) rhs_s[i] = 0.0;
  __syncthreads();
  // matrix entries do not depend on the barrier parameter: for sequences that continue in their stage they were
  // added by chd_k_asm (8 CTAs per sequence) before this kernel; the right-hand side (shared-memory atomics) is
  // always assembled here
  if (pre_refreshed && polish) {   // chd_k_asm put the adaptive weight on the diagonal: top it up
    const double extra = delta_w - I.delta_w;
    for (int i = tid; i < n_act; i += nt)
      if (vk[i] >= 0) chd_kadd(K, vk[i], vk[i], extra);
  }
  if (pre_refreshed) {
    const double* r0 = D.rhs0 + (size_t)b * (D.Na_max + D.nb_max);
    const double* r1 = D.rhs1 + (size_t)b * (D.Na_max + D.nb_max);
    for (int i = tid; i < Na; i += nt) rhs_s[i] = r0[i] + mu * r1[i];
    for (int i = tid; i < nbl; i += nt) rhs_s[8 * D.nbc_max + i] = r0[D.Na_max + i] + mu * r1[D.Na_max + i];
  } else {
    chd_assemble(D, b, K, delta_w, mu, sf, rhs_s, true, tid, nt);
  }
  __syncthreads();
  for (int i = tid; i < K.Np; i += nt) K.bord[((size_t)(i >> 3) * nbt + (NBR >> 3)) * 64 + (NBR & 7) * 8 + (i & 7)] = i < Na ? rhs_s[i] : 0.0;
  for (int i = tid; i < nbl; i += nt) K.corn[(size_t)NBR * nbp8 + i] = rhs_s[8 * D.nbc_max + i];
  CHD_PROF(1);
  // y^+ * Jd^T Jd of the squared-distance rows: in-kernel only on the first iteration of a stage, afterwards
  // chd_k_curv has already added them on the side stream (many CTAs per sequence: the atomics are the cost)
  if (!pre_refreshed) chd_curv_rows(D, b, K, sg, warp, nwarp, lane, win + warp * 192);
  __threadfence_block();
  __syncthreads();
  CHD_PROF(2);

  // ---------------- C. tiled band LDL^T with dense border ----------------
  // corner (incl. rhs row) to shared memory; initial window: band tiles (I, J), 0 <= J <= I <= q and border columns 0..q
  for (int i = tid; i < nbp8 * nbp8; i += nt) cc[i] = K.corn[i];
  // (without the shared-memory window the factorisation runs in place on Kwork: no window copy, no stream-in)
  if (WS) {
    for (int idx = tid; idx < Q * Q * 64; idx += nt) {
      const int e = idx & 63, pr = idx >> 6, J = pr / Q, t = pr % Q;
      if (J + t < Q && J < nbc && J + t < nbc) win[chd_win_slot(J + t, J, Q) * 64 + e] = K.band[((size_t)J * Qs + t) * 64 + e];
    }
    for (int idx = tid; idx < Q * nbt * 64; idx += nt) {
      const int J = idx / (nbt * 64);
      if (J < nbc) bwin[idx] = K.bord[(size_t)J * nbt * 64 + idx % (nbt * 64)];
    }
  }
  __syncthreads();
  // index tables: pair list (gi >= gj) over band groups 0..q-1 and border groups q..q+nbt-1 (built once),
  // window slot of every band group of the current block column (double buffered, no integer division)
  __shared__ int s_rs[2][96];
  __shared__ int s_gnz[2][96];   // per panel group: any non-zero entry in the X tile (zero tiles skip their trailing updates)
  __shared__ unsigned short s_pairs[3000];
  __shared__ unsigned char s_cmp[CHD_KKT_THREADS / 32][64];   // per warp: rank -> id of the non-zero panel groups
  __shared__ __align__(16) double s_winv[2][64];   // inverse of the current / next diagonal tile factor, fragment order
  __shared__ __align__(8) unsigned long long s_mbar;   // completion of the TMA bulk stream-in of one block row
  unsigned mbar_phase = 0;
  const bool use_tma = WS && D.tma;
  if (use_tma && tid == 0) chd_mbar_init(&s_mbar, 1);
  const int GB = K.q, Gm = K.q + nbt_s, npairs = Gm * (Gm + 1) / 2;
  for (int p = tid; p < npairs && p < 3000; p += nt) {
    int gi = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
    while (gi * (gi + 1) / 2 > p) --gi;
    while ((gi + 1) * (gi + 2) / 2 <= p) ++gi;
    s_pairs[p] = (unsigned short)((gi << 8) | (p - gi * (gi + 1) / 2));
  }
  auto tri = [](int a_, int b_) { return a_ > b_ ? a_ * (a_ + 1) / 2 + b_ : b_ * (b_ + 1) / 2 + a_; };
  for (int g = tid; g < GB; g += nt) s_rs[0][g] = (1 + g) % Q;
  for (int g = tid; g < 96; g += nt) s_gnz[0][g] = 0, s_gnz[1][g] = 0;
  if (warp == 0) {
    const bool ok = chd_tile_ldl(WS ? win : K.band, dinv, s_winv[0], lane);   // tile (0,0) sits in slot 0
    if (!ok && lane == 0) s_fail = 1;
  }
  __syncthreads();
  int kslot = 0, cur = 0;  // kslot = Kc % Q
#ifdef CHD_PROFILE
  long long pf_a = 0, pf_b = 0, pf_c = 0, pf_acc[3] = {0, 0, 0};
#endif
  for (int Kc = 0; Kc < nbc; ++Kc, cur ^= 1) {
#ifdef CHD_PROFILE
    pf_a = clock64();
#endif
    const int tq = min(K.q, nbc - 1 - Kc);          // band tiles below the diagonal tile
    const int* rs = s_rs[cur];
    // tile addresses: circular triangular window in shared memory, or in place in the global band / border storage
    auto band_tile = [&](int gi, int gj) -> double* {     // tile (Kc+1+gi, Kc+1+gj), gi >= gj
      return WS ? win + (size_t)tri(rs[gi], rs[gj]) * 64 : K.band + ((size_t)(Kc + 1 + gj) * Qs + (gi - gj)) * 64;
    };
    auto bord_tile = [&](int bi, int gj) -> double* {     // border tile bi of block column Kc+1+gj
      return WS ? bwin + ((size_t)rs[gj] * nbt + bi) * 64 : K.bord + ((size_t)(Kc + 1 + gj) * nbt + bi) * 64;
    };
    double* Tkk = WS ? win + (size_t)tri(kslot, kslot) * 64 : K.band + (size_t)Kc * Qs * 64;
    double* Bk = WS ? bwin + (size_t)kslot * nbt * 64 : K.bord + (size_t)Kc * nbt * 64;
    const double* dv = dinv + 8 * cur;
    // (b) panel: Y = A L0^-T = A W^T (W = L0^-1 from the diagonal-tile factorisation) as one tensor-core product per
    //     8x8 panel tile, X = Y D^-1; both go to the panel buffers in fragment order, X also to global (final L);
    //     the diagonal tile goes to global as well
    {
      const double* wv = s_winv[cur];
      const int r = lane >> 2, k = lane & 3;
      const double2 wf = *reinterpret_cast<const double2*>(wv + 2 * lane);
      const double d0 = dv[2 * k], d1 = dv[2 * k + 1];
      const int f0 = r * 8 + chd_frag_col(2 * k), f1 = r * 8 + chd_frag_col(2 * k + 1);
      for (int g = warp; g < tq + nbt_s; g += nwarp) {
        const bool band_t = g < tq;
        const int pg = band_t ? g : GB + (g - tq);                   // group id inside the panel buffers
        const double* A = band_t ? (WS ? win + (size_t)tri(rs[g], kslot) * 64 : K.band + ((size_t)Kc * Qs + 1 + g) * 64) : Bk + (g - tq) * 64;
        const double ax = A[r * 8 + k], ay = A[r * 8 + k + 4];
        double c0 = 0.0, c1 = 0.0;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1)
                     : "d"(ax), "d"(wf.x));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1)
                     : "d"(ay), "d"(wf.y));
        const double x0 = c0 * d0, x1 = c1 * d1;
        ypan[pg * 64 + f0] = c0, ypan[pg * 64 + f1] = c1;
        xpan[pg * 64 + f0] = x0, xpan[pg * 64 + f1] = x1;
        double* G = band_t ? K.band + ((size_t)Kc * Qs + 1 + g) * 64 : K.bord + ((size_t)Kc * nbt + (g - tq)) * 64;
        *reinterpret_cast<double2*>(G + r * 8 + 2 * k) = make_double2(x0, x1);
        const bool nz = __any_sync(0xffffffffu, x0 != 0.0 || x1 != 0.0);
        if (lane == 0 && nz) s_gnz[cur][pg] = 1;
      }
    }
    if (WS)
      for (int e = tid; e < 64; e += nt) K.band[(size_t)Kc * Qs * 64 + e] = Tkk[e];
    __syncthreads();
#ifdef CHD_PROFILE
    pf_b = clock64();
#endif
    // (c) stream in block row Kc + Q (its slots are dead now), trailing updates on the fp64 tensor core;
    //     warp 0 takes the pair that completes the next diagonal tile and factors it right away
    const int In = Kc + Q;
    // Two stream-in variants (D.tma, environment CHD_TMA at batch creation; measured on the benchmark batch: 1.00 ms per
    // launch with the per-thread cp.async chunks, 1.04 ms with the TMA producer warp, 1.21 ms when a lane of an updating
    // warp issues the bulk copies, 1.11 ms with dynamically dealt update blocks):
    if (WS && !use_tma && In < nbc) {
      // 16-byte cp.async chunks by the warps 1..15; warp 0 goes straight to the diagonal tile
      for (int idx = tid - 32; idx < Q * 32 + nbt_s * 32; idx += nt - 32) {
        if (idx < 0) break;
        const int tile = idx >> 5, off = (idx & 31) * 2;
        if (tile < Q) {
          const int J = tile < GB ? Kc + 1 + tile : In;
          double* dst = tile < GB ? win + (size_t)tri(kslot, rs[tile]) * 64 : Tkk;
          chd_copy16(dst + off, K.band + ((size_t)J * Qs + (In - J)) * 64 + off, WS);
        } else {
          chd_copy16(Bk + (tile - Q) * 64 + off, K.bord + (size_t)In * nbt * 64 + (tile - Q) * 64 + off, WS);
        }
      }
    }
    // TMA variant: the last warp is the producer of the shared-memory window -- it only streams (the trailing updates
    // are dealt to the warps 1 .. nwarp-2), so no warp of the update loop is held up by the serial issue of the copies
    const bool producer = use_tma && warp == nwarp - 1;
    if (producer && In < nbc && lane == 0) {
      // TMA bulk copies (cp.async.bulk + mbarrier): one thread streams the Q band tiles of block row In (512 contiguous
      // bytes each in the block-column storage) and its border tiles (one contiguous piece) into the slots the pivot
      // block row just vacated; the slots were last touched through the generic proxy (panel phase, before the barrier)
      chd_fence_async_smem();
      chd_mbar_expect(&s_mbar, (unsigned)((Q + nbt_s) * 512));
      for (int tile = 0; tile < Q; ++tile) {
        const int J = tile < GB ? Kc + 1 + tile : In;
        double* dst = tile < GB ? win + (size_t)tri(kslot, rs[tile]) * 64 : Tkk;
        chd_bulk_g2s(dst, K.band + ((size_t)J * Qs + (In - J)) * 64, 512u, &s_mbar);
      }
      chd_bulk_g2s(Bk, K.bord + (size_t)In * nbt * 64, (unsigned)(nbt_s * 512), &s_mbar);
    }
    if (warp == 0) {
      if (tq >= 1) {
        double* Tn = band_tile(0, 0);
        chd_tile_sub_xyT(Tn, xpan, ypan, lane);
        __syncwarp();
        const bool ok = chd_tile_ldl(Tn, dinv + 8 * (cur ^ 1), s_winv[cur ^ 1], lane);
        if (!ok && lane == 0) s_fail = 1;
      }
    } else if (!producer) {
      if (warp == 1) {
        for (int g = lane; g < GB; g += 32) {   // slot table of the next block column
          int v = kslot + 2 + g;
          while (v >= Q) v -= Q;
          s_rs[cur ^ 1][g] = v;
        }
        for (int g = lane; g < Gm; g += 32) s_gnz[cur ^ 1][g] = 0;
      }
      // compact list of the groups with a non-zero X tile (every warp builds it redundantly: no extra barrier).
      // the pair table enumerates (i >= j) row by row, so its first na(na+1)/2 entries pair the first na entries.
      const int* gnz = s_gnz[cur];
      int na;
      const unsigned char* cmp = s_cmp[warp];
      {
        // up to 64 groups, two per lane; scatter by rank through a per-warp table (__fns is slow)
        const int g1 = lane + 32;
        const bool act0 = lane < Gm && (lane >= GB || lane < tq) && gnz[lane];
        const bool act1 = g1 < Gm && (g1 >= GB || g1 < tq) && gnz[g1 < 96 ? g1 : 0];
        const unsigned m0 = __ballot_sync(0xffffffffu, act0), m1 = __ballot_sync(0xffffffffu, act1);
        const unsigned below = (1u << lane) - 1u;
        const int n0 = __popc(m0);
        na = n0 + __popc(m1);
        if (act0) s_cmp[warp][__popc(m0 & below)] = (unsigned char)lane;
        if (act1) s_cmp[warp][n0 + __popc(m1 & below)] = (unsigned char)g1;
        __syncwarp();
      }
      const bool compact = Gm <= 64;
      const int np_loop = compact ? na * (na + 1) / 2 : npairs;
      if (!WS && !compact) {
        // more than 64 panel groups, window in global (L2) memory: each warp collects up to four target tiles, issues all their loads, and only
        // then runs the tensor-core updates and the stores.  (With the window in shared memory the simple loop below
        // is faster: measured 707 vs 820 ms of KKT time per benchmark step.)
        const int r8 = (lane >> 2) * 8 + 2 * (lane & 3);
        int p = warp - 1;
        while (p < np_loop) {
          double* Cp[4];
          const double *Xp[4], *Yp[4];
          int nq = 0;
          while (nq < 4 && p < np_loop) {
            const int ai = s_pairs[p] >> 8, aj = s_pairs[p] & 255;
            p += nwarp - 1;
            int gi, gj;
            if (compact) {
              gi = cmp[ai], gj = cmp[aj];
            } else {
              gi = ai, gj = aj;
              if ((gi < GB && gi >= tq) || (gj < GB && gj >= tq) || !gnz[gi] || !gnz[gj]) continue;
            }
            if (gi == 0 && gj == 0) continue;
            const double* X = xpan + gi * 64;
            const double* Y = ypan + gj * 64;
            if (gi < GB) {
              Cp[nq] = band_tile(gi, gj), Xp[nq] = X, Yp[nq] = Y, ++nq;
            } else if (gj < GB) {
              Cp[nq] = bord_tile(gi - GB, gj), Xp[nq] = X, Yp[nq] = Y, ++nq;
            } else {
              const int bi = gi - GB, bj = gj - GB;
              for (int e = lane; e < 64; e += 32) {
                const int r = e >> 3, cq = e & 7;
                double acc = 0.0;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) acc += X[r * 8 + kk] * Y[cq * 8 + kk];
                cc[(bi * 8 + r) * nbp8 + bj * 8 + cq] -= acc;
              }
            }
          }
          double c0[4], c1[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < nq) c0[i] = Cp[i][r8], c1[i] = Cp[i][r8 + 1];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < nq) {
              chd_tile_mma(c0[i], c1[i], Xp[i], Yp[i], lane);
              Cp[i][r8] = c0[i], Cp[i][r8 + 1] = c1[i];
            }
        }
      } else {
        // window in shared memory: two target tiles per loop trip with interleaved tensor-core updates (the dependent
        // chain index -> tile address -> load -> 2 x DMMA -> store of a single tile leaves the pipe idle most of the time)
        auto decode = [&](int p, double*& C, const double*& X, const double*& Y, int& bi, int& bj) -> int {
          const int ai = s_pairs[p] >> 8, aj = s_pairs[p] & 255;
          int gi, gj;
          if (compact) {
            gi = cmp[ai], gj = cmp[aj];
          } else {
            gi = ai, gj = aj;
            if ((gi < GB && gi >= tq) || (gj < GB && gj >= tq) || !gnz[gi] || !gnz[gj]) return 0;
          }
          if (gi == 0 && gj == 0) return 0;   // the next diagonal tile is updated by warp 0
          X = xpan + gi * 64;
          Y = ypan + gj * 64;
          if (gi < GB) {
            C = band_tile(gi, gj);
            return 1;
          }
          if (gj < GB) {
            C = bord_tile(gi - GB, gj);
            return 1;
          }
          bi = gi - GB, bj = gj - GB;
          return 2;
        };
        auto corner = [&](const double* X, const double* Y, int bi, int bj) {   // corner block: same tensor-core update, row stride nbp8
          double* Cc = cc + (size_t)(bi * 8 + (lane >> 2)) * nbp8 + bj * 8 + 2 * (lane & 3);
          double c0 = Cc[0], c1 = Cc[1];
          chd_tile_mma(c0, c1, X, Y, lane);
          Cc[0] = c0, Cc[1] = c1;
        };
        const int step = use_tma ? nwarp - 2 : nwarp - 1, r8 = (lane >> 2) * 8 + 2 * (lane & 3);
        if (compact) {
          // 2 x 2 register blocking over the compacted group list: one trip loads the operand fragments of two panel
          // rows (X) and two panel columns (Y) once and updates up to four target tiles with them -- the loop is
          // shared-memory bandwidth bound, this cuts the bytes per tile update from 2 KB to 1.5 KB and the index work 4x
          const int nb2 = (na + 1) >> 1, nblk = nb2 * (nb2 + 1) / 2;
          for (int p = warp - 1; p < nblk; p += step) {
            const int bi = s_pairs[p] >> 8, bj = s_pairs[p] & 255;
            const int i1 = 2 * bi + 1, j1 = 2 * bj + 1;
            const bool vi1 = i1 < na, vj1 = j1 < na;
            int gi[2], gj[2];
            gi[0] = cmp[2 * bi], gi[1] = cmp[i1 & 63];
            gj[0] = cmp[2 * bj], gj[1] = cmp[j1 & 63];
            double2 xf[2], yf[2];
            xf[0] = *reinterpret_cast<const double2*>(xpan + gi[0] * 64 + 2 * lane);
            yf[0] = *reinterpret_cast<const double2*>(ypan + gj[0] * 64 + 2 * lane);
            xf[1] = vi1 ? *reinterpret_cast<const double2*>(xpan + gi[1] * 64 + 2 * lane) : make_double2(0.0, 0.0);
            yf[1] = vj1 ? *reinterpret_cast<const double2*>(ypan + gj[1] * 64 + 2 * lane) : make_double2(0.0, 0.0);
            double* Cp[4];
            double2 cv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int a = t & 1, c = t >> 1;          // tile (row a, column c) of the block
              bool ok = (a == 0 || vi1) && (c == 0 || vj1) && !(bi == bj && a == 0 && c == 1);
              const int g_i = gi[a], g_j = gj[c];
              ok = ok && !(g_i == 0 && g_j == 0);       // the next diagonal tile is updated by warp 0
              Cp[t] = nullptr;
              if (ok) {
                if (g_i < GB) Cp[t] = band_tile(g_i, g_j);
                else if (g_j < GB) Cp[t] = bord_tile(g_i - GB, g_j);
                else corner(xpan + g_i * 64, ypan + g_j * 64, g_i - GB, g_j - GB);
              }
              if (Cp[t]) cv[t] = *reinterpret_cast<const double2*>(Cp[t] + r8);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (Cp[t]) {
                const double2 xa = xf[t & 1], yb = yf[t >> 1];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(cv[t].x), "+d"(cv[t].y)
                             : "d"(-xa.x), "d"(yb.x));
              }
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (Cp[t]) {
                const double2 xa = xf[t & 1], yb = yf[t >> 1];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(cv[t].x), "+d"(cv[t].y)
                             : "d"(-xa.y), "d"(yb.y));
                *reinterpret_cast<double2*>(Cp[t] + r8) = cv[t];
              }
          }
        } else
        for (int p = warp - 1; p < np_loop; p += 2 * step) {
          double *C1 = nullptr, *C2 = nullptr;
          const double *X1 = nullptr, *Y1 = nullptr, *X2 = nullptr, *Y2 = nullptr;
          int b1i = 0, b1j = 0, b2i = 0, b2j = 0;
          const int k1 = decode(p, C1, X1, Y1, b1i, b1j);
          const int k2 = p + step < np_loop ? decode(p + step, C2, X2, Y2, b2i, b2j) : 0;
          if (k1 == 1 && k2 == 1) {
            double2 c1 = *reinterpret_cast<const double2*>(C1 + r8), c2 = *reinterpret_cast<const double2*>(C2 + r8);
            const double2 xa1 = *reinterpret_cast<const double2*>(X1 + 2 * lane), yb1 = *reinterpret_cast<const double2*>(Y1 + 2 * lane);
            const double2 xa2 = *reinterpret_cast<const double2*>(X2 + 2 * lane), yb2 = *reinterpret_cast<const double2*>(Y2 + 2 * lane);
            asm volatile(
                "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%4}, {%5}, {%0,%1};\n\t"
                "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%2,%3}, {%6}, {%7}, {%2,%3};\n\t"
                "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%8}, {%9}, {%0,%1};\n\t"
                "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%2,%3}, {%10}, {%11}, {%2,%3};"
                : "+d"(c1.x), "+d"(c1.y), "+d"(c2.x), "+d"(c2.y)
                : "d"(-xa1.x), "d"(yb1.x), "d"(-xa2.x), "d"(yb2.x), "d"(-xa1.y), "d"(yb1.y), "d"(-xa2.y), "d"(yb2.y));
            *reinterpret_cast<double2*>(C1 + r8) = c1;
            *reinterpret_cast<double2*>(C2 + r8) = c2;
          } else {
            if (k1 == 1) chd_tile_sub_xyT(C1, X1, Y1, lane);
            if (k2 == 1) chd_tile_sub_xyT(C2, X2, Y2, lane);
          }
          if (k1 == 2) corner(X1, Y1, b1i, b1j);
          if (k2 == 2) corner(X2, Y2, b2i, b2j);
        }
      }
    }
#ifdef CHD_PROFILE
    pf_c = clock64();
#endif
    if (use_tma && In < nbc) {
      chd_mbar_wait(&s_mbar, mbar_phase);
      mbar_phase ^= 1u;
    } else {
      chd_copy_wait(WS);
    }
    __syncthreads();
    kslot = kslot + 1 == Q ? 0 : kslot + 1;
#ifdef CHD_PROFILE
    { const long long t_ = clock64(); pf_acc[0] += pf_b - pf_a, pf_acc[1] += pf_c - pf_b, pf_acc[2] += t_ - pf_c; }
#endif
  }
#ifdef CHD_PROFILE
  if (tid == 32) I.dbg[0] += (double)pf_acc[0], I.dbg[1] += (double)pf_acc[1], I.dbg[2] += (double)pf_acc[2];   // warp 1: panel+barrier | own updates | wait+barrier
  if (tid == 0) I.dbg[3] += (double)pf_acc[0], I.dbg[4] += (double)pf_acc[1], I.dbg[5] += (double)pf_acc[2];    // warp 0: panel+barrier | diagonal tile | wait+barrier
#endif
  CHD_PROF(3);
  // dense LDL^T of the border Schur complement S = cc[0..nbl)^2 and solve S xb = rb (rb = row NBR of cc)
  for (int k = 0; k < nbl; ++k) {
    const double dk = cc[k * nbp8 + k];
    if (tid == 0 && !(dk > 0.0 && isfinite(dk))) s_fail = 1;
    __syncthreads();
    const double ik = 1.0 / dk;
    const int rem = nbl - k - 1;
    for (int idx = tid; idx < rem * rem; idx += nt) {
      const int i = k + 1 + idx / rem, j = k + 1 + idx % rem;
      if (j <= i) cc[i * nbp8 + j] -= cc[i * nbp8 + k] * ik * cc[j * nbp8 + k];
    }
    for (int j = k + 1 + tid; j < nbl; j += nt) cc[NBR * nbp8 + j] -= cc[j * nbp8 + k] * ik * cc[NBR * nbp8 + k];
    __syncthreads();
  }
  double* xb = xs + 8 * D.nbc_max;  // border solution
  if (warp == 0) {
    // column-oriented back-substitution by one warp: x_k = acc_k / d_k, then acc_j -= (L d)[k][j] x_k for j < k
    // (row k of cc holds the unscaled column entries; one division per unknown instead of one per matrix entry)
    for (int j = lane; j < nbl; j += 32) xb[j] = cc[NBR * nbp8 + j];
    __syncwarp();
    for (int k = nbl - 1; k >= 0; --k) {
      const double xk = xb[k] / cc[k * nbp8 + k];
      __syncwarp();
      for (int j = lane; j < k; j += 32) xb[j] -= cc[k * nbp8 + j] * xk;
      if (lane == 0) xb[k] = xk;
      __syncwarp();
    }
  }
  __syncthreads();
  double* sol = D.sol + (size_t)b * (D.Na_max + D.nb_max);
  for (int i = tid; i < nbl; i += nt) sol[Na + i] = xb[i];
  CHD_PROF(4);
  // backward substitution, column oriented: acc = u - Lb^T xb; for K descending: x_K = L0^-T acc_K (warp 0),
  // then acc_J -= L(K,J)^T x_K for the block row K (all threads).  Tiles are staged two block rows ahead.
  {
    double* accv = xpan;                  // panel buffers are free now; accv needs 8*nbc doubles <= (Q+nbt)*64? no -> use xs2
    accv = xs2;
    for (int i = tid; i < K.Np; i += nt) {
      const double* bt = K.bord + (size_t)(i >> 3) * nbt * 64 + (i & 7);
      double v = bt[NBR * 8];
      for (int q2 = 0; q2 < nbl; ++q2) v -= bt[q2 * 8] * xb[q2];
      accv[i] = v;
    }
    // staging buffers in the (now free) window + border window: two chunks of R block rows each,
    // row slot = diag tile | tiles (K, K-1-g), g = 0..q-1.  Warp 0 alone walks through a chunk (the recursion is
    // sequential anyway; without block barriers a block row costs ~0.5 k cycles instead of 1.4 k) while the other
    // warps prefetch the next chunk; one barrier per chunk.
    const int per = Q * 64;
    const int R = max(1, (int)((win_region - (WS ? n_even + xs_len : 0)) / 64) / (2 * Q));
    double* stg = win;
    auto stage_chunk = [&](int Ktop, int buf, int t0, int tstep) {   // block rows Ktop, Ktop-1, ... (R of them)
      for (int idx = t0; idx < R * Q * 32; idx += tstep) {
        const int rr = idx / (Q * 32), rem = idx - rr * (Q * 32);
        const int Kr = Ktop - rr;
        if (Kr < 0) break;
        const int tile = rem >> 5, off = (rem & 31) * 2;
        const int J = tile == 0 ? Kr : Kr - tile;     // tile 0: diagonal; tile g+1: (Kr, Kr-1-g)
        if (J < 0) continue;
        chd_copy16(stg + ((size_t)buf * R + rr) * per + tile * 64 + off, K.band + ((size_t)J * Qs + (Kr - J)) * 64 + off, WS);
      }
    };
    stage_chunk(nbc - 1, 0, tid, nt);
    chd_copy_wait(WS);
    __syncthreads();
    int buf = 0;
    for (int Ktop = nbc - 1; Ktop >= 0; Ktop -= R, buf ^= 1) {
      if (warp != 0) {
        stage_chunk(Ktop - R, buf ^ 1, tid - 32, nt - 32);
      } else {
        for (int rr = 0; rr < R && Ktop - rr >= 0; ++rr) {
          const int Kc = Ktop - rr;
          const double* T0 = stg + ((size_t)buf * R + rr) * per;
          double xk[8];
#pragma unroll
          for (int c = 7; c >= 0; --c) {
            double v = accv[Kc * 8 + c];
#pragma unroll
            for (int p = 7; p > c; --p) v -= T0[p * 8 + c] * xk[p];   // oldest unknown first: only the last link waits for xk[c+1]
            xk[c] = v;
          }
          if (lane < 8) {
            const int gi = Kc * 8 + lane;
            double v = xk[0];
#pragma unroll
            for (int c = 1; c < 8; ++c) v = lane == c ? xk[c] : v;
            xs[gi] = v;
            if (gi < Na) sol[gi] = v;
          }
          const int nrow = min(K.q, Kc);   // tiles (Kc, Kc-1-g), g < nrow
          for (int idx = lane; idx < nrow * 8; idx += 32) {
            const int g = idx >> 3, c = idx & 7;
            const double* T = T0 + (g + 1) * 64;
            double v = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) v += T[r * 8 + c] * xk[r];
            accv[(Kc - 1 - g) * 8 + c] -= v;
          }
          __syncwarp();
        }
      }
      chd_copy_wait(WS);
      __syncthreads();
    }
  }
  CHD_PROF(5);
  if (I.band_ovf) {
    // a coupling left the band (stage 3 moved a polynomial boundary further than the layout allows): the stage fails and
    // the schedule goes on with the fixed-duration stage 4, as the reference does after a failed stage 3
    __syncthreads();
    if (tid == 0) I.status = -2, chd_stage_advance(D, I, -2, sg.snap_after);
    return;
  }
  if (s_fail) {
    // numerical breakdown: raise the primal regularisation and retry next iteration (no step is taken)
    if (tid == 0) {
      I.delta_w = fmin(fmax(I.delta_w * 100.0, 1e-4), CHD_DW_MAX * 10);
      I.a_pr = 0.0, I.a_du = 0.0, I.dphi = 0.0;
      I.ls_fail += 1;
      I.step_ready = 1;
      I.kw_req = 1;
      if (I.delta_w > CHD_DW_MAX) I.status = -2, chd_stage_advance(D, I, -2, sg.snap_after);
    }
    for (int i = tid; i < n; i += nt) D.dx[vo + i] = 0.0;
    for (int r = tid; r < m; r += nt) D.ds[ro + r] = 0.0, D.dy[ro + r] = 0.0, D.dzL[ro + r] = 0.0, D.dzU[ro + r] = 0.0;
    return;
  }
  __syncthreads();

  // ---------------- D. recover the full step, fraction-to-the-boundary, line-search inputs ----------------
  double* dx = D.dx + vo;
  for (int i = tid; i < n; i += nt) {
    const int k = vk[i];
    const double v = (k >= 0 && i < n_act) ? sol[k] : 0.0;
    dx[i] = v;
    vecn[i] = v;   // step in the unknowns (switch times for the durations): what the Jacobian columns refer to
  }
  __syncthreads();
  if (sg.opt_dur) {   // the line search moves x, i.e. phase durations: dd_k = dtau_k - dtau_{k-1}
    for (int ee = 0; ee < h->n_ee; ++ee)
      for (int k = 1 + tid; k < h->n_phases[ee] - 1; k += nt) dx[h->dur_xoff[ee] + k] = vecn[h->dur_xoff[ee] + k] - vecn[h->dur_xoff[ee] + k - 1];
  }
  double a_pr = 1.0, a_du = 1.0, a_dphi = 0.0, a_phi = 0.0;
  for (int i = tid; i < n; i += nt) a_dphi += sf * grad[i] * vecn[i];
  for (int r = tid; r < m; r += nt) {
    const int f = rf[r];
    if (!(f & CHD_ROW_ACTIVE)) continue;
    if (f & CHD_ROW_EQ) {
      D.dy[ro + r] = sol[rk[r]];
      D.ds[ro + r] = 0.0;
      continue;
    }
    const double sc = D.sc[ro + r], s = D.s[ro + r];
    double Jdx = 0.0;
    for (int e = ep[r]; e < ep[r + 1]; ++e) {
      const int col = ec[e];
      if (col >= 0) Jdx += Jv[e] * vecn[col];
    }
    const double riq = sc * D.g[ro + r] - s;
    const double ds = sc * Jdx + riq;
    const bool hl = f & CHD_ROW_HASL, hu = f & CHD_ROW_HASU;
    const double gapL = hl ? s - D.dL[ro + r] : 1.0, gapU = hu ? D.dU[ro + r] - s : 1.0;
    const double zL = D.zL[ro + r], zU = D.zU[ro + r];
    const double sigL = hl ? zL / gapL : 0.0, sigU = hu ? zU / gapU : 0.0;
    const double bvec = (hl ? mu / gapL : 0.0) - (hu ? mu / gapU : 0.0);
    const double dy = (sigL + sigU) * ds - D.y[ro + r] - bvec;
    const double dzL = hl ? mu / gapL - zL - sigL * ds : 0.0;
    const double dzU = hu ? mu / gapU - zU + sigU * ds : 0.0;
    D.ds[ro + r] = ds, D.dy[ro + r] = dy, D.dzL[ro + r] = dzL, D.dzU[ro + r] = dzU;
    if (hl && ds < 0) a_pr = fmin(a_pr, -tau * gapL / ds);
    if (hu && ds > 0) a_pr = fmin(a_pr, tau * gapU / ds);
    if (hl && dzL < 0) a_du = fmin(a_du, -tau * zL / dzL);
    if (hu && dzU < 0) a_du = fmin(a_du, -tau * zU / dzU);
    if (hl) a_dphi -= mu * ds / gapL, a_phi -= mu * log(gapL);
    if (hu) a_dphi += mu * ds / gapU, a_phi -= mu * log(gapU);
  }
  a_pr = chd_block_min(a_pr, red);
  a_du = chd_block_min(a_du, red);
  const double dphi = chd_block_sum(a_dphi, red);
  const double phib = chd_block_sum(a_phi, red);
  if (tid == 0) {
    I.a_pr = a_pr, I.a_du = a_du, I.dphi = dphi;
    I.phi0 = sf * D.cost[2 * b] + phib;
    I.step_ready = 1;
    I.kw_req = 1;
  }
  CHD_PROF(6);
}

// Kwork <- Kbase for the sequences whose KKT kernel asked for it; runs on a side stream, overlapped with the line
// search / evaluation kernels of the next iteration, on the SMs the one-CTA-per-sequence kernels leave idle
__global__ void __launch_bounds__(256) chd_k_kcopy(ChdDev D) {
  const int b = blockIdx.y;
  if (!D.ipm[b].kw_req) return;
  if (blockIdx.x == 0) {   // right-hand side accumulators of chd_k_asm
    const size_t go = (size_t)b * (D.Na_max + D.nb_max);
    for (int i = threadIdx.x; i < D.Na_max + D.nb_max; i += blockDim.x) D.rhs0[go + i] = 0.0, D.rhs1[go + i] = 0.0;
  }
  if (D.stages[D.ipm[b].stage].opt_dur && D.ipm[b].phase == CHD_PH_RUN) {
    // stage 3: the cost Hessian moves with the durations; chd_k_hess_dur rebuilds it into a cleared Kwork after the line search
    chd_hess_clear(D, b, D.Kwork + (size_t)b * D.kstride, blockIdx.x, gridDim.x);
    return;
  }
  const size_t cnt2 = D.kstride / 2, per = (cnt2 + gridDim.x - 1) / gridDim.x;
  const size_t lo = (size_t)blockIdx.x * per, hi = lo + per < cnt2 ? lo + per : cnt2;
  const double2* src = reinterpret_cast<const double2*>(D.Kbase + (size_t)b * D.kstride);
  double2* dst = reinterpret_cast<double2*>(D.Kwork + (size_t)b * D.kstride);
  const size_t nt = blockDim.x;
  size_t i = lo + threadIdx.x;
  for (; i + 3 * nt < hi; i += 4 * nt) {
    const double2 v0 = src[i], v1 = src[i + nt], v2 = src[i + 2 * nt], v3 = src[i + 3 * nt];
    dst[i] = v0, dst[i + nt] = v1, dst[i + 2 * nt] = v2, dst[i + 3 * nt] = v3;
  }
  for (; i < hi; i += nt) dst[i] = src[i];
}

// distance-row curvature terms of the next iteration (needs the iterate the line search just accepted); side stream,
// after chd_k_kcopy, grid (G, B)
__global__ void __launch_bounds__(256) chd_k_curv(ChdDev D) {
  __shared__ double s_ws[8 * 192];
  const int b = blockIdx.y;
  const ChdIpm& I = D.ipm[b];
  if (!I.kw_req || I.phase != CHD_PH_RUN) return;
  ChdKT K;
  chd_kt_init(D, D.seq + b, D.Kwork + (size_t)b * D.kstride, K);
  K.ovf = &D.ipm[b].band_ovf;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  chd_curv_rows(D, b, K, D.stages[I.stage], blockIdx.x * 8 + warp, gridDim.x * 8, lane, s_ws + warp * 192);
}

// matrix part of the assembly for the sequences that continue in their stage (Kwork already refreshed by chd_k_kcopy
// and chd_k_curv): grid (G, B), launched right before chd_k_kkt.  The reductions into L2 are the cost, so G CTAs per
// sequence on the SMs the one-CTA-per-sequence kernels leave idle take 1/G of the time.
__global__ void __launch_bounds__(256) chd_k_asm(ChdDev D) {
  const int b = blockIdx.y;
  const ChdIpm& I = D.ipm[b];
  if (!I.kw_req || I.phase != CHD_PH_RUN) return;
  ChdKT K;
  chd_kt_init(D, D.seq + b, D.Kwork + (size_t)b * D.kstride, K);
  K.ovf = &D.ipm[b].band_ovf;
  const size_t go = (size_t)b * (D.Na_max + D.nb_max);
  chd_assemble(D, b, K, I.delta_w, I.mu, I.sf, nullptr, true, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, D.rhs0 + go, D.rhs1 + go);
}

// fp64 throughput probe for the roofline denominators: mode 0 = DFMA chains, mode 1 = DMMA (mma.sync m8n8k4 f64)
__global__ void __launch_bounds__(256) chd_k_fp64_peak(int mode, int iters, double* sink) {
  const double s = 1.0 + 1e-9 * threadIdx.x;
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.5 + i;
  if (mode == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], s, 1e-3);
    }
  } else {
    const double x = s, y = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(a[2 * i]), "+d"(a[2 * i + 1])
                     : "d"(x), "d"(y));
    }
  }
  double t = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += a[i];
  if (t == 123.456) sink[threadIdx.x] = t;
}

// the elimination window lives in shared memory (WS) or, for very wide bands, in a global scratch buffer
__global__ void __launch_bounds__(CHD_KKT_THREADS) chd_k_kkt(ChdDev D) { chd_kkt_body<true>(D); }
__global__ void __launch_bounds__(CHD_KKT_THREADS) chd_k_kkt_gwin(ChdDev D) { chd_kkt_body<false>(D); }
