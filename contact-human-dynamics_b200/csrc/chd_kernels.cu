// sm_100a kernels of the batched interior-point trajectory optimiser (product code).
// One CTA owns one sequence in every kernel; sequences are independent NLPs.
//
//   chd_k_stage_begin : row activity flags + IPM state reset for a stage
//   chd_k_eval        : f, grad f, g, Jacobian slot values at the current x       (HBM/L2 streaming)
//   chd_k_init        : gradient-based scaling, relaxed bounds, slack / multiplier initialisation
//   chd_k_kkt         : error measures + barrier update, condensed KKT assembly (band + border),
//                       unpivoted band LDL^T with dense border, triangular solves, step recovery
//   chd_k_linesearch  : filter backtracking line search (trial evaluations in shared memory), update
//   chd_k_sample      : SaveSolution sampling (phys_optim.cpp:63-143)
#include <cuda_runtime.h>

#include "chd_eval.cuh"

// ------------------------------------------------------------------ block helpers ----------------
__device__ __forceinline__ double chd_block_sum(double v, double* red) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  double r = red[0];
  __syncthreads();
  return r;
}
__device__ __forceinline__ double chd_block_max(double v, double* red) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + off]);
    __syncthreads();
  }
  double r = red[0];
  __syncthreads();
  return r;
}
__device__ __forceinline__ double chd_block_min(double v, double* red) { return -chd_block_max(-v, red); }

// ------------------------------------------------------------------ stage begin -------------------
__global__ void chd_k_stage_begin(ChdDev D, ChdStageDev sg, int max_iter) {
  const int b = blockIdx.x;
  const ChdSeq* h = D.seq + b;
  int* rf = D.rflag + (size_t)b * D.m_max;
  const double* lo = D.row_lo + (size_t)b * D.m_max;
  const double* hi = D.row_hi + (size_t)b * D.m_max;
  const int* rset = D.row_set + (size_t)b * D.m_max;
  for (int r = threadIdx.x; r < D.m_max; r += blockDim.x) {
    int f = 0;
    if (r < h->m && (sg.set_mask & CHD_MASK(rset[r]))) {
      f = CHD_ROW_ACTIVE;
      if (lo[r] == hi[r]) f |= CHD_ROW_EQ;
      else {
        if (lo[r] > -CHD_INF) f |= CHD_ROW_HASL;
        if (hi[r] < CHD_INF) f |= CHD_ROW_HASU;
      }
    }
    rf[r] = f;
    D.g[(size_t)b * D.m_max + r] = 0.0;
  }
  if (threadIdx.x == 0) {
    ChdIpm& I = D.ipm[b];
    I.status = 1;
    I.iter = 0;
    I.nfilt = 0;
    I.ls_fail = 0;
    I.max_iter = max_iter;
    I.mu = CHD_MU_INIT;
    I.delta_w = CHD_DELTA_W0;
    I.mu_filter = -1.0;
    I.sf = 1.0;
    for (int q = 0; q < 8; ++q) I.prof[q] = 0.0;
  }
}

// ------------------------------------------------------------------ evaluation --------------------
// dynamic shared memory: x[n_max] | grad[n_max] | red[CHD_THREADS]
__global__ void __launch_bounds__(CHD_THREADS) chd_k_eval(ChdDev D, ChdStageDev sg, int only_running) {
  extern __shared__ double sm[];
  const int b = blockIdx.x;
  if (only_running && D.ipm[b].status != 1) return;
  const ChdSeq* h = D.seq + b;
  double* xs = sm;
  double* gs = sm + D.n_max;
  double* red = sm + 2 * D.n_max;
  const double* x = D.x + (size_t)b * D.n_max;
  for (int i = threadIdx.x; i < h->n; i += blockDim.x) xs[i] = x[i], gs[i] = 0.0;
  __syncthreads();
  ChdCtx c;
  chd_make_ctx(D, b, xs, c);
  chd_eval_all<true>(c, sg, D.g + (size_t)b * D.m_max, D.Jv + (size_t)b * D.slots_max, gs, D.cost + 2 * b, red);
  double* grad = D.grad + (size_t)b * D.n_max;
  for (int i = threadIdx.x; i < h->n; i += blockDim.x) grad[i] = gs[i];
}

// ------------------------------------------------------------------ init --------------------------
// IPOPT gradient-based scaling (nlp_scaling_max_gradient 100), relaxed bounds, slack push, z = 1, y = 0.
__global__ void __launch_bounds__(CHD_THREADS) chd_k_init(ChdDev D) {
  __shared__ double red[CHD_THREADS];
  const int b = blockIdx.x;
  const ChdSeq* h = D.seq + b;
  const size_t ro = (size_t)b * D.m_max;
  const int* vk = D.var_kkt + (size_t)b * D.n_max;
  const double* grad = D.grad + (size_t)b * D.n_max;
  double gm = 0.0;
  for (int i = threadIdx.x; i < h->n; i += blockDim.x)
    if (vk[i] >= 0) gm = fmax(gm, fabs(grad[i]));
  gm = chd_block_max(gm, red);
  const double sf = gm > CHD_SCAL_MAX_GRAD ? CHD_SCAL_MAX_GRAD / gm : 1.0;
  const int* ep = D.ent_ptr + (size_t)b * (D.m_max + 1);
  const int* ec = D.ent_col + (size_t)b * D.slots_max;
  const double* Jv = D.Jv + (size_t)b * D.slots_max;
  int nb_l = 0;
  for (int r = threadIdx.x; r < h->m; r += blockDim.x) {
    const int f = D.rflag[ro + r];
    if (!(f & CHD_ROW_ACTIVE)) continue;
    double rm = 0.0;
    for (int e = ep[r]; e < ep[r + 1]; ++e) {
      const int col = ec[e];
      if (col >= 0 && vk[col] >= 0) rm = fmax(rm, fabs(Jv[e]));
    }
    double sc = rm > CHD_SCAL_MAX_GRAD ? CHD_SCAL_MAX_GRAD / rm : 1.0;
    sc = fmax(sc, 1e-8);
    D.sc[ro + r] = sc;
    const double lo = D.row_lo[ro + r] * sc, hi = D.row_hi[ro + r] * sc;
    const double d = sc * D.g[ro + r];
    D.y[ro + r] = 0.0;
    if (f & CHD_ROW_EQ) {
      D.dL[ro + r] = lo, D.dU[ro + r] = hi, D.s[ro + r] = d, D.zL[ro + r] = 0.0, D.zU[ro + r] = 0.0;
      continue;
    }
    const bool hl = f & CHD_ROW_HASL, hu = f & CHD_ROW_HASU;
    const double dL = hl ? lo - CHD_BOUND_RELAX * fmax(1.0, fabs(lo)) : -INFINITY;
    const double dU = hu ? hi + CHD_BOUND_RELAX * fmax(1.0, fabs(hi)) : INFINITY;
    double s = d;
    if (hl) {
      double pL = CHD_KAPPA1 * fmax(1.0, fabs(dL));
      if (hu) pL = fmin(pL, CHD_KAPPA2 * (dU - dL));
      s = fmax(s, dL + pL);
    }
    if (hu) {
      double pU = CHD_KAPPA1 * fmax(1.0, fabs(dU));
      if (hl) pU = fmin(pU, CHD_KAPPA2 * (dU - dL));
      s = fmin(s, dU - pU);
    }
    D.dL[ro + r] = dL, D.dU[ro + r] = dU, D.s[ro + r] = s;
    D.zL[ro + r] = hl ? 1.0 : 0.0;
    D.zU[ro + r] = hu ? 1.0 : 0.0;
    nb_l += (hl ? 1 : 0) + (hu ? 1 : 0);
  }
  const double nbnd = chd_block_sum((double)nb_l, red);
  int act = 0;
  for (int r = threadIdx.x; r < h->m; r += blockDim.x) act += (D.rflag[ro + r] & CHD_ROW_ACTIVE) ? 1 : 0;
  const double nact = chd_block_sum((double)act, red);
  if (threadIdx.x == 0) {
    D.ipm[b].sf = sf;
    D.ipm[b].n_bounds = (int)nbnd;
    D.ipm[b].m_act = (int)nact;
  }
}

// ------------------------------------------------------------------ KKT ---------------------------
struct ChdKkt {
  int Na, w, W, nbp, nbr;  // W = w_max + 1 (band column pitch), nbp = nb_max + 1 (border pitch), nbr = index of the rhs row
  double *band, *bord, *corn;
};
// add v at KKT position (i, j) of the symmetric matrix (lower triangle storage)
__device__ __forceinline__ void chd_kadd(const ChdKkt& K, int i, int j, double v) {
  if (i < j) { int t = i; i = j; j = t; }
  if (i < K.Na) atomicAdd(K.band + (size_t)j * K.W + (i - j), v);
  else if (j < K.Na) atomicAdd(K.bord + (size_t)j * K.nbp + (i - K.Na), v);
  else atomicAdd(K.corn + (size_t)(i - K.Na) * K.nbp + (j - K.Na), v);
}
__device__ __forceinline__ void chd_radd(const ChdKkt& K, int i, double v) {  // rhs
  if (i < K.Na) atomicAdd(K.bord + (size_t)i * K.nbp + K.nbr, v);
  else atomicAdd(K.corn + (size_t)K.nbr * K.nbp + (i - K.Na), v);
}

// Gauss-Newton Hessian of a least-squares cost sample: H += wgt * sum_dim J_dim^T J_dim, where the
// sample is sum over (up to two) located polynomials of sign * B(deriv) node values.
__device__ void chd_hess_sample(const ChdKkt& K, const int* vk, const ChdSpl* P, const double* sign, int np, int deriv, double wgt) {
  for (int a = 0; a < np * 12; ++a) {
    const int pa = a / 12, qa = a % 12, va = P[pa].var[qa];
    if (va < 0) continue;
    const int ia = vk[va];
    if (ia < 0) continue;
    const double wa = sign[pa] * chd_slot_w(P[pa], deriv, qa);
    if (wa == 0.0) continue;
    for (int bq = 0; bq < np * 12; ++bq) {
      const int pb = bq / 12, qb = bq % 12;
      if ((qb % 3) != (qa % 3)) continue;
      const int vb = P[pb].var[qb];
      if (vb < 0) continue;
      const int ib = vk[vb];
      if (ib < 0 || ia < ib) continue;
      const double wb = sign[pb] * chd_slot_w(P[pb], deriv, qb);
      if (wb != 0.0) chd_kadd(K, ia, ib, wgt * wa * wb);
    }
  }
}

// dynamic shared memory layout of chd_k_kkt (doubles):
//   red[CHD_THREADS] | vecn[n_max] (J^T y accumulation, later dx) | corn[nbp*nbp] | lvec[W+nbp] | xw[W+nbp] | win[W*W] | bwin[W*nbp]
__global__ void __launch_bounds__(CHD_THREADS) chd_k_kkt(ChdDev D, ChdStageDev sg) {
  extern __shared__ double sm[];
  const int b = blockIdx.x;
  ChdIpm& I = D.ipm[b];
  if (I.status != 1) return;
  const ChdSeq* h = D.seq + b;
  const int n = h->n, m = h->m, tid = threadIdx.x, nt = blockDim.x;
  const size_t ro = (size_t)b * D.m_max, vo = (size_t)b * D.n_max;
  const int* rf = D.rflag + ro;
  const int* vk = D.var_kkt + vo;
  const int* rk = D.row_kkt + ro;
  const int* ep = D.ent_ptr + (size_t)b * (D.m_max + 1);
  const int* ec = D.ent_col + (size_t)b * D.slots_max;
  const double* Jv = D.Jv + (size_t)b * D.slots_max;
  const double* grad = D.grad + vo;
  double* red = sm;
  double* vecn = sm + CHD_THREADS;
  ChdKkt K;
  K.Na = h->Na, K.w = h->w, K.W = D.w_max + 1, K.nbp = D.nb_max + 1, K.nbr = D.nb_max;
  K.band = D.Kband + (size_t)b * D.Na_max * K.W;
  K.bord = D.Kbord + (size_t)b * D.Na_max * K.nbp;
  K.corn = D.Kcorn + (size_t)b * K.nbp * K.nbp;
  const int NBR = D.nb_max;  // index of the rhs "row" in bord / corn
  const int nbl = h->nb;     // border unknowns of this sequence (<= nb_max)
  double* cc = vecn + D.n_max;
  double* lvec = cc + (size_t)K.nbp * K.nbp;
  double* xw = lvec + (K.W + K.nbp);
  // elimination window: shared memory when it fits, else a per-sequence global scratch (L2 resident)
  double* win = D.win_smem ? xw + (K.W + K.nbp) : D.scratch + (size_t)b * ((size_t)K.W * K.W + (size_t)K.W * K.nbp);
  double* bwin = win + (size_t)K.W * K.W;
  const double sf = I.sf;
  double mu = I.mu;
  long long tk0 = clock64();
#define CHD_PROF(slot) do { __syncthreads(); if (tid == 0) { long long t_ = clock64(); I.prof[slot] += (double)(t_ - tk0); tk0 = t_; } } while (0)

  // ---------------- A. error measures, convergence, barrier update ----------------
  for (int i = tid; i < n; i += nt) vecn[i] = sf * grad[i];
  __syncthreads();
  double a_ysum = 0, a_zsum = 0, a_cviol = 0, a_theta = 0, a_rs = 0, a_cmax = -INFINITY, a_cmin = INFINITY, a_violu = 0;
  for (int r = tid; r < m; r += nt) {
    const int f = rf[r];
    if (!(f & CHD_ROW_ACTIVE)) continue;
    const double sc = D.sc[ro + r], gval = D.g[ro + r], d = sc * gval, y = D.y[ro + r];
    a_ysum += fabs(y);
    a_violu = fmax(a_violu, fmax(D.row_lo[ro + r] - gval, gval - D.row_hi[ro + r]));
    const double ys = sc * y;
    for (int e = ep[r]; e < ep[r + 1]; ++e) {
      const int col = ec[e];
      if (col >= 0 && ys != 0.0) atomicAdd(vecn + col, ys * Jv[e]);
    }
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
  for (int i = tid; i < n; i += nt)
    if (vk[i] >= 0) a_dual = fmax(a_dual, fabs(vecn[i]));
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
    if (!done) {
      I.mu = mu;
      I.tau = fmax(CHD_TAU_MIN, 1.0 - mu);
      if (I.iter == 0) I.theta_max = 1e4 * fmax(1.0, theta), I.theta_min = 1e-4 * fmax(1.0, theta);
      if (mu != I.mu_filter) I.nfilt = 0, I.mu_filter = mu;
      I.theta0 = theta;
    }
  }
  if (done) return;
  const double tau = fmax(CHD_TAU_MIN, 1.0 - mu);
  const double delta_w = I.delta_w;

  CHD_PROF(0);
  // ---------------- B. assemble the condensed KKT system ----------------
  const int Na = K.Na, W = K.W, nbp = K.nbp;
  for (size_t i = tid; i < (size_t)Na * W; i += nt) K.band[i] = 0.0;
  for (size_t i = tid; i < (size_t)Na * nbp; i += nt) K.bord[i] = 0.0;
  for (int i = tid; i < nbp * nbp; i += nt) K.corn[i] = 0.0;
  __syncthreads();
  // diagonals + objective gradient part of the rhs
  for (int i = tid; i < n; i += nt) {
    const int k = vk[i];
    if (k < 0) continue;
    chd_kadd(K, k, k, delta_w);
    chd_radd(K, k, -sf * grad[i]);
  }
  for (int r = tid; r < m; r += nt) {
    const int k = rk[r];
    if (k < 0) continue;
    const int f = rf[r];
    if (!(f & CHD_ROW_ACTIVE)) { chd_kadd(K, k, k, -1.0); continue; }   // row of an inactive set: decoupled dummy unknown
    chd_kadd(K, k, k, -CHD_DELTA_C);
    const double sc = D.sc[ro + r];
    const double re = sc * D.g[ro + r] - D.dL[ro + r];
    chd_radd(K, k, -re);
    const double ys = sc * D.y[ro + r];
    for (int e = ep[r]; e < ep[r + 1]; ++e) {
      const int col = ec[e];
      if (col < 0) continue;
      const int kc = vk[col];
      if (kc < 0) continue;
      const double v = sc * Jv[e];
      if (v == 0.0) continue;
      chd_kadd(K, k, kc, v);
      chd_radd(K, kc, -ys * Jv[e]);
    }
  }
  // inequality rows: J^T Sigma J and rhs  -J^T (Sigma r - bvec)
  for (int r = tid; r < m; r += nt) {
    const int f = rf[r];
    if (!(f & CHD_ROW_ACTIVE) || (f & CHD_ROW_EQ)) continue;
    const double sc = D.sc[ro + r], s = D.s[ro + r];
    const double gapL = (f & CHD_ROW_HASL) ? s - D.dL[ro + r] : 1.0, gapU = (f & CHD_ROW_HASU) ? D.dU[ro + r] - s : 1.0;
    const double sigL = (f & CHD_ROW_HASL) ? D.zL[ro + r] / gapL : 0.0, sigU = (f & CHD_ROW_HASU) ? D.zU[ro + r] / gapU : 0.0;
    const double Sig = sigL + sigU;
    const double bvec = ((f & CHD_ROW_HASL) ? mu / gapL : 0.0) - ((f & CHD_ROW_HASU) ? mu / gapU : 0.0);
    const double riq = sc * D.g[ro + r] - s;
    const double coef = Sig * riq - bvec;
    const int e0 = ep[r], e1 = ep[r + 1];
    for (int ea = e0; ea < e1; ++ea) {
      const int ca = ec[ea];
      if (ca < 0) continue;
      const int ka = vk[ca];
      if (ka < 0) continue;
      const double va = sc * Jv[ea];
      if (va == 0.0) continue;
      chd_radd(K, ka, -va * coef);
      const double sva = Sig * va;
      for (int eb = e0; eb < e1; ++eb) {
        const int cb = ec[eb];
        if (cb < 0) continue;
        const int kb = vk[cb];
        if (kb < 0 || ka < kb) continue;
        const double vb = sc * Jv[eb];
        if (vb != 0.0) chd_kadd(K, ka, kb, sva * vb);
      }
    }
  }
  CHD_PROF(1);
  // Hessian model: Gauss-Newton cost Hessian + y^+ * Jd^T Jd of the squared-distance rows
  {
    ChdCtx c;
    chd_make_ctx(D, b, D.x + vo, c);
    const int n_ee = h->n_ee, nsp = 2 + n_ee, F = h->F, ns = h->n_smooth;
    for (int it = tid; it < nsp * F; it += nt) {
      const int s = it / F, i = it % F, cls = s < 2 ? s : 2;
      ChdSpl P[2];
      double sgn[2] = {1.0, -1.0};
      chd_spl_at(c, s, c.t_data[i], P[1]);
      if (sg.w_data[cls] != 0.0) chd_hess_sample(K, vk, P + 1, sgn, 1, 0, sf * sg.w_data[cls]);
      if (i < ns && (sg.w_vel[cls] != 0.0 || sg.w_acc[cls] != 0.0)) {
        chd_spl_at(c, s, c.t_data[i] + h->dt, P[0]);
        if (sg.w_vel[cls] != 0.0) chd_hess_sample(K, vk, P, sgn, 2, 0, sf * sg.w_vel[cls]);
        if (sg.w_acc[cls] != 0.0) chd_hess_sample(K, vk, P, sgn, 2, 1, sf * sg.w_acc[cls]);
      }
    }
    for (int si = 0; si < h->nsets; ++si) {
      const ChdSet st = c.sets[si];
      if (!(sg.set_mask & CHD_MASK(st.type))) continue;
      if (st.type != CHD_SET_ROM && st.type != CHD_SET_HEEL) continue;
      for (int k = tid; k < st.nitems; k += nt) {
        const int R = st.row0 + k;
        const double yc = D.sc[ro + R] * D.y[ro + R];
        if (!(yc > 0.0)) continue;
        const double t = c.t_rom[k];
        if (st.type == CHD_SET_HEEL) {
          ChdSpl P[2];
          double sgn[2] = {1.0, -1.0};
          chd_spl_at(c, chd_sp_motion(st.a), t, P[0]);
          chd_spl_at(c, chd_sp_motion(st.b), t, P[1]);
          chd_hess_sample(K, vk, P, sgn, 2, 0, yc);
        } else {
          // d = p_ee - R(e) h - c : Jd columns are dim-aligned for the lin / ee blocks and dR_j h for the angular block
          ChdSpl L, A, E;
          chd_spl_at(c, 0, t, L);
          chd_spl_at(c, 1, t, A);
          chd_spl_at(c, chd_sp_motion(st.a), t, E);
          double e[3];
          chd_spl_val(c, A, 0, e);
          ChdTrig tr;
          chd_trig(e, tr);
          const double* hip = chd_hip(c, st.a, t);
          double dRh[3][3];
          for (int j = 0; j < 3; ++j) {
            double Dj[9];
            chd_dR(tr, j, Dj);
            chd_mv(Dj, hip, dRh[j]);
          }
          // 36 columns: block 0 lin (-B e_dim), block 1 ang (-B dRh[dim]), block 2 ee (+B e_dim)
          for (int a = 0; a < 36; ++a) {
            const int ba = a / 12, qa = a % 12, da = qa % 3;
            const ChdSpl& Pa = ba == 0 ? L : (ba == 1 ? A : E);
            const int va = Pa.var[qa];
            if (va < 0) continue;
            const int ia = vk[va];
            if (ia < 0) continue;
            const double wa = (ba == 2 ? 1.0 : -1.0) * chd_slot_w(Pa, 0, qa);
            if (wa == 0.0) continue;
            for (int bq = 0; bq < 36; ++bq) {
              const int bb = bq / 12, qb = bq % 12, db = qb % 3;
              const ChdSpl& Pb = bb == 0 ? L : (bb == 1 ? A : E);
              const int vb = Pb.var[qb];
              if (vb < 0) continue;
              const int ib = vk[vb];
              if (ib < 0 || ia < ib) continue;
              const double wb = (bb == 2 ? 1.0 : -1.0) * chd_slot_w(Pb, 0, qb);
              if (wb == 0.0) continue;
              // column vectors: lin/ee -> e_dim, ang -> dRh[dim]
              double dotv;
              if (ba != 1 && bb != 1) dotv = da == db ? 1.0 : 0.0;
              else if (ba == 1 && bb == 1) dotv = chd_dot(dRh[da], dRh[db]);
              else if (ba == 1) dotv = dRh[da][db];
              else dotv = dRh[db][da];
              if (dotv != 0.0) chd_kadd(K, ia, ib, yc * wa * wb * dotv);
            }
          }
        }
      }
    }
  }
  __threadfence_block();
  __syncthreads();
  CHD_PROF(2);

  // ---------------- C. band LDL^T with dense border (right-looking, shared-memory window) ----------------
  // window column slot = column % W; bwin[slot*nbp + r] border rows (r < nbl) and rhs (r = NBR)
  for (int i = tid; i < nbp * nbp; i += nt) cc[i] = K.corn[i];
  for (int j = 0; j < W && j < Na; ++j) {
    for (int i = tid; i < W; i += nt) win[(size_t)j * W + i] = K.band[(size_t)j * W + i];
    for (int i = tid; i < nbp; i += nt) bwin[(size_t)j * nbp + i] = K.bord[(size_t)j * nbp + i];
  }
  __syncthreads();
  __shared__ int s_fail;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  for (int k = 0; k < Na; ++k) {
    const int slot = k % W;
    double* colk = win + (size_t)slot * W;
    double* bk = bwin + (size_t)slot * nbp;
    const double d = colk[0];
    if (tid == 0 && !(fabs(d) > 1e-300 && isfinite(d))) s_fail = 1;
    const double inv = 1.0 / d;
    const int lim = min(K.w, Na - 1 - k);  // rows k+1 .. k+lim exist
    for (int i = tid; i <= K.w; i += nt) lvec[i] = (i >= 1 && i <= lim) ? colk[i] * inv : 0.0;
    for (int i = tid; i < nbp; i += nt) lvec[W + i] = bk[i] * inv;   // scaled border column (and u_k = z_k / d at NBR)
    __syncthreads();
    // trailing band update: column j = k + cj
    for (int cj = 1 + warp; cj <= lim; cj += nwarp) {
      const double aj = colk[cj];
      if (aj == 0.0) continue;
      double* colj = win + (size_t)((k + cj) % W) * W;
      for (int ci = cj + lane; ci <= lim; ci += 32) colj[ci - cj] -= lvec[ci] * aj;
    }
    // border rows + rhs
    for (int idx = tid; idx < lim * nbp; idx += nt) {
      const int cj = 1 + idx / nbp, r = idx % nbp;
      if (r < nbl || r == NBR) bwin[(size_t)((k + cj) % W) * nbp + r] -= bk[r] * lvec[cj];
    }
    // corner (lower triangle incl. the rhs row NBR)
    for (int idx = tid; idx < nbp * nbl; idx += nt) {
      const int r = idx / nbl, q = idx % nbl;
      if ((r < nbl && q <= r) || r == NBR) cc[r * nbp + q] -= bk[r] * lvec[W + q];
    }
    __syncthreads();
    // write the finished column (d, L) and scaled border / u_k; then stream in column k + W
    for (int i = tid; i <= K.w; i += nt) K.band[(size_t)k * W + i] = i == 0 ? d : lvec[i];
    for (int i = tid; i < nbp; i += nt) K.bord[(size_t)k * nbp + i] = lvec[W + i];
    const int kn = k + W;
    if (kn < Na) {
      for (int i = tid; i < W; i += nt) colk[i] = K.band[(size_t)kn * W + i];
      for (int i = tid; i < nbp; i += nt) bk[i] = K.bord[(size_t)kn * nbp + i];
    }
    __syncthreads();
  }
  CHD_PROF(3);
  // dense Cholesky of the border Schur complement S = cc[0..nbl)^2 and solve S xb = rb (rb = row NBR of cc)
  for (int k = 0; k < nbl; ++k) {
    const double dk = cc[k * nbp + k];
    if (tid == 0 && !(dk > 0.0 && isfinite(dk))) s_fail = 1;
    __syncthreads();
    const double ik = 1.0 / dk;
    // column k of L (unit lower with D): l_i = S_ik / dk ; update S_ij -= l_i S_jk
    for (int idx = tid; idx < (nbl - k - 1) * (nbl - k - 1); idx += nt) {
      const int i = k + 1 + idx / (nbl - k - 1), j = k + 1 + idx % (nbl - k - 1);
      if (j <= i) cc[i * nbp + j] -= cc[i * nbp + k] * ik * cc[j * nbp + k];
    }
    for (int j = k + 1 + tid; j < nbl; j += nt) cc[NBR * nbp + j] -= cc[j * nbp + k] * ik * cc[NBR * nbp + k];  // forward on rhs
    __syncthreads();
  }
  if (tid == 0) {
    // backward: xb = L^-T D^-1 z   (nbl is small)
    for (int k = nbl - 1; k >= 0; --k) {
      double v = cc[NBR * nbp + k] / cc[k * nbp + k];
      for (int i = k + 1; i < nbl; ++i) v -= (cc[i * nbp + k] / cc[k * nbp + k]) * xw[W + i];
      xw[W + k] = v;
    }
  }
  __syncthreads();
  double* sol = D.sol + (size_t)b * (D.Na_max + D.nb_max);
  for (int i = tid; i < nbl; i += nt) sol[Na + i] = xw[W + i];
  CHD_PROF(4);
  // backward substitution on the band: x_k = u_k - sum_i L[k+i][k] x_{k+i} - sum_b Lb[b][k] xb
  // chunks of CH columns are staged in shared memory by all threads, then swept by warp 0
  {
    const int CH = min(32, W);
    for (int k1 = Na - 1; k1 >= 0; k1 -= CH) {
      const int k0 = max(k1 - CH + 1, 0), cnt = k1 - k0 + 1;
      for (int idx = tid; idx < cnt * W; idx += nt) win[idx] = K.band[(size_t)(k0 + idx / W) * W + idx % W];
      for (int idx = tid; idx < cnt * nbp; idx += nt) bwin[idx] = K.bord[(size_t)(k0 + idx / nbp) * nbp + idx % nbp];
      __syncthreads();
      if (warp == 0) {
        for (int k = k1; k >= k0; --k) {
          const double* Lc = win + (size_t)(k - k0) * W;
          const double* Bc = bwin + (size_t)(k - k0) * nbp;
          double acc = 0.0;
          const int lim = min(K.w, Na - 1 - k);
          for (int i = 1 + lane; i <= lim; i += 32) acc += Lc[i] * xw[(k + i) % W];
          for (int q = lane; q < nbl; q += 32) acc += Bc[q] * xw[W + q];
          for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
          const double xk = Bc[NBR] - acc;
          if (lane == 0) xw[k % W] = xk, sol[k] = xk;
          __syncwarp();
        }
      }
      __syncthreads();
    }
  }
  if (s_fail) {
    // numerical breakdown: raise the primal regularisation and retry next iteration (no step is taken)
    if (tid == 0) {
      I.delta_w = fmin(fmax(I.delta_w * 100.0, 1e-4), CHD_DW_MAX * 10);
      I.a_pr = 0.0, I.a_du = 0.0, I.dphi = 0.0;
      I.ls_fail += 1;
      if (I.delta_w > CHD_DW_MAX) I.status = -2;
    }
    for (int i = tid; i < n; i += nt) D.dx[vo + i] = 0.0;
    for (int r = tid; r < m; r += nt) D.ds[ro + r] = 0.0, D.dy[ro + r] = 0.0, D.dzL[ro + r] = 0.0, D.dzU[ro + r] = 0.0;
    return;
  }
  __syncthreads();

  CHD_PROF(5);
  // ---------------- D. recover the full step, fraction-to-the-boundary, line-search inputs ----------------
  double* dx = D.dx + vo;
  for (int i = tid; i < n; i += nt) {
    const int k = vk[i];
    const double v = k >= 0 ? sol[k] : 0.0;
    dx[i] = v;
    vecn[i] = v;
  }
  __syncthreads();
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
  }
  CHD_PROF(6);
}

// ------------------------------------------------------------------ line search -------------------
// dynamic shared memory: xt[n_max] | red[CHD_THREADS]
__global__ void __launch_bounds__(CHD_THREADS) chd_k_linesearch(ChdDev D, ChdStageDev sg) {
  extern __shared__ double sm[];
  __shared__ int s_ok, s_ftype;
  __shared__ double s_cost;
  const int b = blockIdx.x;
  ChdIpm& I = D.ipm[b];
  if (I.status != 1) return;
  const ChdSeq* h = D.seq + b;
  const int n = h->n, m = h->m, tid = threadIdx.x, nt = blockDim.x;
  const size_t ro = (size_t)b * D.m_max, vo = (size_t)b * D.n_max;
  const int* rf = D.rflag + ro;
  double* xt = sm;
  double* red = sm + D.n_max;
  const double* x = D.x + vo;
  const double* dx = D.dx + vo;
  double* gt = D.gt + ro;
  const double mu = I.mu, sf = I.sf, theta = I.theta0, phi0 = I.phi0, dphi = I.dphi, a_pr = I.a_pr, a_du = I.a_du;
  if (a_pr == 0.0 && dphi == 0.0) {  // failed factorisation: nothing to do this iteration
    if (tid == 0) I.iter += 1;
    return;
  }
  ChdCtx c;
  chd_make_ctx(D, b, xt, c);
  double alpha = a_pr;
  int ls = 0;
  bool accepted = false, ftype = false;
  for (ls = 0; ls < CHD_MAX_BACKTRACK; ++ls) {
    for (int i = tid; i < n; i += nt) xt[i] = x[i] + alpha * dx[i];
    __syncthreads();
    chd_eval_all<false>(c, sg, gt, nullptr, nullptr, &s_cost, red);
    double a_th = 0.0, a_bar = 0.0;
    for (int r = tid; r < m; r += nt) {
      const int f = rf[r];
      if (!(f & CHD_ROW_ACTIVE)) continue;
      const double d = D.sc[ro + r] * gt[r];
      if (f & CHD_ROW_EQ) a_th += fabs(d - D.dL[ro + r]);
      else {
        const double st = D.s[ro + r] + alpha * D.ds[ro + r];
        a_th += fabs(d - st);
        if (f & CHD_ROW_HASL) a_bar -= mu * log(st - D.dL[ro + r]);
        if (f & CHD_ROW_HASU) a_bar -= mu * log(D.dU[ro + r] - st);
      }
    }
    const double theta_t = chd_block_sum(a_th, red);
    const double bar_t = chd_block_sum(a_bar, red);
    if (tid == 0) {
      const double phit = sf * s_cost + bar_t;
      bool ok = isfinite(phit) && isfinite(theta_t) && theta_t <= I.theta_max;
      for (int q = 0; ok && q < I.nfilt; ++q)
        if (theta_t >= I.filt[2 * q] && phit >= I.filt[2 * q + 1]) ok = false;
      bool acc = false, ft = false;
      if (ok) {
        const bool switching = dphi < 0 && alpha * pow(-dphi, CHD_S_PHI) > pow(theta, CHD_S_THETA);
        const bool armijo = phit <= phi0 + CHD_ETA_PHI * alpha * dphi;
        if (theta <= I.theta_min && switching) {
          if (armijo) acc = true, ft = true;
        } else if (theta_t <= (1 - CHD_GAMMA_THETA) * theta || phit <= phi0 - CHD_GAMMA_PHI * theta) {
          acc = true;
          ft = switching && armijo;
        }
      }
      s_ok = acc, s_ftype = ft;
    }
    __syncthreads();
    accepted = s_ok, ftype = s_ftype;
    __syncthreads();
    if (accepted) break;
    alpha *= 0.5;
  }
  if (!accepted) {
    alpha *= 2.0;  // the last (smallest) trial step is taken, as in oracle/ipm_proto.py
    ls = CHD_MAX_BACKTRACK;
  }
  // x was left at the accepted trial point in xt
  double* xg = D.x + vo;
  for (int i = tid; i < n; i += nt) xg[i] = xt[i];
  for (int r = tid; r < m; r += nt) {
    const int f = rf[r];
    if (!(f & CHD_ROW_ACTIVE)) continue;
    D.y[ro + r] += alpha * D.dy[ro + r];
    if (f & CHD_ROW_EQ) continue;
    const double s = D.s[ro + r] + alpha * D.ds[ro + r];
    D.s[ro + r] = s;
    if (f & CHD_ROW_HASL) {
      const double gap = s - D.dL[ro + r];
      double z = D.zL[ro + r] + a_du * D.dzL[ro + r];
      z = fmin(fmax(z, mu / (CHD_KAPPA_SIGMA * gap)), CHD_KAPPA_SIGMA * mu / gap);
      D.zL[ro + r] = z;
    }
    if (f & CHD_ROW_HASU) {
      const double gap = D.dU[ro + r] - s;
      double z = D.zU[ro + r] + a_du * D.dzU[ro + r];
      z = fmin(fmax(z, mu / (CHD_KAPPA_SIGMA * gap)), CHD_KAPPA_SIGMA * mu / gap);
      D.zU[ro + r] = z;
    }
  }
  if (tid == 0) {
    if (accepted && !ftype && I.nfilt < CHD_FILT_MAX) {
      I.filt[2 * I.nfilt] = (1 - CHD_GAMMA_THETA) * theta;
      I.filt[2 * I.nfilt + 1] = phi0 - CHD_GAMMA_PHI * theta;
      I.nfilt += 1;
    }
    if (!accepted) I.ls_fail += 1;
    // Levenberg-Marquardt style adaptation of the primal regularisation
    if (ls == 0) I.delta_w = fmax(I.delta_w / CHD_DW_DEC, CHD_DW_MIN);
    else I.delta_w = fmin(I.delta_w * pow(CHD_DW_INC, (double)min(ls, 3)), CHD_DW_MAX);
    I.iter += 1;
  }
}

// ------------------------------------------------------------------ sampling ----------------------
// SaveSolution (phys_optim.cpp:63-143): t accumulates dt while t <= T + 1e-5.
// out: B x fo_max x (6 + 7 n_ee_max)
__global__ void chd_k_sample(ChdDev D, double* out, int* frames_out) {
  const int b = blockIdx.x;
  const ChdSeq* h = D.seq + b;
  ChdCtx c;
  chd_make_ctx(D, b, D.x + (size_t)b * D.n_max, c);
  const int n_ee = h->n_ee, stride = 6 + 7 * D.n_ee_max;
  double tot = 0.0;
  for (int k = 0; k < h->sp_npoly[0]; ++k) tot += c.poly_T[k];  // Spline::GetTotalTime of base_linear
  const int nf = (int)((tot + 1e-5) / h->dt) + 1;
  if (threadIdx.x == 0 && frames_out) frames_out[b] = nf;
  for (int i = threadIdx.x; i < nf; i += blockDim.x) {
    double t = 0.0;
    for (int q = 0; q < i; ++q) t += h->dt;  // same accumulation as the reference loop
    if (!(t <= tot + 1e-5)) continue;
    double* o = out + ((size_t)b * D.fo_max + i) * stride;
    ChdSpl P;
    double v[3];
    chd_spl_at(c, 0, t, P);
    chd_spl_val(c, P, 0, v);
    o[0] = v[0], o[1] = v[1], o[2] = v[2];
    chd_spl_at(c, 1, t, P);
    chd_spl_val(c, P, 0, v);
    for (int d = 0; d < 3; ++d) o[3 + d] = v[d] / M_PI * 180;
    for (int ee = 0; ee < n_ee; ++ee) {
      chd_spl_at(c, chd_sp_motion(ee), t, P);
      chd_spl_val(c, P, 0, v);
      for (int d = 0; d < 3; ++d) o[6 + 3 * ee + d] = v[d];
      chd_spl_at(c, chd_sp_force(n_ee, ee), t, P);
      chd_spl_val(c, P, 0, v);
      for (int d = 0; d < 3; ++d) o[6 + 3 * D.n_ee_max + 3 * ee + d] = v[d];
      // PhaseDurations::IsContactPhase (phys_optim.cpp:135): phase id parity against the start flag
      {
        const double* pt = D.phase_tend + ((size_t)b * D.n_ee_max + ee) * D.Ph_max;
        double tl;
        const int ph = chd_locate(pt, h->n_phases[ee], t, &tl);
        o[6 + 6 * D.n_ee_max + ee] = ((ph % 2 == 0) ? h->start_contact[ee] : !h->start_contact[ee]) ? 1.0 : 0.0;
      }
    }
  }
}
