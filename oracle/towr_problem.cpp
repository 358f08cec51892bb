// TEST INFRASTRUCTURE ONLY -- see towr_oracle.hpp.
// NLP assembly (ifopt stacking), constraint sets, cost terms, staged schedule and the C API that
// tests / bench.py's cpu_baseline leg load through ctypes.
#include <algorithm>
#include <cassert>
#include <cstdio>
#include <functional>
#include <numeric>

#include "towr_oracle.hpp"

namespace oracle {

struct Row {
  std::vector<std::pair<int, double>> e;
  void add(int c, double v) { e.emplace_back(c, v); }
  // row += sign * d^T J   (J columns shifted by off)
  void add_dT_J(const SJac& J, int off, const Vec3& d, double sign) {
    for (size_t i = 0; i < J.col.size(); ++i) add(J.col[i] + off, sign * dot(d, J.val[i]));
  }
};

struct Problem;

struct Constraint {
  Problem* P = nullptr;
  int rows = 0;
  std::string name;
  virtual ~Constraint() = default;
  virtual void values(double* g) const = 0;
  virtual void bounds(double* lo, double* hi) const = 0;
  virtual void jac(Row* r) const = 0;
  // sum_i y_i * (second derivative model of row i), full symmetric triplets.  Only the terms chd-ipm
  // keeps: exact bilinear f x (c - p) of the dynamics rows, and y * Jd^T Jd of the
  // squared-distance rows (leg length, toe-heel distance).  Linear rows contribute nothing.
  virtual void lag_hessian(const double* y, std::vector<std::array<double, 3>>& trip) const {}
  // time stamp of every row (layout diagnostics: KKT ordering studies in tests)
  virtual void row_times(double* t) const = 0;
};
struct Cost {
  Problem* P = nullptr;
  std::string name;
  virtual ~Cost() = default;
  virtual double value() const = 0;
  virtual void grad(double* g) const = 0;                       // accumulates into the stacked gradient
  virtual void gn_hessian(std::vector<std::array<double, 3>>& trip) const = 0;  // (i, j, v) full symmetric
};

// towr TimeDiscretizationConstraint: 0, dt, 2dt, ... (accumulated) plus T itself.
static std::vector<double> discretize(double T, double dt) {
  double t = 0.0;
  std::vector<double> dts = {t};
  for (int i = 0; i < floor(T / dt); ++i) {
    t += dt;
    dts.push_back(t);
  }
  dts.push_back(T);
  return dts;
}

static void add_outer(std::vector<std::array<double, 3>>& trip, const std::vector<std::pair<int, Vec3>>& cols, double w) {
  for (auto& a : cols)
    for (auto& b : cols) {
      double v = w * dot(a.second, b.second);
      if (v != 0.0) trip.push_back({(double)a.first, (double)b.first, v});
    }
}
static void gather(std::vector<std::pair<int, Vec3>>& out, const SJac& J, int off, double s) {
  for (size_t i = 0; i < J.col.size(); ++i) {
    int c = J.col[i] + off;
    bool found = false;
    for (auto& o : out)
      if (o.first == c) {
        o.second = o.second + s * J.val[i];
        found = true;
        break;
      }
    if (!found) out.emplace_back(c, s * J.val[i]);
  }
}


struct Problem {
  // ---- inputs (phys_optim.cpp:380-417) ----
  int F = 0, n_ee = 0;
  double dt = 0, T = 0;
  std::vector<Vec3> hip_l, hip_r;
  double max_leg = 0, max_heel = 0, heel_dist = 0, mass = 0;
  std::vector<std::array<double, 6>> inertia;
  std::vector<Vec3> lin_data, ang_data;
  std::vector<std::vector<Vec3>> ee_data;  // solver order: Ltoe, Rtoe, Lheel, Rheel
  Vec3 normal{0, 0, 1}, point{0, 0, 0};
  std::vector<int> start_contact;
  std::vector<std::vector<double>> init_dur;
  double w_com_lin = 0.4, w_com_ang = 1.7, w_ee = 0.3, w_smooth = 0.1, w_dur = 0.1;
  // ---- Parameters (parameters.cpp:46-70) ----
  double base_poly_dur = 0.1, add_polys_after = 2.0;
  int polys_per_swing = 6, polys_per_stance = 6;
  double force_limit = 1000, dt_rom = 0.08, dt_height = 0.1, dt_dyn = 0.1;
  double dur_lo = 0.0, dur_hi = 500.0;
  double mu = 0.5;          // towr HeightMap default friction (assumed)
  double g = 9.80665;       // towr DynamicModel gravity (assumed)
  // ---- variables ----
  NodesVars base_lin, base_ang;
  std::vector<NodesVars> ee_motion, ee_force;
  std::vector<PhaseDurations> dur;
  Spline s_lin, s_ang;
  std::vector<Spline> s_motion, s_force;
  Euler euler;
  // ---- stage ----
  int stage = -1;
  bool opt_dur = false;
  std::vector<std::unique_ptr<Constraint>> cons;
  std::vector<std::unique_ptr<Cost>> costs;
  int n = 0, m = 0;

  double terrain_height(double x, double y) const {  // ground_plane.cpp:18-26
    double z = -normal[1] * (y - point[1]) - normal[0] * (x - point[0]);
    z /= normal[2];
    z += point[2];
    return z;
  }
  double dhdx() const { return -normal[0] / normal[2]; }  // ground_plane.cpp:28-40
  double dhdy() const { return -normal[1] / normal[2]; }
  Vec3 gvec() const {  // humanoid_rigid_body_dynamics.cpp:208-211 with g = -floor_normal (phys_optim.cpp:437)
    double nn = sqrt(dot(normal, normal));
    return {-normal[0] / nn, -normal[1] / nn, -normal[2] / nn};
  }
  Mat3 inertia_at(double t) const {  // humanoid_rigid_body_dynamics.cpp:81-87
    int idx = (int)((t / T) * F);
    if (idx == F) idx -= 1;
    const auto& I = inertia[idx];
    Mat3 M = {{{I[0], I[3], I[4]}, {I[3], I[1], I[5]}, {I[4], I[5], I[2]}}};
    return M;
  }
  Vec3 hip_at(int ee, double t) const {  // leg_length_constraint.cpp:40-44, humanoid.h:45-48
    int idx = (int)((t / T) * F);
    if (idx == F) idx -= 1;
    return (ee == 0 || ee == 2) ? hip_l[idx] : hip_r[idx];
  }

  void build_variables();
  void set_stage(int st);
  void get_x(double* x) const;
  void set_x(const double* x);
  void update_splines() {
    for (auto& s : s_motion) s.update_durations();
    for (auto& s : s_force) s.update_durations();
  }
};

// phys_optim.cpp:289-312
static std::vector<int> polys_changing_phase(bool start_constant, const std::vector<double>& durations, double max_dur,
                                             int n_polys_per_change) {
  std::vector<int> out;
  bool is_constant = start_constant;
  double per_s = n_polys_per_change / max_dur;
  for (double d : durations) {
    if (!is_constant) {
      int np = n_polys_per_change;
      if (d > max_dur) np += (int)ceil((d - max_dur) * per_s);
      out.push_back(np);
    }
    is_constant = !is_constant;
  }
  return out;
}

void Problem::build_variables() {
  // total time: Parameters::GetTotalTime takes the first foot as reference (parameters.cpp:138-153)
  T = std::accumulate(init_dur[0].begin(), init_dur[0].end(), 0.0);
  // base polynomial durations (parameters.cpp:109-125)
  std::vector<double> base_dur;
  {
    double t_left = T, eps = 1e-10;
    while (t_left > eps) {
      base_dur.push_back(t_left > base_poly_dur ? base_poly_dur : t_left);
      t_left -= base_poly_dur;
    }
  }
  int n_nodes = (int)base_dur.size() + 1;
  // initial / final base state (phys_optim.cpp:442-489): positions from frame 0 / F-1, linear velocity =
  // mean of the first / last 5 finite differences; angular velocity is computed but never applied.
  const int avg = 5;
  Vec3 v0{0, 0, 0}, vf{0, 0, 0};
  for (int i = 0; i < avg; ++i) {
    v0 = v0 + (1.0 / dt) * (lin_data[i + 1] - lin_data[i]);
    vf = vf + (1.0 / dt) * (lin_data[F - 1 - i] - lin_data[F - 2 - i]);
  }
  v0 = (1.0 / avg) * v0;
  vf = (1.0 / avg) * vf;
  // nlp_formulation.cpp:106-131
  base_lin = make_nodes_all(n_nodes, "base-lin");
  base_lin.set_by_linear_interpolation(lin_data[0], lin_data[F - 1], T);
  base_lin.add_bound(0, kVel, v0);
  base_lin.add_bound(n_nodes - 1, kVel, vf);
  base_ang = make_nodes_all(n_nodes, "base-ang");
  base_ang.set_by_linear_interpolation(ang_data[0], ang_data[F - 1], T);

  ee_motion.clear();
  ee_force.clear();
  dur.clear();
  ee_motion.reserve(n_ee);
  ee_force.reserve(n_ee);
  dur.reserve(n_ee);
  for (int ee = 0; ee < n_ee; ++ee) {
    // phys_optim.cpp:516-540
    auto swing_polys = polys_changing_phase(start_contact[ee], init_dur[ee], add_polys_after, polys_per_swing);
    auto stance_polys = polys_changing_phase(!start_contact[ee], init_dur[ee], add_polys_after, polys_per_stance);
    // nlp_formulation.cpp:133-162: initialise on the line from the first-frame foot position to the
    // final base XY projected on the floor
    NodesVars mv = make_ee_motion((int)init_dur[ee].size(), start_contact[ee], "ee-motion_" + std::to_string(ee), swing_polys);
    double fx = lin_data[F - 1][0], fy = lin_data[F - 1][1];
    mv.set_by_linear_interpolation(ee_data[ee][0], Vec3{fx, fy, terrain_height(fx, fy)}, T);
    ee_motion.push_back(std::move(mv));
    // nlp_formulation.cpp:164-186
    NodesVars fv = make_ee_force((int)init_dur[ee].size(), start_contact[ee], "ee-force_" + std::to_string(ee), stance_polys);
    Vec3 fs{0, 0, mass * g / n_ee};
    fv.set_by_linear_interpolation(fs, fs, T);
    ee_force.push_back(std::move(fv));
    // nlp_formulation.cpp:188-203
    PhaseDurations pdur;
    pdur.durations = init_dur[ee];
    pdur.t_total = std::accumulate(init_dur[ee].begin(), init_dur[ee].end(), 0.0);
    pdur.initial_contact = start_contact[ee];
    pdur.lo = dur_lo;
    pdur.hi = dur_hi;
    dur.push_back(pdur);
  }
  // shared stance variables: ifopt hands GetValues() back through SetVariables(), so both nodes of a
  // stance phase end up at the value of the last NodeValueInfo.
  for (auto* nv : {&base_lin, &base_ang}) {
    std::vector<double> x(nv->rows());
    nv->get_values(x.data());
    nv->set_values(x.data());
  }
  for (auto& nv : ee_motion) {
    std::vector<double> x(nv.rows());
    nv.get_values(x.data());
    nv.set_values(x.data());
  }
  // splines (towr SplineHolder)
  s_lin.nv = &base_lin;
  s_lin.poly_dur = base_dur;
  s_ang.nv = &base_ang;
  s_ang.poly_dur = base_dur;
  s_motion.assign(n_ee, Spline());
  s_force.assign(n_ee, Spline());
  for (int ee = 0; ee < n_ee; ++ee) {
    s_motion[ee].nv = &ee_motion[ee];
    s_motion[ee].pd = &dur[ee];
    s_force[ee].nv = &ee_force[ee];
    s_force[ee].pd = &dur[ee];
  }
  update_splines();
  euler.s = &s_ang;
  // ifopt stacking (nlp_formulation.cpp:84-91): base_lin, base_ang, ee motions, ee forces [, durations]
  int off = 0;
  base_lin.offset = off;
  off += base_lin.rows();
  base_ang.offset = off;
  off += base_ang.rows();
  for (auto& v : ee_motion) {
    v.offset = off;
    off += v.rows();
  }
  for (auto& v : ee_force) {
    v.offset = off;
    off += v.rows();
  }
  for (auto& d : dur) {
    d.offset = off;
    off += d.rows();
  }
}

void Problem::get_x(double* x) const {
  base_lin.get_values(x + base_lin.offset);
  base_ang.get_values(x + base_ang.offset);
  for (auto& v : ee_motion) v.get_values(x + v.offset);
  for (auto& v : ee_force) v.get_values(x + v.offset);
  if (opt_dur)
    for (auto& d : dur)
      for (int i = 0; i < d.rows(); ++i) x[d.offset + i] = d.durations[i];
}
void Problem::set_x(const double* x) {
  base_lin.set_values(x + base_lin.offset);
  base_ang.set_values(x + base_ang.offset);
  for (auto& v : ee_motion) v.set_values(x + v.offset);
  for (auto& v : ee_force) v.set_values(x + v.offset);
  if (opt_dur)
    for (auto& d : dur) d.set_values(x + d.offset);
  update_splines();
}

// ============================================================ constraints =======================
// towr SplineAccConstraint (used nlp_formulation.cpp:349-360)
struct SplineAcc : Constraint {
  const Spline* s;
  const NodesVars* nv;
  int nj;
  SplineAcc(Problem* p, const Spline* sp, const NodesVars* v) {
    P = p;
    s = sp;
    nv = v;
    nj = (int)s->poly_dur.size() - 1;
    rows = 3 * nj;
    name = "splineacc-" + v->name;
  }
  void values(double* g) const override {
    for (int j = 0; j < nj; ++j) {
      Vec3 a0 = s->point(j, s->poly_dur[j]).a, a1 = s->point(j + 1, 0.0).a;
      for (int d = 0; d < 3; ++d) g[3 * j + d] = a0[d] - a1[d];
    }
  }
  void bounds(double* lo, double* hi) const override {
    for (int i = 0; i < rows; ++i) lo[i] = hi[i] = 0.0;
  }
  void row_times(double* t) const override {
    double acc = 0;
    for (int j = 0; j < nj; ++j) {
      acc += s->poly_dur[j];
      for (int d = 0; d < 3; ++d) t[3 * j + d] = acc;
    }
  }
  void jac(Row* r) const override {
    for (int j = 0; j < nj; ++j) {
      SJac a0 = s->jac_wrt_nodes(j, s->poly_dur[j], kAcc), a1 = s->jac_wrt_nodes(j + 1, 0.0, kAcc);
      a0.axpy(-1.0, a1);
      for (size_t c = 0; c < a0.col.size(); ++c)
        for (int d = 0; d < 3; ++d)  // a node value only feeds the row of its own dimension
          if (nv->index_map[a0.col[c]][0].dim == d) r[3 * j + d].add(a0.col[c] + nv->offset, a0.val[c][d]);
    }
  }
};

// towr TerrainConstraint (used nlp_formulation.cpp:321-331); node 0 is skipped.
struct Terrain : Constraint {
  int ee;
  Terrain(Problem* p, int e) {
    P = p;
    ee = e;
    rows = (int)P->ee_motion[ee].nodes.size() - 1;
    name = "terrain-ee-motion_" + std::to_string(ee);
  }
  void values(double* g) const override {
    const auto& nv = P->ee_motion[ee];
    for (int id = 1; id < (int)nv.nodes.size(); ++id) {
      const Vec3& p = nv.nodes[id].p;
      g[id - 1] = p[2] - P->terrain_height(p[0], p[1]);
    }
  }
  void bounds(double* lo, double* hi) const override {
    const auto& nv = P->ee_motion[ee];
    for (int id = 1; id < (int)nv.nodes.size(); ++id) {
      lo[id - 1] = 0.0;
      hi[id - 1] = nv.is_constant_node(id) ? 0.0 : 1e20;
    }
  }
  void row_times(double* t) const override {
    const auto& sp = P->s_motion[ee];
    double acc = 0;
    for (int id = 1; id < (int)sp.nv->nodes.size(); ++id) {
      acc += sp.poly_dur[id - 1];
      t[id - 1] = acc;
    }
  }
  void jac(Row* r) const override {
    const auto& nv = P->ee_motion[ee];
    for (int id = 1; id < (int)nv.nodes.size(); ++id) {
      r[id - 1].add(nv.rev[id][kPos * 3 + Z] + nv.offset, 1.0);
      r[id - 1].add(nv.rev[id][kPos * 3 + X] + nv.offset, -P->dhdx());
      r[id - 1].add(nv.rev[id][kPos * 3 + Y] + nv.offset, -P->dhdy());
    }
  }
};

// towr ForceConstraint (used nlp_formulation.cpp:334-346): 5 rows per non-constant force node.
struct Force : Constraint {
  int ee;
  std::vector<int> ids;
  Vec3 n, t1, t2;
  Force(Problem* p, int e) {
    P = p;
    ee = e;
    ids = P->ee_force[ee].non_constant_nodes();
    rows = 5 * (int)ids.size();
    name = "force-ee-force_" + std::to_string(ee);
    // towr HeightMap::GetNormalizedBasis for a plane (position independent)
    double hx = P->dhdx(), hy = P->dhdy();
    n = {-hx, -hy, 1.0};
    t1 = {1.0, 0.0, hx};
    t2 = {0.0, 1.0, hy};
    n = (1.0 / sqrt(dot(n, n))) * n;
    t1 = (1.0 / sqrt(dot(t1, t1))) * t1;
    t2 = (1.0 / sqrt(dot(t2, t2))) * t2;
  }
  void dirs(Vec3 d[5]) const {
    double mu = P->mu;
    d[0] = n;
    d[1] = t1 - mu * n;
    d[2] = t1 + mu * n;
    d[3] = t2 - mu * n;
    d[4] = t2 + mu * n;
  }
  void values(double* g) const override {
    Vec3 d[5];
    dirs(d);
    int row = 0;
    for (int id : ids) {
      const Vec3& f = P->ee_force[ee].nodes[id].p;
      for (int k = 0; k < 5; ++k) g[row++] = dot(f, d[k]);
    }
  }
  void bounds(double* lo, double* hi) const override {
    int row = 0;
    for (size_t i = 0; i < ids.size(); ++i) {
      lo[row] = 0.0, hi[row++] = P->force_limit;  // unilateral
      lo[row] = -1e20, hi[row++] = 0.0;           // t1 <  mu n
      lo[row] = 0.0, hi[row++] = 1e20;            // t1 > -mu n
      lo[row] = -1e20, hi[row++] = 0.0;
      lo[row] = 0.0, hi[row++] = 1e20;
    }
  }
  void row_times(double* t) const override {
    const auto& sp = P->s_force[ee];
    std::vector<double> nt(sp.nv->nodes.size(), 0.0);
    for (size_t i = 1; i < nt.size(); ++i) nt[i] = nt[i - 1] + sp.poly_dur[i - 1];
    int row = 0;
    for (int id : ids)
      for (int k = 0; k < 5; ++k) t[row++] = nt[id];
  }
  void jac(Row* r) const override {
    Vec3 d[5];
    dirs(d);
    const auto& nv = P->ee_force[ee];
    int row = 0;
    for (int id : ids)
      for (int k = 0; k < 5; ++k) {
        for (int dim = 0; dim < 3; ++dim) r[row].add(nv.rev[id][kPos * 3 + dim] + nv.offset, d[k][dim]);
        row++;
      }
  }
};

// humanoid_dynamic_constraint.cpp + humanoid_rigid_body_dynamics.cpp
struct Dynamic : Constraint {
  std::vector<double> dts;
  Dynamic(Problem* p) {
    P = p;
    dts = discretize(P->T, P->dt_dyn);
    rows = 6 * (int)dts.size();
    name = "dynamic";
  }
  struct Model {
    State com;
    Mat3 R, Ib, Iw;
    Vec3 w, wd;
    std::vector<Vec3> f, p;
  };
  Model update(double t) const {  // humanoid_dynamic_constraint.cpp:124-143
    Model M;
    M.com = P->s_lin.point(t);
    M.R = P->euler.rot(t);
    M.w = P->euler.ang_vel(t);
    M.wd = P->euler.ang_acc(t);
    for (int ee = 0; ee < P->n_ee; ++ee) {
      M.f.push_back(P->s_force[ee].point(t).p);
      M.p.push_back(P->s_motion[ee].point(t).p);
    }
    M.Ib = P->inertia_at(t);
    M.Iw = mul(mul(M.R, M.Ib), transpose(M.R));
    return M;
  }
  void values(double* g) const override {  // humanoid_rigid_body_dynamics.cpp:89-115
    for (size_t k = 0; k < dts.size(); ++k) {
      Model M = update(dts[k]);
      Vec3 fsum{0, 0, 0}, tau{0, 0, 0};
      for (int ee = 0; ee < P->n_ee; ++ee) {
        tau = tau + cross(M.f[ee], M.com.p - M.p[ee]);
        fsum = fsum + M.f[ee];
      }
      Vec3 Iww = mul(M.Iw, M.w);
      Vec3 ang = mul(M.Iw, M.wd) + cross(M.w, Iww) - tau;
      Vec3 lin = P->mass * M.com.a - fsum - (P->mass * P->g) * P->gvec();
      for (int d = 0; d < 3; ++d) {
        g[6 * k + d] = ang[d];
        g[6 * k + 3 + d] = lin[d];
      }
    }
  }
  void bounds(double* lo, double* hi) const override {
    for (int i = 0; i < rows; ++i) lo[i] = hi[i] = 0.0;
  }
  void row_times(double* t) const override {
    for (size_t k = 0; k < dts.size(); ++k)
      for (int d = 0; d < 6; ++d) t[6 * k + d] = dts[k];
  }
  static void put(Row* r, const SJac& J, int off) {  // 3 rows
    for (size_t c = 0; c < J.col.size(); ++c)
      for (int d = 0; d < 3; ++d) r[d].add(J.col[c] + off, J.val[c][d]);
  }
  void jac(Row* r) const override {
    for (size_t k = 0; k < dts.size(); ++k) {
      double t = dts[k];
      Model M = update(t);
      Row* ra = r + 6 * k;      // AX..AZ
      Row* rl = r + 6 * k + 3;  // LX..LZ
      // base linear (humanoid_rigid_body_dynamics.cpp:117-135)
      SJac Jp = P->s_lin.jac_wrt_nodes(t, kPos), Ja = P->s_lin.jac_wrt_nodes(t, kAcc);
      SJac tau_sum;
      for (const Vec3& f : M.f) tau_sum.axpy(1.0, Jp.lmul(skew(f)));
      SJac neg;
      neg.axpy(-1.0, tau_sum);
      put(ra, neg, P->base_lin.offset);
      SJac ml;
      ml.axpy(P->mass, Ja);
      put(rl, ml, P->base_lin.offset);
      // base angular (humanoid_rigid_body_dynamics.cpp:137-179)
      {
        Mat3 Rt = transpose(M.R), RIb = mul(M.R, M.Ib);
        Vec3 v11 = mul(M.Ib, mul(Rt, M.wd));
        SJac j1 = P->euler.deriv_rot_vec_mult(t, v11, false);
        j1.axpy(1.0, P->euler.deriv_rot_vec_mult(t, M.wd, true).lmul(RIb));
        j1.axpy(1.0, P->euler.deriv_ang_acc(t).lmul(M.Iw));
        Vec3 v21 = mul(M.Ib, mul(Rt, M.w));
        SJac jw = P->euler.deriv_ang_vel(t);
        SJac j2 = P->euler.deriv_rot_vec_mult(t, v21, false);
        j2.axpy(1.0, P->euler.deriv_rot_vec_mult(t, M.w, true).lmul(RIb));
        j2.axpy(1.0, jw.lmul(M.Iw));
        SJac j = j1;
        j.axpy(1.0, j2.lmul(skew(M.w)));
        j.axpy(-1.0, jw.lmul(skew(mul(M.Iw, M.w))));
        put(ra, j, P->base_ang.offset);
      }
      for (int ee = 0; ee < P->n_ee; ++ee) {
        Vec3 rr = M.com.p - M.p[ee];
        // force nodes (humanoid_rigid_body_dynamics.cpp:181-193)
        SJac Jf = P->s_force[ee].jac_wrt_nodes(t, kPos);
        put(ra, Jf.lmul(skew(rr)), P->ee_force[ee].offset);
        SJac nf;
        nf.axpy(-1.0, Jf);
        put(rl, nf, P->ee_force[ee].offset);
        // ee position nodes (humanoid_rigid_body_dynamics.cpp:195-206)
        SJac Jx = P->s_motion[ee].jac_wrt_nodes(t, kPos);
        put(ra, Jx.lmul(skew(M.f[ee])), P->ee_motion[ee].offset);
        if (P->opt_dur) {  // humanoid_dynamic_constraint.cpp:112-118
          SJac JfT = P->s_force[ee].jac_pos_wrt_durations(t);
          put(ra, JfT.lmul(skew(rr)), P->dur[ee].offset);
          SJac nfT;
          nfT.axpy(-1.0, JfT);
          put(rl, nfT, P->dur[ee].offset);
          SJac JxT = P->s_motion[ee].jac_pos_wrt_durations(t);
          put(ra, JxT.lmul(skew(M.f[ee])), P->dur[ee].offset);
        }
      }
    }
  }
  // second derivative of  -y_ang . sum_ee f x (c - p)  (bilinear, exact)
  void lag_hessian(const double* y, std::vector<std::array<double, 3>>& trip) const override {
    for (size_t k = 0; k < dts.size(); ++k) {
      double t = dts[k];
      Vec3 ya{y[6 * k], y[6 * k + 1], y[6 * k + 2]};
      if (ya[0] == 0.0 && ya[1] == 0.0 && ya[2] == 0.0) continue;
      SJac Jc = P->s_lin.jac_wrt_nodes(t, kPos);
      for (int ee = 0; ee < P->n_ee; ++ee) {
        SJac Jf = P->s_force[ee].jac_wrt_nodes(t, kPos), Jp = P->s_motion[ee].jac_wrt_nodes(t, kPos);
        for (size_t i = 0; i < Jf.col.size(); ++i) {
          double uf = Jf.col[i] + P->ee_force[ee].offset;
          for (size_t j = 0; j < Jc.col.size(); ++j) {
            double v = -dot(ya, cross(Jf.val[i], Jc.val[j]));
            if (v == 0.0) continue;
            double uc = Jc.col[j] + P->base_lin.offset;
            trip.push_back({uf, uc, v});
            trip.push_back({uc, uf, v});
          }
          for (size_t j = 0; j < Jp.col.size(); ++j) {
            double v = dot(ya, cross(Jf.val[i], Jp.val[j]));
            if (v == 0.0) continue;
            double up = Jp.col[j] + P->ee_motion[ee].offset;
            trip.push_back({uf, up, v});
            trip.push_back({up, uf, v});
          }
        }
      }
    }
  }
};

// leg_length_constraint.cpp
struct LegLength : Constraint {
  int ee;
  std::vector<double> dts;
  double max_len;
  LegLength(Problem* p, int e) {
    P = p;
    ee = e;
    dts = discretize(P->T, P->dt_rom);
    rows = (int)dts.size();
    max_len = (ee == 0 || ee == 1) ? P->max_leg : P->max_heel;  // :21-27
    name = "leg-length-" + std::to_string(ee);
  }
  Vec3 hip_to_ee(double t, Vec3* hip = nullptr) const {  // :37-60
    Vec3 h = P->hip_at(ee, t);
    if (hip) *hip = h;
    Vec3 base = P->s_lin.point(t).p, pe = P->s_motion[ee].point(t).p;
    return pe - (mul(P->euler.rot(t), h) + base);
  }
  void values(double* g) const override {
    for (size_t k = 0; k < dts.size(); ++k) {
      Vec3 d = hip_to_ee(dts[k]);
      g[k] = 0.5 * dot(d, d);
    }
  }
  void row_times(double* t) const override {
    for (size_t k = 0; k < dts.size(); ++k) t[k] = dts[k];
  }
  void bounds(double* lo, double* hi) const override {
    for (int k = 0; k < rows; ++k) lo[k] = 0.0, hi[k] = 0.5 * max_len * max_len;
  }
  void jac(Row* r) const override {  // :62-111
    for (size_t k = 0; k < dts.size(); ++k) {
      double t = dts[k];
      Vec3 h, d = hip_to_ee(t, &h);
      r[k].add_dT_J(P->s_lin.jac_wrt_nodes(t, kPos), P->base_lin.offset, d, -1.0);
      r[k].add_dT_J(P->euler.deriv_rot_vec_mult(t, h, false), P->base_ang.offset, d, -1.0);
      r[k].add_dT_J(P->s_motion[ee].jac_wrt_nodes(t, kPos), P->ee_motion[ee].offset, d, 1.0);
      if (P->opt_dur) r[k].add_dT_J(P->s_motion[ee].jac_pos_wrt_durations(t), P->dur[ee].offset, d, 1.0);
    }
  }
  void lag_hessian(const double* y, std::vector<std::array<double, 3>>& trip) const override {
    for (size_t k = 0; k < dts.size(); ++k) {
      if (y[k] == 0.0) continue;
      double t = dts[k];
      Vec3 h = P->hip_at(ee, t);
      std::vector<std::pair<int, Vec3>> cols;
      gather(cols, P->s_lin.jac_wrt_nodes(t, kPos), P->base_lin.offset, -1.0);
      gather(cols, P->euler.deriv_rot_vec_mult(t, h, false), P->base_ang.offset, -1.0);
      gather(cols, P->s_motion[ee].jac_wrt_nodes(t, kPos), P->ee_motion[ee].offset, 1.0);
      if (P->opt_dur) gather(cols, P->s_motion[ee].jac_pos_wrt_durations(t), P->dur[ee].offset, 1.0);
      add_outer(trip, cols, y[k]);
    }
  }
};

// ee_dist_constraint.cpp
struct EEDist : Constraint {
  int e1, e2;
  std::vector<double> dts;
  EEDist(Problem* p, int a, int b) {
    P = p;
    e1 = a;
    e2 = b;
    dts = discretize(P->T, P->dt_rom);
    rows = (int)dts.size();
    name = "ee-dist-" + std::to_string(a) + "-" + std::to_string(b);
  }
  void values(double* g) const override {
    for (size_t k = 0; k < dts.size(); ++k) {
      Vec3 d = P->s_motion[e1].point(dts[k]).p - P->s_motion[e2].point(dts[k]).p;
      g[k] = 0.5 * dot(d, d);
    }
  }
  void row_times(double* t) const override {
    for (size_t k = 0; k < dts.size(); ++k) t[k] = dts[k];
  }
  void bounds(double* lo, double* hi) const override {
    for (int k = 0; k < rows; ++k) lo[k] = hi[k] = 0.5 * P->heel_dist * P->heel_dist;
  }
  void jac(Row* r) const override {
    for (size_t k = 0; k < dts.size(); ++k) {
      double t = dts[k];
      Vec3 d = P->s_motion[e1].point(t).p - P->s_motion[e2].point(t).p;
      r[k].add_dT_J(P->s_motion[e1].jac_wrt_nodes(t, kPos), P->ee_motion[e1].offset, d, 1.0);
      r[k].add_dT_J(P->s_motion[e2].jac_wrt_nodes(t, kPos), P->ee_motion[e2].offset, d, -1.0);
      if (P->opt_dur) {
        r[k].add_dT_J(P->s_motion[e1].jac_pos_wrt_durations(t), P->dur[e1].offset, d, 1.0);
        r[k].add_dT_J(P->s_motion[e2].jac_pos_wrt_durations(t), P->dur[e2].offset, d, -1.0);
      }
    }
  }
  void lag_hessian(const double* y, std::vector<std::array<double, 3>>& trip) const override {
    for (size_t k = 0; k < dts.size(); ++k) {
      if (y[k] == 0.0) continue;
      double t = dts[k];
      std::vector<std::pair<int, Vec3>> cols;
      gather(cols, P->s_motion[e1].jac_wrt_nodes(t, kPos), P->ee_motion[e1].offset, 1.0);
      gather(cols, P->s_motion[e2].jac_wrt_nodes(t, kPos), P->ee_motion[e2].offset, -1.0);
      if (P->opt_dur) {
        gather(cols, P->s_motion[e1].jac_pos_wrt_durations(t), P->dur[e1].offset, 1.0);
        gather(cols, P->s_motion[e2].jac_pos_wrt_durations(t), P->dur[e2].offset, -1.0);
      }
      add_outer(trip, cols, y[k]);
    }
  }
};

// height_constraint.cpp
struct Height : Constraint {
  int ee;
  std::vector<double> dts;
  Height(Problem* p, int e) {
    P = p;
    ee = e;
    dts = discretize(P->T, P->dt_height);
    rows = (int)dts.size();
    name = "height-terrain-ee-" + std::to_string(e);
  }
  void values(double* g) const override {
    for (size_t k = 0; k < dts.size(); ++k) g[k] = dot(P->normal, P->s_motion[ee].point(dts[k]).p - P->point);
  }
  void row_times(double* t) const override {
    for (size_t k = 0; k < dts.size(); ++k) t[k] = dts[k];
  }
  void bounds(double* lo, double* hi) const override {
    for (int k = 0; k < rows; ++k) lo[k] = 0.0, hi[k] = 1e20;
  }
  void jac(Row* r) const override {
    for (size_t k = 0; k < dts.size(); ++k) {
      r[k].add_dT_J(P->s_motion[ee].jac_wrt_nodes(dts[k], kPos), P->ee_motion[ee].offset, P->normal, 1.0);
      if (P->opt_dur) r[k].add_dT_J(P->s_motion[ee].jac_pos_wrt_durations(dts[k]), P->dur[ee].offset, P->normal, 1.0);
    }
  }
};

// total_duration_constraint.cpp
struct TotalDuration : Constraint {
  int ee;
  TotalDuration(Problem* p, int e) {
    P = p;
    ee = e;
    rows = 1;
    name = "contactduration-" + std::to_string(e);
  }
  void values(double* g) const override {
    double s = 0;
    for (int i = 0; i < P->dur[ee].rows(); ++i) s += P->dur[ee].durations[i];
    g[0] = s;
  }
  void row_times(double* t) const override { t[0] = P->T; }
  void bounds(double* lo, double* hi) const override {
    lo[0] = std::max(0.0, P->T - P->dur_hi);
    hi[0] = P->T - P->dur_lo;
  }
  void jac(Row* r) const override {
    for (int i = 0; i < P->dur[ee].rows(); ++i) r[0].add(P->dur[ee].offset + i, 1.0);
  }
};

// ============================================================ costs =============================
// data_cost.cpp
struct DataCost : Cost {
  const Spline* s;
  const NodesVars* nv;
  const std::vector<Vec3>* data;
  double w;
  int ee;  // -1 for base splines
  DataCost(Problem* p, const Spline* sp, const NodesVars* v, const std::vector<Vec3>* d, double weight, int ee_id) {
    P = p, s = sp, nv = v, data = d, w = weight, ee = ee_id;
    name = v->name + "-data";
  }
  double value() const override {  // :40-54
    double cost = 0, t = 0;
    for (size_t i = 0; i < data->size(); ++i) {
      Vec3 diff = (*data)[i] - s->point(t).p;
      cost += dot(diff, diff);
      t += 1 * P->dt;
    }
    return 0.5 * w * cost;
  }
  void grad(double* g) const override {  // :56-96
    double t = 0;
    for (size_t i = 0; i < data->size(); ++i) {
      Vec3 diff = (*data)[i] - s->point(t).p;
      SJac J = s->jac_wrt_nodes(t, kPos);
      for (size_t c = 0; c < J.col.size(); ++c) g[J.col[c] + nv->offset] += -1.0 * dot(J.val[c], diff) * w;
      if (ee >= 0 && P->opt_dur) {
        SJac JT = s->jac_pos_wrt_durations(t);
        for (size_t c = 0; c < JT.col.size(); ++c) g[JT.col[c] + P->dur[ee].offset] += -1.0 * dot(JT.val[c], diff) * w;
      }
      t += P->dt * 1;
    }
  }
  void gn_hessian(std::vector<std::array<double, 3>>& trip) const override {
    double t = 0;
    for (size_t i = 0; i < data->size(); ++i) {
      std::vector<std::pair<int, Vec3>> cols;
      gather(cols, s->jac_wrt_nodes(t, kPos), nv->offset, 1.0);
      if (ee >= 0 && P->opt_dur) gather(cols, s->jac_pos_wrt_durations(t), P->dur[ee].offset, 1.0);
      add_outer(trip, cols, w);
      t += P->dt * 1;
    }
  }
};

// vel_smooth_cost.cpp (deriv = kPos: "velocity smoothing", kVel: "acceleration smoothing")
struct SmoothCost : Cost {
  const Spline* s;
  const NodesVars* nv;
  int deriv;
  double w;
  int ee;
  SmoothCost(Problem* p, const Spline* sp, const NodesVars* v, int d, double weight, int ee_id) {
    P = p, s = sp, nv = v, deriv = d, w = weight, ee = ee_id;
    name = v->name + "-deriv" + std::to_string(d + 1) + "-smooth";
  }
  double value() const override {  // :37-50
    double cost = 0;
    double dt = P->dt;
    for (double t = 0.0; t < (s->total_time() - dt); t += dt) {
      Vec3 diff = s->point(t + dt).at(deriv) - s->point(t).at(deriv);
      cost += dot(diff, diff);
    }
    return 0.5 * w * cost;
  }
  void grad(double* g) const override {  // :52-100
    double dt = P->dt;
    for (double t = 0.0; t < (s->total_time() - dt); t += dt) {
      Vec3 diff = s->point(t + dt).at(deriv) - s->point(t).at(deriv);
      SJac J = s->jac_wrt_nodes(t + dt, deriv);
      J.axpy(-1.0, s->jac_wrt_nodes(t, deriv));
      for (size_t c = 0; c < J.col.size(); ++c) g[J.col[c] + nv->offset] += dot(J.val[c], diff) * w;
      if (ee >= 0 && P->opt_dur && deriv == kPos) {
        SJac JT = s->jac_pos_wrt_durations(t + dt);
        JT.axpy(-1.0, s->jac_pos_wrt_durations(t));
        for (size_t c = 0; c < JT.col.size(); ++c) g[JT.col[c] + P->dur[ee].offset] += dot(JT.val[c], diff) * w;
      }
      // deriv == kVel together with duration optimisation throws in the reference (:72-79);
      // set_stage never builds that combination.
    }
  }
  void gn_hessian(std::vector<std::array<double, 3>>& trip) const override {
    double dt = P->dt;
    for (double t = 0.0; t < (s->total_time() - dt); t += dt) {
      std::vector<std::pair<int, Vec3>> cols;
      gather(cols, s->jac_wrt_nodes(t + dt, deriv), nv->offset, 1.0);
      gather(cols, s->jac_wrt_nodes(t, deriv), nv->offset, -1.0);
      if (ee >= 0 && P->opt_dur && deriv == kPos) {
        gather(cols, s->jac_pos_wrt_durations(t + dt), P->dur[ee].offset, 1.0);
        gather(cols, s->jac_pos_wrt_durations(t), P->dur[ee].offset, -1.0);
      }
      add_outer(trip, cols, w);
    }
  }
};

// duration_cost.cpp
struct DurationCost : Cost {
  int ee;
  double w;
  DurationCost(Problem* p, int e, double weight) {
    P = p, ee = e, w = weight;
    name = "ee-schedule_" + std::to_string(e) + "-duration";
  }
  double value() const override {
    double c = 0;
    const auto& d = P->dur[ee].durations;
    for (size_t i = 0; i + 1 < d.size(); ++i) c += (P->init_dur[ee][i] - d[i]) * (P->init_dur[ee][i] - d[i]);
    return 0.5 * w * c;
  }
  void grad(double* g) const override {
    const auto& d = P->dur[ee].durations;
    for (size_t i = 0; i + 1 < d.size(); ++i) g[P->dur[ee].offset + i] += w * (-(P->init_dur[ee][i] - d[i]));
  }
  void gn_hessian(std::vector<std::array<double, 3>>& trip) const override {
    for (int i = 0; i < P->dur[ee].rows(); ++i) trip.push_back({(double)(P->dur[ee].offset + i), (double)(P->dur[ee].offset + i), w});
  }
};

// ============================================================ staged schedule ====================
// phys_optim.cpp:554-749 / SURVEY Appendix B.  stage ids: 0=1.1, 1=1.2, 2=2.1, 3=2.2, 4=3, 5=4.
void Problem::set_stage(int st) {
  stage = st;
  opt_dur = (st == 4);
  cons.clear();
  costs.clear();
  auto base_acc = [&] {
    cons.emplace_back(new SplineAcc(this, &s_lin, &base_lin));
    cons.emplace_back(new SplineAcc(this, &s_ang, &base_ang));
  };
  auto leg = [&] {
    for (int ee = 0; ee < n_ee; ++ee) cons.emplace_back(new Terrain(this, ee));
    for (int ee = 0; ee < n_ee; ++ee) cons.emplace_back(new LegLength(this, ee));
  };
  auto heel = [&] {
    if (n_ee >= 4) {  // nlp_formulation.cpp:243-262 pairs (0,2),(1,3)
      cons.emplace_back(new EEDist(this, 0, 2));
      cons.emplace_back(new EEDist(this, 1, 3));
    }
  };
  auto dynamics = [&] {
    cons.emplace_back(new Dynamic(this));
    for (int ee = 0; ee < n_ee; ++ee) cons.emplace_back(new Force(this, ee));
  };
  auto height = [&] {
    for (int ee = 0; ee < n_ee; ++ee) cons.emplace_back(new Height(this, ee));
  };
  auto data_costs = [&](double wl, double wa, double we) {  // phys_optim.cpp:314-333
    costs.emplace_back(new DataCost(this, &s_lin, &base_lin, &lin_data, wl, -1));
    costs.emplace_back(new DataCost(this, &s_ang, &base_ang, &ang_data, wa, -1));
    for (int ee = 0; ee < n_ee; ++ee) costs.emplace_back(new DataCost(this, &s_motion[ee], &ee_motion[ee], &ee_data[ee], we, ee));
  };
  auto smooth_costs = [&](int deriv, double wl, double wa, double we) {  // :335-373
    costs.emplace_back(new SmoothCost(this, &s_lin, &base_lin, deriv, wl, -1));
    costs.emplace_back(new SmoothCost(this, &s_ang, &base_ang, deriv, wa, -1));
    for (int ee = 0; ee < n_ee; ++ee) costs.emplace_back(new SmoothCost(this, &s_motion[ee], &ee_motion[ee], deriv, we, ee));
  };
  switch (st) {
    case 0:
      base_acc();
      break;
    case 1:
      base_acc(), leg(), heel();
      break;
    case 2:
      base_acc(), leg(), dynamics(), heel();
      break;
    case 3:
      base_acc(), leg(), dynamics(), heel(), height();
      break;
    case 4:
      base_acc(), leg(), dynamics(), height(), heel();
      for (int ee = 0; ee < n_ee; ++ee) cons.emplace_back(new TotalDuration(this, ee));
      break;
    case 5:
      base_acc(), leg(), dynamics(), height(), heel();
      break;
  }
  if (st <= 1) {
    data_costs(1.0, 1.0, 1.0);            // :560-562
    smooth_costs(kPos, 0.1, 0.1, 0.1);    // :564
  } else {
    data_costs(w_com_lin, w_com_ang, w_ee);      // :631-633
    smooth_costs(kPos, 0.001, 0.001, w_smooth);  // :635
    if (st != 4) smooth_costs(kVel, 0.0001, 0.0001, 0.0001);  // :637, none in stage 3 (:693)
    if (st == 4)
      for (int ee = 0; ee < n_ee; ++ee) costs.emplace_back(new DurationCost(this, ee, w_dur));  // :696-703
  }
  n = base_lin.rows() + base_ang.rows();
  for (auto& v : ee_motion) n += v.rows();
  for (auto& v : ee_force) n += v.rows();
  if (opt_dur)
    for (auto& d : dur) n += d.rows();
  m = 0;
  for (auto& c : cons) m += c->rows;
}

}  // namespace oracle

// ================================================================ C API ==========================
using namespace oracle;
extern "C" {

void* chdo_create(int F, int n_ee, double dt, const double* hip_l, const double* hip_r, double max_leg, double max_heel,
                  double heel_dist, double mass, const double* inertia, const double* lin, const double* ang,
                  const double* ee_data, const double* normal, const double* point, const int* start_contact,
                  const int* n_phases, const double* durations, const double* weights /*5*/) {
  Problem* P = new Problem();
  P->F = F, P->n_ee = n_ee, P->dt = dt;
  auto v3 = [](const double* p, int i) { return Vec3{p[3 * i], p[3 * i + 1], p[3 * i + 2]}; };
  for (int i = 0; i < F; ++i) {
    P->hip_l.push_back(v3(hip_l, i));
    P->hip_r.push_back(v3(hip_r, i));
    P->inertia.push_back({inertia[6 * i], inertia[6 * i + 1], inertia[6 * i + 2], inertia[6 * i + 3], inertia[6 * i + 4], inertia[6 * i + 5]});
    P->lin_data.push_back(v3(lin, i));
    P->ang_data.push_back(v3(ang, i));
  }
  P->ee_data.resize(n_ee);
  for (int ee = 0; ee < n_ee; ++ee)
    for (int i = 0; i < F; ++i) P->ee_data[ee].push_back(v3(ee_data + (size_t)ee * F * 3, i));
  P->max_leg = max_leg, P->max_heel = max_heel, P->heel_dist = heel_dist, P->mass = mass;
  P->normal = {normal[0], normal[1], normal[2]};
  P->point = {point[0], point[1], point[2]};
  int off = 0;
  for (int ee = 0; ee < n_ee; ++ee) {
    P->start_contact.push_back(start_contact[ee]);
    P->init_dur.emplace_back(durations + off, durations + off + n_phases[ee]);
    off += n_phases[ee];
  }
  P->w_com_lin = weights[0], P->w_com_ang = weights[1], P->w_ee = weights[2], P->w_smooth = weights[3], P->w_dur = weights[4];
  P->build_variables();
  P->set_stage(0);
  return P;
}
void chdo_destroy(void* h) { delete (Problem*)h; }
void chdo_set_stage(void* h, int st) { ((Problem*)h)->set_stage(st); }
int chdo_n(void* h) { return ((Problem*)h)->n; }
int chdo_m(void* h) { return ((Problem*)h)->m; }
double chdo_total_time(void* h) { return ((Problem*)h)->T; }
void chdo_get_x(void* h, double* x) { ((Problem*)h)->get_x(x); }
void chdo_set_x(void* h, const double* x) { ((Problem*)h)->set_x(x); }
// sizes of the variable sets in stacking order: base_lin, base_ang, motion[n_ee], force[n_ee], dur[n_ee]
void chdo_var_set_sizes(void* h, int* out) {
  Problem* P = (Problem*)h;
  int k = 0;
  out[k++] = P->base_lin.rows();
  out[k++] = P->base_ang.rows();
  for (auto& v : P->ee_motion) out[k++] = v.rows();
  for (auto& v : P->ee_force) out[k++] = v.rows();
  for (auto& d : P->dur) out[k++] = d.rows();
}
void chdo_var_bounds(void* h, double* lo, double* hi) {
  Problem* P = (Problem*)h;
  for (int i = 0; i < P->n; ++i) lo[i] = -1e20, hi[i] = 1e20;
  for (int i = 0; i < P->base_lin.rows(); ++i) lo[P->base_lin.offset + i] = P->base_lin.lo[i], hi[P->base_lin.offset + i] = P->base_lin.hi[i];
  if (P->opt_dur)
    for (auto& d : P->dur)
      for (int i = 0; i < d.rows(); ++i) lo[d.offset + i] = d.lo, hi[d.offset + i] = d.hi;
}
// duration variable blocks of the current stage (stage 3 only): offsets and sizes of the PhaseDurations sets in x;
// returns the number of blocks (0 when the durations are not optimised)
int chdo_dur_blocks(void* h, int* off, int* cnt) {
  Problem* P = (Problem*)h;
  if (!P->opt_dur) return 0;
  int k = 0;
  for (auto& d : P->dur) {
    if (off) off[k] = d.offset;
    if (cnt) cnt[k] = d.rows();
    ++k;
  }
  return k;
}
int chdo_n_ee(void* h) { return ((Problem*)h)->n_ee; }
void chdo_init_durations(void* h, double* d) {
  Problem* P = (Problem*)h;
  int k = 0;
  for (auto& v : P->init_dur)
    for (size_t i = 0; i + 1 < v.size(); ++i) d[k++] = v[i];
}
int chdo_num_constraint_sets(void* h) { return (int)((Problem*)h)->cons.size(); }
int chdo_constraint_set_rows(void* h, int i) { return ((Problem*)h)->cons[i]->rows; }
const char* chdo_constraint_set_name(void* h, int i) { return ((Problem*)h)->cons[i]->name.c_str(); }
void chdo_con_bounds(void* h, double* lo, double* hi) {
  Problem* P = (Problem*)h;
  int r = 0;
  for (auto& c : P->cons) {
    c->bounds(lo + r, hi + r);
    r += c->rows;
  }
}
double chdo_cost(void* h) {
  double f = 0;
  for (auto& c : ((Problem*)h)->costs) f += c->value();
  return f;
}
int chdo_num_costs(void* h) { return (int)((Problem*)h)->costs.size(); }
double chdo_cost_term(void* h, int i) { return ((Problem*)h)->costs[i]->value(); }
void chdo_grad(void* h, double* g) {
  Problem* P = (Problem*)h;
  std::fill(g, g + P->n, 0.0);
  for (auto& c : P->costs) c->grad(g);
}
void chdo_cons(void* h, double* g) {
  Problem* P = (Problem*)h;
  int r = 0;
  for (auto& c : P->cons) {
    c->values(g + r);
    r += c->rows;
  }
}
// Sparse Jacobian as sorted, duplicate-summed triplets.  Call with vals == nullptr to get nnz.
static std::vector<std::vector<std::pair<int, double>>> jac_rows(Problem* P) {
  std::vector<Row> rows(P->m);
  int r = 0;
  for (auto& c : P->cons) {
    c->jac(rows.data() + r);
    r += c->rows;
  }
  std::vector<std::vector<std::pair<int, double>>> out(P->m);
  for (int i = 0; i < P->m; ++i) {
    auto& e = rows[i].e;
    std::stable_sort(e.begin(), e.end(), [](auto& a, auto& b) { return a.first < b.first; });
    for (auto& kv : e) {
      if (!out[i].empty() && out[i].back().first == kv.first)
        out[i].back().second += kv.second;
      else
        out[i].push_back(kv);
    }
  }
  return out;
}
int chdo_jac(void* h, int* ri, int* ci, double* vals) {
  Problem* P = (Problem*)h;
  auto rows = jac_rows(P);
  int nnz = 0;
  for (int i = 0; i < P->m; ++i)
    for (auto& kv : rows[i]) {
      if (vals) ri[nnz] = i, ci[nnz] = kv.first, vals[nnz] = kv.second;
      nnz++;
    }
  return nnz;
}
// Gauss-Newton Hessian of the cost (exact when durations are fixed: every cost is a convex quadratic
// in the node variables).  Sorted duplicate-summed full-symmetric triplets; vals == nullptr -> nnz.
int chdo_cost_hessian(void* h, int* ri, int* ci, double* vals) {
  Problem* P = (Problem*)h;
  std::vector<std::array<double, 3>> trip;
  for (auto& c : P->costs) c->gn_hessian(trip);
  std::sort(trip.begin(), trip.end(), [](auto& a, auto& b) { return a[0] != b[0] ? a[0] < b[0] : a[1] < b[1]; });
  int nnz = 0;
  int pi = -1, pj = -1;
  for (auto& t : trip) {
    int i = (int)t[0], j = (int)t[1];
    if (i == pi && j == pj) {
      if (vals) vals[nnz - 1] += t[2];
    } else {
      if (vals) ri[nnz] = i, ci[nnz] = j, vals[nnz] = t[2];
      nnz++;
      pi = i, pj = j;
    }
  }
  return nnz;
}
// Constraint-curvature model sum_i y_i * H_i (see Constraint::lag_hessian).  y has m entries (unscaled rows).
static int sum_triplets(std::vector<std::array<double, 3>>& trip, int* ri, int* ci, double* vals) {
  std::sort(trip.begin(), trip.end(), [](auto& a, auto& b) { return a[0] != b[0] ? a[0] < b[0] : a[1] < b[1]; });
  int nnz = 0, pi = -1, pj = -1;
  for (auto& t : trip) {
    int i = (int)t[0], j = (int)t[1];
    if (i == pi && j == pj) {
      if (vals) vals[nnz - 1] += t[2];
    } else {
      if (vals) ri[nnz] = i, ci[nnz] = j, vals[nnz] = t[2];
      nnz++;
      pi = i, pj = j;
    }
  }
  return nnz;
}
int chdo_lag_hessian(void* h, const double* y, int* ri, int* ci, double* vals) {
  Problem* P = (Problem*)h;
  std::vector<std::array<double, 3>> trip;
  int r = 0;
  for (auto& c : P->cons) {
    c->lag_hessian(y + r, trip);
    r += c->rows;
  }
  return sum_triplets(trip, ri, ci, vals);
}
void chdo_row_times(void* h, double* t) {
  Problem* P = (Problem*)h;
  int r = 0;
  for (auto& c : P->cons) {
    c->row_times(t + r);
    r += c->rows;
  }
}
// per variable: [t_first, t_last] of the nodes it parameterises (durations: [0, T])
void chdo_var_times(void* h, double* t0, double* t1) {
  Problem* P = (Problem*)h;
  auto fill = [&](const NodesVars& nv, const Spline& sp) {
    std::vector<double> nt(nv.nodes.size(), 0.0);
    for (size_t i = 1; i < nt.size(); ++i) nt[i] = nt[i - 1] + sp.poly_dur[i - 1];
    for (int idx = 0; idx < nv.rows(); ++idx) {
      double a = 1e300, b = -1e300;
      for (auto& nvi : nv.index_map[idx]) a = std::min(a, nt[nvi.node]), b = std::max(b, nt[nvi.node]);
      t0[nv.offset + idx] = a, t1[nv.offset + idx] = b;
    }
  };
  fill(P->base_lin, P->s_lin);
  fill(P->base_ang, P->s_ang);
  for (int ee = 0; ee < P->n_ee; ++ee) fill(P->ee_motion[ee], P->s_motion[ee]), fill(P->ee_force[ee], P->s_force[ee]);
  if (P->opt_dur)
    for (auto& d : P->dur)
      for (int i = 0; i < d.rows(); ++i) t0[d.offset + i] = 0, t1[d.offset + i] = P->T;
}
// SaveSolution (phys_optim.cpp:63-143): sample all splines at t = 0, dt, ... while t <= T + 1e-5.
// out layout per frame: base_lin(3) base_ang_deg(3) ee_pos(3*n_ee) ee_force(3*n_ee) contact(n_ee).
int chdo_sample(void* h, double* out) {
  Problem* P = (Problem*)h;
  double tot = P->s_lin.total_time();
  int nf = (int)((tot + 1e-5) / P->dt) + 1;
  if (!out) return nf;
  int stride = 6 + 7 * P->n_ee;
  double t = 0.0;
  int i = 0;
  while (t <= tot + 1e-5 && i < nf) {
    double* o = out + (size_t)i * stride;
    Vec3 p = P->s_lin.point(t).p, a = P->s_ang.point(t).p;
    for (int d = 0; d < 3; ++d) o[d] = p[d], o[3 + d] = a[d] / M_PI * 180;
    for (int ee = 0; ee < P->n_ee; ++ee) {
      Vec3 q = P->s_motion[ee].point(t).p, f = P->s_force[ee].point(t).p;
      for (int d = 0; d < 3; ++d) o[6 + 3 * ee + d] = q[d], o[6 + 3 * P->n_ee + 3 * ee + d] = f[d];
      o[6 + 6 * P->n_ee + ee] = P->dur[ee].is_contact_phase(t) ? 1.0 : 0.0;
    }
    t += P->dt;
    i++;
  }
  return i;
}
// ---- function-level probes used by the known-answer tests ----
// spline id: 0 base_lin, 1 base_ang, 2+ee motion, 2+n_ee+ee force.  out = p,v,a (9 doubles).
static const Spline* pick(Problem* P, int sid) {
  if (sid == 0) return &P->s_lin;
  if (sid == 1) return &P->s_ang;
  if (sid < 2 + P->n_ee) return &P->s_motion[sid - 2];
  return &P->s_force[sid - 2 - P->n_ee];
}
void chdo_spline_point(void* h, int sid, double t, double* out) {
  State s = pick((Problem*)h, sid)->point(t);
  for (int d = 0; d < 3; ++d) out[d] = s.p[d], out[3 + d] = s.v[d], out[6 + d] = s.a[d];
}
int chdo_spline_num_polys(void* h, int sid) { return (int)pick((Problem*)h, sid)->poly_dur.size(); }
void chdo_spline_poly_durations(void* h, int sid, double* out) {
  auto& d = pick((Problem*)h, sid)->poly_dur;
  std::copy(d.begin(), d.end(), out);
}
// dense 3 x rows Jacobian of the spline value (deriv) wrt its own variable set
void chdo_spline_jac(void* h, int sid, double t, int deriv, double* out /*3*rows*/) {
  const Spline* s = pick((Problem*)h, sid);
  int rows = s->nv->rows();
  std::fill(out, out + 3 * rows, 0.0);
  SJac J = s->jac_wrt_nodes(t, deriv);
  for (size_t c = 0; c < J.col.size(); ++c)
    for (int d = 0; d < 3; ++d) out[d * rows + J.col[c]] = J.val[c][d];
}
void chdo_euler(void* h, double t, double* R9, double* w3, double* wd3) {
  Problem* P = (Problem*)h;
  Mat3 R = P->euler.rot(t);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R9[3 * i + j] = R.m[i][j];
  Vec3 w = P->euler.ang_vel(t), wd = P->euler.ang_acc(t);
  for (int d = 0; d < 3; ++d) w3[d] = w[d], wd3[d] = wd[d];
}
}  // extern "C"
