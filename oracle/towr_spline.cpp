// TEST INFRASTRUCTURE ONLY -- see towr_oracle.hpp.  Spline algebra, variable sets, Euler kinematics.
#include "towr_oracle.hpp"

#include <cassert>
#include <numeric>

namespace oracle {

// ------------------------------------------------------------------ NodesVars ------------------
void NodesVars::finalize() {
  rev.assign(nodes.size(), {-1, -1, -1, -1, -1, -1});
  for (int idx = 0; idx < rows(); ++idx)
    for (const auto& nvi : index_map[idx]) rev[nvi.node][nvi.deriv * 3 + nvi.dim] = idx;
  lo.assign(rows(), -1e20);
  hi.assign(rows(), 1e20);
}

// towr NodesVariables::GetValues -- when several node values share one optimisation index the
// LAST NodeValueInfo wins (stance phases: the phase's end node).
void NodesVars::get_values(double* x) const {
  for (int idx = 0; idx < rows(); ++idx)
    for (const auto& nvi : index_map[idx]) x[idx] = nodes[nvi.node].at(nvi.deriv, nvi.dim);
}
void NodesVars::set_values(const double* x) {
  for (int idx = 0; idx < rows(); ++idx)
    for (const auto& nvi : index_map[idx]) nodes[nvi.node].at(nvi.deriv, nvi.dim) = x[idx];
}

// towr NodesVariables::SetByLinearInterpolation (used nlp_formulation.cpp:120,126,156,181):
// optimised positions lie on the line a->b at fraction node/(N-1); optimised velocities = (b-a)/T.
void NodesVars::set_by_linear_interpolation(const Vec3& a, const Vec3& b, double T) {
  Vec3 dp = b - a;
  Vec3 avg = (1.0 / T) * dp;
  int num_nodes = (int)nodes.size();
  for (int idx = 0; idx < rows(); ++idx)
    for (const auto& nvi : index_map[idx]) {
      if (nvi.deriv == kPos) {
        double frac = nvi.node / static_cast<double>(num_nodes - 1);
        nodes[nvi.node].p[nvi.dim] = a[nvi.dim] + frac * dp[nvi.dim];
      } else {
        nodes[nvi.node].v[nvi.dim] = avg[nvi.dim];
      }
    }
}

// towr NodesVariables::AddStartBound / AddFinalBound (nlp_formulation.cpp:121-122): an equality
// bound; IPOPT's default fixed_variable_treatment turns it into a constant held at the bound.
void NodesVars::add_bound(int node, int deriv, const Vec3& val) {
  for (int dim = 0; dim < 3; ++dim) {
    int idx = rev[node][deriv * 3 + dim];
    if (idx < 0) continue;
    lo[idx] = hi[idx] = val[dim];
    nodes[node].at(deriv, dim) = val[dim];
  }
}

bool NodesVars::is_constant_node(int node) const {
  // towr NodesVariablesPhaseBased::IsConstantNode: constant if either adjacent poly is constant.
  int last = (int)nodes.size() - 1;
  std::vector<int> adj;
  if (node == 0)
    adj = {0};
  else if (node == last)
    adj = {last - 1};
  else
    adj = {node - 1, node};
  for (int p : adj)
    if (poly_info[p].is_constant) return true;
  return false;
}
std::vector<int> NodesVars::non_constant_nodes() const {
  std::vector<int> r;
  for (int i = 0; i < (int)nodes.size(); ++i)
    if (!is_constant_node(i)) r.push_back(i);
  return r;
}
int NodesVars::phase_of_node(int node) const {
  // towr NodesVariablesPhaseBased::GetPhase (valid for non-constant nodes): phase of first adjacent poly
  int last = (int)nodes.size() - 1;
  int poly = node == 0 ? 0 : (node == last ? last - 1 : node - 1);
  return poly_info[poly].phase;
}
std::vector<double> NodesVars::phase_to_poly_durations(const std::vector<double>& phase_dur) const {
  std::vector<double> d;
  d.reserve(poly_info.size());
  for (const auto& pi : poly_info) d.push_back(phase_dur.at(pi.phase) / pi.n_polys_in_phase);
  return d;
}

// towr NodesVariablesAll: index -> node floor(idx/6), [px,py,pz,vx,vy,vz].
NodesVars make_nodes_all(int n_nodes, const std::string& name) {
  NodesVars nv;
  nv.name = name;
  nv.nodes.assign(n_nodes, Node());
  for (int idx = 0; idx < n_nodes * 6; ++idx) {
    int internal = idx % 6;
    nv.index_map.push_back({NVI{idx / 6, internal < 3 ? kPos : kVel, internal % 3}});
  }
  nv.finalize();
  return nv;
}

// nodes_variables_dynamic_phase_based.cpp:10-34 (BuildDynamicPolyInfos)
static std::vector<PolyInfo> build_dynamic_poly_infos(int phase_count, bool first_phase_constant,
                                                      const std::vector<int>& n_polys_changing) {
  std::vector<PolyInfo> info;
  bool constant = first_phase_constant;
  int change_count = 0;
  for (int i = 0; i < phase_count; ++i) {
    if (constant)
      info.push_back({i, 0, 1, true});
    else {
      int np = n_polys_changing.at(change_count);
      for (int j = 0; j < np; ++j) info.push_back({i, j, np, false});
      change_count++;
    }
    constant = !constant;
  }
  return info;
}

// nodes_variables_dynamic_phase_based.cpp:58-106: swing nodes optimise [px,vx,py,vy,pz,vz];
// a stance phase contributes one xyz position shared by its two nodes, velocities pinned to 0.
NodesVars make_ee_motion(int phase_count, bool in_contact_at_start, const std::string& name,
                         const std::vector<int>& n_polys_swing) {
  NodesVars nv;
  nv.name = name;
  nv.poly_info = build_dynamic_poly_infos(phase_count, in_contact_at_start, n_polys_swing);
  nv.nodes.assign(nv.poly_info.size() + 1, Node());
  for (int node = 0; node < (int)nv.nodes.size(); ++node) {
    if (!nv.is_constant_node(node)) {
      for (int dim = 0; dim < 3; ++dim) {
        nv.index_map.push_back({NVI{node, kPos, dim}});
        nv.index_map.push_back({NVI{node, kVel, dim}});
      }
    } else {
      nv.nodes[node].v = {0, 0, 0};
      nv.nodes[node + 1].v = {0, 0, 0};
      for (int dim = 0; dim < 3; ++dim) nv.index_map.push_back({NVI{node, kPos, dim}, NVI{node + 1, kPos, dim}});
      node += 1;
    }
  }
  nv.finalize();
  return nv;
}

// nodes_variables_dynamic_phase_based.cpp:108-151: mirrored -- stance nodes free, swing nodes 0.
NodesVars make_ee_force(int phase_count, bool in_contact_at_start, const std::string& name,
                        const std::vector<int>& n_polys_stance) {
  NodesVars nv;
  nv.name = name;
  nv.poly_info = build_dynamic_poly_infos(phase_count, !in_contact_at_start, n_polys_stance);
  nv.nodes.assign(nv.poly_info.size() + 1, Node());
  for (int node = 0; node < (int)nv.nodes.size(); ++node) {
    if (!nv.is_constant_node(node)) {
      for (int dim = 0; dim < 3; ++dim) {
        nv.index_map.push_back({NVI{node, kPos, dim}});
        nv.index_map.push_back({NVI{node, kVel, dim}});
      }
    } else {
      nv.nodes[node] = Node();
      nv.nodes[node + 1] = Node();
      node += 1;
    }
  }
  nv.finalize();
  return nv;
}

// ------------------------------------------------------------------ PhaseDurations -------------
void PhaseDurations::set_values(const double* x) {
  double sum = 0;
  for (int i = 0; i < rows(); ++i) {
    durations[i] = x[i];
    sum += x[i];
  }
  durations.back() = t_total - sum;  // last phase fills up to the total time
}
bool PhaseDurations::is_contact_phase(double t) const {
  int phase = Spline::segment_id(t, durations);
  return phase % 2 == 0 ? initial_contact : !initial_contact;
}

// ------------------------------------------------------------------ Spline ---------------------
// towr Spline::GetSegmentID: first segment whose accumulated end time >= t - 1e-10.
int Spline::segment_id(double t_global, const std::vector<double>& durations) {
  const double eps = 1e-10;
  double t = 0;
  int i = 0;
  for (double d : durations) {
    t += d;
    if (t >= t_global - eps) return i;
    i++;
  }
  return (int)durations.size() - 1;  // towr asserts here; clamp to the last segment instead
}
std::pair<int, double> Spline::local_time(double t_global) const {
  int id = segment_id(t_global, poly_dur);
  double tl = t_global;
  for (int i = 0; i < id; ++i) tl -= poly_dur[i];
  return {id, tl};
}
double Spline::total_time() const { return std::accumulate(poly_dur.begin(), poly_dur.end(), 0.0); }

// towr CubicHermitePolynomial::UpdateCoeff + Polynomial::GetPoint
State Spline::point(int id, double t) const {
  const Node& n0 = nv->nodes[id];
  const Node& n1 = nv->nodes[id + 1];
  double T = poly_dur[id];
  State s;
  for (int d = 0; d < 3; ++d) {
    double a = n0.p[d], b = n0.v[d];
    double c = -(3 * (n0.p[d] - n1.p[d]) + T * (2 * n0.v[d] + n1.v[d])) / (T * T);
    double dd = (2 * (n0.p[d] - n1.p[d]) + T * (n0.v[d] + n1.v[d])) / (T * T * T);
    s.p[d] = a + b * t + c * t * t + dd * t * t * t;
    s.v[d] = b + 2 * c * t + 3 * dd * t * t;
    s.a[d] = 2 * c + 6 * dd * t;
  }
  return s;
}

// towr CubicHermitePolynomial::GetDerivativeWrt{Start,End}Node
static double basis(int dxdt, bool end_node, int node_deriv, double t, double T) {
  double t2 = t * t, t3 = t2 * t, T2 = T * T, T3 = T2 * T;
  if (!end_node) {
    if (dxdt == kPos) return node_deriv == kPos ? (2 * t3) / T3 - (3 * t2) / T2 + 1 : t - (2 * t2) / T + t3 / T2;
    if (dxdt == kVel) return node_deriv == kPos ? (6 * t2) / T3 - (6 * t) / T2 : (3 * t2) / T2 - (4 * t) / T + 1;
    return node_deriv == kPos ? (12 * t) / T3 - 6 / T2 : (6 * t) / T2 - 4 / T;
  }
  if (dxdt == kPos) return node_deriv == kPos ? (3 * t2) / T2 - (2 * t3) / T3 : t3 / T2 - t2 / T;
  if (dxdt == kVel) return node_deriv == kPos ? (6 * t) / T2 - (6 * t2) / T3 : (3 * t2) / T2 - (2 * t) / T;
  return node_deriv == kPos ? 6 / T2 - (12 * t) / T3 : (6 * t) / T2 - 2 / T;
}

// towr NodeSpline::FillJacobianWrtNodes: optimisation indices shared by the start and end node
// of the active polynomial receive the SUM of both sensitivities.
SJac Spline::jac_wrt_nodes(int id, double tl, int dxdt) const {
  SJac J;
  double T = poly_dur[id];
  for (int side = 0; side < 2; ++side) {
    int node = id + side;
    for (int deriv = 0; deriv < 2; ++deriv)
      for (int dim = 0; dim < 3; ++dim) {
        int idx = nv->rev[node][deriv * 3 + dim];
        if (idx < 0) continue;
        J.add(idx, dim, basis(dxdt, side == 1, deriv, tl, T));
      }
  }
  return J;
}

// towr PhaseSpline::GetJacobianOfPosWrtDurations + PhaseDurations::GetJacobianOfPos
SJac Spline::jac_pos_wrt_durations(double t_global) const {
  auto lt = local_time(t_global);
  int id = lt.first;
  double t = lt.second, T = poly_dur[id];
  const Node& n0 = nv->nodes[id];
  const Node& n1 = nv->nodes[id + 1];
  State st = point(id, t);
  Vec3 dxdT;
  double t2 = t * t, t3 = t2 * t, T2 = T * T, T3 = T2 * T, T4 = T3 * T;
  for (int d = 0; d < 3; ++d) {
    double x0 = n0.p[d], x1 = n1.p[d], v0 = n0.v[d], v1 = n1.v[d];
    dxdT[d] = (t3 * (v0 + v1)) / T3 - (t2 * (2 * v0 + v1)) / T2 - (3 * t3 * (2 * x0 - 2 * x1 + T * v0 + T * v1)) / T4 +
              (2 * t2 * (3 * x0 - 3 * x1 + 2 * T * v0 + T * v1)) / T3;
  }
  const PolyInfo& pi = nv->poly_info[id];
  double inner = 1.0 / pi.n_polys_in_phase;
  Vec3 dx_dT = inner * (dxdT - (double)pi.poly_in_phase * st.v);

  int current_phase = segment_id(t_global, pd->durations);
  int P = (int)pd->durations.size();
  bool in_last = current_phase == P - 1;
  SJac J;
  if (!in_last) J.addcol(current_phase, dx_dT);
  for (int phase = 0; phase < current_phase; ++phase) {
    J.addcol(phase, -1.0 * st.v);
    if (in_last) J.addcol(phase, -1.0 * dx_dT);
  }
  return J;
}

// ------------------------------------------------------------------ Euler ----------------------
Mat3 Euler::R(const Vec3& e) {
  double sx = sin(e[X]), cx = cos(e[X]), sy = sin(e[Y]), cy = cos(e[Y]), sz = sin(e[Z]), cz = cos(e[Z]);
  Mat3 m = {{{cy * cz, cz * sx * sy - cx * sz, sx * sz + cx * cz * sy},
             {cy * sz, cx * cz + sx * sy * sz, cx * sy * sz - cz * sx},
             {-sy, cy * sx, cx * cy}}};
  return m;
}
void Euler::dR(const Vec3& e, Mat3 out[3]) {
  double sx = sin(e[X]), cx = cos(e[X]), sy = sin(e[Y]), cy = cos(e[Y]), sz = sin(e[Z]), cz = cos(e[Z]);
  out[X] = {{{0, cz * cx * sy + sx * sz, cx * sz - sx * cz * sy},
             {0, -sx * cz + cx * sy * sz, -sx * sy * sz - cz * cx},
             {0, cy * cx, -sx * cy}}};
  out[Y] = {{{-sy * cz, cz * sx * cy, cx * cz * cy}, {-sy * sz, sx * cy * sz, cx * cy * sz}, {-cy, -sy * sx, -cx * sy}}};
  out[Z] = {{{-cy * sz, -sz * sx * sy - cx * cz, sx * cz - cx * sz * sy},
             {cy * cz, -cx * sz + sx * sy * cz, cx * sy * cz + sz * sx},
             {0, 0, 0}}};
}
Mat3 Euler::M(const Vec3& e) {
  double sy = sin(e[Y]), cy = cos(e[Y]), sz = sin(e[Z]), cz = cos(e[Z]);
  Mat3 m = {{{cy * cz, -sz, 0}, {cy * sz, cz, 0}, {-sy, 0, 1}}};
  return m;
}
void Euler::dM(const Vec3& e, Mat3 out[3]) {
  double sy = sin(e[Y]), cy = cos(e[Y]), sz = sin(e[Z]), cz = cos(e[Z]);
  out[X] = {{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}};
  out[Y] = {{{-sy * cz, 0, 0}, {-sy * sz, 0, 0}, {-cy, 0, 0}}};
  out[Z] = {{{-cy * sz, -cz, 0}, {cy * cz, -sz, 0}, {0, 0, 0}}};
}
void Euler::d2M(const Vec3& e, Mat3 out[3][3]) {
  double sy = sin(e[Y]), cy = cos(e[Y]), sz = sin(e[Z]), cz = cos(e[Z]);
  Mat3 zero = {{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out[i][j] = zero;
  out[Y][Y] = {{{-cy * cz, 0, 0}, {-cy * sz, 0, 0}, {sy, 0, 0}}};
  out[Y][Z] = out[Z][Y] = {{{sy * sz, 0, 0}, {-sy * cz, 0, 0}, {0, 0, 0}}};
  out[Z][Z] = {{{-cy * cz, sz, 0}, {-cy * sz, -cz, 0}, {0, 0, 0}}};
}
Mat3 Euler::Mdot(const Vec3& e, const Vec3& ed) {
  Mat3 d[3];
  dM(e, d);
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = d[Y].m[i][j] * ed[Y] + d[Z].m[i][j] * ed[Z];
  return r;
}
Vec3 Euler::ang_vel(double t) const {
  State o = s->point(t);
  return mul(M(o.p), o.v);
}
Vec3 Euler::ang_acc(double t) const {
  State o = s->point(t);
  return mul(Mdot(o.p, o.v), o.v) + mul(M(o.p), o.a);
}

// d(R v)/d(nodes)  (inverse: d(R^T v)/d(nodes)); towr EulerConverter::DerivOfRotVecMult.
SJac Euler::deriv_rot_vec_mult(double t, const Vec3& v, bool inverse) const {
  State o = s->point(t);
  Mat3 d[3];
  dR(o.p, d);
  SJac Jp = s->jac_wrt_nodes(t, kPos);
  SJac J;
  for (size_t c = 0; c < Jp.col.size(); ++c) {
    Vec3 acc{0, 0, 0};
    for (int k = 0; k < 3; ++k) {
      double w = Jp.val[c][k];  // d e_k / d var
      if (w == 0.0) continue;
      Mat3 dk = inverse ? transpose(d[k]) : d[k];
      acc = acc + w * mul(dk, v);
    }
    J.addcol(Jp.col[c], acc);
  }
  return J;
}

// d(omega)/d(nodes), omega = M(e) edot; towr EulerConverter::GetDerivOfAngVelWrtEulerNodes.
SJac Euler::deriv_ang_vel(double t) const {
  State o = s->point(t);
  Mat3 m = M(o.p), d[3];
  dM(o.p, d);
  SJac Jp = s->jac_wrt_nodes(t, kPos), Jv = s->jac_wrt_nodes(t, kVel);
  SJac J;
  for (size_t c = 0; c < Jp.col.size(); ++c) {
    Vec3 acc{0, 0, 0};
    for (int k = 0; k < 3; ++k)
      if (Jp.val[c][k] != 0.0) acc = acc + Jp.val[c][k] * mul(d[k], o.v);
    J.addcol(Jp.col[c], acc);
  }
  J.axpy(1.0, Jv.lmul(m));
  return J;
}

// d(omega_dot)/d(nodes), omega_dot = Mdot edot + M eddot; towr ...::GetDerivOfAngAccWrtEulerNodes.
SJac Euler::deriv_ang_acc(double t) const {
  State o = s->point(t);
  Mat3 m = M(o.p), md = Mdot(o.p, o.v), d[3], d2[3][3];
  dM(o.p, d);
  d2M(o.p, d2);
  SJac Jp = s->jac_wrt_nodes(t, kPos), Jv = s->jac_wrt_nodes(t, kVel), Ja = s->jac_wrt_nodes(t, kAcc);
  SJac J;
  for (size_t c = 0; c < Jp.col.size(); ++c) {
    Vec3 acc{0, 0, 0};
    for (int j = 0; j < 3; ++j) {
      double w = Jp.val[c][j];
      if (w == 0.0) continue;
      // d(Mdot)/d e_j = sum_k d2M[j][k] * edot_k
      Mat3 dmd;
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) dmd.m[a][b] = d2[j][Y].m[a][b] * o.v[Y] + d2[j][Z].m[a][b] * o.v[Z];
      acc = acc + w * (mul(dmd, o.v) + mul(d[j], o.a));
    }
    J.addcol(Jp.col[c], acc);
  }
  for (size_t c = 0; c < Jv.col.size(); ++c) {
    Vec3 acc{0, 0, 0};
    for (int j = 0; j < 3; ++j) {
      double w = Jv.val[c][j];
      if (w == 0.0) continue;
      Vec3 mdcol{md.m[0][j], md.m[1][j], md.m[2][j]};
      acc = acc + w * (mul(d[j], o.v) + mdcol);
    }
    J.addcol(Jv.col[c], acc);
  }
  J.axpy(1.0, Ja.lmul(m));
  return J;
}

}  // namespace oracle
