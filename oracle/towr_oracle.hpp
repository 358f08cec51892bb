// TEST INFRASTRUCTURE ONLY -- CPU oracle for the phys-optim hot path.
//
// A from-scratch, single-threaded C++17 restatement of the NLP that the reference's
// `towr_phys_optim/phys_optim.cpp` hands to IPOPT: the spline algebra, variable sets,
// constraint sets, cost terms and their analytic first derivatives.  Everything that lives
// in the reference repo is restated from the cited file:line; everything that lives in the
// un-vendored third-party libraries (davrempe/towr fork of ethz-adrl/towr v1.4, ifopt 2.0.1)
// is restated from their published algorithm as recorded in SURVEY.md section 8(c).
//
// PARITY UNPINNED: the reference ships no tests / golden vectors for this path and its
// binary cannot be built offline (TOWR, ifopt, IPOPT, HSL, Eigen, gflags absent).  This oracle
// is therefore validated by closed forms and finite differences (tests/test_oracle_*.py), not
// against reference output.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this code.  The product path (contact-human-dynamics_b200/) never does.
#pragma once
#include <array>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace oracle {

using Vec3 = std::array<double, 3>;
struct Mat3 {
  double m[3][3];
};

enum Dx { kPos = 0, kVel = 1, kAcc = 2 };
enum { X = 0, Y = 1, Z = 2 };

inline Vec3 operator+(const Vec3& a, const Vec3& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline Vec3 operator*(double s, const Vec3& a) { return {s * a[0], s * a[1], s * a[2]}; }
inline double dot(const Vec3& a, const Vec3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
inline Vec3 mul(const Mat3& A, const Vec3& v) {
  Vec3 r;
  for (int i = 0; i < 3; ++i) r[i] = A.m[i][0] * v[0] + A.m[i][1] * v[1] + A.m[i][2] * v[2];
  return r;
}
inline Mat3 mul(const Mat3& A, const Mat3& B) {
  Mat3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
  return C;
}
inline Mat3 transpose(const Mat3& A) {
  Mat3 T;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T.m[i][j] = A.m[j][i];
  return T;
}
inline Mat3 skew(const Vec3& v) {  // humanoid_rigid_body_dynamics.cpp:55-65 (Cross)
  Mat3 S = {{{0, -v[2], v[1]}, {v[2], 0, -v[0]}, {-v[1], v[0], 0}}};
  return S;
}

struct Node {
  Vec3 p{0, 0, 0}, v{0, 0, 0};
  double& at(int deriv, int dim) { return deriv == kPos ? p[dim] : v[dim]; }
  double at(int deriv, int dim) const { return deriv == kPos ? p[dim] : v[dim]; }
};
struct State {
  Vec3 p, v, a;
  const Vec3& at(int d) const { return d == kPos ? p : (d == kVel ? v : a); }
};

// towr NodesVariables::NodeValueInfo
struct NVI {
  int node, deriv, dim;
};
// towr NodesVariablesPhaseBased::PolyInfo
struct PolyInfo {
  int phase, poly_in_phase, n_polys_in_phase;
  bool is_constant;
};

// A column-sparse 3 x n Jacobian (the role Eigen::SparseMatrix<double,RowMajor> plays in towr).
struct SJac {
  std::vector<int> col;
  std::vector<Vec3> val;
  void add(int c, int dim, double v) {
    for (size_t i = 0; i < col.size(); ++i)
      if (col[i] == c) {
        val[i][dim] += v;
        return;
      }
    col.push_back(c);
    Vec3 z{0, 0, 0};
    z[dim] = v;
    val.push_back(z);
  }
  void addcol(int c, const Vec3& v) {
    for (int d = 0; d < 3; ++d) add(c, d, v[d]);
  }
  void axpy(double s, const SJac& o) {
    for (size_t i = 0; i < o.col.size(); ++i) addcol(o.col[i], s * o.val[i]);
  }
  SJac lmul(const Mat3& A) const {  // A * this
    SJac r;
    r.col = col;
    r.val.resize(val.size());
    for (size_t i = 0; i < col.size(); ++i) r.val[i] = mul(A, val[i]);
    return r;
  }
  SJac shifted(int off) const {
    SJac r = *this;
    for (auto& c : r.col) c += off;
    return r;
  }
};

// ---------------------------------------------------------------------------------------------
// Variable sets (towr NodesVariables / NodesVariablesAll / NodesVariablesPhaseBased + the
// reference's nodes_variables_dynamic_phase_based.cpp) and PhaseDurations.
// ---------------------------------------------------------------------------------------------
struct NodesVars {
  std::string name;
  std::vector<Node> nodes;
  std::vector<std::vector<NVI>> index_map;   // opt index -> node value infos
  std::vector<std::array<int, 6>> rev;       // node -> [deriv*3+dim] -> opt index or -1
  std::vector<double> lo, hi;                // variable bounds (+-1e20 = none)
  std::vector<PolyInfo> poly_info;           // phase based only
  int offset = 0;                            // column offset inside the stacked NLP x

  int rows() const { return (int)index_map.size(); }
  void finalize();
  void get_values(double* x) const;
  void set_values(const double* x);
  void set_by_linear_interpolation(const Vec3& a, const Vec3& b, double T);
  void add_bound(int node, int deriv, const Vec3& val);  // equality bound on all 3 dims
  // phase based helpers
  bool is_constant_node(int node) const;
  std::vector<int> non_constant_nodes() const;
  int phase_of_node(int node) const;
  std::vector<double> phase_to_poly_durations(const std::vector<double>& phase_dur) const;
};

NodesVars make_nodes_all(int n_nodes, const std::string& name);
NodesVars make_ee_motion(int phase_count, bool in_contact_at_start, const std::string& name,
                         const std::vector<int>& n_polys_swing);
NodesVars make_ee_force(int phase_count, bool in_contact_at_start, const std::string& name,
                        const std::vector<int>& n_polys_stance);

struct PhaseDurations {
  std::vector<double> durations;  // all P phases
  double t_total = 0;
  bool initial_contact = false;
  double lo = 0, hi = 500;
  int offset = 0;
  int rows() const { return (int)durations.size() - 1; }
  void set_values(const double* x);
  bool is_contact_phase(double t) const;
};

struct Spline {
  const NodesVars* nv = nullptr;
  const PhaseDurations* pd = nullptr;  // phase splines only
  std::vector<double> poly_dur;
  void update_durations() {
    if (pd) poly_dur = nv->phase_to_poly_durations(pd->durations);
  }
  static int segment_id(double t, const std::vector<double>& d);
  std::pair<int, double> local_time(double t) const;
  double total_time() const;
  State point(int id, double tl) const;
  State point(double t) const {
    auto lt = local_time(t);
    return point(lt.first, lt.second);
  }
  SJac jac_wrt_nodes(int id, double tl, int dxdt) const;  // local (set-relative) columns
  SJac jac_wrt_nodes(double t, int dxdt) const {
    auto lt = local_time(t);
    return jac_wrt_nodes(lt.first, lt.second, dxdt);
  }
  SJac jac_pos_wrt_durations(double t) const;  // columns = free phase durations
};

// Euler ZYX kinematics (towr EulerConverter restated).
struct Euler {
  const Spline* s = nullptr;
  static Mat3 R(const Vec3& e);
  static void dR(const Vec3& e, Mat3 out[3]);
  static Mat3 M(const Vec3& e);
  static void dM(const Vec3& e, Mat3 out[3]);
  static void d2M(const Vec3& e, Mat3 out[3][3]);
  static Mat3 Mdot(const Vec3& e, const Vec3& ed);
  Mat3 rot(double t) const { return R(s->point(t).p); }
  Vec3 ang_vel(double t) const;
  Vec3 ang_acc(double t) const;
  SJac deriv_rot_vec_mult(double t, const Vec3& v, bool inverse) const;
  SJac deriv_ang_vel(double t) const;
  SJac deriv_ang_acc(double t) const;
};

}  // namespace oracle
