// Host-side (FP64, C++17) controller design for libmpcb200: small dense linear algebra, the discrete Riccati
// solver, condensing of the reference's MPC problem and construction of the stacked ADMM operators.
// Runs once per controller (the reference's cold path, src/sub/design_mpc.jl:54-129); nothing here is on the
// per-solve hot path.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

#include "../../include/mpcb200.h"

namespace mpcb {

// Column-major dense matrix (Julia layout).
struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() = default;
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)j * r + i]; }
  double operator()(int i, int j) const { return a[(size_t)j * r + i]; }
  static Mat eye(int n) {
    Mat m(n, n);
    for (int i = 0; i < n; i++) m(i, i) = 1.0;
    return m;
  }
  static Mat from(const double* p, int r_, int c_) {
    Mat m(r_, c_);
    if (p) m.a.assign(p, p + (size_t)r_ * c_);
    return m;
  }
};

Mat matmul(const Mat& A, const Mat& B);        // A*B
Mat matmul_tn(const Mat& A, const Mat& B);     // A'*B
Mat transpose(const Mat& A);
Mat add(const Mat& A, const Mat& B, double sb = 1.0);  // A + sb*B
bool cholesky_lower(Mat& A);                   // in place, lower triangle; false if not PD
bool spd_inverse(const Mat& A, Mat& Ainv);     // via Cholesky
bool lu_solve(Mat A, Mat& B);                  // solves A X = B in place (partial pivoting)
bool sym_extreme_eigs(const Mat& A, double& lmin, double& lmax);  // A symmetric positive definite
bool dare_sda(const Mat& A, const Mat& B, const Mat& Q, const Mat& R, Mat& P, std::string& err);

// Everything the kernels need, still on the host.
struct Design {
  int nx = 0, nu = 0, H = 0, nz = 0, mg = 0, nt = 0, np = 0;  // np = 2nx+nu parameter length [x0; xref; uref]
  int nball = 0;   // contractive terminal set: the first nball general rows form one ball |G v - b(p)|_2 <= sqrt(0.9) |x0 - xref|_2
  Mat A, B, Q, R, S, P;
  bool use_R = true, use_S = false;
  Mat Pc;   // nz x nz   Hessian in absolute-input coordinates (includes the S term)
  Mat Lq;   // nz x np   q(p) = Lq p
  Mat Lv;   // nz x np   unconstrained optimum v_unc(p) = Lv p = -Pc^-1 Lq p (settings.cold_init; empty otherwise)
  Mat G;    // mg x nz   general rows
  Mat Lb;   // mg x np   bound offset b(p) = Lb p
  std::vector<double> lo, hi;        // nt: constant part of the bounds ([umin..|lg..], [umax..|ug..])
  std::vector<unsigned char> is_eq;  // nt
  std::vector<double> rho_vec;       // nt
  Mat T;    // nt x nt   [I;G] K^-1 [I,G']
  Mat C;    // nt x nt   [[Pc, G'],[G, 0]]
  Mat Ac;   // nt x nz   [I;G]  (certificate pass)
  double rho = 0, lmin = 0, lmax = 0;
  // Stage-wise (Riccati) form of the x-update K x~ = r for box-only problems without the S term (admm_riccati.cu): per stage
  // K_k, K_k', Lambda_k^-1, Acl_k, Acl_k' in the kernel's layout (ric_stage_doubles(nx, nu) doubles per stage); empty when
  // the form does not apply (general rows, S term).
  std::vector<double> ric_stage;
};

// ineq_scale multiplies the step size of the INEQUALITY general rows (state box): the second rung of the rho ladder (mpcb_api.cu)
// cached factors of the LQ problem behind K = Pc + (sigma + rho) I (box-only, no S term): fills D.ric_stage
bool riccati_factors(Design& D, double sigma);

int build_design(const mpcb_linear_desc& d, const mpcb_settings& s, Design& out, std::string& err, double ineq_scale = 1.0);

}  // namespace mpcb
