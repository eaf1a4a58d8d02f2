// Host-side controller design (see host_design.hpp).  Reference anchors, relative to /root/reference:
//   weights            src/sub/design_mpc.jl:264-283
//   terminal cost P    src/sub/design_mpc.jl:327           (ControlSystems.are -> dare_sda here)
//   terminal equality  src/sub/design_mpc.jl:330-331
//   cost               src/sub/design_mpc.jl:405-465       (no 1/2 factor; S-term only with R != 0 and S != 0)
//   constraints        src/sub/model_modeler_implementation/linear/mpc_modeler_implementation_linear.jl:58-87
#include "host_design.hpp"

#include <algorithm>
#include <cmath>
#include <thread>

namespace mpcb {

namespace {
// Columns [0, n) of a product split over host threads when the product is large enough to pay for them (the design of a
// controller with hundreds of decision variables is O(nz^3): condensing, K^-1, the stacked operator).  Every column is computed
// by exactly one thread with the same operation order as the serial loop: results do not depend on the thread count.
template <class F>
void parallel_columns(int n, double work_per_column, F&& body) {
  unsigned nt = std::thread::hardware_concurrency();
  nt = std::min<unsigned>(nt == 0 ? 1 : nt, 16);
  if (work_per_column * n < 4e6 || nt < 2 || n < 2) { body(0, n); return; }
  nt = std::min<unsigned>(nt, (unsigned)n);
  std::vector<std::thread> th;
  const int chunk = (n + (int)nt - 1) / (int)nt;
  for (unsigned t = 1; t < nt; t++) {
    const int j0 = (int)t * chunk, j1 = std::min(n, j0 + chunk);
    if (j0 < j1) th.emplace_back([&body, j0, j1]() { body(j0, j1); });
  }
  body(0, std::min(n, chunk));
  for (auto& x : th) x.join();
}
}  // namespace

Mat matmul(const Mat& A, const Mat& B) {
  Mat C(A.r, B.c);
  const int m = A.r, k = A.c, n = B.c;
  parallel_columns(n, (double)m * k, [&](int j0, int j1) {
    for (int j = j0; j < j1; j++) {
      double* cj = &C.a[(size_t)j * m];
      for (int l = 0; l < k; l++) {
        const double b = B(l, j);
        if (b == 0.0) continue;
        const double* al = &A.a[(size_t)l * m];
        for (int i = 0; i < m; i++) cj[i] += al[i] * b;
      }
    }
  });
  return C;
}

Mat matmul_tn(const Mat& A, const Mat& B) {  // A' * B, A is k x m, B is k x n
  Mat C(A.c, B.c);
  const int m = A.c, k = A.r, n = B.c;
  parallel_columns(n, (double)m * k, [&](int j0, int j1) {
    for (int j = j0; j < j1; j++) {
      const double* bj = &B.a[(size_t)j * k];
      for (int i = 0; i < m; i++) {
        const double* ai = &A.a[(size_t)i * k];
        double s = 0.0;
        for (int l = 0; l < k; l++) s += ai[l] * bj[l];
        C(i, j) = s;
      }
    }
  });
  return C;
}

Mat transpose(const Mat& A) {
  Mat T(A.c, A.r);
  for (int j = 0; j < A.c; j++)
    for (int i = 0; i < A.r; i++) T(j, i) = A(i, j);
  return T;
}

Mat add(const Mat& A, const Mat& B, double sb) {
  Mat C = A;
  for (size_t i = 0; i < C.a.size(); i++) C.a[i] += sb * B.a[i];
  return C;
}

bool cholesky_lower(Mat& A) {
  const int n = A.r;
  for (int j = 0; j < n; j++) {
    double d = A(j, j);
    for (int k = 0; k < j; k++) d -= A(j, k) * A(j, k);
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    A(j, j) = d;
    // column j below the diagonal: A(i,j) = (A(i,j) - sum_k A(i,k) A(j,k)) / d, done as axpys over columns k
    for (int k = 0; k < j; k++) {
      const double ljk = A(j, k);
      if (ljk == 0.0) continue;
      double* cj = &A.a[(size_t)j * n];
      const double* ck = &A.a[(size_t)k * n];
      for (int i = j + 1; i < n; i++) cj[i] -= ck[i] * ljk;
    }
    const double inv = 1.0 / d;
    for (int i = j + 1; i < n; i++) A(i, j) *= inv;
    for (int i = 0; i < j; i++) A(i, j) = 0.0;
  }
  return true;
}

bool spd_inverse(const Mat& A, Mat& Ainv) {
  const int n = A.r;
  Mat L = A;
  if (!cholesky_lower(L)) return false;
  // Linv = L^-1 (lower), column by column
  Mat Li(n, n);
  const Mat Lt = transpose(L);      // row i of L contiguous
  parallel_columns(n, (double)n * n / 6.0, [&](int j0, int j1) {
    for (int j = j0; j < j1; j++) {
      Li(j, j) = 1.0 / L(j, j);
      double* lj = &Li.a[(size_t)j * n];
      for (int i = j + 1; i < n; i++) {
        const double* li = &Lt.a[(size_t)i * n];
        double s = 0.0;
        for (int k = j; k < i; k++) s += li[k] * lj[k];
        lj[i] = -s / li[i];
      }
    }
  });
  Ainv = matmul_tn(Li, Li);  // L^-T L^-1
  for (int j = 0; j < n; j++)
    for (int i = 0; i < j; i++) {
      const double v = 0.5 * (Ainv(i, j) + Ainv(j, i));
      Ainv(i, j) = v;
      Ainv(j, i) = v;
    }
  return true;
}

bool lu_solve(Mat A, Mat& B) {
  const int n = A.r, m = B.c;
  for (int k = 0; k < n; k++) {
    int piv = k;
    double best = std::fabs(A(k, k));
    for (int i = k + 1; i < n; i++)
      if (std::fabs(A(i, k)) > best) { best = std::fabs(A(i, k)); piv = i; }
    if (best == 0.0) return false;
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(A(k, j), A(piv, j));
      for (int j = 0; j < m; j++) std::swap(B(k, j), B(piv, j));
    }
    const double inv = 1.0 / A(k, k);
    for (int i = k + 1; i < n; i++) {
      const double f = A(i, k) * inv;
      if (f == 0.0) continue;
      for (int j = k + 1; j < n; j++) A(i, j) -= f * A(k, j);
      for (int j = 0; j < m; j++) B(i, j) -= f * B(k, j);
    }
  }
  for (int j = 0; j < m; j++)
    for (int i = n - 1; i >= 0; i--) {
      double s = B(i, j);
      for (int k = i + 1; k < n; k++) s -= A(i, k) * B(k, j);
      B(i, j) = s / A(i, i);
    }
  return true;
}

bool sym_extreme_eigs(const Mat& A, double& lmin, double& lmax) {
  const int n = A.r;
  // rho = sqrt(lmin lmax) is a step-size heuristic: large matrices stop at 1e-10 relative change (each sweep is O(n^2))
  const double tol = n <= 256 ? 1e-15 : 1e-9;
  const int max_sweeps = n <= 256 ? 3000 : 400;
  std::vector<double> v(n), w(n);
  auto normalize = [&](std::vector<double>& x) {
    double s = 0;
    for (double t : x) s += t * t;
    s = std::sqrt(s);
    for (double& t : x) t /= s;
    return s;
  };
  // largest: power iteration from a fixed, non-degenerate start
  for (int i = 0; i < n; i++) v[i] = 1.0 + 0.37 * std::sin(1.0 + 1.7 * i);
  normalize(v);
  double lam = 0;
  for (int it = 0; it < max_sweeps; it++) {
    for (int i = 0; i < n; i++) w[i] = 0;
    for (int j = 0; j < n; j++) {
      const double vj = v[j];
      const double* aj = &A.a[(size_t)j * n];
      for (int i = 0; i < n; i++) w[i] += aj[i] * vj;
    }
    double nl = 0;
    for (int i = 0; i < n; i++) nl += w[i] * v[i];
    normalize(w);
    v.swap(w);
    if (it > 5 && std::fabs(nl - lam) <= tol * std::fabs(nl)) { lam = nl; break; }
    lam = nl;
  }
  lmax = lam;
  // smallest: inverse iteration through the Cholesky factor
  Mat L = A;
  if (!cholesky_lower(L)) return false;
  for (int i = 0; i < n; i++) v[i] = 1.0 + 0.41 * std::cos(0.3 + 2.3 * i);
  normalize(v);
  double mu = 0;
  for (int it = 0; it < max_sweeps; it++) {
    w = v;
    for (int i = 0; i < n; i++) {  // L y = v
      double s = w[i];
      for (int k = 0; k < i; k++) s -= L(i, k) * w[k];
      w[i] = s / L(i, i);
    }
    for (int i = n - 1; i >= 0; i--) {  // L' x = y
      double s = w[i];
      for (int k = i + 1; k < n; k++) s -= L(k, i) * w[k];
      w[i] = s / L(i, i);
    }
    double nm = 0;
    for (int i = 0; i < n; i++) nm += w[i] * v[i];  // Rayleigh quotient of A^-1
    normalize(w);
    v.swap(w);
    if (it > 5 && std::fabs(nm - mu) <= tol * std::fabs(nm)) { mu = nm; break; }
    mu = nm;
  }
  lmin = 1.0 / mu;
  return lmin > 0 && lmax > 0;
}

// Structure-preserving doubling algorithm for X = A'X(I + G X)^-1 A + Q,  G = B R^-1 B'.
bool dare_sda(const Mat& A, const Mat& B, const Mat& Q, const Mat& R, Mat& P, std::string& err) {
  const int n = A.r;
  Mat Rinv_Bt = transpose(B);
  if (!lu_solve(R, Rinv_Bt)) { err = "dare: R is singular"; return false; }
  Mat Gk = matmul(B, Rinv_Bt);
  Mat Ak = A, Hk = Q;
  for (int it = 0; it < 100; it++) {
    Mat W = add(Mat::eye(n), matmul(Gk, Hk));
    Mat V1 = Ak, V2 = Gk;
    if (!lu_solve(W, V1) || !lu_solve(W, V2)) { err = "dare: singular I + G H"; return false; }
    Mat Akt = transpose(Ak);
    Mat An = matmul(Ak, V1);
    Mat Gn = add(Gk, matmul(matmul(Ak, V2), Akt));
    Mat Hn = add(Hk, matmul(matmul(Akt, Hk), V1));
    double diff = 0, nrm = 0;
    for (size_t i = 0; i < Hn.a.size(); i++) {
      diff = std::max(diff, std::fabs(Hn.a[i] - Hk.a[i]));
      nrm = std::max(nrm, std::fabs(Hn.a[i]));
    }
    Ak = An; Gk = Gn; Hk = Hn;
    if (!std::isfinite(nrm)) { err = "dare: diverged"; return false; }
    if (diff <= 1e-15 * std::max(1.0, nrm)) {
      P = Hk;
      for (int j = 0; j < n; j++)
        for (int i = 0; i < j; i++) { const double v = 0.5 * (P(i, j) + P(j, i)); P(i, j) = v; P(j, i) = v; }
      return true;
    }
  }
  err = "dare: doubling did not converge (is (A,B) stabilisable?)";
  return false;
}

// K x~ = r with K = Pc + (sigma + rho) I is the optimality system of
//     min  sum_{k=1..H} 1/2 e_k' W_k e_k + sum_{k<H} (1/2 u_k' Rh u_k - r_k' u_k),   e_0 = 0,  e_{k+1} = A e_k + B u_k,
// W_k = 2Q (k < H), 2P (k = H), Rh = 2R + (sigma + rho) I  (Pc of build_design below: the factor 2 is the reference's missing
// 1/2, design_mpc.jl:436-465).  Backward Riccati recursion from Pi_H = W_H:
//     Lam_k = Rh + B' Pi_{k+1} B,  K_k = Lam_k^-1 B' Pi_{k+1} A,  Acl_k = A - B K_k,  Pi_k = W_k + A' Pi_{k+1} Acl_k   (W_0 = 0)
// The kernel needs K_k, K_k', Lam_k^-1, Acl_k, Acl_k' per stage (layout of admm_riccati.cu: every matrix row-major, padded to
// an even number of doubles).
bool riccati_factors(Design& D, double sigma) {
  D.ric_stage.clear();
  if (D.mg != 0 || D.use_S) return false;
  const int nx = D.nx, nu = D.nu, H = D.H;
  auto ev = [](int n) { return (n + 1) & ~1; };
  const int oK = 0, oKT = oK + ev(nu * nx), oLI = oKT + ev(nx * nu), oACL = oLI + ev(nu * nu), oACLT = oACL + ev(nx * nx), SZ = oACLT + ev(nx * nx);
  D.ric_stage.assign((size_t)H * SZ, 0.0);
  Mat Rh(nu, nu);
  for (int j = 0; j < nu; j++)
    for (int i = 0; i < nu; i++) Rh(i, j) = (D.use_R ? 2.0 * D.R(i, j) : 0.0) + (i == j ? sigma + D.rho : 0.0);
  Mat Pi(nx, nx);
  for (size_t i = 0; i < Pi.a.size(); i++) Pi.a[i] = 2.0 * D.P.a[i];
  const Mat At = transpose(D.A), Bt = transpose(D.B);
  for (int k = H - 1; k >= 0; k--) {
    const Mat PB = matmul(Pi, D.B);                         // nx x nu
    const Mat Lam = add(Rh, matmul(Bt, PB));
    Mat Li;
    if (!spd_inverse(Lam, Li)) { D.ric_stage.clear(); return false; }
    const Mat K = matmul(Li, matmul(transpose(PB), D.A));   // nu x nx   (Pi symmetric: B' Pi = (Pi B)')
    const Mat Acl = add(D.A, matmul(D.B, K), -1.0);
    double* st = &D.ric_stage[(size_t)k * SZ];
    for (int i = 0; i < nu; i++)
      for (int m = 0; m < nx; m++) { st[oK + i * nx + m] = K(i, m); st[oKT + m * nu + i] = K(i, m); }
    for (int i = 0; i < nu; i++)
      for (int l = 0; l < nu; l++) st[oLI + i * nu + l] = Li(i, l);
    for (int m = 0; m < nx; m++)
      for (int n = 0; n < nx; n++) { st[oACL + m * nx + n] = Acl(m, n); st[oACLT + n * nx + m] = Acl(m, n); }
    Mat Pn = matmul(At, matmul(Pi, Acl));
    if (k > 0)
      for (size_t i = 0; i < Pn.a.size(); i++) Pn.a[i] += 2.0 * D.Q.a[i];
    for (int j = 0; j < nx; j++)
      for (int i = 0; i <= j; i++) { const double v = 0.5 * (Pn(i, j) + Pn(j, i)); Pi(i, j) = v; Pi(j, i) = v; }
  }
  return true;
}

int build_design(const mpcb_linear_desc& d, const mpcb_settings& s, Design& D, std::string& err, double ineq_scale) {
  if (d.nx <= 0 || d.nu <= 0 || d.horizon <= 0) { err = "nx, nu, horizon must be positive"; return MPCB_ERR_INVALID; }
  if (!d.A || !d.B || !d.Q || !d.R || !d.umin || !d.umax) { err = "A, B, Q, R, umin, umax are required"; return MPCB_ERR_INVALID; }
  if (d.terminal_mode != MPCB_TERMINAL_NONE && d.terminal_mode != MPCB_TERMINAL_EQUALITY && d.terminal_mode != MPCB_TERMINAL_CONTRACTIVE) {
    err = "terminal_mode: none / equality / contractive are supported (neighborhood is unimplemented in the reference)";
    return MPCB_ERR_INVALID;
  }
  if (d.state_constraint && (!d.xmin || !d.xmax)) { err = "state_constraint needs xmin/xmax"; return MPCB_ERR_INVALID; }
  const int nx = d.nx, nu = d.nu, H = d.horizon, nz = nu * H, np = 2 * nx + nu;
  D.nx = nx; D.nu = nu; D.H = H; D.nz = nz; D.np = np;
  D.A = Mat::from(d.A, nx, nx); D.B = Mat::from(d.B, nx, nu);
  D.Q = Mat::from(d.Q, nx, nx); D.R = Mat::from(d.R, nu, nu); D.S = Mat::from(d.S, nu, nu);
  D.use_R = D.R(0, 0) != 0.0;                       // design_mpc.jl:436,449
  D.use_S = D.use_R && D.S(0, 0) != 0.0;            // design_mpc.jl:436 (S-term only in the R != 0 branch)
  for (int i = 0; i < nu; i++)
    if (!(d.umin[i] <= d.umax[i])) { err = "umin > umax"; return MPCB_ERR_INVALID; }
  if (d.P) D.P = Mat::from(d.P, nx, nx);
  else if (!dare_sda(D.A, D.B, D.Q, D.R, D.P, err)) return MPCB_ERR_NUMERIC;

  // prediction matrices: e_k = Phi_k e0 + Gam_k eps,  Gam_{k+1} = A Gam_k + [.. B at block k ..]
  std::vector<Mat> Phi(H + 1), Gam(H + 1);
  Phi[0] = Mat::eye(nx); Gam[0] = Mat(nx, nz);
  for (int k = 0; k < H; k++) {
    Phi[k + 1] = matmul(D.A, Phi[k]);
    Gam[k + 1] = matmul(D.A, Gam[k]);
    for (int j = 0; j < nu; j++)
      for (int i = 0; i < nx; i++) Gam[k + 1](i, k * nu + j) += D.B(i, j);
  }
  // Hessian / gradient map in deviation inputs: J = eps' M eps + 2 (Fe e0)' eps + const, Pc_dev = 2M
  // M = sum_k Gam_k' W_k Gam_k and Fe = sum_k Gam_k' W_k Phi_k as ONE product each over the stacked rows (k, r): column c of Gam_k is
  // zero unless k > c / nu (an input acts on later states only), so the dot product of columns i and j starts at stage max(i, j) / nu + 1,
  // and M is symmetric: a sixth of the flops of the plain products, split over host threads by column.
  Mat M(nz, nz), Fe(nz, nx);
  {
    const int rows = nx * (H + 1);
    Mat Gall(rows, nz), WGall(rows, nz), Phall(rows, nx);
    for (int k = 0; k <= H; k++) {
      const Mat& W = (k == H) ? D.P : D.Q;
      const Mat WG = matmul(W, Gam[k]);                 // nx x nz
      for (int c = 0; c < nz; c++)
        for (int r = 0; r < nx; r++) { Gall(k * nx + r, c) = Gam[k](r, c); WGall(k * nx + r, c) = WG(r, c); }
      for (int c = 0; c < nx; c++)
        for (int r = 0; r < nx; r++) Phall(k * nx + r, c) = Phi[k](r, c);
    }
    parallel_columns(nz, (double)rows * nz / 6.0, [&](int j0, int j1) {
      for (int j = j0; j < j1; j++) {
        const double* wj = &WGall.a[(size_t)j * rows];
        for (int i = 0; i <= j; i++) {
          const double* gi = &Gall.a[(size_t)i * rows];
          double acc = 0.0;
          for (int l = (j / nu + 1) * nx; l < rows; l++) acc += gi[l] * wj[l];
          M(i, j) = acc;
        }
        for (int c = 0; c < nx; c++) {
          const double* ph = &Phall.a[(size_t)c * rows];
          double acc = 0.0;
          for (int l = (j / nu + 1) * nx; l < rows; l++) acc += wj[l] * ph[l];
          Fe(j, c) = acc;
        }
      }
    });
    for (int j = 0; j < nz; j++)
      for (int i = 0; i < j; i++) M(j, i) = M(i, j);
  }
  if (D.use_R)
    for (int k = 0; k < H; k++)
      for (int j = 0; j < nu; j++)
        for (int i = 0; i < nu; i++) M(k * nu + i, k * nu + j) += D.R(i, j);
  Mat Pdev(nz, nz);
  for (size_t i = 0; i < M.a.size(); i++) Pdev.a[i] = 2.0 * M.a[i];
  D.Pc = Pdev;
  if (D.use_S) {  // sum_{k<H-1} (u_k - u_{k+1})' S (u_k - u_{k+1})   (design_mpc.jl:423-446)
    for (int k = 0; k + 1 < H; k++)
      for (int j = 0; j < nu; j++)
        for (int i = 0; i < nu; i++) {
          const double v = 2.0 * D.S(i, j);
          D.Pc(k * nu + i, k * nu + j) += v;
          D.Pc((k + 1) * nu + i, (k + 1) * nu + j) += v;
          D.Pc(k * nu + i, (k + 1) * nu + j) -= v;
          D.Pc((k + 1) * nu + i, k * nu + j) -= v;
        }
  }
  for (int j = 0; j < nz; j++)
    for (int i = 0; i < j; i++) { const double v = 0.5 * (D.Pc(i, j) + D.Pc(j, i)); D.Pc(i, j) = v; D.Pc(j, i) = v; }
  // q(p) = 2 Fe (x0 - xref) - Pdev (1 (x) uref)
  D.Lq = Mat(nz, np);
  for (int i = 0; i < nz; i++) {
    for (int j = 0; j < nx; j++) { D.Lq(i, j) = 2.0 * Fe(i, j); D.Lq(i, nx + j) = -2.0 * Fe(i, j); }
    for (int j = 0; j < nu; j++) {
      double sacc = 0;
      for (int k = 0; k < H; k++) sacc += Pdev(i, k * nu + j);
      D.Lq(i, 2 * nx + j) = -sacc;
    }
  }
  // general rows
  const bool term_rows = d.terminal_mode == MPCB_TERMINAL_EQUALITY || d.terminal_mode == MPCB_TERMINAL_CONTRACTIVE;
  const int mg = (term_rows ? nx : 0) + (d.state_constraint ? nx * H : 0);
  D.nball = d.terminal_mode == MPCB_TERMINAL_CONTRACTIVE ? nx : 0;
  D.mg = mg; D.nt = nz + mg;
  D.G = Mat(mg, nz); D.Lb = Mat(mg, np);
  D.lo.assign(D.nt, 0.0); D.hi.assign(D.nt, 0.0); D.is_eq.assign(D.nt, 0);
  // infinite bounds are legal for a Hyperrectangle; like OSQP (OSQP_INFTY) they are held as +-1e30 so that the support term of
  // the infeasibility certificate never forms inf * 0
  auto fin = [](double v) { return v > 1e30 ? 1e30 : (v < -1e30 ? -1e30 : v); };
  for (int k = 0; k < H; k++)
    for (int j = 0; j < nu; j++) { D.lo[k * nu + j] = fin(d.umin[j]); D.hi[k * nu + j] = fin(d.umax[j]); }
  int row = 0;
  auto add_rows = [&](int k, bool equality) {
    // Gam_k v in [lo,hi] + b,  b = -Phi_k x0 + (Phi_k - I*[!equality]) xref + Gam_k (1 (x) uref)
    for (int i = 0; i < nx; i++, row++) {
      for (int j = 0; j < nz; j++) D.G(row, j) = Gam[k](i, j);
      for (int j = 0; j < nx; j++) {
        D.Lb(row, j) = -Phi[k](i, j);
        D.Lb(row, nx + j) = Phi[k](i, j) - ((!equality && i == j) ? 1.0 : 0.0);
      }
      for (int j = 0; j < nu; j++) {
        double sacc = 0;
        for (int kk = 0; kk < H; kk++) sacc += Gam[k](i, kk * nu + j);
        D.Lb(row, 2 * nx + j) = sacc;
      }
      D.lo[nz + row] = equality ? 0.0 : fin(d.xmin[i]);
      D.hi[nz + row] = equality ? 0.0 : fin(d.xmax[i]);
      D.is_eq[nz + row] = equality ? 1 : 0;
    }
  };
  if (term_rows) add_rows(H, true);      // contractive: the same rows e_H = G v - b(p); their set is a ball, not {0} (bounds unused)
  for (int i = 0; i < D.nball; i++) { D.lo[nz + i] = -1e30; D.hi[nz + i] = 1e30; D.is_eq[nz + i] = 0; }
  if (d.state_constraint)
    for (int k = 1; k <= H; k++) add_rows(k, false);

  // rho
  if (!sym_extreme_eigs(D.Pc, D.lmin, D.lmax)) { err = "condensed Hessian is not positive definite (need R != 0 or full-rank Q)"; return MPCB_ERR_NUMERIC; }
  D.rho = s.rho > 0 ? s.rho : std::sqrt(D.lmin * D.lmax);
  // General rows get rho / |G_i|^2: the per-row step size that row equilibration (OSQP's Ruiz scaling of A) would give.
  // The rows of Gamma_k are tiny for slow plants (quadruple tank: |G_i|^2 ~ 1e-4..1e-3), and with a common rho the state
  // box / terminal equality rows converge 10-200x slower (oracle experiment recorded in DESIGN.md section 3).
  D.rho_vec.assign(D.nt, D.rho);
  for (int i = 0; i < mg; i++) {
    double rn = 0.0;
    for (int j = 0; j < nz; j++) rn += D.G(i, j) * D.G(i, j);
    D.rho_vec[nz + i] = (D.is_eq[nz + i] ? s.rho_eq_scale * D.rho : ineq_scale * D.rho) / std::max(rn, 1e-12);
  }
  if (D.nball) {      // the projection onto a ball is closed-form only for one common step size on its rows: rho / mean |G_i|^2
    double mean = 0.0;
    for (int i = 0; i < D.nball; i++) { double rn = 0.0; for (int j = 0; j < nz; j++) rn += D.G(i, j) * D.G(i, j); mean += rn / D.nball; }
    for (int i = 0; i < D.nball; i++) D.rho_vec[nz + i] = D.rho / std::max(mean, 1e-12);
  }
  // K = Pc + (sigma + rho) I + G' diag(rho_g) G ;  T = [I;G] K^-1 [I,G'] ;  C = [[Pc,G'],[G,0]]
  Mat K = D.Pc;
  for (int i = 0; i < nz; i++) K(i, i) += s.sigma + D.rho;
  if (mg) {
    Mat RG = D.G;
    for (int j = 0; j < nz; j++)
      for (int i = 0; i < mg; i++) RG(i, j) *= D.rho_vec[nz + i];
    K = add(K, matmul_tn(D.G, RG));
  }
  Mat Kinv;
  if (!spd_inverse(K, Kinv)) { err = "K = Pc + (sigma+rho) I + G' rho G is not positive definite"; return MPCB_ERR_NUMERIC; }
  D.Ac = Mat(D.nt, nz);
  for (int i = 0; i < nz; i++) D.Ac(i, i) = 1.0;
  for (int j = 0; j < nz; j++)
    for (int i = 0; i < mg; i++) D.Ac(nz + i, j) = D.G(i, j);
  D.T = matmul(matmul(D.Ac, Kinv), transpose(D.Ac));
  for (int j = 0; j < D.nt; j++)
    for (int i = 0; i < j; i++) { const double v = 0.5 * (D.T(i, j) + D.T(j, i)); D.T(i, j) = v; D.T(j, i) = v; }
  D.C = Mat(D.nt, D.nt);
  for (int j = 0; j < nz; j++) {
    for (int i = 0; i < nz; i++) D.C(i, j) = D.Pc(i, j);
    for (int i = 0; i < mg; i++) { D.C(nz + i, j) = D.G(i, j); D.C(j, nz + i) = D.G(i, j); }
  }
  if (nx <= 8 && nu <= 4) riccati_factors(D, s.sigma);      // stage-wise form of the same x-update (admm_riccati.cu), where it applies
  // cold-start map (settings.cold_init): v_unc(p) = Lv p = -Pc^-1 Lq p, the unconstrained optimum
  D.Lv = Mat(0, 0);
  if (s.cold_init) {
    Mat Pinv;
    if (!spd_inverse(D.Pc, Pinv)) { err = "condensed Hessian is not positive definite"; return MPCB_ERR_NUMERIC; }
    D.Lv = matmul(Pinv, D.Lq);
    for (double& v : D.Lv.a) v = -v;
  }
  return MPCB_OK;
}

}  // namespace mpcb
