/*
 * osqp_ref.c -- CPU restatement of the OSQP 0.6 ADMM solver.  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED (see oracle/mpc_oracle.py header): the reference reaches OSQP through
 *   JuMP.optimize!(C.tuning.modeler)              /root/reference/src/main/computation_mpc.jl:41
 *   JuMP.Model(optimizer_with_attributes(OSQP.Optimizer))   /root/reference/src/sub/solver_selection.jl:94
 * i.e. libosqp 0.6.x (OSQP.jl "0.8", Project.toml:28) with library defaults; libosqp is not vendored and not
 * installed here, so this file restates the published algorithm (Stellato, Banjac, Goulart, Bemporad, Boyd,
 * "OSQP: an operator splitting solver for quadratic programs", Math. Prog. Comp. 2020) with the 0.6 defaults:
 *   rho=0.1 (x1e3 on equality rows, 1e-6 on unbounded rows), sigma=1e-6, alpha=1.6, eps_abs=eps_rel=1e-3,
 *   eps_prim_inf=eps_dual_inf=1e-4, max_iter=4000, Ruiz scaling x10, check_termination=25, adaptive_rho on
 *   (tolerance 5), polish off, warm_start on, scaled_termination off.
 * Deviation, stated: OSQP picks adaptive_rho_interval from wall-clock timing (non-deterministic); this port uses
 * the library's own no-profiling fallback, ADAPTIVE_RHO_MULTIPLE_TERMINATION(4) x check_termination = 100.
 * The KKT system is factored with an up-looking sparse LDL' (the algorithm QDLDL implements, T. Davis "LDL")
 * under a fill-reducing permutation supplied by the caller.
 *
 * Build:  gcc -O3 -march=native -fopenmp -shared -fPIC oracle/osqp_ref.c -o oracle/libosqp_ref.so -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OSQP_INFTY 1e30
#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4
#define RHO_MIN 1e-6
#define RHO_MAX 1e6
#define RHO_TOL 1e-4
#define RHO_EQ_OVER_RHO_INEQ 1e3

typedef struct {
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance;
  int max_iter, scaling, check_termination, adaptive_rho, adaptive_rho_interval, warm_start, reset_rho_each_solve;
} osqp_ref_settings;

void osqp_ref_default_settings(osqp_ref_settings *s) {
  s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6; s->eps_abs = 1e-3; s->eps_rel = 1e-3;
  s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4; s->adaptive_rho_tolerance = 5.0;
  s->max_iter = 4000; s->scaling = 10; s->check_termination = 25; s->adaptive_rho = 1;
  s->adaptive_rho_interval = 100; s->warm_start = 1; s->reset_rho_each_solve = 0;
}

typedef struct {
  int n, m;
  /* scaled problem data (CSC); P holds the upper triangle only */
  int *Pp, *Pi; double *Px; int *Ap, *Ai; double *Ax;
  double *q, *l, *u;                 /* scaled */
  double *D, *E, *Dinv, *Einv; double c, cinv;
  /* rho */
  double rho; double *rho_vec, *rho_inv_vec; int *constr_type; /* -1 loose, 0 ineq, 1 eq */
  /* KKT (permuted, upper-tri CSC) + LDL' */
  int N; int *perm, *pinv; int *Kp, *Ki; double *Kx; int *rho_pos; /* position of -1/rho_i diag in Kx */
  int *etree, *Lnz, *Lp, *Li; double *Lx, *Dg, *Dginv; int *iw, *flag, *pattern; double *yw, *bw;
  /* iterates */
  double *x, *y, *z, *xz_tilde, *x_prev, *z_prev, *delta_x, *delta_y, *Ax_, *Px_, *Aty, *tmpn, *tmpm;
  osqp_ref_settings st;
} osqp_ref_work;

static double vmaxabs(const double *v, int n) { double r = 0; for (int i = 0; i < n; i++) { double a = fabs(v[i]); if (a > r) r = a; } return r; }
static double limit_scaling(double v) { v = v < MIN_SCALING ? 1.0 : v; return v > MAX_SCALING ? MAX_SCALING : v; }

/* y = A x (CSC) */
static void csc_mv(int m, int n, const int *p, const int *i, const double *x_, const double *v, double *y) {
  memset(y, 0, sizeof(double) * m);
  for (int j = 0; j < n; j++) { double vj = v[j]; if (vj == 0) continue; for (int k = p[j]; k < p[j + 1]; k++) y[i[k]] += x_[k] * vj; }
}
/* y = A' x */
static void csc_tmv(int n, const int *p, const int *i, const double *x_, const double *v, double *y) {
  for (int j = 0; j < n; j++) { double s = 0; for (int k = p[j]; k < p[j + 1]; k++) s += x_[k] * v[i[k]]; y[j] = s; }
}
/* y = P x, P symmetric stored as upper triangle */
static void sym_mv(int n, const int *p, const int *i, const double *x_, const double *v, double *y) {
  memset(y, 0, sizeof(double) * n);
  for (int j = 0; j < n; j++) for (int k = p[j]; k < p[j + 1]; k++) {
    int r = i[k]; y[r] += x_[k] * v[j]; if (r != j) y[j] += x_[k] * v[r];
  }
}

/* ---------------- Ruiz equilibration (osqp scaling.c: scale_data) ---------------- */
static void scale_data(osqp_ref_work *w) {
  int n = w->n, m = w->m;
  for (int i = 0; i < n; i++) { w->D[i] = 1; w->Dinv[i] = 1; }
  for (int i = 0; i < m; i++) { w->E[i] = 1; w->Einv[i] = 1; }
  w->c = 1; w->cinv = 1;
  double *Dt = w->tmpn, *Et = w->tmpm;
  for (int it = 0; it < w->st.scaling; it++) {
    /* column inf-norms of KKT = [P A'; A 0] */
    for (int j = 0; j < n; j++) Dt[j] = 0;
    for (int j = 0; j < n; j++) for (int k = w->Pp[j]; k < w->Pp[j + 1]; k++) {
      double a = fabs(w->Px[k]); int r = w->Pi[k];
      if (a > Dt[j]) Dt[j] = a; if (a > Dt[r]) Dt[r] = a;
    }
    for (int i = 0; i < m; i++) Et[i] = 0;
    for (int j = 0; j < n; j++) for (int k = w->Ap[j]; k < w->Ap[j + 1]; k++) {
      double a = fabs(w->Ax[k]); if (a > Dt[j]) Dt[j] = a; if (a > Et[w->Ai[k]]) Et[w->Ai[k]] = a;
    }
    for (int j = 0; j < n; j++) Dt[j] = 1.0 / sqrt(limit_scaling(Dt[j]));
    for (int i = 0; i < m; i++) Et[i] = 1.0 / sqrt(limit_scaling(Et[i]));
    for (int j = 0; j < n; j++) for (int k = w->Pp[j]; k < w->Pp[j + 1]; k++) w->Px[k] *= Dt[j] * Dt[w->Pi[k]];
    for (int j = 0; j < n; j++) for (int k = w->Ap[j]; k < w->Ap[j + 1]; k++) w->Ax[k] *= Dt[j] * Et[w->Ai[k]];
    for (int j = 0; j < n; j++) { w->q[j] *= Dt[j]; w->D[j] *= Dt[j]; }
    for (int i = 0; i < m; i++) w->E[i] *= Et[i];
    /* cost normalisation */
    for (int j = 0; j < n; j++) Dt[j] = 0;
    for (int j = 0; j < n; j++) for (int k = w->Pp[j]; k < w->Pp[j + 1]; k++) {
      double a = fabs(w->Px[k]); int r = w->Pi[k];
      if (a > Dt[j]) Dt[j] = a; if (a > Dt[r]) Dt[r] = a;
    }
    double mean = 0; for (int j = 0; j < n; j++) mean += Dt[j]; mean /= n;
    double nq = limit_scaling(vmaxabs(w->q, n));
    double ct = mean > nq ? mean : nq; ct = 1.0 / limit_scaling(ct);
    for (int k = 0; k < w->Pp[n]; k++) w->Px[k] *= ct;
    for (int j = 0; j < n; j++) w->q[j] *= ct;
    w->c *= ct;
  }
  w->cinv = 1.0 / w->c;
  for (int j = 0; j < n; j++) w->Dinv[j] = 1.0 / w->D[j];
  for (int i = 0; i < m; i++) { w->Einv[i] = 1.0 / w->E[i]; w->l[i] *= w->E[i]; w->u[i] *= w->E[i]; }
}

/* ---------------- rho vector (osqp auxil.c: set_rho_vec) ---------------- */
static void set_rho_vec(osqp_ref_work *w) {
  double rho = w->rho; rho = rho < RHO_MIN ? RHO_MIN : (rho > RHO_MAX ? RHO_MAX : rho); w->rho = rho;
  for (int i = 0; i < w->m; i++) {
    if (w->l[i] < -OSQP_INFTY * MIN_SCALING && w->u[i] > OSQP_INFTY * MIN_SCALING) { w->constr_type[i] = -1; w->rho_vec[i] = RHO_MIN; }
    else if (w->u[i] - w->l[i] < RHO_TOL) { w->constr_type[i] = 1; w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * rho; }
    else { w->constr_type[i] = 0; w->rho_vec[i] = rho; }
    w->rho_inv_vec[i] = 1.0 / w->rho_vec[i];
  }
}

/* ---------------- KKT assembly + sparse LDL' ---------------- */
typedef struct { int r, c; double v; int tag; } trip;
static int trip_cmp(const void *a, const void *b) {
  const trip *x = (const trip *)a, *y = (const trip *)b;
  if (x->c != y->c) return x->c - y->c; return x->r - y->r;
}
static int build_kkt(osqp_ref_work *w) {
  int n = w->n, m = w->m, N = n + m;
  int nnzmax = w->Pp[n] + n + w->Ap[n] + m;
  trip *t = (trip *)malloc(sizeof(trip) * nnzmax); int nt = 0;
  char *hasdiag = (char *)calloc(n, 1);
  for (int j = 0; j < n; j++) for (int k = w->Pp[j]; k < w->Pp[j + 1]; k++) {
    int r = w->Pi[k]; double v = w->Px[k]; if (r == j) { v += w->st.sigma; hasdiag[j] = 1; }
    t[nt].r = r; t[nt].c = j; t[nt].v = v; t[nt].tag = -1; nt++;
  }
  for (int j = 0; j < n; j++) if (!hasdiag[j]) { t[nt].r = j; t[nt].c = j; t[nt].v = w->st.sigma; t[nt].tag = -1; nt++; }
  free(hasdiag);
  for (int j = 0; j < n; j++) for (int k = w->Ap[j]; k < w->Ap[j + 1]; k++) { t[nt].r = j; t[nt].c = n + w->Ai[k]; t[nt].v = w->Ax[k]; t[nt].tag = -1; nt++; }
  for (int i = 0; i < m; i++) { t[nt].r = n + i; t[nt].c = n + i; t[nt].v = -w->rho_inv_vec[i]; t[nt].tag = i; nt++; }
  for (int k = 0; k < nt; k++) {
    int r = w->pinv[t[k].r], c = w->pinv[t[k].c]; if (r > c) { int s = r; r = c; c = s; }
    t[k].r = r; t[k].c = c;
  }
  qsort(t, nt, sizeof(trip), trip_cmp);
  w->N = N; w->Kp = (int *)calloc(N + 1, sizeof(int)); w->Ki = (int *)malloc(sizeof(int) * nt); w->Kx = (double *)malloc(sizeof(double) * nt);
  w->rho_pos = (int *)malloc(sizeof(int) * (m > 0 ? m : 1));
  for (int k = 0; k < nt; k++) { w->Kp[t[k].c + 1]++; w->Ki[k] = t[k].r; w->Kx[k] = t[k].v; if (t[k].tag >= 0) w->rho_pos[t[k].tag] = k; }
  for (int j = 0; j < N; j++) w->Kp[j + 1] += w->Kp[j];
  free(t);
  /* symbolic: elimination tree + column counts */
  w->etree = (int *)malloc(sizeof(int) * N); w->Lnz = (int *)malloc(sizeof(int) * N); w->Lp = (int *)malloc(sizeof(int) * (N + 1));
  w->iw = (int *)malloc(sizeof(int) * N); w->flag = (int *)malloc(sizeof(int) * N); w->pattern = (int *)malloc(sizeof(int) * N);
  w->yw = (double *)calloc(N, sizeof(double)); w->bw = (double *)malloc(sizeof(double) * N);
  w->Dg = (double *)malloc(sizeof(double) * N); w->Dginv = (double *)malloc(sizeof(double) * N);
  for (int i = 0; i < N; i++) { w->etree[i] = -1; w->Lnz[i] = 0; w->flag[i] = -1; }
  for (int j = 0; j < N; j++) {
    w->flag[j] = j;
    for (int k = w->Kp[j]; k < w->Kp[j + 1]; k++) {
      int i = w->Ki[k];
      while (i < j && w->flag[i] != j) { if (w->etree[i] == -1) w->etree[i] = j; w->Lnz[i]++; w->flag[i] = j; i = w->etree[i]; }
    }
  }
  w->Lp[0] = 0; for (int i = 0; i < N; i++) w->Lp[i + 1] = w->Lp[i] + w->Lnz[i];
  w->Li = (int *)malloc(sizeof(int) * (w->Lp[N] > 0 ? w->Lp[N] : 1)); w->Lx = (double *)malloc(sizeof(double) * (w->Lp[N] > 0 ? w->Lp[N] : 1));
  return 0;
}
static int factor_kkt(osqp_ref_work *w) {
  int N = w->N; int *Lnz = w->iw; double *Y = w->yw;
  for (int k = 0; k < N; k++) {
    int top = N; w->flag[k] = k; Lnz[k] = 0; Y[k] = 0;
    for (int p = w->Kp[k]; p < w->Kp[k + 1]; p++) {
      int i = w->Ki[p]; Y[i] += w->Kx[p];
      int len = 0;
      for (; w->flag[i] != k; i = w->etree[i]) { w->pattern[len++] = i; w->flag[i] = k; }
      while (len > 0) w->pattern[--top] = w->pattern[--len];
    }
    double dk = Y[k]; Y[k] = 0;
    for (; top < N; top++) {
      int i = w->pattern[top]; double yi = Y[i]; Y[i] = 0;
      int p2 = w->Lp[i] + Lnz[i], p;
      for (p = w->Lp[i]; p < p2; p++) Y[w->Li[p]] -= w->Lx[p] * yi;
      double lki = yi * w->Dginv[i]; dk -= lki * yi; w->Li[p] = k; w->Lx[p] = lki; Lnz[i]++;
    }
    if (dk == 0.0) return -1;
    w->Dg[k] = dk; w->Dginv[k] = 1.0 / dk;
  }
  return 0;
}
static void solve_kkt(osqp_ref_work *w, double *b) { /* b (length N) overwritten with solution */
  int N = w->N; double *x = w->bw;
  for (int i = 0; i < N; i++) x[i] = b[w->perm[i]];
  for (int j = 0; j < N; j++) { double xj = x[j]; for (int p = w->Lp[j]; p < w->Lp[j + 1]; p++) x[w->Li[p]] -= w->Lx[p] * xj; }
  for (int j = 0; j < N; j++) x[j] *= w->Dginv[j];
  for (int j = N - 1; j >= 0; j--) { double s = x[j]; for (int p = w->Lp[j]; p < w->Lp[j + 1]; p++) s -= w->Lx[p] * x[w->Li[p]]; x[j] = s; }
  for (int i = 0; i < N; i++) b[w->perm[i]] = x[i];
}
static int update_rho(osqp_ref_work *w, double rho_new) {
  w->rho = rho_new; set_rho_vec(w);
  for (int i = 0; i < w->m; i++) w->Kx[w->rho_pos[i]] = -w->rho_inv_vec[i];
  return factor_kkt(w);
}

/* ---------------- setup / cleanup ---------------- */
static double *dalloc(int n) { return (double *)calloc(n > 0 ? n : 1, sizeof(double)); }
osqp_ref_work *osqp_ref_setup(int n, int m, const int *Pp, const int *Pi, const double *Px, const double *q,
                              const int *Ap, const int *Ai, const double *Ax, const double *l, const double *u,
                              const int *perm, const osqp_ref_settings *st) {
  osqp_ref_work *w = (osqp_ref_work *)calloc(1, sizeof(osqp_ref_work));
  w->n = n; w->m = m; w->st = *st;
  int pnz = Pp[n], anz = Ap[n];
  w->Pp = (int *)malloc(sizeof(int) * (n + 1)); memcpy(w->Pp, Pp, sizeof(int) * (n + 1));
  w->Pi = (int *)malloc(sizeof(int) * (pnz > 0 ? pnz : 1)); memcpy(w->Pi, Pi, sizeof(int) * pnz);
  w->Px = dalloc(pnz); memcpy(w->Px, Px, sizeof(double) * pnz);
  w->Ap = (int *)malloc(sizeof(int) * (n + 1)); memcpy(w->Ap, Ap, sizeof(int) * (n + 1));
  w->Ai = (int *)malloc(sizeof(int) * (anz > 0 ? anz : 1)); memcpy(w->Ai, Ai, sizeof(int) * anz);
  w->Ax = dalloc(anz); memcpy(w->Ax, Ax, sizeof(double) * anz);
  w->q = dalloc(n); memcpy(w->q, q, sizeof(double) * n);
  w->l = dalloc(m); w->u = dalloc(m);
  for (int i = 0; i < m; i++) { w->l[i] = l[i] < -OSQP_INFTY ? -OSQP_INFTY : l[i]; w->u[i] = u[i] > OSQP_INFTY ? OSQP_INFTY : u[i]; }
  w->D = dalloc(n); w->Dinv = dalloc(n); w->E = dalloc(m); w->Einv = dalloc(m);
  w->rho_vec = dalloc(m); w->rho_inv_vec = dalloc(m); w->constr_type = (int *)calloc(m > 0 ? m : 1, sizeof(int));
  w->x = dalloc(n); w->y = dalloc(m); w->z = dalloc(m); w->xz_tilde = dalloc(n + m); w->x_prev = dalloc(n); w->z_prev = dalloc(m);
  w->delta_x = dalloc(n); w->delta_y = dalloc(m); w->Ax_ = dalloc(m); w->Px_ = dalloc(n); w->Aty = dalloc(n); w->tmpn = dalloc(n); w->tmpm = dalloc(m);
  if (st->scaling > 0) scale_data(w); else { for (int i = 0; i < n; i++) { w->D[i] = w->Dinv[i] = 1; } for (int i = 0; i < m; i++) { w->E[i] = w->Einv[i] = 1; } w->c = w->cinv = 1; }
  w->rho = st->rho; set_rho_vec(w);
  w->perm = (int *)malloc(sizeof(int) * (n + m)); w->pinv = (int *)malloc(sizeof(int) * (n + m));
  for (int i = 0; i < n + m; i++) { w->perm[i] = perm ? perm[i] : i; w->pinv[w->perm[i]] = i; }
  build_kkt(w);
  if (factor_kkt(w) != 0) return NULL;
  return w;
}
void osqp_ref_cleanup(osqp_ref_work *w) {
  if (!w) return;
  free(w->Pp); free(w->Pi); free(w->Px); free(w->Ap); free(w->Ai); free(w->Ax); free(w->q); free(w->l); free(w->u);
  free(w->D); free(w->E); free(w->Dinv); free(w->Einv); free(w->rho_vec); free(w->rho_inv_vec); free(w->constr_type);
  free(w->perm); free(w->pinv); free(w->Kp); free(w->Ki); free(w->Kx); free(w->rho_pos); free(w->etree); free(w->Lnz); free(w->Lp);
  free(w->Li); free(w->Lx); free(w->Dg); free(w->Dginv); free(w->iw); free(w->flag); free(w->pattern); free(w->yw); free(w->bw);
  free(w->x); free(w->y); free(w->z); free(w->xz_tilde); free(w->x_prev); free(w->z_prev); free(w->delta_x); free(w->delta_y);
  free(w->Ax_); free(w->Px_); free(w->Aty); free(w->tmpn); free(w->tmpm); free(w);
}
/* osqp_update_bounds for a subset of rows (unscaled values in) */
void osqp_ref_update_bounds(osqp_ref_work *w, int k, const int *rows, const double *lnew, const double *unew) {
  for (int j = 0; j < k; j++) { int i = rows[j]; w->l[i] = lnew[j] * w->E[i]; w->u[i] = unew[j] * w->E[i]; }
}
int osqp_ref_kkt_nnz_L(const osqp_ref_work *w) { return w->Lp[w->N]; }
double osqp_ref_rho(const osqp_ref_work *w) { return w->rho; }
/* expose one KKT solve for the tests (b has length n+m, overwritten) */
void osqp_ref_kkt_solve(osqp_ref_work *w, double *b) { solve_kkt(w, b); }

/* ---------------- residuals / termination (osqp auxil.c) ---------------- */
typedef struct { double pri_res, dua_res, obj; } info_t;
static void compute_Ax_Px_Aty(osqp_ref_work *w) {
  csc_mv(w->m, w->n, w->Ap, w->Ai, w->Ax, w->x, w->Ax_);
  sym_mv(w->n, w->Pp, w->Pi, w->Px, w->x, w->Px_);
  csc_tmv(w->n, w->Ap, w->Ai, w->Ax, w->y, w->Aty);
}
static double scaled_norm(const double *s, const double *v, int n) { double r = 0; for (int i = 0; i < n; i++) { double a = fabs(s[i] * v[i]); if (a > r) r = a; } return r; }
static void update_info(osqp_ref_work *w, info_t *inf) {
  int n = w->n, m = w->m;
  compute_Ax_Px_Aty(w);
  double pr = 0; for (int i = 0; i < m; i++) { double a = fabs(w->Einv[i] * (w->Ax_[i] - w->z[i])); if (a > pr) pr = a; }
  double dr = 0; for (int j = 0; j < n; j++) { double a = fabs(w->Dinv[j] * (w->Px_[j] + w->q[j] + w->Aty[j])); if (a > dr) dr = a; }
  inf->pri_res = pr; inf->dua_res = dr * w->cinv;
  double o = 0; for (int j = 0; j < n; j++) o += w->x[j] * (0.5 * w->Px_[j] + w->q[j]);
  inf->obj = o * w->cinv;
}
static int check_termination(osqp_ref_work *w, const info_t *inf, int approximate) {
  int n = w->n, m = w->m; const osqp_ref_settings *s = &w->st;
  double ea = s->eps_abs, er = s->eps_rel, epi = s->eps_prim_inf, edi = s->eps_dual_inf;
  if (approximate) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
  int prim_ok = 0, dual_ok = 0, prim_inf = 0, dual_inf = 0;
  if (m == 0) prim_ok = 1;
  else {
    double na = scaled_norm(w->Einv, w->Ax_, m), nz = scaled_norm(w->Einv, w->z, m);
    double eps_prim = ea + er * (na > nz ? na : nz);
    if (inf->pri_res < eps_prim) prim_ok = 1;
    else {
      /* project delta_y onto the polar of the recession cone of [l,u] (osqp auxil.c: is_primal_infeasible) */
      for (int i = 0; i < m; i++) {
        if (w->u[i] > OSQP_INFTY * MIN_SCALING) {
          if (w->l[i] < -OSQP_INFTY * MIN_SCALING) w->delta_y[i] = 0.0;
          else if (w->delta_y[i] > 0.0) w->delta_y[i] = 0.0;
        } else if (w->l[i] < -OSQP_INFTY * MIN_SCALING) { if (w->delta_y[i] < 0.0) w->delta_y[i] = 0.0; }
      }
      double ndy = scaled_norm(w->E, w->delta_y, m);
      if (ndy > epi) {
        double lhs = 0;
        for (int i = 0; i < m; i++) {
          double dy = w->delta_y[i];
          if (dy > 0) lhs += w->u[i] * dy; else if (dy < 0) lhs += w->l[i] * dy;
        }
        if (lhs < -epi * ndy) {
          csc_tmv(n, w->Ap, w->Ai, w->Ax, w->delta_y, w->tmpn);
          if (scaled_norm(w->Dinv, w->tmpn, n) < epi * ndy) prim_inf = 1;
        }
      }
    }
  }
  {
    double a = scaled_norm(w->Dinv, w->Px_, n), b = scaled_norm(w->Dinv, w->Aty, n), c = scaled_norm(w->Dinv, w->q, n);
    double mx = a > b ? a : b; mx = mx > c ? mx : c;
    double eps_dual = ea + er * mx * w->cinv;
    if (inf->dua_res < eps_dual) dual_ok = 1;
    else {
      double ndx = scaled_norm(w->D, w->delta_x, n);
      if (ndx > edi) {
        double qdx = 0; for (int j = 0; j < n; j++) qdx += w->q[j] * w->delta_x[j];
        if (qdx < -w->c * edi * ndx) {
          sym_mv(n, w->Pp, w->Pi, w->Px, w->delta_x, w->tmpn);
          if (scaled_norm(w->Dinv, w->tmpn, n) < w->c * edi * ndx) {
            csc_mv(m, n, w->Ap, w->Ai, w->Ax, w->delta_x, w->tmpm);
            int ok = 1;
            for (int i = 0; i < m && ok; i++) {
              double v = w->Einv[i] * w->tmpm[i];
              if ((w->u[i] < OSQP_INFTY * MIN_SCALING && v > edi * ndx) || (w->l[i] > -OSQP_INFTY * MIN_SCALING && v < -edi * ndx)) ok = 0;
            }
            dual_inf = ok;
          }
        }
      }
    }
  }
  if (prim_ok && dual_ok) return approximate ? 2 : 1;
  if (prim_inf) return approximate ? 3 : -3;
  if (dual_inf) return approximate ? 4 : -4;
  return 0;
}
static double compute_rho_estimate(osqp_ref_work *w) {
  int n = w->n, m = w->m;
  compute_Ax_Px_Aty(w);
  double pr = 0; for (int i = 0; i < m; i++) { double a = fabs(w->Ax_[i] - w->z[i]); if (a > pr) pr = a; }
  double dr = 0; for (int j = 0; j < n; j++) { double a = fabs(w->Px_[j] + w->q[j] + w->Aty[j]); if (a > dr) dr = a; }
  double a = vmaxabs(w->z, m), b = vmaxabs(w->Ax_, m); pr /= ((a > b ? a : b) + 1e-10);
  a = vmaxabs(w->q, n); b = vmaxabs(w->Aty, n); double c = vmaxabs(w->Px_, n); a = a > b ? a : b; a = a > c ? a : c; dr /= (a + 1e-10);
  double r = w->rho * sqrt(pr / (dr + 1e-10));
  return r < RHO_MIN ? RHO_MIN : (r > RHO_MAX ? RHO_MAX : r);
}

/* ---------------- osqp_solve ---------------- */
/* status codes = OSQP's: 1 solved, 2 solved inaccurate, -2 max iter, -3 primal infeasible, 3 prim. inf. inaccurate,
 * -4 dual infeasible, 4 dual inf. inaccurate */
int osqp_ref_solve(osqp_ref_work *w, int cold_start, double *x_out, double *y_out, int *iters_out,
                   double *pri_out, double *dua_out, double *obj_out, int *rho_updates_out) {
  int n = w->n, m = w->m; const osqp_ref_settings *s = &w->st;
  if (s->reset_rho_each_solve && w->rho != s->rho) { if (update_rho(w, s->rho)) return -10; }
  if (cold_start || !s->warm_start) { memset(w->x, 0, sizeof(double) * n); memset(w->z, 0, sizeof(double) * m); memset(w->y, 0, sizeof(double) * m); }
  int status = 0, iter, rho_updates = 0; info_t inf = {0, 0, 0};
  for (iter = 1; iter <= s->max_iter; iter++) {
    memcpy(w->x_prev, w->x, sizeof(double) * n); memcpy(w->z_prev, w->z, sizeof(double) * m);
    /* update_xz_tilde */
    for (int j = 0; j < n; j++) w->xz_tilde[j] = s->sigma * w->x_prev[j] - w->q[j];
    for (int i = 0; i < m; i++) w->xz_tilde[n + i] = w->z_prev[i] - w->rho_inv_vec[i] * w->y[i];
    solve_kkt(w, w->xz_tilde);
    for (int i = 0; i < m; i++) w->xz_tilde[n + i] = w->z_prev[i] + w->rho_inv_vec[i] * (w->xz_tilde[n + i] - w->y[i]);
    /* update_x */
    for (int j = 0; j < n; j++) { w->x[j] = s->alpha * w->xz_tilde[j] + (1.0 - s->alpha) * w->x_prev[j]; w->delta_x[j] = w->x[j] - w->x_prev[j]; }
    /* update_z */
    for (int i = 0; i < m; i++) {
      double v = s->alpha * w->xz_tilde[n + i] + (1.0 - s->alpha) * w->z_prev[i] + w->rho_inv_vec[i] * w->y[i];
      w->z[i] = v < w->l[i] ? w->l[i] : (v > w->u[i] ? w->u[i] : v);
    }
    /* update_y */
    for (int i = 0; i < m; i++) {
      double dy = w->rho_vec[i] * (s->alpha * w->xz_tilde[n + i] + (1.0 - s->alpha) * w->z_prev[i] - w->z[i]);
      w->delta_y[i] = dy; w->y[i] += dy;
    }
    int can_check = s->check_termination && (iter % s->check_termination == 0);
    if (can_check) { update_info(w, &inf); status = check_termination(w, &inf, 0); if (status) break; }
    if (s->adaptive_rho && s->adaptive_rho_interval && (iter % s->adaptive_rho_interval == 0)) {
      double rn = compute_rho_estimate(w);
      if (rn > w->rho * s->adaptive_rho_tolerance || rn < w->rho / s->adaptive_rho_tolerance) { if (update_rho(w, rn)) return -10; rho_updates++; }
    }
  }
  if (!status) {
    if (iter > s->max_iter) iter = s->max_iter;
    update_info(w, &inf); status = check_termination(w, &inf, 0);
    if (!status) status = check_termination(w, &inf, 1);
    if (!status) status = -2;
  }
  /* unscale solution (osqp scaling.c: unscale_solution) */
  if (x_out) for (int j = 0; j < n; j++) x_out[j] = w->D[j] * w->x[j];
  if (y_out) for (int i = 0; i < m; i++) y_out[i] = w->cinv * w->E[i] * w->y[i];
  if (iters_out) *iters_out = iter; if (pri_out) *pri_out = inf.pri_res; if (dua_out) *dua_out = inf.dua_res;
  if (obj_out) *obj_out = inf.obj; if (rho_updates_out) *rho_updates_out = rho_updates;
  return status;
}

/* ---------------- batch driver: one workspace ("JuMP model") per thread, as the reference would run it -------------- */
/* upd_rows[k]: rows whose l=u are replaced per problem by upd_vals[b*k + j] (the JuMP.fix of x[:,1] and of the references).
 * sel_cols[ns]: columns of the primal solution copied out per problem (x_sel is nb x ns).  Returns 0 or <0. */
int osqp_ref_solve_batch(int n, int m, const int *Pp, const int *Pi, const double *Px, const double *q,
                         const int *Ap, const int *Ai, const double *Ax, const double *l, const double *u,
                         const int *perm, const osqp_ref_settings *st, int nb, int k, const int *upd_rows,
                         const double *upd_vals, int cold_start, int ns, const int *sel_cols, double *x_sel,
                         int *status, int *iters, double *pri, double *dua, double *obj, int nthreads) {
  int err = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    osqp_ref_work *w = osqp_ref_setup(n, m, Pp, Pi, Px, q, Ap, Ai, Ax, l, u, perm, st);
    double *xf = (double *)malloc(sizeof(double) * n);
    if (!w) {
#pragma omp atomic write
      err = -1;
    } else {
#pragma omp for schedule(dynamic, 16)
      for (int b = 0; b < nb; b++) {
        if (k > 0) osqp_ref_update_bounds(w, k, upd_rows, upd_vals + (size_t)b * k, upd_vals + (size_t)b * k);
        int it = 0; double pr = 0, du = 0, ob = 0;
        int s = osqp_ref_solve(w, cold_start, xf, NULL, &it, &pr, &du, &ob, NULL);
        status[b] = s; iters[b] = it; if (pri) pri[b] = pr; if (dua) dua[b] = du; if (obj) obj[b] = ob;
        for (int j = 0; j < ns; j++) x_sel[(size_t)b * ns + j] = xf[sel_cols[j]];
      }
    }
    free(xf); osqp_ref_cleanup(w);
  }
  return err;
}
int osqp_ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
