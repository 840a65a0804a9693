/* ORACLE -- test infrastructure, never shipped, never on the product path.
 *
 * Plain-C CPU restatement of the reference's multiple-shooting NLP and of a CPU solve of it.
 *   NLP definition      src/Mahi/Mpc/ModelGenerator.cpp:23-233   (V, p, g, J layouts; Euler step :33-34;
 *                                                                 linearised step :45-48,58)
 *   run-time marshalling src/Mahi/Mpc/ModelControl.cpp:116-172    (p assembly, x0 pinning, bounds)
 *   solver options       src/Mahi/Mpc/ModelControl.cpp:54-59      (tol 1e-5, max_iter 200)
 *
 * The reference solves with CasADi-generated callbacks + IPOPT 3.14.3 + MUMPS; none of the three is
 * in /root/reference or installable here, so PARITY IS UNPINNED against the true
 * reference solver.  This file restates the same KKT conditions and solves them with a full-space
 * Newton / primal-dual interior-point iteration on the *interleaved banded KKT matrix* with a
 * partial-pivoting band LU -- deliberately a different linear-algebra route from the GPU's Riccati
 * sweep.  Model callbacks are sympy-generated (oracle/gen_c.py -> oracle/gen/<model>_gen.c).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "gen/double_pendulum_gen.c"
#include "arm4_jets.c" /* builder-defined 4-DOF arm (BASELINE config 3): second-order forward-mode AD, no generated code */

#define MAXNX 8
#define MAXNU 4
#define MAXS (MAXNX + MAXNU)
#define INF_THRESH 1e19 /* IPOPT nlp_*_bound_inf [upstream] */

typedef struct {
  const char* name;
  int nx, nu;
  void (*f)(const double*, const double*, double*);
  void (*fAB)(const double*, const double*, double*, double*, double*);
  void (*fABH)(const double*, const double*, const double*, double*, double*, double*, double*);
} omodel_t;

static const omodel_t g_models[] = {
    {"double_pendulum", 4, 2, double_pendulum_f, double_pendulum_fAB, double_pendulum_fABH},
    {"arm4", 8, 4, arm4_f, arm4_fAB, arm4_fABH},
};

int oracle_model_index(const char* name) {
  for (unsigned i = 0; i < sizeof(g_models) / sizeof(g_models[0]); ++i)
    if (!strcmp(name, g_models[i].name)) return (int)i;
  return -1;
}
int oracle_model_nx(int m) { return g_models[m].nx; }
int oracle_model_nu(int m) { return g_models[m].nu; }

typedef struct {
  const omodel_t* m;
  int N, is_linear, nx, nu, s, NV, M, np;
  double h;
  /* views into p (ModelGenerator.cpp:129-187) */
  const double *r, *Q, *R, *Rm, *linA, *linB, *xdot_init, *x_init, *u_init;
} oprob_t;

static void prob_init(oprob_t* P, int model, int N, double h, int is_linear, const double* p) {
  P->m = &g_models[model];
  P->N = N; P->h = h; P->is_linear = is_linear;
  int nx = P->nx = P->m->nx, nu = P->nu = P->m->nu;
  P->s = nx + nu; P->NV = nx * (N + 1) + nu * N; P->M = N * nx;
  P->np = N * nx + nx + 3 * nu + (is_linear ? nx * nx + nx * nu + 2 * nx : 0);
  if (p) {
    P->r = p; p += N * nx;
    P->Q = p; p += nx; P->R = p; p += nu; P->Rm = p; p += nu;
    if (is_linear) { P->linA = p; p += nx * nx; P->linB = p; p += nx * nu; P->xdot_init = p; p += nx; P->x_init = p; p += nx; }
    P->u_init = p;
  }
}
int oracle_np(int model, int N, int is_linear) { oprob_t P; prob_init(&P, model, N, 0.0, is_linear, 0); return P.np; }

/* one shooting node: F, Fx = dF/dx (row-major nx*nx), Fu (nx*nu), Hs = sum_i w_i d2F_i/dz2 (s*s) */
static void node(const oprob_t* P, const double* x, const double* u, const double* w, int want,
                 double* F, double* Fx, double* Fu, double* Hs) {
  int nx = P->nx, nu = P->nu, s = P->s; double h = P->h;
  double f[MAXNX], A[MAXNX * MAXNX], B[MAXNX * MAXNU], H[MAXS * MAXS];
  if (P->is_linear) { /* ModelGenerator.cpp:47-48: A,B column-major in p */
    for (int i = 0; i < nx; ++i) {
      double acc = P->xdot_init[i];
      for (int j = 0; j < nx; ++j) { A[i * nx + j] = P->linA[j * nx + i]; acc += A[i * nx + j] * (x[j] - P->x_init[j]); }
      for (int j = 0; j < nu; ++j) { B[i * nu + j] = P->linB[j * nx + i]; acc += B[i * nu + j] * (u[j] - P->u_init[j]); }
      f[i] = acc;
    }
    if (want >= 2) memset(H, 0, sizeof(double) * s * s);
  } else if (want == 0) P->m->f(x, u, f);
  else if (want == 1) P->m->fAB(x, u, f, A, B);
  else P->m->fABH(x, u, w, f, A, B, H);
  for (int i = 0; i < nx; ++i) F[i] = x[i] + h * f[i]; /* ModelGenerator.cpp:33 */
  if (want >= 1) {
    for (int i = 0; i < nx; ++i) {
      for (int j = 0; j < nx; ++j) Fx[i * nx + j] = (i == j ? 1.0 : 0.0) + h * A[i * nx + j];
      for (int j = 0; j < nu; ++j) Fu[i * nu + j] = h * B[i * nu + j];
    }
  }
  if (want >= 2) for (int i = 0; i < s * s; ++i) Hs[i] = h * H[i];
}

/* BASELINE config 5 / SURVEY 8a-2,a-5,a-7: n independent node evaluations, AoS inputs.
 * g = F(x,u) - xnext; w may be NULL (then H is not written). */
void oracle_node_eval(int model, int n, double h, const double* x, const double* u, const double* xnext,
                      const double* w, double* g, double* Fx, double* Fu, double* H) {
  oprob_t P; prob_init(&P, model, 1, h, 0, 0);
  int nx = P.nx, nu = P.nu, s = P.s;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    double F[MAXNX];
    node(&P, x + (size_t)i * nx, u + (size_t)i * nu, w ? w + (size_t)i * nx : 0, w ? 2 : 1, F,
         Fx + (size_t)i * nx * nx, Fu + (size_t)i * nx * nu, w ? H + (size_t)i * s * s : 0);
    for (int j = 0; j < nx; ++j) g[(size_t)i * nx + j] = F[j] - xnext[(size_t)i * nx + j];
  }
}

/* <name>_get_x_dot_init / _get_A / _get_B (ModelGenerator.cpp:51-53); A,B returned column-major
 * the way ModelControl.cpp:127-133 flattens the DM into p. */
void oracle_linearise(int model, const double* x, const double* u, double* f, double* A_colmajor, double* B_colmajor) {
  const omodel_t* m = &g_models[model];
  double A[MAXNX * MAXNX], B[MAXNX * MAXNU];
  m->fAB(x, u, f, A, B);
  for (int i = 0; i < m->nx; ++i) {
    for (int j = 0; j < m->nx; ++j) A_colmajor[j * m->nx + i] = A[i * m->nx + j];
    for (int j = 0; j < m->nu; ++j) B_colmajor[j * m->nx + i] = B[i * m->nu + j];
  }
}

/* ---------------------------------------------------------------------------------------------
 * whole-NLP evaluation in stage blocks.
 *   g[M], f, gradf[NV], Fx[N][nx*nx], Fu[N][nx*nu], Hzz[N][s*s] (diagonal stage blocks of
 *   hess_l = lam_f*f + lam_g'g incl. R/Rm terms; the (u_{k},u_{k+1}) cross block is -2*lam_f*R)
 * ------------------------------------------------------------------------------------------- */
static double nlp_eval(const oprob_t* P, const double* V, const double* lam, double lam_f, int want,
                       double* g, double* gradf, double* Fx, double* Fu, double* Hzz) {
  int nx = P->nx, nu = P->nu, s = P->s, N = P->N;
  double fval = 0.0;
  if (gradf) memset(gradf, 0, sizeof(double) * P->NV);
  for (int k = 0; k < N; ++k) {
    const double* x = V + k * s; const double* u = x + nx; const double* xn = V + (k + 1) * s;
    const double* up = k ? u - s : P->u_init;
    double F[MAXNX], e[MAXNX], w[MAXNX], fx[MAXNX * MAXNX], fu[MAXNX * MAXNU], hs[MAXS * MAXS];
    if (want >= 2) { /* need e for w: evaluate F first */
      node(P, x, u, 0, 0, F, 0, 0, 0);
      for (int i = 0; i < nx; ++i) w[i] = 2.0 * lam_f * P->Q[i] * (F[i] - P->r[k * nx + i]) + lam[k * nx + i];
    }
    node(P, x, u, w, want, F, fx, fu, hs);
    for (int i = 0; i < nx; ++i) {
      e[i] = F[i] - P->r[k * nx + i];
      if (g) g[k * nx + i] = F[i] - xn[i];
      fval += P->Q[i] * e[i] * e[i];
    }
    for (int j = 0; j < nu; ++j) { double du = u[j] - up[j]; fval += P->R[j] * du * du + P->Rm[j] * u[j] * u[j]; }
    if (want >= 1) {
      if (Fx) memcpy(Fx + (size_t)k * nx * nx, fx, sizeof(double) * nx * nx);
      if (Fu) memcpy(Fu + (size_t)k * nx * nu, fu, sizeof(double) * nx * nu);
      if (gradf) {
        for (int j = 0; j < nx; ++j) { double a = 0; for (int i = 0; i < nx; ++i) a += fx[i * nx + j] * 2.0 * P->Q[i] * e[i]; gradf[k * s + j] += a; }
        for (int j = 0; j < nu; ++j) {
          double a = 0; for (int i = 0; i < nx; ++i) a += fu[i * nu + j] * 2.0 * P->Q[i] * e[i];
          double du = u[j] - up[j];
          gradf[k * s + nx + j] += a + 2.0 * P->R[j] * du + 2.0 * P->Rm[j] * u[j];
          if (k) gradf[(k - 1) * s + nx + j] -= 2.0 * P->R[j] * du;
        }
      }
    }
    if (want >= 2 && Hzz) {
      double* H = Hzz + (size_t)k * s * s;
      for (int a = 0; a < s; ++a) for (int b = 0; b < s; ++b) {
        double acc = hs[a * s + b];
        for (int i = 0; i < nx; ++i) {
          double fa = a < nx ? fx[i * nx + a] : fu[i * nu + a - nx];
          double fb = b < nx ? fx[i * nx + b] : fu[i * nu + b - nx];
          acc += 2.0 * lam_f * P->Q[i] * fa * fb;
        }
        H[a * s + b] = acc;
      }
      for (int j = 0; j < nu; ++j) H[(nx + j) * s + nx + j] += 2.0 * lam_f * (P->R[j] + P->Rm[j]) + (k < N - 1 ? 2.0 * lam_f * P->R[j] : 0.0);
    }
  }
  return fval;
}

/* exported whole-NLP evaluators (nlp_f / nlp_g / nlp_grad_f / nlp_jac_g / nlp_hess_l in block form) */
double oracle_nlp_eval(int model, int N, double h, int is_linear, const double* V, const double* p,
                       const double* lam_g, double lam_f, double* g, double* gradf, double* Fx, double* Fu, double* Hzz) {
  oprob_t P; prob_init(&P, model, N, h, is_linear, p);
  int want = Hzz ? 2 : (Fx || Fu || gradf) ? 1 : 0;
  return nlp_eval(&P, V, lam_g, lam_f, want, g, gradf, Fx, Fu, Hzz);
}

/* ---------------------------------------------------------------------------------------------
 * band LU with partial pivoting (LAPACK dgbtf2/dgbtrs-style storage: A(i,j) at ab[kl+ku+i-j + j*ld])
 * ------------------------------------------------------------------------------------------- */
static int band_solve(int n, int kl, int ku, double* ab, int ld, int* ipiv, double* b) {
  int kv = kl + ku, ju = 0;
  for (int j = 0; j < n; ++j) {
    int km = kl < n - 1 - j ? kl : n - 1 - j;
    double* col = ab + (size_t)j * ld + kv; /* col[i-j] = A(i,j) for i>=j */
    int jp = 0; double best = fabs(col[0]);
    for (int i = 1; i <= km; ++i) if (fabs(col[i]) > best) { best = fabs(col[i]); jp = i; }
    ipiv[j] = j + jp;
    if (best == 0.0 || !isfinite(best)) return j + 1;
    int t = j + ku + jp; if (t > n - 1) t = n - 1; if (t > ju) ju = t;
    if (jp) for (int c = j; c <= ju; ++c) { /* swap rows j and j+jp */
      double* cc = ab + (size_t)c * ld + kv + j - c;
      double tmp = cc[0]; cc[0] = cc[jp]; cc[jp] = tmp;
    }
    double inv = 1.0 / col[0];
    for (int i = 1; i <= km; ++i) col[i] *= inv;
    for (int c = j + 1; c <= ju; ++c) {
      double* cc = ab + (size_t)c * ld + kv + j - c;
      double a = cc[0];
      if (a != 0.0) for (int i = 1; i <= km; ++i) cc[i] -= col[i] * a;
    }
  }
  for (int j = 0; j < n; ++j) { /* L solve */
    int km = kl < n - 1 - j ? kl : n - 1 - j;
    int p = ipiv[j];
    if (p != j) { double t = b[j]; b[j] = b[p]; b[p] = t; }
    const double* col = ab + (size_t)j * ld + kv;
    for (int i = 1; i <= km; ++i) b[j + i] -= col[i] * b[j];
  }
  for (int j = n - 1; j >= 0; --j) { /* U solve */
    const double* col = ab + (size_t)j * ld + kv;
    b[j] /= col[0];
    int lo = j - kv > 0 ? j - kv : 0;
    for (int i = lo; i < j; ++i) b[i] -= col[i - j] * b[j];
  }
  return 0;
}

/* index maps of the interleaved KKT ordering [u0 l0 | x1 u1 l1 | ... | x_{N-1} u_{N-1} l_{N-1} | x_N] */
static inline int ix_x(const oprob_t* P, int k) { return (P->nu + P->nx) + (k - 1) * (P->s + P->nx); }
static inline int ix_u(const oprob_t* P, int k) { return k ? ix_x(P, k) + P->nx : 0; }
static inline int ix_l(const oprob_t* P, int k) { return ix_u(P, k) + P->nu; }

typedef struct {
  double f, kkt_error, mu;
  int iters, status; /* 0 converged, 1 max_iter, 2 numerical failure */
} ostats_t;

typedef struct {
  double *g, *gradf, *Fx, *Fu, *Hzz, *ab, *rhs, *d, *lam_new, *Vt, *gt, *zL, *zU, *dzL, *dzU, *Sig, *gphi;
  int* ipiv;
} owork_t;

static void work_alloc(owork_t* W, const oprob_t* P) {
  int nx = P->nx, nu = P->nu, s = P->s, N = P->N, n = P->NV - nx + P->M;
  int kl = s + nx, ld = 3 * kl + 1;
  W->g = malloc(sizeof(double) * P->M); W->gradf = malloc(sizeof(double) * P->NV);
  W->Fx = malloc(sizeof(double) * N * nx * nx); W->Fu = malloc(sizeof(double) * N * nx * nu);
  W->Hzz = malloc(sizeof(double) * N * s * s); W->ab = malloc(sizeof(double) * (size_t)ld * n);
  W->rhs = malloc(sizeof(double) * n); W->d = malloc(sizeof(double) * P->NV); W->lam_new = malloc(sizeof(double) * P->M);
  W->Vt = malloc(sizeof(double) * P->NV); W->gt = malloc(sizeof(double) * P->M);
  W->zL = calloc(P->NV, sizeof(double)); W->zU = calloc(P->NV, sizeof(double));
  W->dzL = calloc(P->NV, sizeof(double)); W->dzU = calloc(P->NV, sizeof(double));
  W->Sig = calloc(P->NV, sizeof(double)); W->gphi = calloc(P->NV, sizeof(double));
  W->ipiv = malloc(sizeof(int) * n);
}
static void work_free(owork_t* W) {
  free(W->g); free(W->gradf); free(W->Fx); free(W->Fu); free(W->Hzz); free(W->ab); free(W->rhs); free(W->d);
  free(W->lam_new); free(W->Vt); free(W->gt); free(W->zL); free(W->zU); free(W->dzL); free(W->dzU); free(W->Sig);
  free(W->gphi); free(W->ipiv);
}

/* IPOPT's scaled optimality error E_mu (Waechter & Biegler 2006 eq. 5) [upstream]; uses W->g,
 * gradf, Fx, Fu that must be current for (V). */
static double kkt_err(const oprob_t* P, const owork_t* W, const double* V, const double* lam, const double* lb,
                      const double* ub, double mu) {
  int nx = P->nx, nu = P->nu, s = P->s, N = P->N;
  double ed = 0, ep = 0, ec = 0, lsum = 0, zsum = 0;
  for (int i = 0; i < P->M; ++i) { double a = fabs(W->g[i]); if (a > ep) ep = a; lsum += fabs(lam[i]); }
  for (int k = 0; k <= N; ++k) {
    int nz = k < N ? s : nx;
    for (int a = (k == 0 ? nx : 0); a < nz; ++a) {
      int iv = k * s + a;
      double r = W->gradf[iv] - W->zL[iv] + W->zU[iv];
      if (k < N) for (int i = 0; i < nx; ++i)
        r += (a < nx ? W->Fx[(size_t)k * nx * nx + i * nx + a] : W->Fu[(size_t)k * nx * nu + i * nu + a - nx]) * lam[k * nx + i];
      if (a < nx && k > 0) r -= lam[(k - 1) * nx + a];
      if (fabs(r) > ed) ed = fabs(r);
      zsum += fabs(W->zL[iv]) + fabs(W->zU[iv]);
      if (lb[iv] > -INF_THRESH) { double c = fabs((V[iv] - lb[iv]) * W->zL[iv] - mu); if (c > ec) ec = c; }
      if (ub[iv] < INF_THRESH) { double c = fabs((ub[iv] - V[iv]) * W->zU[iv] - mu); if (c > ec) ec = c; }
    }
  }
  int n = P->NV - nx, m = P->M;
  double sd = (lsum + zsum) / (m + n); sd = (sd > 100.0 ? sd : 100.0) / 100.0;
  double sc = zsum / n; sc = (sc > 100.0 ? sc : 100.0) / 100.0;
  double e = ed / sd; if (ep > e) e = ep; if (ec / sc > e) e = ec / sc;
  (void)nu;
  return e;
}

static double barrier(const oprob_t* P, const double* V, const double* lb, const double* ub, double mu) {
  double b = 0;
  if (mu == 0.0) return 0.0;
  for (int i = P->nx; i < P->NV; ++i) {
    if (lb[i] > -INF_THRESH) b -= mu * log(V[i] - lb[i]);
    if (ub[i] < INF_THRESH) b -= mu * log(ub[i] - V[i]);
  }
  return b;
}

/* One solve.  V: in = start point (zeros = the reference's cold start, ModelControl.cpp:28-50,68), out =
 * solution.  lam: in/out constraint multipliers.  lb/ub: length NV (entries [0,nx) ignored: x0 is pinned
 * to `state`, ModelControl.cpp:144-145); NULL = unbounded. */
static int solve_one(const oprob_t* P, owork_t* W, const double* state, double* V, double* lam, const double* lb_in,
                     const double* ub_in, double tol, int max_iter, double mu_init, ostats_t* st) {
  int nx = P->nx, nu = P->nu, s = P->s, N = P->N, NV = P->NV, M = P->M;
  int n = NV - nx + M, kl = s + nx, ku = s + nx, ld = 2 * kl + ku + 1, kv = kl + ku;
  double* lb = malloc(sizeof(double) * NV); double* ub = malloc(sizeof(double) * NV);
  int bounded = 0;
  for (int i = 0; i < NV; ++i) {
    lb[i] = lb_in ? lb_in[i] : -HUGE_VAL; ub[i] = ub_in ? ub_in[i] : HUGE_VAL;
    if (i >= nx && (lb[i] > -INF_THRESH || ub[i] < INF_THRESH)) bounded = 1;
  }
  memcpy(V, state, sizeof(double) * nx);
  double mu = bounded ? mu_init : 0.0;
  memset(W->zL, 0, sizeof(double) * NV); memset(W->zU, 0, sizeof(double) * NV);
  if (bounded) { /* bound_push = bound_frac = 1e-2 [upstream IPOPT defaults] */
    for (int i = nx; i < NV; ++i) {
      int hl = lb[i] > -INF_THRESH, hu = ub[i] < INF_THRESH;
      if (hl) { double pl = 1e-2 * fmax(1.0, fabs(lb[i])); if (hu) pl = fmin(pl, 1e-2 * (ub[i] - lb[i])); if (V[i] < lb[i] + pl) V[i] = lb[i] + pl; W->zL[i] = 1.0; }
      if (hu) { double pu = 1e-2 * fmax(1.0, fabs(ub[i])); if (hl) pu = fmin(pu, 1e-2 * (ub[i] - lb[i])); if (V[i] > ub[i] - pu) V[i] = ub[i] - pu; W->zU[i] = 1.0; }
    }
  }
  double nu_merit = 1.0, err = HUGE_VAL, fval = 0;
  int it = 0, status = 1;
  for (it = 0; it <= max_iter; ++it) {
    fval = nlp_eval(P, V, lam, 1.0, 2, W->g, W->gradf, W->Fx, W->Fu, W->Hzz);
    err = kkt_err(P, W, V, lam, lb, ub, 0.0);
    if (err <= tol) { status = 0; break; }
    if (it == max_iter) break;
    if (bounded) /* monotone mu (kappa_eps 10, kappa_mu 0.2, theta_mu 1.5) [upstream] */
      while (mu > tol / 10.0 && kkt_err(P, W, V, lam, lb, ub, mu) <= 10.0 * mu) mu = fmax(tol / 10.0, fmin(0.2 * mu, pow(mu, 1.5)));
    for (int i = 0; i < NV; ++i) {
      double sg = 0, gp = W->gradf[i];
      if (i >= nx && lb[i] > -INF_THRESH) { double dl = V[i] - lb[i]; sg += W->zL[i] / dl; gp -= mu / dl; }
      if (i >= nx && ub[i] < INF_THRESH) { double du = ub[i] - V[i]; sg += W->zU[i] / du; gp += mu / du; }
      W->Sig[i] = sg; W->gphi[i] = gp;
    }
    double delta = 0.0; int ok = 0;
    for (int attempt = 0; attempt < 30 && !ok; ++attempt) {
      memset(W->ab, 0, sizeof(double) * (size_t)ld * n);
#define BAND(i, j) W->ab[(size_t)(j) * ld + kv + (i) - (j)]
      for (int k = 0; k <= N; ++k) {
        int nz = k < N ? s : nx;
        for (int a = (k == 0 ? nx : 0); a < nz; ++a) {
          int ra = a < nx ? ix_x(P, k) + a : ix_u(P, k) + a - nx;
          W->rhs[ra] = -W->gphi[k * s + a];
          if (k < N) {
            for (int b = (k == 0 ? nx : 0); b < s; ++b) {
              int cb = b < nx ? ix_x(P, k) + b : ix_u(P, k) + b - nx;
              BAND(ra, cb) += W->Hzz[(size_t)k * s * s + a * s + b];
            }
            for (int i = 0; i < nx; ++i) { /* J' and J */
              double v = a < nx ? W->Fx[(size_t)k * nx * nx + i * nx + a] : W->Fu[(size_t)k * nx * nu + i * nu + a - nx];
              BAND(ra, ix_l(P, k) + i) += v; BAND(ix_l(P, k) + i, ra) += v;
            }
            if (a >= nx && k < N - 1) { int cu = ix_u(P, k + 1) + a - nx; BAND(ra, cu) -= 2.0 * P->R[a - nx]; BAND(cu, ra) -= 2.0 * P->R[a - nx]; }
          }
          BAND(ra, ra) += W->Sig[k * s + a] + delta;
          if (a < nx && k > 0) { int cl = ix_l(P, k - 1) + a; BAND(ra, cl) -= 1.0; BAND(cl, ra) -= 1.0; }
        }
        if (k < N) for (int i = 0; i < nx; ++i) W->rhs[ix_l(P, k) + i] = -W->g[k * nx + i];
      }
#undef BAND
      if (band_solve(n, kl, ku, W->ab, ld, W->ipiv, W->rhs) == 0) {
        ok = 1;
        memset(W->d, 0, sizeof(double) * NV);
        for (int k = 0; k <= N; ++k) {
          int nz = k < N ? s : nx;
          for (int a = (k == 0 ? nx : 0); a < nz; ++a) { double v = W->rhs[a < nx ? ix_x(P, k) + a : ix_u(P, k) + a - nx]; if (!isfinite(v)) ok = 0; W->d[k * s + a] = v; }
          if (k < N) for (int i = 0; i < nx; ++i) W->lam_new[k * nx + i] = W->rhs[ix_l(P, k) + i];
        }
        if (ok) { /* curvature test d'(H+Sigma+delta)d > 0 as the inertia surrogate */
          double curv = 0, dmax = 0;
          for (int k = 0; k < N; ++k) {
            const double* dz = W->d + k * s;
            for (int a = 0; a < s; ++a) { double acc = 0; for (int b = 0; b < s; ++b) acc += W->Hzz[(size_t)k * s * s + a * s + b] * dz[b]; curv += dz[a] * acc; }
            if (k < N - 1) for (int j = 0; j < nu; ++j) curv -= 4.0 * P->R[j] * dz[nx + j] * dz[s + nx + j];
          }
          for (int i = 0; i < NV; ++i) { curv += (W->Sig[i] + delta) * W->d[i] * W->d[i]; if (fabs(W->d[i]) > dmax) dmax = fabs(W->d[i]); }
          if (!(curv > 0.0) && dmax > 0.0) ok = 0;
        }
      }
      if (!ok) delta = delta == 0.0 ? 1e-4 : delta * 8.0;
    }
    if (!ok) { status = 2; break; }
    double amax = 1.0, az = 1.0;
    if (bounded) {
      double tau = fmax(0.99, 1.0 - mu);
      for (int i = nx; i < NV; ++i) {
        double d = W->d[i];
        if (lb[i] > -INF_THRESH) { double dl = V[i] - lb[i]; W->dzL[i] = (mu - W->zL[i] * d) / dl - W->zL[i]; if (d < 0) amax = fmin(amax, -tau * dl / d); if (W->dzL[i] < 0) az = fmin(az, -tau * W->zL[i] / W->dzL[i]); }
        if (ub[i] < INF_THRESH) { double du = ub[i] - V[i]; W->dzU[i] = (mu + W->zU[i] * d) / du - W->zU[i]; if (d > 0) amax = fmin(amax, tau * du / d); if (W->dzU[i] < 0) az = fmin(az, -tau * W->zU[i] / W->dzU[i]); }
      }
    }
    double lmax = 0, theta0 = 0, D = 0;
    for (int i = 0; i < M; ++i) { if (fabs(W->lam_new[i]) > lmax) lmax = fabs(W->lam_new[i]); theta0 += fabs(W->g[i]); }
    if (1.1 * lmax + 1e-8 > nu_merit) nu_merit = 1.1 * lmax + 1e-8;
    for (int i = 0; i < NV; ++i) D += W->gphi[i] * W->d[i];
    D -= nu_merit * theta0;
    double m0 = fval + barrier(P, V, lb, ub, mu) + nu_merit * theta0, a = amax;
    for (int ls = 0; ls < 40; ++ls) {
      for (int i = 0; i < NV; ++i) W->Vt[i] = V[i] + a * W->d[i];
      double ft = nlp_eval(P, W->Vt, 0, 1.0, 0, W->gt, 0, 0, 0, 0), th = 0;
      for (int i = 0; i < M; ++i) th += fabs(W->gt[i]);
      if (ft + barrier(P, W->Vt, lb, ub, mu) + nu_merit * th <= m0 + 1e-4 * a * D + 1e-12 * fabs(m0)) break;
      a *= 0.5;
    }
    for (int i = nx; i < NV; ++i) V[i] += a * W->d[i];
    for (int i = 0; i < M; ++i) lam[i] += a * (W->lam_new[i] - lam[i]);
    if (bounded) for (int i = nx; i < NV; ++i) { /* kappa_sigma = 1e10 safeguard (eq. 16) [upstream] */
      if (lb[i] > -INF_THRESH) { double dl = V[i] - lb[i], z = W->zL[i] + az * W->dzL[i]; W->zL[i] = fmin(fmax(z, mu / (1e10 * dl)), 1e10 * mu / dl); }
      if (ub[i] < INF_THRESH) { double du = ub[i] - V[i], z = W->zU[i] + az * W->dzU[i]; W->zU[i] = fmin(fmax(z, mu / (1e10 * du)), 1e10 * mu / du); }
    }
  }
  st->f = fval; st->kkt_error = err; st->iters = it; st->status = status; st->mu = mu;
  free(lb); free(ub);
  return status;
}

/* Batch of independent solves, OpenMP over instances (this is the "restated CPU baseline" of BASELINE.md 3).
 *   x0[B*nx], p[B*np], V[B*NV] (in: warm start, out), lam[B*M] (in/out), lb/ub[NV] shared or NULL,
 *   stats[B] */
int oracle_solve_batch(int model, int N, double h, int is_linear, int B, const double* x0, const double* p,
                       double* V, double* lam, const double* lb, const double* ub, double tol, int max_iter,
                       double mu_init, ostats_t* stats, int nthreads) {
  int bad = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : bad)
  {
    oprob_t P; prob_init(&P, model, N, h, is_linear, 0);
    owork_t W; work_alloc(&W, &P);
#pragma omp for schedule(dynamic, 4)
    for (int b = 0; b < B; ++b) {
      oprob_t Pb; prob_init(&Pb, model, N, h, is_linear, p + (size_t)b * P.np);
      bad += solve_one(&Pb, &W, x0 + (size_t)b * P.nx, V + (size_t)b * P.NV, lam + (size_t)b * P.M, lb, ub, tol, max_iter,
                       mu_init, &stats[b]) != 0;
    }
    work_free(&W);
  }
  return bad;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
