/* ORACLE -- test infrastructure only.
 *
 * 4-DOF arm of BASELINE config 3, restated for the oracle with second-order forward-mode automatic differentiation
 * ("jets": value, gradient, packed Hessian w.r.t. the 12 variables q, qdot, tau).  The reference has no runnable
 * 4-DOF model (only the symbolic mass matrix of src/inverseTest.cpp:59-83); the model is the
 * builder-defined planar 4-link chain with tip point masses documented in DESIGN.md.  This restatement is independent
 * of the product's: the mass matrix, its derivatives (for the Christoffel symbols) and the gravity vector come from the
 * closed-form double sums over link pairs
 *      M_ij = sum_{p>=max(i,j)} m_p sum_{a=i..p} sum_{b=j..p} l_a l_b cos(th_a - th_b),   th_a = q_0 + ... + q_a
 * (the product assembles sum_p m_p J_p' J_p from symbolic tip Jacobians), derivatives are propagated numerically by
 * the jets (the product differentiates symbolically), and M qddot = rhs is solved by Gaussian elimination on jets
 * (the product emits a symbolic Cholesky).
 */
#include <math.h>
#include <string.h>

enum { AJ_N = 12, AJ_H = 78 };
typedef struct { double v, g[AJ_N], h[AJ_H]; } jet;

static inline int aj_idx(int i, int j) { return i * AJ_N - (i * (i - 1)) / 2 + (j - i); } /* i <= j */

static jet aj_const(double c) { jet r; memset(&r, 0, sizeof r); r.v = c; return r; }
static jet aj_var(double c, int k) { jet r = aj_const(c); r.g[k] = 1.0; return r; }
static jet aj_add(jet a, jet b) { for (int i = 0; i < AJ_N; ++i) a.g[i] += b.g[i]; for (int i = 0; i < AJ_H; ++i) a.h[i] += b.h[i]; a.v += b.v; return a; }
static jet aj_sub(jet a, jet b) { for (int i = 0; i < AJ_N; ++i) a.g[i] -= b.g[i]; for (int i = 0; i < AJ_H; ++i) a.h[i] -= b.h[i]; a.v -= b.v; return a; }
static jet aj_scale(jet a, double s) { for (int i = 0; i < AJ_N; ++i) a.g[i] *= s; for (int i = 0; i < AJ_H; ++i) a.h[i] *= s; a.v *= s; return a; }
static jet aj_mul(jet a, jet b) {
  jet r;
  r.v = a.v * b.v;
  for (int i = 0; i < AJ_N; ++i) r.g[i] = a.v * b.g[i] + b.v * a.g[i];
  for (int i = 0; i < AJ_N; ++i)
    for (int j = i; j < AJ_N; ++j) {
      const int k = aj_idx(i, j);
      r.h[k] = a.v * b.h[k] + b.v * a.h[k] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
    }
  return r;
}
/* r = f(a) given f, f', f'' at a.v */
static jet aj_unary(jet a, double f, double f1, double f2) {
  jet r;
  r.v = f;
  for (int i = 0; i < AJ_N; ++i) r.g[i] = f1 * a.g[i];
  for (int i = 0; i < AJ_N; ++i)
    for (int j = i; j < AJ_N; ++j) { const int k = aj_idx(i, j); r.h[k] = f1 * a.h[k] + f2 * a.g[i] * a.g[j]; }
  return r;
}
static jet aj_sin(jet a) { return aj_unary(a, sin(a.v), cos(a.v), -sin(a.v)); }
static jet aj_cos(jet a) { return aj_unary(a, cos(a.v), -sin(a.v), -cos(a.v)); }
static jet aj_recip(jet a) { const double r = 1.0 / a.v; return aj_unary(a, r, -r * r, 2.0 * r * r * r); }
static jet aj_div(jet a, jet b) { return aj_mul(a, aj_recip(b)); }

static const double ARM4_L[4] = {0.30, 0.25, 0.20, 0.15};
static const double ARM4_M[4] = {2.0, 1.5, 1.0, 0.5};
static const double ARM4_D[4] = {0.5, 0.4, 0.3, 0.2};
static const double ARM4_G = 9.81;

/* qdd[4] jets at (x = [q, qd], u = tau) */
static void arm4_qdd_jets(const double* x, const double* u, jet* qdd) {
  enum { n = 4 };
  jet q[n], qd[n], tau[n], th[n];
  for (int i = 0; i < n; ++i) { q[i] = aj_var(x[i], i); qd[i] = aj_var(x[n + i], n + i); tau[i] = aj_var(u[i], 2 * n + i); }
  th[0] = q[0];
  for (int i = 1; i < n; ++i) th[i] = aj_add(th[i - 1], q[i]);
  jet cd[n][n], sd[n][n], cth[n];
  for (int a = 0; a < n; ++a) {
    cth[a] = aj_cos(th[a]);
    for (int b = 0; b < n; ++b) { jet d = aj_sub(th[a], th[b]); cd[a][b] = aj_cos(d); sd[a][b] = aj_sin(d); }
  }
  jet M[n][n], dM[n][n][n], G[n];
  for (int i = 0; i < n; ++i) {
    G[i] = aj_const(0.0);
    for (int p = i; p < n; ++p)
      for (int a = i; a <= p; ++a) G[i] = aj_add(G[i], aj_scale(cth[a], ARM4_G * ARM4_M[p] * ARM4_L[a]));
    for (int j = 0; j < n; ++j) {
      M[i][j] = aj_const(0.0);
      for (int k = 0; k < n; ++k) dM[i][j][k] = aj_const(0.0);
      const int p0 = i > j ? i : j;
      for (int p = p0; p < n; ++p)
        for (int a = i; a <= p; ++a)
          for (int b = j; b <= p; ++b) {
            const double w = ARM4_M[p] * ARM4_L[a] * ARM4_L[b];
            M[i][j] = aj_add(M[i][j], aj_scale(cd[a][b], w));
            for (int k = 0; k < n; ++k) {
              const int s = (k <= a) - (k <= b); /* d(th_a - th_b)/dq_k */
              if (s) dM[i][j][k] = aj_add(dM[i][j][k], aj_scale(sd[a][b], -w * s));
            }
          }
    }
  }
  jet rhs[n];
  for (int i = 0; i < n; ++i) {
    jet c = aj_const(0.0);
    for (int j = 0; j < n; ++j)
      for (int k = 0; k < n; ++k) {
        jet gam = aj_scale(aj_sub(aj_add(dM[i][j][k], dM[i][k][j]), dM[j][k][i]), 0.5);
        c = aj_add(c, aj_mul(gam, aj_mul(qd[j], qd[k])));
      }
    rhs[i] = aj_sub(aj_sub(aj_sub(tau[i], c), G[i]), aj_scale(qd[i], ARM4_D[i]));
  }
  /* Gaussian elimination without pivoting (M is SPD) */
  for (int c = 0; c < n; ++c) {
    jet inv = aj_recip(M[c][c]);
    for (int r = c + 1; r < n; ++r) {
      jet f = aj_mul(M[r][c], inv);
      for (int k = c; k < n; ++k) M[r][k] = aj_sub(M[r][k], aj_mul(f, M[c][k]));
      rhs[r] = aj_sub(rhs[r], aj_mul(f, rhs[c]));
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    jet s = rhs[r];
    for (int k = r + 1; k < n; ++k) s = aj_sub(s, aj_mul(M[r][k], qdd[k]));
    qdd[r] = aj_div(s, M[r][r]);
  }
}

void arm4_f(const double* x, const double* u, double* f) {
  jet qdd[4];
  arm4_qdd_jets(x, u, qdd);
  for (int i = 0; i < 4; ++i) { f[i] = x[4 + i]; f[4 + i] = qdd[i].v; }
}

/* A row-major 8x8, B row-major 8x4 */
void arm4_fAB(const double* x, const double* u, double* f, double* A, double* B) {
  jet qdd[4];
  arm4_qdd_jets(x, u, qdd);
  memset(A, 0, sizeof(double) * 64);
  memset(B, 0, sizeof(double) * 32);
  for (int i = 0; i < 4; ++i) {
    f[i] = x[4 + i]; f[4 + i] = qdd[i].v;
    A[i * 8 + 4 + i] = 1.0;
    for (int j = 0; j < 8; ++j) A[(4 + i) * 8 + j] = qdd[i].g[j];
    for (int j = 0; j < 4; ++j) B[(4 + i) * 4 + j] = qdd[i].g[8 + j];
  }
}

/* H = sum_i w_i d2 f_i / dz2, dense row-major 12x12 over z = (x, u) */
void arm4_fABH(const double* x, const double* u, const double* w, double* f, double* A, double* B, double* H) {
  jet qdd[4];
  arm4_qdd_jets(x, u, qdd);
  memset(A, 0, sizeof(double) * 64);
  memset(B, 0, sizeof(double) * 32);
  memset(H, 0, sizeof(double) * 144);
  for (int i = 0; i < 4; ++i) {
    f[i] = x[4 + i]; f[4 + i] = qdd[i].v;
    A[i * 8 + 4 + i] = 1.0;
    for (int j = 0; j < 8; ++j) A[(4 + i) * 8 + j] = qdd[i].g[j];
    for (int j = 0; j < 4; ++j) B[(4 + i) * 4 + j] = qdd[i].g[8 + j];
    for (int a = 0; a < 12; ++a)
      for (int b = a; b < 12; ++b) {
        const double v = w[4 + i] * qdd[i].h[aj_idx(a, b)];
        H[a * 12 + b] += v;
        if (a != b) H[b * 12 + a] += v;
      }
  }
}
