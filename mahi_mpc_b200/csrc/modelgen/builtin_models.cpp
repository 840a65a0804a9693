// mahi-mpc-b200 :: built-in model definitions, written against the SX stand-in exactly the way a
// user of the reference writes them against casadi::SX.
#include <cmath>

#include "codegen.hpp"

namespace mahi {
namespace mpc {
namespace codegen {

using sym::SX;

// Double pendulum of examples/ex_model_generate.cpp:24-43 (L = m = 1, g = 9.81): same rational
// function of sin/cos(qA), sin/cos(qB), cos(qA+qB), grouped by torque / velocity / gravity terms.
static ModelExpr double_pendulum() {
    const double L = 1.0, m = 1.0, g = 9.81;
    SX qA = SX::sym("qA"), qB = SX::sym("qB"), qA_dot = SX::sym("qA_dot"), qB_dot = SX::sym("qB_dot");
    SX TA = SX::sym("TA"), TB = SX::sym("TB");
    const double mL2 = m * L * L, gmL = g * m * L;
    SX cA = cos(qA), cB = cos(qB), sB = sin(qB), cAB = cos(qA + qB);
    SX den = mL2 * (cB * cB - 2.0);
    SX vA2 = qA_dot * qA_dot, vB2 = qB_dot * qB_dot, vAB = 2.0 * qA_dot * qB_dot;
    SX numA = TA - TB * (1.0 + cB) + mL2 * sB * (vA2 * (1.0 + cB) + vB2 + vAB) - gmL * (2.0 * cA - cAB * cB);
    SX numB = TA * (1.0 + cB) - TB * (3.0 + 2.0 * cB) + mL2 * sB * (vA2 * (3.0 + 2.0 * cB) + (vB2 + vAB) * (1.0 + cB)) +
              gmL * (cAB * (2.0 + cB) - 2.0 * cA * (1.0 + cB));
    ModelExpr e;
    e.name = "double_pendulum";
    e.x = SX::vertcat({qA, qB, qA_dot, qB_dot});
    e.x_dot = SX::vertcat({qA_dot, qB_dot, -numA / den, numB / den});
    e.u = SX::vertcat({TA, TB});
    return e;
}

// 4-DOF arm of BASELINE config 3.  The reference holds no runnable 4-DOF model (only the symbolic mass matrix of
// src/inverseTest.cpp:59-83 with 40 un-valued inertial symbols), so this is the builder-defined
// stand-in documented in DESIGN.md: a planar serial chain of four links with point masses at the link tips,
// gravity along -y, viscous joint damping; x = [q; qdot], u = tau,
//     qddot = M(q)^-1 (tau - C(q,qdot) qdot - G(q) - D qdot),   C from the Christoffel symbols of M.
// M is assembled from the tip Jacobians (sum_i m_i J_i' J_i); the oracle restates it from the closed-form
// double sum over link pairs, so the two derivations are independent.
static ModelExpr arm4() {
    const int n = 4;
    const double len[n] = {0.30, 0.25, 0.20, 0.15}, mass[n] = {2.0, 1.5, 1.0, 0.5}, damp[n] = {0.5, 0.4, 0.3, 0.2};
    const double g = 9.81;
    SX q = SX::sym("q", n), qd = SX::sym("qd", n), tau = SX::sym("tau", n);
    std::vector<SX> th(n), px(n), py(n);
    SX acc_th = SX(0.0), cx = SX(0.0), cy = SX(0.0);
    for (int i = 0; i < n; ++i) {
        acc_th = acc_th + q(i);
        th[i] = acc_th;
        cx = cx + len[i] * cos(th[i]);
        cy = cy + len[i] * sin(th[i]);
        px[i] = cx;
        py[i] = cy;
    }
    SX M(n, n);
    SX V = SX(0.0);
    for (int i = 0; i < n; ++i) {
        SX Jx = sym::jacobian(px[i], q), Jy = sym::jacobian(py[i], q);  // 1 x n
        M = M + mass[i] * (mtimes(Jx.T(), Jx) + mtimes(Jy.T(), Jy));
        V = V + (mass[i] * g) * py[i];
    }
    SX G = sym::gradient(V, q);
    // Coriolis / centrifugal vector from the Christoffel symbols
    std::vector<SX> dM(n);  // dM[k] = dM/dq_k  (n x n)
    for (int k = 0; k < n; ++k) {
        SX d(n, n);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) d.id(i, j) = sym::graph().forward({M.id(i, j)}, q.ids()[k])[0];
        dM[k] = d;
    }
    SX rhs(n, 1);
    for (int i = 0; i < n; ++i) {
        SX c = SX(0.0);
        for (int j = 0; j < n; ++j)
            for (int k = 0; k < n; ++k) c = c + 0.5 * (dM[k](i, j) + dM[j](i, k) - dM[i](j, k)) * qd(j) * qd(k);
        rhs.id(i, 0) = (tau(i) - c - G(i) - damp[i] * qd(i)).ids()[0];
    }
    // solve M qdd = rhs by a symbolic Cholesky (M is SPD)
    SX L(n, n);
    for (int j = 0; j < n; ++j) {
        SX d = M(j, j);
        for (int k = 0; k < j; ++k) d = d - L(j, k) * L(j, k);
        SX ljj = sqrt(d);
        L.id(j, j) = ljj.ids()[0];
        for (int i = j + 1; i < n; ++i) {
            SX s2 = M(i, j);
            for (int k = 0; k < j; ++k) s2 = s2 - L(i, k) * L(j, k);
            L.id(i, j) = (s2 / ljj).ids()[0];
        }
    }
    std::vector<SX> y(n), qdd(n);
    for (int i = 0; i < n; ++i) {
        SX s2 = rhs(i);
        for (int k = 0; k < i; ++k) s2 = s2 - L(i, k) * y[k];
        y[i] = s2 / L(i, i);
    }
    for (int i = n - 1; i >= 0; --i) {
        SX s2 = y[i];
        for (int k = i + 1; k < n; ++k) s2 = s2 - L(k, i) * qdd[k];
        qdd[i] = s2 / L(i, i);
    }
    ModelExpr e;
    e.name = "arm4";
    e.x = SX::vertcat({q, qd});
    e.x_dot = SX::vertcat({qd, qdd[0], qdd[1], qdd[2], qdd[3]});
    e.u = tau;
    return e;
}

bool builtin_model(const std::string& which, ModelExpr* out) {
    if (which == "double_pendulum") {
        *out = double_pendulum();
        return true;
    }
    if (which == "arm4") {
        *out = arm4();
        return true;
    }
    return false;
}

}  // namespace codegen
}  // namespace mpc
}  // namespace mahi
