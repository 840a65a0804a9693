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

bool builtin_model(const std::string& which, ModelExpr* out) {
    if (which == "double_pendulum") {
        *out = double_pendulum();
        return true;
    }
    return false;
}

}  // namespace codegen
}  // namespace mpc
}  // namespace mahi
