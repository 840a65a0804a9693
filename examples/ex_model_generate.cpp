// ex_model_generate -- builds the model artefacts of the double pendulum (L = m = 1, g = 9.81, 4 states, 2
// torques, 2 ms step, 25 shooting nodes), i.e. what the reference's examples/ex_model_generate.cpp does, through
// the same ModelParameters / ModelGenerator calls.  Output in the working directory:
//   <name>.json, <name>.cu -> <name>.so (CUDA kernels, sm_100a), <name>_linear_functions.c -> .so (CPU, CasADi ABI)
// where <name> = nonlinear_double_pendulum, or linear_double_pendulum with -l / --linear.
//   usage: ex_model_generate [-l] [-n NODES] [-d STEP_MS] [--name NAME]
// --name overrides <name>: the reference's control drivers load "double_pendulum" / "linear_double_pendulum"
// (examples/model_control_example.cpp:17-18), not the names its generate driver writes.
#include <Mahi/Mpc/ModelGenerator.hpp>
#include <Mahi/Util.hpp>

using namespace casadi;
using namespace mahi::mpc;

namespace {
struct Pendulum {
    SX x, x_dot, u;
};

// Equations of motion of the planar double pendulum with joint torques (the closed form the reference driver
// spells out term by term, grouped here by physical origin).
Pendulum double_pendulum(double L, double m, double g) {
    SX qA = SX::sym("qA"), qB = SX::sym("qB"), wA = SX::sym("qA_dot"), wB = SX::sym("qB_dot");
    SX TA = SX::sym("TA"), TB = SX::sym("TB");
    const double I = m * L * L, G = g * m * L;
    SX cB = cos(qB), sB = sin(qB), cA = cos(qA), cAB = cos(qA + qB);
    SX centripetal_A = I * sB * (wA * wA * (1.0 + cB) + wB * wB + 2.0 * wA * wB);
    SX centripetal_B = I * sB * (wA * wA * (3.0 + 2.0 * cB) + (wB * wB + 2.0 * wA * wB) * (1.0 + cB));
    SX gravity_A = G * (cAB * cB - 2.0 * cA);
    SX gravity_B = G * (cAB * (2.0 + cB) - 2.0 * cA * (1.0 + cB));
    SX den = I * (cB * cB - 2.0);
    SX aA = -(TA - TB * (1.0 + cB) + centripetal_A + gravity_A) / den;
    SX aB = (TA * (1.0 + cB) - TB * (3.0 + 2.0 * cB) + centripetal_B + gravity_B) / den;
    Pendulum p;
    p.x = SX::vertcat({qA, qB, wA, wB});
    p.x_dot = SX::vertcat({wA, wB, aA, aB});
    p.u = SX::vertcat({TA, TB});
    return p;
}
}  // namespace

int main(int argc, char* argv[]) {
    mahi::util::Options options("ex_model_generate", "Generates the double-pendulum MPC model for mahi-mpc-b200");
    options.add_options()("l,linear", "Generates linearized model.")("n,nodes", "number of shooting nodes (default 25)", true)(
        "d,dt", "step size in milliseconds (default 2)", true)("name", "model name (default [non]linear_double_pendulum)", true)("h,help", "print help");
    auto result = options.parse(argc, argv);
    if (result.count("help")) {
        std::cout << options.help();
        return 0;
    }
    const bool linear = result.count("linear") > 0;
    const int num_shooting_nodes = result.count("nodes") ? std::stoi(result.values("nodes")[0]) : 25;
    const mahi::util::Time time_step = mahi::util::milliseconds(result.count("dt") ? std::stoi(result.values("dt")[0]) : 2);

    const Pendulum pend = double_pendulum(1.0, 1.0, 9.81);
    const std::string model_name = result.count("name") ? result.values("name")[0] : std::string(linear ? "linear_" : "nonlinear_") + "double_pendulum";

    ModelParameters model_parameters(model_name, pend.x.size1(), pend.u.size1(), time_step, num_shooting_nodes, linear);
    ModelGenerator my_generator(model_parameters, pend.x, pend.x_dot, pend.u);
    my_generator.create_model();
    my_generator.generate_c_code();
    my_generator.compile_model();
    return 0;
}
