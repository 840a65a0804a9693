// ex_model_control -- closed-loop MPC of the generated double pendulum (BASELINE config 1): each 2 ms cycle the
// controller re-plans over the horizon against a sinusoidal reference, the first control of the plan is applied and
// the plant is advanced with one RK4 step -- the loop of the reference's examples/model_generate_example.cpp:156-230
// (its ex_model_control.cpp is a truncated, non-compiling copy of it), written against ModelControl.
//   usage: ex_model_control [-l] [-t SIM_SECONDS] [-e TOL] [-q Q0,Q1,..] [-r R0,R1,..]
// Writes <name>_results.csv (time, states, controls) and prints the average solve time.
#include <Mahi/Mpc/ModelControl.hpp>
#include <Mahi/Util.hpp>

#include <cmath>
#include <fstream>

using mahi::util::PI;
using namespace mahi::mpc;

namespace {
// reference trajectory over the horizon starting at t0: positions +-a sin(2 pi f t), velocities their derivatives
// (examples/model_control_example.cpp:60-68)
std::vector<double> sinusoid_reference(double t0, double dt, int nodes, int nx, double amp, double freq) {
    std::vector<double> traj;
    traj.reserve(static_cast<size_t>(nodes) * nx);
    const int half = nx / 2;
    for (int k = 0; k < nodes; ++k) {
        const double t = t0 + k * dt, w = 2 * PI * freq;
        for (int j = 0; j < nx; ++j) {
            const int jj = j < half ? j : j - half;
            const double sign = jj % 2 == 0 ? 1.0 : -1.0;
            traj.push_back(j < half ? sign * amp * std::sin(w * t) : sign * amp * w * std::cos(w * t));
        }
    }
    return traj;
}

std::vector<double> plant_rhs(const casadi::Function& f, const std::vector<double>& x, const std::vector<double>& u) {
    return std::vector<double>(f(std::vector<casadi::DM>{casadi::DM(x), casadi::DM(u)})[0]);
}

// classical RK4 step of the plant (examples/model_generate_example.cpp:207-213)
std::vector<double> rk4(const casadi::Function& f, const std::vector<double>& x, const std::vector<double>& u, double dt) {
    auto axpy = [](const std::vector<double>& a, double s, const std::vector<double>& b) {
        std::vector<double> r(a.size());
        for (size_t i = 0; i < a.size(); ++i) r[i] = a[i] + s * b[i];
        return r;
    };
    const auto k1 = plant_rhs(f, x, u), k2 = plant_rhs(f, axpy(x, dt / 2, k1), u), k3 = plant_rhs(f, axpy(x, dt / 2, k2), u),
               k4 = plant_rhs(f, axpy(x, dt, k3), u);
    std::vector<double> r(x.size());
    for (size_t i = 0; i < x.size(); ++i) r[i] = x[i] + dt / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    return r;
}
}  // namespace

int main(int argc, char* argv[]) {
    mahi::util::Options options("ex_model_control", "Closed-loop double-pendulum MPC on mahi-mpc-b200");
    options.add_options()("l,linear", "Use the linearized model.")("t,time", "simulated seconds (default 0.2)", true)(
        "e,tol", "solver tolerance (default 1e-5, the reference's ipopt.tol)", true)("q,qvec", "state weights, comma separated", true)("r,rvec", "control-rate weights, comma separated", true)("h,help", "print help");
    auto result = options.parse(argc, argv);
    if (result.count("help")) {
        std::cout << options.help();
        return 0;
    }
    const std::string model_name = std::string(result.count("linear") ? "linear_" : "nonlinear_") + "double_pendulum";
    const double sim_seconds = result.count("time") ? std::stod(result.values("time")[0]) : 0.2;
    std::vector<double> Q = {10, 1, 5, 5}, R = {5, 5}, Rm = {1, 1};  // thread_model_control_example.cpp:24-25
    if (result.count("qvec")) Q = result.doubles("qvec");
    if (result.count("rvec")) R = result.doubles("rvec");

    std::cout << "Loading " << model_name << std::endl;
    casadi::Dict solver_opts;
    if (result.count("tol")) solver_opts["ipopt.tol"] = std::stod(result.values("tol")[0]);
    ModelControl model_control(model_name, Q, R, Rm, solver_opts);
    const int nx = model_control.model_parameters.num_x, nu = model_control.model_parameters.num_u;
    const int nodes = model_control.model_parameters.num_shooting_nodes;
    const double dt = model_control.model_parameters.step_size.as_seconds();
    auto plant = casadi::external(model_name + "_get_x_dot_init", model_name + "_linear_functions.so");

    std::vector<double> state(nx, 0.0), control(nu, 0.0);
    std::ofstream csv(model_name + "_results.csv");
    csv << "t";
    for (int i = 0; i < nx; ++i) csv << ",x" << i;
    for (int j = 0; j < nu; ++j) csv << ",u" << j;
    csv << ",iterations\n";

    double t = 0.0, solve_ms_sum = 0.0;
    int cycles = 0;
    mahi::util::Clock sim_clock;
    while (t < sim_seconds - 1e-12) {
        mahi::util::Clock cycle_clock;
        model_control.calc_u(mahi::util::seconds(t), state, control, sinusoid_reference(t, dt, nodes, nx, 1.0, 1.0));
        solve_ms_sum += cycle_clock.get_elapsed_time().as_microseconds() * 1e-3;
        control = model_control.control_results[0].u;  // first move of the plan
        csv.precision(17);
        csv << t;
        for (double v : state) csv << "," << v;
        for (double v : control) csv << "," << v;
        csv << "," << model_control.last_solve_info().iterations << "\n";
        state = rk4(plant, state, control, dt);
        t += dt;
        ++cycles;
    }
    mahi::util::print("sim time: {:.2f} s, avg solve time: {:.3f} ms over {} cycles", sim_clock.get_elapsed_time().as_seconds(),
                      solve_ms_sum / cycles, cycles);
    return 0;
}
