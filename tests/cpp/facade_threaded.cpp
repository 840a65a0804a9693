// GPU test program for the C++ facade (run by tests/test_facade.py in the directory that holds the generated model):
// the free-running solver thread of ModelControl (start_calc / set_state / control_at_time / stop_calc,
// reference: src/Mahi/Mpc/ModelControl.cpp:75-114,192-197), update_weights, update_control_limits, and the batched form.
#include <Mahi/Mpc.hpp>

#include <cmath>
#include <cstdio>
#include <cstdlib>

using namespace mahi::mpc;
using mahi::util::PI;

#define CHECK(c)                                                              \
    do {                                                                      \
        if (!(c)) {                                                           \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);        \
            return 1;                                                         \
        }                                                                     \
    } while (0)

static std::vector<double> reference(double t0, double dt, int nodes) {
    std::vector<double> r;
    for (int k = 0; k < nodes; ++k) {
        const double t = t0 + k * dt, w = 2 * PI;
        r.push_back(std::sin(w * t)); r.push_back(-std::sin(w * t)); r.push_back(w * std::cos(w * t)); r.push_back(-w * std::cos(w * t));
    }
    return r;
}

int main() {
    const std::string name = "nonlinear_double_pendulum";
    casadi::Dict opts;
    opts["ipopt.tol"] = 1e-9;
    ModelControl mc(name, {10, 1, 5, 5}, {5, 5}, {1, 1}, opts);
    const int nx = mc.model_parameters.num_x, nu = mc.model_parameters.num_u, N = mc.model_parameters.num_shooting_nodes;
    const double dt = mc.model_parameters.step_size.as_seconds();
    CHECK(nx == 4 && nu == 2 && N == 25);

    // synchronous solve, plan unpacking and the zero-order-hold lookup
    std::vector<double> state(nx, 0.0), control(nu, 0.0);
    mc.calc_u(mahi::util::seconds(0.1), state, control, reference(0.1, dt, N));
    CHECK(mc.last_solve_info().status == 0);
    CHECK((int)mc.control_results.size() == N);
    CHECK(std::fabs(mc.control_results[3].time.as_seconds() - (0.1 + 3 * dt)) < 1e-9);
    const auto plan = mc.control_results;
    CHECK(mc.control_at_time(mahi::util::seconds(0.0)).u == plan[0].u);                 // before the plan: first entry
    CHECK(mc.control_at_time(mahi::util::seconds(0.1 + 2.5 * dt)).u == plan[2].u);      // last entry with time < t
    CHECK(mc.control_at_time(mahi::util::seconds(10.0)).u == plan[N - 1].u);            // after the plan: last entry
    const std::vector<double> u_free = plan[0].u;

    // control limits: the plan respects them and the limit is active
    const double lim = 0.5 * std::fmax(std::fabs(u_free[0]), std::fabs(u_free[1]));
    mc.update_control_limits({-lim, -lim}, {lim, lim});
    mc.calc_u(mahi::util::seconds(0.1), state, control, reference(0.1, dt, N));
    CHECK(mc.last_solve_info().status == 0);
    double umax = 0;
    for (const auto& r : mc.control_results) umax = std::fmax(umax, std::fmax(std::fabs(r.u[0]), std::fabs(r.u[1])));
    CHECK(umax <= lim && umax > lim * (1 - 1e-6));
    mc.update_control_limits({-10e30, -10e30}, {10e30, 10e30});
    mc.calc_u(mahi::util::seconds(0.1), state, control, reference(0.1, dt, N));
    CHECK(std::fabs(mc.control_results[0].u[0] - u_free[0]) < 1e-6 && std::fabs(mc.control_results[0].u[1] - u_free[1]) < 1e-6);

    // weights: a heavier control-magnitude weight shrinks the first control
    mc.update_weights({}, {}, {100, 100});
    mc.calc_u(mahi::util::seconds(0.1), state, control, reference(0.1, dt, N));
    CHECK(std::fabs(mc.control_results[0].u[0]) < std::fabs(u_free[0]));
    mc.update_weights({}, {}, {1, 1});

    // free-running solver thread feeding a 1 kHz loop (examples/thread_model_control_example.cpp:66-109)
    mc.set_state(mahi::util::seconds(0.0), state, control, reference(0.0, dt, N));
    mc.start_calc();
    mahi::util::sleep(mahi::util::milliseconds(30));
    int lookups = 0;
    for (int cyc = 0; cyc < 50; ++cyc) {
        const double t = cyc * 1e-3;
        mc.set_state(mahi::util::seconds(t), state, control, reference(t, dt, N));
        const auto r = mc.control_at_time(mahi::util::seconds(t));
        CHECK((int)r.u.size() == nu && std::isfinite(r.u[0]) && std::isfinite(r.u[1]));
        control = r.u;
        ++lookups;
        mahi::util::sleep(mahi::util::milliseconds(1));
    }
    mc.stop_calc();
    CHECK(lookups == 50);

    // batched form: B copies of the same problem give B copies of the same plan
    const int B = 37;
    BatchModelControl bmc(name, B, {10, 1, 5, 5}, {5, 5}, {1, 1}, opts);
    std::vector<double> states(B * nx, 0.0), controls(B * nu, 0.0), trajs;
    const auto ref1 = reference(0.1, dt, N);
    for (int b = 0; b < B; ++b) trajs.insert(trajs.end(), ref1.begin(), ref1.end());
    bmc.calc_u(states, controls, trajs);
    for (int b = 0; b < B; ++b) {
        CHECK(bmc.status()[b] == 0);
        const auto u0 = bmc.first_control(b);
        CHECK(std::fabs(u0[0] - u_free[0]) < 1e-6 && std::fabs(u0[1] - u_free[1]) < 1e-6);
    }
    // the same batch sharded over "several" devices (two shards on device 0: the parity box has one GPU): same plans
    BatchModelControl sharded(name, B, std::vector<int>{0, 0}, {10, 1, 5, 5}, {5, 5}, {1, 1}, opts);
    const std::vector<double>& Vs = sharded.calc_u(states, controls, trajs);
    const std::vector<double>& V1 = bmc.calc_u(states, controls, trajs);  // warm start from the first call: converges at once
    CHECK(Vs.size() == V1.size());
    for (size_t q = 0; q < Vs.size(); ++q) CHECK(std::fabs(Vs[q] - V1[q]) < 1e-6);
    for (int b = 0; b < B; ++b) CHECK(sharded.status()[b] == 0);
    sharded.update_weights({10, 1, 5, 5});
    sharded.reset_warm_start();
    bool threw = false;
    try { bmc.calc_u(states, controls, ref1); } catch (const std::invalid_argument&) { threw = true; }
    CHECK(threw);
    threw = false;
    try { ModelControl missing("no_such_model"); } catch (const std::exception&) { threw = true; }
    CHECK(threw);
    std::printf("FACADE-THREADED-OK\n");
    return 0;
}
