// mahi-mpc-b200 :: the MATLAB prototype's closed loop (reference: workingMPC_8_30_21.m) through the C ABI.
//
// The prototype differs from the C++ library's NLP in three places, all available as solver options here:
//   integrator   classical RK4 with 4 sub-steps per shooting interval (workingMPC_8_30_21.m:46-59)      -> MMPC_RK4, substeps = 4
//   Hessian      CasADi differentiates twice through that integrator (:109-112, opti + sqpmethod)       -> MMPC_HESSIAN_EXACT
//   objective    1000 (x_k - r_k)' Q (x_k - r_k) + u_k' R u_k on the decision variables (:90-106)        -> MMPC_COST_NODE,
//                                                                                                          Q <- 1000 Q, R <- 0, Rm <- R
// T = 1 s, N = 10 (dt = 0.1 s), reference r(t) = pi/2 (sin t, sin t, cos t, cos t), start at rest, 10 N closed-loop steps with the
// first control of every plan applied to the same RK4 x 4 plant (:130-150, without the prototype's random disturbance).
// Prints the tracking error of the first joint angle along the loop.      usage: ex_prototype_mpc [cycles]
#include <mahi_mpc_b200.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

static void check(int rc, const char* what) {
    if (rc != MMPC_OK) {
        std::fprintf(stderr, "%s: %s\n", what, mmpc_last_error());
        std::exit(1);
    }
}

int main(int argc, char** argv) {
    const int N = 10, nx = 4, nu = 2, cycles = argc > 1 ? std::atoi(argv[1]) : 10 * N;
    const double T = 1.0, dt = T / N, half_pi = 1.5707963267948966;
    mmpc_model_params mp{};
    mp.name = "double_pendulum";
    mp.timespan_us = (long long)(T * 1e6);
    mp.step_size_us = (long long)(dt * 1e6);
    mp.num_x = nx; mp.num_u = nu; mp.num_shooting_nodes = N;
    mp.dll_filepath = "builtin:double_pendulum";  // the same equations of motion as workingMPC_8_30_21.m:27-38
    mmpc_model_t model;
    check(mmpc_model_create(&mp, &model), "mmpc_model_create");
    mmpc_options o;
    check(mmpc_options_default(&o), "mmpc_options_default");
    o.integrator = MMPC_RK4; o.substeps = 4; o.hessian = MMPC_HESSIAN_EXACT; o.cost = MMPC_COST_NODE;
    o.tol = 1e-8;
    mmpc_solver_t solver;
    check(mmpc_solver_create(model, 0, 1, &o, &solver), "mmpc_solver_create");
    const double Q[4] = {1000, 1000, 1000, 1000}, R[2] = {0, 0}, Rm[2] = {1, 1};
    check(mmpc_set_weights(solver, Q, R, Rm, 0), "mmpc_set_weights");
    int NV = 0;
    check(mmpc_model_dims(model, nullptr, nullptr, nullptr, &NV, nullptr, nullptr), "mmpc_model_dims");
    std::vector<double> x(nx, 0.0), u(nu, 0.0), traj(N * nx), V(NV), zero(nx, 0.0), F(nx), Fx(nx * nx), Fu(nx * nu);
    auto ref = [&](double t, double* r) { r[0] = r[1] = half_pi * std::sin(t); r[2] = r[3] = half_pi * std::cos(t); };
    double t0 = 0.0, err_max = 0.0;
    int it_sum = 0;
    for (int c = 0; c < cycles; ++c) {
        for (int k = 0; k < N; ++k) ref(t0 + (k + 1) * dt, &traj[k * nx]);  // r_k is the reference of x_{k+1} (MMPC_COST_NODE)
        mmpc_stats st;
        check(mmpc_solve(solver, x.data(), u.data(), traj.data(), V.data(), nullptr, &st, MMPC_MEM_HOST, nullptr), "mmpc_solve");
        if (st.status != MMPC_SOLVE_CONVERGED) std::fprintf(stderr, "cycle %d: status %d after %d iterations\n", c, st.status, st.iterations);
        it_sum += st.iterations;
        u[0] = V[nx]; u[1] = V[nx + 1];  // first control of the plan
        // plant: the same RK4 x 4 interval map, evaluated by the node-evaluation entry point (g = F(x, u) - 0)
        check(mmpc_eval_nodes(model, 0, &o, 1, x.data(), u.data(), zero.data(), nullptr, F.data(), Fx.data(), Fu.data(), nullptr, MMPC_MEM_HOST, nullptr),
              "mmpc_eval_nodes");
        x = F;
        t0 += dt;
        double r[4];
        ref(t0, r);
        err_max = std::fmax(err_max, c >= N ? std::fabs(x[0] - r[0]) : 0.0);
        if (c % N == N - 1) std::printf("t = %5.2f s   qA = %+.4f (reference %+.4f)   u = (%+.3f, %+.3f)   Newton iterations %d\n", t0, x[0], r[0], u[0], u[1], st.iterations);
    }
    std::printf("closed loop of %d cycles: mean Newton iterations %.2f, max |qA - reference| after the first second %.3e rad\n", cycles,
                (double)it_sum / cycles, err_max);
    mmpc_solver_free(solver);
    mmpc_model_free(model);
    return 0;
}
