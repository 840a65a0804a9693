// mahi::mpc::ModelControl / BatchModelControl: the closed-loop host driver over the C ABI
// (reference: src/Mahi/Mpc/ModelControl.cpp:7-211).  The reference marshals std::vectors into casadi::nlpsol
// each cycle; here the same vectors go to mmpc_solve, which runs the CUDA kernels.
#include <Mahi/Mpc/ModelControl.hpp>

#include <fstream>
#include <sstream>
#include <stdexcept>

#include "../../../include/mahi_mpc_b200.h"

namespace mahi {
namespace mpc {

namespace {
void check(int rc, const char* what) {
    if (rc != MMPC_OK) throw std::runtime_error(std::string(what) + ": " + mmpc_last_error());
}

ModelParameters read_parameters(const std::string& model_name) {
    mahi::util::json j;  // same statements as the reference (ModelControl.cpp:23-26)
    std::ifstream file(model_name + ".json");
    if (!file) throw std::runtime_error("cannot open " + model_name + ".json");
    file >> j;
    return j["model"].get<ModelParameters>();
}

mmpc_options options_from(const casadi::Dict& d) {
    mmpc_options o;
    mmpc_options_default(&o);  // tol 1e-5, max_iter 200: the reference's hard-coded values (ModelControl.cpp:54-55)
    // The reference stores the constructor's solver_opts and never uses them (ModelControl.hpp:26,46 vs
    // ModelControl.cpp:52-62); here the IPOPT-named options that have a meaning for this solver are honoured.
    for (const auto& kv : d) {
        if (kv.first == "ipopt.tol" || kv.first == "tol") o.tol = kv.second.to_double();
        else if (kv.first == "ipopt.max_iter" || kv.first == "max_iter") o.max_iter = kv.second.to_int();
        else if (kv.first == "ipopt.mu_init" || kv.first == "mu_init") o.mu_init = kv.second.to_double();
        else if (kv.first == "warm_start_duals") o.warm_start_duals = kv.second.to_int();
        else if (kv.first == "kernel_variant") o.kernel_variant = kv.second.to_int();
        else if (kv.first == "device") continue;  // read by the constructor
        // extensions beyond the library's NLP: the MATLAB prototypes' integrator, Hessian and objective (mahi_mpc_b200.h)
        else if (kv.first == "integrator") o.integrator = kv.second.to_int();      // MMPC_EULER / MMPC_RK4
        else if (kv.first == "substeps") o.substeps = kv.second.to_int();
        else if (kv.first == "hessian") o.hessian = kv.second.to_int();            // MMPC_HESSIAN_EXACT / _GAUSS_NEWTON
        else if (kv.first == "cost") o.cost = kv.second.to_int();                  // MMPC_COST_SHOOTING / _NODE
    }
    return o;
}

void default_weights(const ModelParameters& p, std::vector<double>& Q, std::vector<double>& R, std::vector<double>& Rm) {
    // The generator initialises Q, R, Rm to identity before overwriting them from p (ModelGenerator.cpp:122-124);
    // the shipped drivers pass none or only Q, R, so missing weights fall back to ones.
    if (Q.empty()) Q.assign(p.num_x, 1.0);
    if (R.empty()) R.assign(p.num_u, 1.0);
    if (Rm.empty()) Rm.assign(p.num_u, 1.0);
    if (static_cast<int>(Q.size()) != p.num_x || static_cast<int>(R.size()) != p.num_u || static_cast<int>(Rm.size()) != p.num_u)
        throw std::invalid_argument("ModelControl: Q/R/Rm lengths must be num_x/num_u/num_u");
}
}  // namespace

// ------------------------------------------------------------------------------------------------
ModelControl::ModelControl(std::string model_name, std::vector<double> Q, std::vector<double> R, std::vector<double> Rm, casadi::Dict solver_opts)
    : m_solver_opts(std::move(solver_opts)), m_Q(std::move(Q)), m_R(std::move(R)), m_Rm(std::move(Rm)) {
    load_model(model_name);
}

ModelControl::~ModelControl() {
    stop_calc();
    release();
}

void ModelControl::release() {
    if (m_solver) mmpc_solver_free(m_solver);
    if (m_model) mmpc_model_free(m_model);
    m_solver = nullptr;
    m_model = nullptr;
}

void ModelControl::load_model(const std::string& model_name) {
    stop_calc();
    release();
    model_parameters = read_parameters(model_name);
    check(mmpc_model_load(model_name.c_str(), &m_model), "mmpc_model_load");
    const mmpc_options o = options_from(m_solver_opts);
    int device = 0;  // solver_opts["device"]: which GPU this controller solves on (one ModelControl per device is fine)
    for (const auto& kv : m_solver_opts) if (kv.first == "device") device = kv.second.to_int();
    check(mmpc_solver_create(m_model, device, 1, &o, &m_solver), "mmpc_solver_create");
    default_weights(model_parameters, m_Q, m_R, m_Rm);
    m_weights_dirty = true;
    int nv = 0;
    check(mmpc_model_dims(m_model, nullptr, nullptr, nullptr, &nv, nullptr, nullptr), "mmpc_model_dims");
    m_V.assign(nv, 0.0);
}

void ModelControl::set_state(mahi::util::Time time, const std::vector<double>& state, const std::vector<double>& control, std::vector<double> traj) {
    std::lock_guard<std::mutex> lg(m_state_mutex);
    m_time = time;
    m_state = state;
    m_control = control;
    m_traj = std::move(traj);
}

void ModelControl::start_calc() {
    stop_calc();
    m_stop = false;
    m_worker = std::thread([this] {  // the reference detaches this thread and spins in the destructor (ModelControl.cpp:16-19,87-111)
        std::vector<double> calc_times;
        while (!m_stop) {
            mahi::util::Clock timing_clock;
            mahi::util::Time local_time;
            std::vector<double> local_state, local_control, local_traj;
            {
                std::lock_guard<std::mutex> lg(m_state_mutex);
                local_time = m_time;
                local_state = m_state;
                local_control = m_control;
                local_traj = m_traj;
            }
            if (local_state.empty()) {  // set_state has not been called yet
                mahi::util::sleep(mahi::util::microseconds(100));
                continue;
            }
            try {
                calc_u(local_time, local_state, local_control, std::move(local_traj));
            } catch (const std::exception& e) {  // an exception escaping a std::thread body would call std::terminate
                {
                    std::lock_guard<std::mutex> lg(m_output_mutex);
                    m_worker_error = e.what();
                }
                std::cerr << "ModelControl: solver thread stopped: " << e.what() << std::endl;
                m_stop = true;
                break;
            }
            calc_times.push_back(timing_clock.get_elapsed_time().as_microseconds() * 1e-3);
        }
        std::cout << "Average calc time: " << mahi::util::mean(calc_times) << " ms" << std::endl;
    });
}

void ModelControl::stop_calc() {
    m_stop = true;
    if (m_worker.joinable()) m_worker.join();
}

void ModelControl::calc_u(mahi::util::Time control_time, const std::vector<double>& state, const std::vector<double>& control, std::vector<double> traj) {
    const int nx = model_parameters.num_x, nu = model_parameters.num_u, N = model_parameters.num_shooting_nodes;
    if (static_cast<int>(state.size()) != nx || static_cast<int>(control.size()) != nu || static_cast<int>(traj.size()) != N * nx)
        throw std::invalid_argument("ModelControl::calc_u: state/control/traj must have num_x / num_u / num_shooting_nodes*num_x entries");
    curr_time = control_time;
    {
        std::lock_guard<std::mutex> lg(m_weights_mutex);
        if (m_weights_dirty) {  // Q, R, Rm travel in p every call in the reference (ModelControl.cpp:120-122)
            check(mmpc_set_weights(m_solver, m_Q.data(), m_R.data(), m_Rm.data(), 0), "mmpc_set_weights");
            m_weights_dirty = false;
        }
    }
    {
        std::lock_guard<std::mutex> lg(m_control_limits_mutex);  // u bounds are re-stamped every call (ModelControl.cpp:148-154)
        if (m_limits_dirty) {
            check(mmpc_set_bounds(m_solver, nullptr, nullptr, model_parameters.u_min.data(), model_parameters.u_max.data()), "mmpc_set_bounds");
            m_limits_dirty = false;
        }
    }
    mmpc_stats st;
    // x_0 pinned to `state`, u_init = `control`, A/B/xdot for linear models evaluated on the device (ModelControl.cpp:125-145)
    check(mmpc_solve(m_solver, state.data(), control.data(), traj.data(), m_V.data(), nullptr, &st, MMPC_MEM_HOST, nullptr), "mmpc_solve");
    {
        std::lock_guard<std::mutex> lg(m_output_mutex);  // read by last_solve_info() from other threads
        m_info.iterations = st.iterations;
        m_info.status = st.status;
        m_info.objective = st.objective;
        m_info.kkt_error = st.kkt_error;
        m_V_out = m_V;
    }
    format_outputs(m_V);  // the primal warm start (ModelControl.cpp:161) is kept inside the solver handle
}

void ModelControl::format_outputs(const std::vector<double>& opt_output) {
    std::vector<ControlResult> local;
    const int nx = model_parameters.num_x, nu = model_parameters.num_u;
    for (int i = 0; i < model_parameters.num_shooting_nodes; ++i) {
        const auto first = opt_output.begin() + static_cast<std::ptrdiff_t>(i) * (nx + nu);
        local.emplace_back(mahi::util::seconds(curr_time.as_seconds() + model_parameters.step_size.as_seconds() * i),
                           std::vector<double>(first, first + nx), std::vector<double>(first + nx, first + nx + nu));
    }
    std::lock_guard<std::mutex> lg(m_output_mutex);
    control_results = std::move(local);
}

ModelControl::ControlResult ModelControl::control_at_time(mahi::util::Time time) {
    std::lock_guard<std::mutex> lg(m_output_mutex);
    if (control_results.empty()) throw std::runtime_error("ModelControl::control_at_time: no plan yet (call calc_u first)");
    // zero-order hold over the last plan: the last entry whose time is < `time`, the first entry if none
    // (ModelControl.cpp:192-197, with the size check made before the element access)
    size_t i = 0;
    while (i < control_results.size() && control_results[i].time < time) ++i;
    return i == 0 ? control_results[0] : control_results[i - 1];
}

void ModelControl::update_weights(std::vector<double> Q, std::vector<double> R, std::vector<double> Rm) {
    std::lock_guard<std::mutex> lg(m_weights_mutex);
    if (!Q.empty()) m_Q = std::move(Q);
    if (!R.empty()) m_R = std::move(R);
    if (!Rm.empty()) m_Rm = std::move(Rm);
    default_weights(model_parameters, m_Q, m_R, m_Rm);
    m_weights_dirty = true;
}

void ModelControl::update_control_limits(std::vector<double> u_min, std::vector<double> u_max) {
    if (static_cast<int>(u_min.size()) != model_parameters.num_u || static_cast<int>(u_max.size()) != model_parameters.num_u)
        throw std::invalid_argument("update_control_limits: need num_u entries");
    std::lock_guard<std::mutex> lg(m_control_limits_mutex);
    model_parameters.u_min = std::move(u_min);
    model_parameters.u_max = std::move(u_max);
    m_limits_dirty = true;
}

// ------------------------------------------------------------------------------------------------
BatchModelControl::BatchModelControl(std::string model_name, int batch, std::vector<double> Q, std::vector<double> R, std::vector<double> Rm,
                                     casadi::Dict solver_opts, int device)
    : m_batch(batch) {
    model_parameters = read_parameters(model_name);
    check(mmpc_model_load(model_name.c_str(), &m_model), "mmpc_model_load");
    const mmpc_options o = options_from(solver_opts);
    const int rc = mmpc_solver_create(m_model, device, batch, &o, &m_solver);
    if (rc != MMPC_OK) {
        mmpc_model_free(m_model);
        check(rc, "mmpc_solver_create");
    }
    default_weights(model_parameters, Q, R, Rm);
    check(mmpc_set_weights(m_solver, Q.data(), R.data(), Rm.data(), 0), "mmpc_set_weights");
    check(mmpc_model_dims(m_model, nullptr, nullptr, nullptr, &m_nv, nullptr, nullptr), "mmpc_model_dims");
    m_V.assign(static_cast<size_t>(batch) * m_nv, 0.0);
    m_iters.assign(batch, 0);
    m_status.assign(batch, 0);
}

BatchModelControl::BatchModelControl(std::string model_name, int batch, std::vector<int> devices, std::vector<double> Q, std::vector<double> R,
                                     std::vector<double> Rm, casadi::Dict solver_opts)
    : m_batch(batch) {
    model_parameters = read_parameters(model_name);
    check(mmpc_model_load(model_name.c_str(), &m_model), "mmpc_model_load");
    const mmpc_options o = options_from(solver_opts);
    const int rc = mmpc_group_create(m_model, devices.data(), static_cast<int>(devices.size()), batch, &o, &m_group);
    if (rc != MMPC_OK) {
        mmpc_model_free(m_model);
        check(rc, "mmpc_group_create");
    }
    default_weights(model_parameters, Q, R, Rm);
    check(mmpc_group_set_weights(m_group, Q.data(), R.data(), Rm.data()), "mmpc_group_set_weights");
    check(mmpc_model_dims(m_model, nullptr, nullptr, nullptr, &m_nv, nullptr, nullptr), "mmpc_model_dims");
    m_V.assign(static_cast<size_t>(batch) * m_nv, 0.0);
    m_iters.assign(batch, 0);
    m_status.assign(batch, 0);
}

BatchModelControl::~BatchModelControl() {
    if (m_group) mmpc_group_free(m_group);
    if (m_solver) mmpc_solver_free(m_solver);
    if (m_model) mmpc_model_free(m_model);
}

const std::vector<double>& BatchModelControl::calc_u(const std::vector<double>& states, const std::vector<double>& controls,
                                                     const std::vector<double>& trajs) {
    const size_t B = m_batch, nx = model_parameters.num_x, nu = model_parameters.num_u, N = model_parameters.num_shooting_nodes;
    if (states.size() != B * nx || controls.size() != B * nu || trajs.size() != B * N * nx)
        throw std::invalid_argument("BatchModelControl::calc_u: need batch*num_x, batch*num_u, batch*num_shooting_nodes*num_x entries");
    std::vector<mmpc_stats> st(B);
    if (m_group) check(mmpc_group_solve(m_group, states.data(), controls.data(), trajs.data(), m_V.data(), nullptr, st.data()), "mmpc_group_solve");
    else check(mmpc_solve(m_solver, states.data(), controls.data(), trajs.data(), m_V.data(), nullptr, st.data(), MMPC_MEM_HOST, nullptr), "mmpc_solve");
    for (size_t b = 0; b < B; ++b) {
        m_iters[b] = st[b].iterations;
        m_status[b] = st[b].status;
    }
    return m_V;
}

std::vector<double> BatchModelControl::first_control(int b) const {
    const size_t nx = model_parameters.num_x, nu = model_parameters.num_u;
    const auto first = m_V.begin() + static_cast<std::ptrdiff_t>(b) * m_nv + nx;
    return std::vector<double>(first, first + nu);
}

void BatchModelControl::update_weights(std::vector<double> Q, std::vector<double> R, std::vector<double> Rm) {
    const double *q = Q.empty() ? nullptr : Q.data(), *r = R.empty() ? nullptr : R.data(), *rm = Rm.empty() ? nullptr : Rm.data();
    if (m_group) check(mmpc_group_set_weights(m_group, q, r, rm), "mmpc_group_set_weights");
    else check(mmpc_set_weights(m_solver, q, r, rm, 0), "mmpc_set_weights");
}

void BatchModelControl::update_control_limits(std::vector<double> u_min, std::vector<double> u_max) {
    if (static_cast<int>(u_min.size()) != model_parameters.num_u || static_cast<int>(u_max.size()) != model_parameters.num_u)
        throw std::invalid_argument("update_control_limits: need num_u entries");
    if (m_group) check(mmpc_group_set_bounds(m_group, nullptr, nullptr, u_min.data(), u_max.data()), "mmpc_group_set_bounds");
    else check(mmpc_set_bounds(m_solver, nullptr, nullptr, u_min.data(), u_max.data()), "mmpc_set_bounds");
    model_parameters.u_min = std::move(u_min);
    model_parameters.u_max = std::move(u_max);
}

void BatchModelControl::reset_warm_start() {
    if (!m_group) { check(mmpc_reset_warm_start(m_solver), "mmpc_reset_warm_start"); return; }
    int n = 0;
    check(mmpc_group_size(m_group, &n, nullptr), "mmpc_group_size");
    for (int i = 0; i < n; ++i) {
        mmpc_solver* s = nullptr;
        check(mmpc_group_shard(m_group, i, nullptr, nullptr, nullptr, &s), "mmpc_group_shard");
        check(mmpc_reset_warm_start(s), "mmpc_reset_warm_start");
    }
}

}  // namespace mpc
}  // namespace mahi
