// mahi-mpc-b200 :: mahi::mpc::ModelControl -- same public interface as the reference
// (include/Mahi/Mpc/ModelControl.hpp:13-80); every solve goes through the C ABI (include/mahi_mpc_b200.h) to the
// CUDA kernels.  BatchModelControl is the batched form of the same call (B independent calc_u problems per
// launch), which is what the GPU is for.
#pragma once
#include <Mahi/Mpc/ModelParameters.hpp>
#include <Mahi/Util.hpp>
#include <casadi/casadi.hpp>

#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct mmpc_model;
struct mmpc_solver;
struct mmpc_group;

namespace mahi {
namespace mpc {

class ModelControl {
public:
    struct ControlResult {
        ControlResult(mahi::util::Time time_, std::vector<double> x_est_, std::vector<double> u_) : time(time_), x_est(x_est_), u(u_) {}
        mahi::util::Time time;
        std::vector<double> x_est;
        std::vector<double> u;
    };
    /// Per-solve diagnostics the reference never looks at (it ignores IPOPT's return status, ModelControl.cpp:159-160).
    struct SolveInfo {
        int iterations = 0, status = 0;
        double objective = 0.0, kkt_error = 0.0;
    };

    ModelControl(std::string model_name, std::vector<double> Q = {}, std::vector<double> R = {}, std::vector<double> Rm = {},
                 casadi::Dict solver_opts = casadi::Dict());
    ~ModelControl();
    ModelControl(const ModelControl&) = delete;
    ModelControl& operator=(const ModelControl&) = delete;

    ModelParameters model_parameters;
    std::vector<ControlResult> control_results;

    void calc_u(mahi::util::Time time, const std::vector<double>& state, const std::vector<double>& control, std::vector<double> traj);
    void load_model(const std::string& model_name);
    ControlResult control_at_time(mahi::util::Time time);
    void start_calc();
    void stop_calc();
    void set_state(mahi::util::Time time, const std::vector<double>& state, const std::vector<double>& control, std::vector<double> traj);
    void update_weights(std::vector<double> Q = {}, std::vector<double> R = {}, std::vector<double> Rm = {});
    void update_control_limits(std::vector<double> u_min, std::vector<double> u_max);

    SolveInfo last_solve_info() const { std::lock_guard<std::mutex> lg(m_output_mutex); return m_info; }
    /// the full decision vector V of the last solve, reference layout (ModelGenerator.cpp:76-112); a copy, because the
    /// solver thread of start_calc() overwrites it every cycle
    std::vector<double> last_solution() const { std::lock_guard<std::mutex> lg(m_output_mutex); return m_V_out; }
    /// message of the exception that stopped the start_calc() thread, empty if it is healthy (the reference's thread
    /// would have called std::terminate)
    std::string worker_error() const { std::lock_guard<std::mutex> lg(m_output_mutex); return m_worker_error; }

private:
    void format_outputs(const std::vector<double>& opt_output);
    void release();

    casadi::Dict m_solver_opts;
    mahi::util::Time curr_time;
    mmpc_model* m_model = nullptr;
    mmpc_solver* m_solver = nullptr;
    std::vector<double> m_Q, m_R, m_Rm, m_V, m_V_out;
    std::string m_worker_error;
    bool m_weights_dirty = true, m_limits_dirty = false;
    SolveInfo m_info;

    std::atomic<bool> m_stop{false};
    std::thread m_worker;
    mahi::util::Time m_time;
    std::vector<double> m_state, m_control, m_traj;
    mutable std::mutex m_state_mutex, m_output_mutex, m_control_limits_mutex, m_weights_mutex;
};

/// B independent ModelControl::calc_u problems per call (one kernel launch); instance-major flat vectors.
class BatchModelControl {
public:
    BatchModelControl(std::string model_name, int batch, std::vector<double> Q = {}, std::vector<double> R = {},
                      std::vector<double> Rm = {}, casadi::Dict solver_opts = casadi::Dict(), int device = 0);
    /// The batch sharded over several GPUs of one box (contiguous shards, one solver handle per device, no collective on
    /// the solve path; calc_u gathers every shard's plans into the same flat vector).
    BatchModelControl(std::string model_name, int batch, std::vector<int> devices, std::vector<double> Q = {}, std::vector<double> R = {},
                      std::vector<double> Rm = {}, casadi::Dict solver_opts = casadi::Dict());
    ~BatchModelControl();
    BatchModelControl(const BatchModelControl&) = delete;
    BatchModelControl& operator=(const BatchModelControl&) = delete;

    ModelParameters model_parameters;
    int batch() const { return m_batch; }
    /// states[B*nx], controls[B*nu], trajs[B*N*nx] -> V[B*NV] (reference layout per instance)
    const std::vector<double>& calc_u(const std::vector<double>& states, const std::vector<double>& controls, const std::vector<double>& trajs);
    /// first control of instance b's plan: u_0 = V[b*NV + nx .. nx+nu)
    std::vector<double> first_control(int b) const;
    const std::vector<int>& iterations() const { return m_iters; }
    const std::vector<int>& status() const { return m_status; }
    void update_weights(std::vector<double> Q = {}, std::vector<double> R = {}, std::vector<double> Rm = {});
    void update_control_limits(std::vector<double> u_min, std::vector<double> u_max);
    void reset_warm_start();

private:
    mmpc_model* m_model = nullptr;
    mmpc_solver* m_solver = nullptr;
    mmpc_group* m_group = nullptr;  // several devices: shards behind one group handle
    int m_batch, m_nv = 0;
    std::vector<double> m_V;
    std::vector<int> m_iters, m_status;
};

}  // namespace mpc
}  // namespace mahi
