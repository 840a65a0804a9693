// mahi-mpc-b200 :: mahi::mpc::ModelParameters -- field-for-field the reference's struct
// (include/Mahi/Mpc/ModelParameters.hpp:11-28) and its JSON artefact (src/Mahi/Mpc/ModelParameters.cpp:37-72).
#pragma once
#include <Mahi/Util.hpp>
#include <string>
#include <vector>

namespace mahi {
namespace mpc {

struct ModelParameters {
    ModelParameters(std::string name_, int num_x_, int num_u_, mahi::util::Time step_size_, size_t num_shooting_nodes_, bool is_linear_,
                    std::vector<double> u_min_ = {}, std::vector<double> u_max_ = {}, std::vector<double> x_min_ = {},
                    std::vector<double> x_max_ = {});
    ModelParameters() {}

    std::string name;            // model name; artefacts are "<name>.json", "<name>.so", "<name>_linear_functions.so"
    mahi::util::Time timespan;   // step_size * num_shooting_nodes
    mahi::util::Time step_size;  // shooting interval
    int num_x = 0;
    int num_u = 0;
    int num_shooting_nodes = 0;
    std::vector<double> x_min, u_min, x_max, u_max;  // +-10e30 = "no bound"
    std::string dll_filepath;    // the generated model library (here: CUDA kernels for sm_100a)
    bool is_linear = false;
};

struct TrajectoryParameters {  // declared by the reference, unused by the library (ModelParameters.hpp:30-41)
    TrajectoryParameters(std::string name_t_, int num_x_t_, int num_u_t_, mahi::util::Time step_size_t_, size_t num_shooting_nodes_);
    TrajectoryParameters() {}
    std::string name_t;
    mahi::util::Time timespan_t, step_size_t;
    int num_x_t = 0, num_u_t = 0, num_shooting_nodes_t = 0;
};

/// nlohmann-style (de)serialisation hooks, found by ADL from mahi::util::json (reference: ModelParameters.hpp:43-45,
/// ModelParameters.cpp:37-72): `j["model"] = params;` and `j["model"].get<ModelParameters>()`.
void to_json(mahi::util::json& j, const ModelParameters& p);
void from_json(const mahi::util::json& j, ModelParameters& p);

/// {"model": {...}} text in the reference's schema; times in microseconds.
std::string to_json_string(const ModelParameters& p);
/// Parses the text produced by to_json_string / by the reference (nlohmann). Throws std::runtime_error.
ModelParameters model_parameters_from_json_string(const std::string& text);

}  // namespace mpc
}  // namespace mahi
