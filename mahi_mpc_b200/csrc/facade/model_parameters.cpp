// mahi::mpc::ModelParameters: constructor defaults and the "<name>.json" artefact
// (reference: src/Mahi/Mpc/ModelParameters.cpp:7-72).
#include <Mahi/Mpc/ModelParameters.hpp>

#include <sstream>
#include <stdexcept>

#include "../abi/json_min.hpp"

namespace mahi {
namespace mpc {

static const double kNoBound = 10e30;  // the reference's "unbounded" sentinel

ModelParameters::ModelParameters(std::string name_, int num_x_, int num_u_, mahi::util::Time step_size_, size_t num_shooting_nodes_,
                                 bool is_linear_, std::vector<double> u_min_, std::vector<double> u_max_, std::vector<double> x_min_,
                                 std::vector<double> x_max_)
    : name(std::move(name_)), step_size(step_size_), num_x(num_x_), num_u(num_u_), num_shooting_nodes(static_cast<int>(num_shooting_nodes_)),
      x_min(std::move(x_min_)), u_min(std::move(u_min_)), x_max(std::move(x_max_)), u_max(std::move(u_max_)), is_linear(is_linear_) {
    timespan = mahi::util::microseconds(step_size.as_microseconds() * num_shooting_nodes);
    if (x_min.empty()) x_min.assign(num_x, -kNoBound);
    if (x_max.empty()) x_max.assign(num_x, kNoBound);
    if (u_min.empty()) u_min.assign(num_u, -kNoBound);
    if (u_max.empty()) u_max.assign(num_u, kNoBound);
}

TrajectoryParameters::TrajectoryParameters(std::string name_t_, int num_x_t_, int num_u_t_, mahi::util::Time step_size_t_, size_t nodes)
    : name_t(std::move(name_t_)), step_size_t(step_size_t_), num_x_t(num_x_t_), num_u_t(num_u_t_), num_shooting_nodes_t(static_cast<int>(nodes)) {
    timespan_t = mahi::util::microseconds(step_size_t.as_microseconds() * num_shooting_nodes_t);
}

void to_json(mahi::util::json& j, const ModelParameters& p) {
    j = mahi::util::json{{"name", p.name},
                         {"timespan", p.timespan.as_microseconds()},
                         {"step_size", p.step_size.as_microseconds()},
                         {"num_x", p.num_x},
                         {"num_u", p.num_u},
                         {"num_shooting_nodes", p.num_shooting_nodes},
                         {"x_min", p.x_min},
                         {"u_min", p.u_min},
                         {"x_max", p.x_max},
                         {"u_max", p.u_max},
                         {"dll_filepath", p.dll_filepath},
                         {"is_linear", p.is_linear}};
}

void from_json(const mahi::util::json& j, ModelParameters& p) {
    j.at("name").get_to(p.name);
    p.timespan = mahi::util::microseconds(j.at("timespan").get<std::int64_t>());
    p.step_size = mahi::util::microseconds(j.at("step_size").get<std::int64_t>());
    j.at("num_x").get_to(p.num_x);
    j.at("num_u").get_to(p.num_u);
    j.at("num_shooting_nodes").get_to(p.num_shooting_nodes);
    j.at("x_min").get_to(p.x_min);
    j.at("u_min").get_to(p.u_min);
    j.at("x_max").get_to(p.x_max);
    j.at("u_max").get_to(p.u_max);
    j.at("dll_filepath").get_to(p.dll_filepath);
    j.at("is_linear").get_to(p.is_linear);
    // the reference maps the sentinel to +-inf for x only (ModelParameters.cpp:66-69)
    for (size_t i = 0; i < p.x_min.size(); ++i) {
        if (p.x_min[i] == -kNoBound) p.x_min[i] = -std::numeric_limits<double>::infinity();
        if (p.x_max[i] == kNoBound) p.x_max[i] = std::numeric_limits<double>::infinity();
    }
}

static std::string vec_json(const std::vector<double>& v) {
    std::ostringstream os;
    os << "[";
    for (size_t i = 0; i < v.size(); ++i) os << (i ? "," : "") << mmpc_json::number(v[i]);
    os << "]";
    return os.str();
}

std::string to_json_string(const ModelParameters& p) {
    std::ostringstream os;  // keys in the order nlohmann emits them (sorted)
    os << "{\"model\":{"
       << "\"dll_filepath\":" << mmpc_json::quote(p.dll_filepath) << ",\"is_linear\":" << (p.is_linear ? "true" : "false")
       << ",\"name\":" << mmpc_json::quote(p.name) << ",\"num_shooting_nodes\":" << p.num_shooting_nodes << ",\"num_u\":" << p.num_u
       << ",\"num_x\":" << p.num_x << ",\"step_size\":" << p.step_size.as_microseconds() << ",\"timespan\":" << p.timespan.as_microseconds()
       << ",\"u_max\":" << vec_json(p.u_max) << ",\"u_min\":" << vec_json(p.u_min) << ",\"x_max\":" << vec_json(p.x_max)
       << ",\"x_min\":" << vec_json(p.x_min) << "}}";
    return os.str();
}

ModelParameters model_parameters_from_json_string(const std::string& text) {
    mmpc_json::Value root;
    std::string err;
    if (!mmpc_json::parse(text, &root, &err)) throw std::runtime_error("model json: " + err);
    const mmpc_json::Value* j = root.get("model");
    if (!j) throw std::runtime_error("model json: missing \"model\"");
    auto need = [&](const char* k) -> const mmpc_json::Value& {
        const mmpc_json::Value* v = j->get(k);
        if (!v) throw std::runtime_error(std::string("model json: key '") + k + "' not found");
        return *v;
    };
    auto vec = [&](const char* k) {
        std::vector<double> out;
        for (const auto& e : need(k).array) out.push_back(e.number);
        return out;
    };
    ModelParameters p;
    p.name = need("name").string;
    p.timespan = mahi::util::microseconds(static_cast<std::int64_t>(need("timespan").number));
    p.step_size = mahi::util::microseconds(static_cast<std::int64_t>(need("step_size").number));
    p.num_x = static_cast<int>(need("num_x").number);
    p.num_u = static_cast<int>(need("num_u").number);
    p.num_shooting_nodes = static_cast<int>(need("num_shooting_nodes").number);
    p.x_min = vec("x_min");
    p.u_min = vec("u_min");
    p.x_max = vec("x_max");
    p.u_max = vec("u_max");
    p.dll_filepath = need("dll_filepath").string;
    p.is_linear = need("is_linear").boolean;
    // the reference maps the sentinel to +-inf for x only (ModelParameters.cpp:66-69)
    for (size_t i = 0; i < p.x_min.size(); ++i) {
        if (p.x_min[i] == -kNoBound) p.x_min[i] = -std::numeric_limits<double>::infinity();
        if (p.x_max[i] == kNoBound) p.x_max[i] = std::numeric_limits<double>::infinity();
    }
    return p;
}

}  // namespace mpc
}  // namespace mahi
