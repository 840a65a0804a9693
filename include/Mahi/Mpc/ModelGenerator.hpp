// mahi-mpc-b200 :: mahi::mpc::ModelGenerator -- same class, constructor and method names as the reference
// (include/Mahi/Mpc/ModelGenerator.hpp:10-31).  What each step does here:
//   create_model()      derives A = d x_dot/dx, B = d x_dot/du and the Hessian contraction symbolically and
//                       emits + compiles <name>_linear_functions.so (CPU, CasADi ABI), as ModelGenerator.cpp:23-233,55
//   generate_c_code()   writes <name>.cu: the generated device functor + the sm_100a kernel instantiation
//                       (the reference writes <name>.c through CasADi, ModelGenerator.cpp:235-239)
//   compile_model()     nvcc -gencode arch=compute_100a,code=sm_100a -> <name>.so, then <name>.json
//                       (the reference runs gcc -O1, ModelGenerator.cpp:253-259)
#pragma once
#include <Mahi/Mpc/ModelParameters.hpp>
#include <Mahi/Util.hpp>
#include <casadi/casadi.hpp>

namespace mahi {
namespace mpc {

class ModelGenerator {
public:
    ModelGenerator(ModelParameters model_parameters, casadi::SX x, casadi::SX x_dot, casadi::SX u);
    ~ModelGenerator();
    void create_model();
    void generate_c_code();
    /// emits + compiles <name>_linear_functions.so from the three symbolic functions (reference signature,
    /// ModelGenerator.hpp:27; create_model() calls it with jacobian(x_dot, x), jacobian(x_dot, u) and x_dot)
    void generate_linear_functions(casadi::Function get_A, casadi::Function get_B, casadi::Function get_x_dot_init);
    void compile_model();
    void save_param_file();

    const ModelParameters& parameters() const { return m_model_parameters; }

private:
    ModelParameters m_model_parameters;
    casadi::SX m_x, m_x_dot, m_u;
    std::string m_c_file_filepath;
    bool m_created = false;
};

}  // namespace mpc
}  // namespace mahi
