// mahi-mpc-b200 :: model code generation
//
// Replaces what ModelGenerator asks CasADi for (src/Mahi/Mpc/ModelGenerator.cpp:235-251):
//   generate_dependencies("<name>.c")  ->  emit_cuda_model(): a device functor with f, (f,A,B) and
//                                          (f,A,B,sum_i w_i grad^2 f_i) as straight-line fp64 code
//   CodeGenerator("<name>_linear_functions.c") -> emit_linear_functions_c(): <name>_get_A/_get_B/
//                                          _get_x_dot_init as CPU C with CasADi's generated-code ABI
//                                          (src/codegen_usage.cpp:61-70,118-132)
#pragma once
#include <string>
#include <vector>

#include "../expr/sx.hpp"

namespace mahi {
namespace mpc {
namespace codegen {

struct ModelExpr {
    std::string name;
    sym::SX x, x_dot, u;
};

/// CUDA/C++ header text defining `struct mmpc_model_<name>` (see csrc/kernels/model_concept.md).
std::string emit_cuda_model(const ModelExpr& m);

/// Translation unit that instantiates every kernel for the model and exports its vtable.
std::string emit_cuda_unit(const std::string& name, const std::string& model_header);

/// A named symbolic function (what casadi::Function(name, {in...}, {out...}, {in names}, {out names}) holds).
struct SymFunction {
    std::string name;
    std::vector<sym::SX> in, out;
    std::vector<std::string> in_names, out_names;
};

/// C source exporting every function of `fns` with CasADi's generated-code ABI (what casadi::CodeGenerator emits,
/// ModelGenerator.cpp:241-246): dense inputs, outputs as the structural non-zeros in CCS order.
std::string emit_casadi_abi_c(const std::string& banner, const std::vector<SymFunction>& fns);

/// C source of <name>_linear_functions (CasADi ABI).
std::string emit_linear_functions_c(const ModelExpr& m);

/// Built-in model definitions (double_pendulum: examples/ex_model_generate.cpp:24-43; arm4: builder-defined).
bool builtin_model(const std::string& which, ModelExpr* out);

}  // namespace codegen
}  // namespace mpc
}  // namespace mahi
