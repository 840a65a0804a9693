// mahi-mpc-b200 :: the slice of the CasADi C++ API that user code of mahi::mpc touches, mapped onto the
// framework's own symbolic engine (csrc/expr/sx.hpp).  CasADi itself (pinned fadc8644..., CMakeLists.txt:44)
// is not installable offline.  Covered: SX (sym, vertcat, arithmetic, sin/cos/..., size1), jacobian, mtimes,
// inf, Dict, DM, and `external(name, lib)` for the generated <name>_linear_functions.so
// (examples/model_control_example.cpp:46,81-82).
#pragma once
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "../../mahi_mpc_b200/csrc/expr/sx.hpp"

namespace casadi {

using SX = mahi::mpc::sym::SX;
using mahi::mpc::sym::cos;
using mahi::mpc::sym::sin;
using mahi::mpc::sym::tan;
using mahi::mpc::sym::exp;
using mahi::mpc::sym::log;
using mahi::mpc::sym::sqrt;
using mahi::mpc::sym::tanh;
using mahi::mpc::sym::atan;
using mahi::mpc::sym::fabs;
using mahi::mpc::sym::pow;
using mahi::mpc::sym::sq;
using mahi::mpc::sym::mtimes;
using mahi::mpc::sym::jacobian;
using mahi::mpc::sym::gradient;
using mahi::mpc::sym::hessian;
using mahi::mpc::sym::substitute;

constexpr double inf = std::numeric_limits<double>::infinity();

/// Option value: number or string (casadi::GenericType stand-in for the solver options actually used,
/// ModelControl.cpp:54-59).
class GenericType {
public:
    GenericType() : m_num(0), m_is_str(false) {}
    GenericType(double v) : m_num(v), m_is_str(false) {}  // NOLINT
    GenericType(int v) : m_num(v), m_is_str(false) {}     // NOLINT
    GenericType(const char* s) : m_num(0), m_str(s), m_is_str(true) {}  // NOLINT
    GenericType(const std::string& s) : m_num(0), m_str(s), m_is_str(true) {}  // NOLINT
    bool is_string() const { return m_is_str; }
    double to_double() const { return m_num; }
    int to_int() const { return static_cast<int>(m_num); }
    const std::string& to_string() const { return m_str; }
private:
    double m_num;
    std::string m_str;
    bool m_is_str;
};
using Dict = std::map<std::string, GenericType>;

/// Dense numeric matrix, column-major (casadi::DM stand-in); converts to/from std::vector<double>.
class DM {
public:
    DM() : m_rows(0), m_cols(0) {}
    DM(double v) : m_rows(1), m_cols(1), m_d{v} {}  // NOLINT
    DM(const std::vector<double>& v) : m_rows(static_cast<int>(v.size())), m_cols(1), m_d(v) {}  // NOLINT
    DM(int rows, int cols) : m_rows(rows), m_cols(cols), m_d(static_cast<size_t>(rows) * cols, 0.0) {}
    int size1() const { return m_rows; }
    int size2() const { return m_cols; }
    double& operator()(int i, int j = 0) { return m_d[static_cast<size_t>(j) * m_rows + i]; }
    double operator()(int i, int j = 0) const { return m_d[static_cast<size_t>(j) * m_rows + i]; }
    operator std::vector<double>() const { return m_d; }  // NOLINT: dense, column-major (ModelControl.cpp:127-129)
    const std::vector<double>& nonzeros() const { return m_d; }
private:
    int m_rows, m_cols;
    std::vector<double> m_d;
};

/// casadi::Function stand-in.  Two kinds, as the reference uses them:
///  * symbolic: Function(name, {in...}, {out...}, {in names}, {out names}) over SX expressions
///    (ModelGenerator.cpp:51-53 builds <name>_get_A/_get_B/_get_x_dot_init this way and hands them to
///    generate_linear_functions); callable numerically and convertible to generated C (codegen::emit_casadi_abi_c);
///  * external: loaded from a shared library with CasADi's generated-code C ABI (src/codegen_usage.cpp:61-70), see external().
class Function {
public:
    Function() = default;
    Function(const std::string& name, const std::vector<SX>& in, const std::vector<SX>& out, const std::vector<std::string>& in_names = {},
             const std::vector<std::string>& out_names = {})
        : m_name(name), m_in(in), m_out(out), m_in_names(in_names), m_out_names(out_names) {}
    std::vector<DM> operator()(const std::vector<DM>& args) const;
    const std::string& name() const { return m_name; }
    bool is_null() const { return m_fn == nullptr && m_out.empty(); }
    bool is_symbolic() const { return !m_out.empty(); }
    int n_in() const { return static_cast<int>(m_in.size()); }
    int n_out() const { return static_cast<int>(m_out.size()); }
    const std::vector<SX>& sx_in() const { return m_in; }
    const std::vector<SX>& sx_out() const { return m_out; }
    const std::vector<std::string>& name_in() const { return m_in_names; }
    const std::vector<std::string>& name_out() const { return m_out_names; }
private:
    std::vector<SX> m_in, m_out;
    std::vector<std::string> m_in_names, m_out_names;
    friend Function external(const std::string& name, const std::string& lib);
    using eval_t = int (*)(const double**, double**, long long*, double*, int);
    using sp_t = const long long* (*)(long long);
    std::string m_name;
    eval_t m_fn = nullptr;
    sp_t m_sp_out = nullptr;
    void* m_dl = nullptr;
};
Function external(const std::string& name, const std::string& lib);

}  // namespace casadi
