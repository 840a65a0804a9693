// mahi-mpc-b200 :: symbolic scalar-expression engine
//
// Stands in for the slice of casadi::SX the reference uses to *define* a model and to obtain its
// derivatives (examples/ex_model_generate.cpp:19-43; src/Mahi/Mpc/ModelGenerator.cpp:45-53:
// jacobian(x_dot,x), jacobian(x_dot,u)), plus what CasADi's nlpsol derives internally for
// nlp_hess_l (second derivatives of the dynamics contracted with a weight vector).
//
// Design: one process-wide hash-consed DAG (common sub-expressions are shared by construction),
// node ids are topologically ordered, forward- and reverse-mode symbolic differentiation produce
// new nodes in the same DAG, emitters walk the live sub-DAG in id order and print straight-line
// code (CUDA device functors for the sm_100a kernels, C with CasADi's generated-code ABI for
// <name>_linear_functions.so).
#pragma once

#include <cstdint>
#include <initializer_list>
#include <string>
#include <unordered_map>
#include <vector>

namespace mahi {
namespace mpc {
namespace sym {

enum class Op : std::uint8_t { Const, Sym, Add, Sub, Mul, Div, Neg, Sin, Cos, Tan, Exp, Log, Sqrt, Pow, Tanh, Atan, Fabs, Sign };

struct Node {
    Op op;
    int a, b;        // operand ids (-1 if unused)
    double val;      // Const value
    std::string name;  // Sym name
};

/// The expression DAG. `graph()` is the process-wide instance used by SX.
class Graph {
public:
    int constant(double v);
    int symbol(const std::string& name);  // always a fresh node
    int unary(Op op, int a);
    int binary(Op op, int a, int b);
    const Node& node(int id) const { return m_nodes[id]; }
    int size() const { return static_cast<int>(m_nodes.size()); }
    bool is_const(int id) const { return m_nodes[id].op == Op::Const; }
    bool is_zero(int id) const { return is_const(id) && m_nodes[id].val == 0.0; }
    bool is_one(int id) const { return is_const(id) && m_nodes[id].val == 1.0; }

    /// d(outs)/d(var) by forward mode; `var` must be a Sym node.
    std::vector<int> forward(const std::vector<int>& outs, int var);
    /// gradient of scalar `out` w.r.t. each of `vars` by one reverse sweep.
    std::vector<int> reverse(int out, const std::vector<int>& vars);
    /// replace symbols `from[i]` by expressions `to[i]` in `outs`.
    std::vector<int> substitute(const std::vector<int>& outs, const std::vector<int>& from, const std::vector<int>& to);
    /// numeric evaluation: `values` maps Sym id -> value.
    std::vector<double> eval(const std::vector<int>& outs, const std::unordered_map<int, double>& values) const;
    /// ids reachable from outs, ascending (= a valid evaluation order).
    std::vector<int> live(const std::vector<int>& outs) const;

private:
    int intern(Op op, int a, int b);
    std::vector<Node> m_nodes;
    std::unordered_map<std::uint64_t, std::vector<int>> m_memo;
    std::unordered_map<std::uint64_t, int> m_consts;
};

Graph& graph();

/// casadi::SX look-alike: a dense column-major matrix of scalar expressions (a scalar is 1x1).
class SX {
public:
    SX() : m_rows(0), m_cols(0) {}
    SX(double v);  // NOLINT: implicit like casadi
    SX(int rows, int cols);  // zeros
    static SX sym(const std::string& name, int rows = 1, int cols = 1);
    static SX zeros(int rows, int cols = 1) { return SX(rows, cols); }
    static SX vertcat(const std::vector<SX>& parts);
    static SX horzcat(const std::vector<SX>& parts);
    static SX from_ids(int rows, int cols, const std::vector<int>& ids);

    int size1() const { return m_rows; }
    int size2() const { return m_cols; }
    int numel() const { return m_rows * m_cols; }
    bool is_scalar() const { return m_rows == 1 && m_cols == 1; }
    bool is_empty() const { return numel() == 0; }
    int id(int i, int j = 0) const { return m_ids[static_cast<size_t>(j) * m_rows + i]; }
    int& id(int i, int j = 0) { return m_ids[static_cast<size_t>(j) * m_rows + i]; }
    const std::vector<int>& ids() const { return m_ids; }
    SX operator()(int i, int j = 0) const;
    SX T() const;
    SX operator-() const;
    std::string str() const;  // human-readable (debug)

private:
    int m_rows, m_cols;
    std::vector<int> m_ids;
};

SX operator+(const SX& a, const SX& b);
SX operator-(const SX& a, const SX& b);
SX operator*(const SX& a, const SX& b);  // element-wise (casadi semantics); use mtimes for matrix product
SX operator/(const SX& a, const SX& b);
SX sin(const SX& a);
SX cos(const SX& a);
SX tan(const SX& a);
SX exp(const SX& a);
SX log(const SX& a);
SX sqrt(const SX& a);
SX tanh(const SX& a);
SX atan(const SX& a);
SX fabs(const SX& a);
SX pow(const SX& a, const SX& b);
SX sq(const SX& a);
SX mtimes(const SX& a, const SX& b);
SX jacobian(const SX& expr, const SX& wrt);          // (numel(expr) x numel(wrt))
SX gradient(const SX& scalar_expr, const SX& wrt);   // column
SX hessian(const SX& scalar_expr, const SX& wrt);    // symmetric
SX substitute(const SX& expr, const SX& from, const SX& to);
std::vector<double> evaluate(const SX& expr, const SX& vars, const std::vector<double>& values);

// ---------------------------------------------------------------------------------------------
// code emission
// ---------------------------------------------------------------------------------------------
struct EmitInput {
    std::string cname;  // C array name
    SX syms;            // symbols bound to cname[i]
};
struct EmitOutput {
    std::string cname;
    std::vector<int> ids;  // flat list written to cname[i]
    bool skip_zero = false;  // do not emit stores of structural zeros (caller pre-zeroes)
};

/// Emit the body (declarations of temporaries + output stores) of a straight-line function.
std::string emit_body(const std::vector<EmitInput>& ins, const std::vector<EmitOutput>& outs, const std::string& indent,
                      bool cuda);

/// Count of arithmetic nodes (add/sub/mul/div/neg) and transcendental nodes in the live DAG.
struct OpCount {
    int arith = 0, trans = 0;
};
OpCount count_ops(const std::vector<int>& outs);

}  // namespace sym
}  // namespace mpc
}  // namespace mahi
