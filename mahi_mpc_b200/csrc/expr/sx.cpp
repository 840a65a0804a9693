// mahi-mpc-b200 :: symbolic scalar-expression engine (see sx.hpp)
#include "sx.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <stdexcept>

namespace mahi {
namespace mpc {
namespace sym {

Graph& graph() {
    static Graph g;
    return g;
}

static std::uint64_t key_of(Op op, int a, int b) {
    return (static_cast<std::uint64_t>(op) << 58) | (static_cast<std::uint64_t>(static_cast<std::uint32_t>(a + 1)) << 29) |
           static_cast<std::uint64_t>(static_cast<std::uint32_t>(b + 1));
}

int Graph::constant(double v) {
    if (v == 0.0) v = 0.0;  // fold -0.0
    std::uint64_t bits;
    std::memcpy(&bits, &v, sizeof(bits));
    auto it = m_consts.find(bits);
    if (it != m_consts.end()) return it->second;
    m_nodes.push_back(Node{Op::Const, -1, -1, v, {}});
    int id = size() - 1;
    m_consts.emplace(bits, id);
    return id;
}

int Graph::symbol(const std::string& name) {
    m_nodes.push_back(Node{Op::Sym, -1, -1, 0.0, name});
    return size() - 1;
}

int Graph::intern(Op op, int a, int b) {
    std::uint64_t k = key_of(op, a, b);
    auto& bucket = m_memo[k];
    if (!bucket.empty()) return bucket[0];
    m_nodes.push_back(Node{op, a, b, 0.0, {}});
    bucket.push_back(size() - 1);
    return size() - 1;
}

static double apply_unary(Op op, double x) {
    switch (op) {
        case Op::Neg: return -x;
        case Op::Sin: return std::sin(x);
        case Op::Cos: return std::cos(x);
        case Op::Tan: return std::tan(x);
        case Op::Exp: return std::exp(x);
        case Op::Log: return std::log(x);
        case Op::Sqrt: return std::sqrt(x);
        case Op::Tanh: return std::tanh(x);
        case Op::Atan: return std::atan(x);
        case Op::Fabs: return std::fabs(x);
        case Op::Sign: return (x > 0) - (x < 0);
        default: throw std::logic_error("apply_unary: not unary");
    }
}
static double apply_binary(Op op, double x, double y) {
    switch (op) {
        case Op::Add: return x + y;
        case Op::Sub: return x - y;
        case Op::Mul: return x * y;
        case Op::Div: return x / y;
        case Op::Pow: return std::pow(x, y);
        default: throw std::logic_error("apply_binary: not binary");
    }
}

int Graph::unary(Op op, int a) {
    if (is_const(a)) return constant(apply_unary(op, m_nodes[a].val));
    if (op == Op::Neg) {
        if (m_nodes[a].op == Op::Neg) return m_nodes[a].a;                              // -(-x) = x
        if (m_nodes[a].op == Op::Sub) return binary(Op::Sub, m_nodes[a].b, m_nodes[a].a);  // -(x-y) = y-x
    }
    if (op == Op::Cos && m_nodes[a].op == Op::Neg) return unary(Op::Cos, m_nodes[a].a);  // even
    if (op == Op::Fabs && m_nodes[a].op == Op::Neg) return unary(Op::Fabs, m_nodes[a].a);
    if ((op == Op::Sin || op == Op::Tan || op == Op::Tanh || op == Op::Atan) && m_nodes[a].op == Op::Neg)
        return unary(Op::Neg, unary(op, m_nodes[a].a));  // odd
    return intern(op, a, -1);
}

int Graph::binary(Op op, int a, int b) {
    if (is_const(a) && is_const(b)) return constant(apply_binary(op, m_nodes[a].val, m_nodes[b].val));
    const bool na = m_nodes[a].op == Op::Neg, nb = m_nodes[b].op == Op::Neg;
    switch (op) {
        case Op::Add:
            if (is_zero(a)) return b;
            if (is_zero(b)) return a;
            if (nb) return binary(Op::Sub, a, m_nodes[b].a);  // x + (-y) = x - y
            if (na) return binary(Op::Sub, b, m_nodes[a].a);  // (-x) + y = y - x
            if (a == b) return binary(Op::Mul, constant(2.0), a);
            if (a > b) std::swap(a, b);
            break;
        case Op::Sub:
            if (is_zero(b)) return a;
            if (is_zero(a)) return unary(Op::Neg, b);
            if (a == b) return constant(0.0);
            if (nb) return binary(Op::Add, a, m_nodes[b].a);  // x - (-y) = x + y
            break;
        case Op::Mul:
            if (is_zero(a) || is_zero(b)) return constant(0.0);
            if (is_one(a)) return b;
            if (is_one(b)) return a;
            if (is_const(a) && m_nodes[a].val == -1.0) return unary(Op::Neg, b);
            if (is_const(b) && m_nodes[b].val == -1.0) return unary(Op::Neg, a);
            if (na && nb) return binary(Op::Mul, m_nodes[a].a, m_nodes[b].a);
            if (na) return unary(Op::Neg, binary(Op::Mul, m_nodes[a].a, b));  // float negation outward
            if (nb) return unary(Op::Neg, binary(Op::Mul, a, m_nodes[b].a));
            if (is_const(a) && m_nodes[a].val < 0.0 && !is_const(b))
                return unary(Op::Neg, binary(Op::Mul, constant(-m_nodes[a].val), b));
            if (is_const(b) && m_nodes[b].val < 0.0 && !is_const(a))
                return unary(Op::Neg, binary(Op::Mul, a, constant(-m_nodes[b].val)));
            if (a > b) std::swap(a, b);
            break;
        case Op::Div:
            if (is_zero(a)) return constant(0.0);
            if (is_one(b)) return a;
            if (a == b) return constant(1.0);
            if (na && nb) return binary(Op::Div, m_nodes[a].a, m_nodes[b].a);
            if (na) return unary(Op::Neg, binary(Op::Div, m_nodes[a].a, b));
            if (nb) return unary(Op::Neg, binary(Op::Div, a, m_nodes[b].a));
            if (is_const(b)) return binary(Op::Mul, a, constant(1.0 / m_nodes[b].val));
            break;
        case Op::Pow:
            if (is_const(b)) {
                double e = m_nodes[b].val;
                if (e == 0.0) return constant(1.0);
                if (e == 1.0) return a;
                if (e == 2.0) return binary(Op::Mul, a, a);
                if (e == 3.0) return binary(Op::Mul, a, binary(Op::Mul, a, a));
                if (e == 4.0) { int s = binary(Op::Mul, a, a); return binary(Op::Mul, s, s); }
                if (e == -1.0) return binary(Op::Div, constant(1.0), a);
                if (e == 0.5) return unary(Op::Sqrt, a);
            }
            break;
        default: throw std::logic_error("binary: bad op");
    }
    return intern(op, a, b);
}

std::vector<int> Graph::live(const std::vector<int>& outs) const {
    std::vector<char> mark(m_nodes.size(), 0);
    std::vector<int> stack(outs.begin(), outs.end());
    while (!stack.empty()) {
        int id = stack.back();
        stack.pop_back();
        if (id < 0 || mark[id]) continue;
        mark[id] = 1;
        if (m_nodes[id].a >= 0) stack.push_back(m_nodes[id].a);
        if (m_nodes[id].b >= 0) stack.push_back(m_nodes[id].b);
    }
    std::vector<int> order;
    for (int i = 0; i < size(); ++i)
        if (mark[i]) order.push_back(i);
    return order;
}

std::vector<int> Graph::forward(const std::vector<int>& outs, int var) {
    if (m_nodes[var].op != Op::Sym) throw std::invalid_argument("forward: differentiation variable must be a symbol");
    const std::vector<int> order = live(outs);  // snapshot: nodes created below are not part of it
    std::unordered_map<int, int> d;
    const int zero = constant(0.0), one = constant(1.0);
    for (int id : order) {
        const Node n = m_nodes[id];  // copy: m_nodes may reallocate
        int r = zero;
        const int da = n.a >= 0 ? d[n.a] : zero;
        const int db = n.b >= 0 ? d[n.b] : zero;
        switch (n.op) {
            case Op::Const: r = zero; break;
            case Op::Sym: r = (id == var) ? one : zero; break;
            case Op::Add: r = binary(Op::Add, da, db); break;
            case Op::Sub: r = binary(Op::Sub, da, db); break;
            case Op::Mul: r = binary(Op::Add, binary(Op::Mul, da, n.b), binary(Op::Mul, n.a, db)); break;
            case Op::Div:  // (da - (a/b) db)/b
                if (is_zero(da) && is_zero(db)) r = zero;
                else r = binary(Op::Div, binary(Op::Sub, da, binary(Op::Mul, id, db)), n.b);
                break;
            case Op::Neg: r = unary(Op::Neg, da); break;
            case Op::Sin: r = binary(Op::Mul, unary(Op::Cos, n.a), da); break;
            case Op::Cos: r = unary(Op::Neg, binary(Op::Mul, unary(Op::Sin, n.a), da)); break;
            case Op::Tan: r = is_zero(da) ? zero : binary(Op::Mul, binary(Op::Add, one, binary(Op::Mul, id, id)), da); break;
            case Op::Exp: r = binary(Op::Mul, id, da); break;
            case Op::Log: r = binary(Op::Div, da, n.a); break;
            case Op::Sqrt: r = is_zero(da) ? zero : binary(Op::Div, da, binary(Op::Mul, constant(2.0), id)); break;
            case Op::Tanh: r = is_zero(da) ? zero : binary(Op::Mul, binary(Op::Sub, one, binary(Op::Mul, id, id)), da); break;
            case Op::Atan: r = is_zero(da) ? zero : binary(Op::Div, da, binary(Op::Add, one, binary(Op::Mul, n.a, n.a))); break;
            case Op::Fabs: r = is_zero(da) ? zero : binary(Op::Mul, unary(Op::Sign, n.a), da); break;
            case Op::Sign: r = zero; break;
            case Op::Pow: {
                int t = zero;
                if (!is_zero(da)) t = binary(Op::Mul, binary(Op::Mul, n.b, binary(Op::Pow, n.a, binary(Op::Sub, n.b, one))), da);
                if (!is_zero(db)) t = binary(Op::Add, t, binary(Op::Mul, binary(Op::Mul, id, unary(Op::Log, n.a)), db));
                r = t;
            } break;
        }
        d[id] = r;
    }
    std::vector<int> res;
    for (int o : outs) res.push_back(d[o]);
    return res;
}

std::vector<int> Graph::reverse(int out, const std::vector<int>& vars) {
    const std::vector<int> order = live({out});
    std::unordered_map<int, int> adj;
    const int zero = constant(0.0), one = constant(1.0);
    for (int id : order) adj[id] = zero;
    adj[out] = one;
    auto acc = [&](int id, int v) { adj[id] = binary(Op::Add, adj[id], v); };
    for (auto it = order.rbegin(); it != order.rend(); ++it) {
        const int id = *it;
        const int w = adj[id];
        if (is_zero(w)) continue;
        const Node n = m_nodes[id];
        switch (n.op) {
            case Op::Const: case Op::Sym: case Op::Sign: break;
            case Op::Add: acc(n.a, w); acc(n.b, w); break;
            case Op::Sub: acc(n.a, w); acc(n.b, unary(Op::Neg, w)); break;
            case Op::Mul: acc(n.a, binary(Op::Mul, w, n.b)); acc(n.b, binary(Op::Mul, w, n.a)); break;
            case Op::Div: {
                int q = binary(Op::Div, w, n.b);
                acc(n.a, q);
                acc(n.b, unary(Op::Neg, binary(Op::Mul, q, id)));
            } break;
            case Op::Neg: acc(n.a, unary(Op::Neg, w)); break;
            case Op::Sin: acc(n.a, binary(Op::Mul, w, unary(Op::Cos, n.a))); break;
            case Op::Cos: acc(n.a, unary(Op::Neg, binary(Op::Mul, w, unary(Op::Sin, n.a)))); break;
            case Op::Tan: acc(n.a, binary(Op::Mul, w, binary(Op::Add, one, binary(Op::Mul, id, id)))); break;
            case Op::Exp: acc(n.a, binary(Op::Mul, w, id)); break;
            case Op::Log: acc(n.a, binary(Op::Div, w, n.a)); break;
            case Op::Sqrt: acc(n.a, binary(Op::Div, w, binary(Op::Mul, constant(2.0), id))); break;
            case Op::Tanh: acc(n.a, binary(Op::Mul, w, binary(Op::Sub, one, binary(Op::Mul, id, id)))); break;
            case Op::Atan: acc(n.a, binary(Op::Div, w, binary(Op::Add, one, binary(Op::Mul, n.a, n.a)))); break;
            case Op::Fabs: acc(n.a, binary(Op::Mul, w, unary(Op::Sign, n.a))); break;
            case Op::Pow:
                acc(n.a, binary(Op::Mul, w, binary(Op::Mul, n.b, binary(Op::Pow, n.a, binary(Op::Sub, n.b, one)))));
                if (!is_const(n.b)) acc(n.b, binary(Op::Mul, w, binary(Op::Mul, id, unary(Op::Log, n.a))));
                break;
        }
    }
    std::vector<int> res;
    for (int v : vars) {
        auto it = adj.find(v);
        res.push_back(it == adj.end() ? zero : it->second);
    }
    return res;
}

std::vector<int> Graph::substitute(const std::vector<int>& outs, const std::vector<int>& from, const std::vector<int>& to) {
    const std::vector<int> order = live(outs);
    std::unordered_map<int, int> m;
    for (size_t i = 0; i < from.size(); ++i) m[from[i]] = to[i];
    for (int id : order) {
        if (m.count(id)) continue;
        const Node n = m_nodes[id];
        if (n.op == Op::Const || n.op == Op::Sym) { m[id] = id; continue; }
        if (n.b >= 0) m[id] = binary(n.op, m[n.a], m[n.b]);
        else m[id] = unary(n.op, m[n.a]);
    }
    std::vector<int> res;
    for (int o : outs) res.push_back(m[o]);
    return res;
}

std::vector<double> Graph::eval(const std::vector<int>& outs, const std::unordered_map<int, double>& values) const {
    const std::vector<int> order = live(outs);
    std::unordered_map<int, double> v;
    for (int id : order) {
        const Node& n = m_nodes[id];
        if (n.op == Op::Const) v[id] = n.val;
        else if (n.op == Op::Sym) {
            auto it = values.find(id);
            if (it == values.end()) throw std::invalid_argument("eval: unbound symbol " + n.name);
            v[id] = it->second;
        } else if (n.b >= 0) v[id] = apply_binary(n.op, v[n.a], v[n.b]);
        else v[id] = apply_unary(n.op, v[n.a]);
    }
    std::vector<double> res;
    for (int o : outs) res.push_back(v[o]);
    return res;
}

// ---------------------------------------------------------------------------------------------
// SX
// ---------------------------------------------------------------------------------------------
SX::SX(double v) : m_rows(1), m_cols(1), m_ids{graph().constant(v)} {}
SX::SX(int rows, int cols) : m_rows(rows), m_cols(cols), m_ids(static_cast<size_t>(rows) * cols, graph().constant(0.0)) {}

SX SX::sym(const std::string& name, int rows, int cols) {
    SX r(rows, cols);
    for (int j = 0; j < cols; ++j)
        for (int i = 0; i < rows; ++i) {
            std::string nm = name;
            if (rows * cols > 1) nm += "_" + std::to_string(j * rows + i);
            r.id(i, j) = graph().symbol(nm);
        }
    return r;
}

SX SX::from_ids(int rows, int cols, const std::vector<int>& ids) {
    if (static_cast<int>(ids.size()) != rows * cols) throw std::invalid_argument("from_ids: size mismatch");
    SX r(rows, cols);
    r.m_ids = ids;
    return r;
}

SX SX::vertcat(const std::vector<SX>& parts) {
    int cols = -1, rows = 0;
    for (const SX& p : parts) {
        if (p.is_empty()) continue;
        if (cols < 0) cols = p.size2();
        if (p.size2() != cols) throw std::invalid_argument("vertcat: column mismatch");
        rows += p.size1();
    }
    if (cols < 0) return SX();
    SX r(rows, cols);
    int off = 0;
    for (const SX& p : parts) {
        for (int j = 0; j < p.size2(); ++j)
            for (int i = 0; i < p.size1(); ++i) r.id(off + i, j) = p.id(i, j);
        off += p.size1();
    }
    return r;
}

SX SX::horzcat(const std::vector<SX>& parts) {
    std::vector<SX> t;
    for (const SX& p : parts) t.push_back(p.T());
    return vertcat(t).T();
}

SX SX::operator()(int i, int j) const { return from_ids(1, 1, {id(i, j)}); }

SX SX::T() const {
    SX r(m_cols, m_rows);
    for (int j = 0; j < m_cols; ++j)
        for (int i = 0; i < m_rows; ++i) r.id(j, i) = id(i, j);
    return r;
}

template <class F>
static SX map2(const SX& a, const SX& b, F f) {
    if (a.is_scalar() && !b.is_scalar()) {
        SX r(b.size1(), b.size2());
        for (int k = 0; k < b.numel(); ++k) r.id(k % b.size1(), k / b.size1()) = f(a.ids()[0], b.ids()[k]);
        return r;
    }
    if (b.is_scalar() && !a.is_scalar()) {
        SX r(a.size1(), a.size2());
        for (int k = 0; k < a.numel(); ++k) r.id(k % a.size1(), k / a.size1()) = f(a.ids()[k], b.ids()[0]);
        return r;
    }
    if (a.size1() != b.size1() || a.size2() != b.size2()) throw std::invalid_argument("SX: dimension mismatch in element-wise op");
    SX r(a.size1(), a.size2());
    for (int k = 0; k < a.numel(); ++k) r.id(k % a.size1(), k / a.size1()) = f(a.ids()[k], b.ids()[k]);
    return r;
}
template <class F>
static SX map1(const SX& a, F f) {
    SX r(a.size1(), a.size2());
    for (int k = 0; k < a.numel(); ++k) r.id(k % a.size1(), k / a.size1()) = f(a.ids()[k]);
    return r;
}

SX SX::operator-() const { return map1(*this, [](int x) { return graph().unary(Op::Neg, x); }); }
SX operator+(const SX& a, const SX& b) { return map2(a, b, [](int x, int y) { return graph().binary(Op::Add, x, y); }); }
SX operator-(const SX& a, const SX& b) { return map2(a, b, [](int x, int y) { return graph().binary(Op::Sub, x, y); }); }
SX operator*(const SX& a, const SX& b) { return map2(a, b, [](int x, int y) { return graph().binary(Op::Mul, x, y); }); }
SX operator/(const SX& a, const SX& b) { return map2(a, b, [](int x, int y) { return graph().binary(Op::Div, x, y); }); }
SX pow(const SX& a, const SX& b) { return map2(a, b, [](int x, int y) { return graph().binary(Op::Pow, x, y); }); }
SX sq(const SX& a) { return a * a; }
#define MMPC_UNARY(fn, OP) \
    SX fn(const SX& a) { return map1(a, [](int x) { return graph().unary(Op::OP, x); }); }
MMPC_UNARY(sin, Sin)
MMPC_UNARY(cos, Cos)
MMPC_UNARY(tan, Tan)
MMPC_UNARY(exp, Exp)
MMPC_UNARY(log, Log)
MMPC_UNARY(sqrt, Sqrt)
MMPC_UNARY(tanh, Tanh)
MMPC_UNARY(atan, Atan)
MMPC_UNARY(fabs, Fabs)
#undef MMPC_UNARY

SX mtimes(const SX& a, const SX& b) {
    if (a.is_scalar() || b.is_scalar()) return a * b;
    if (a.size2() != b.size1()) throw std::invalid_argument("mtimes: inner dimension mismatch");
    SX r(a.size1(), b.size2());
    Graph& g = graph();
    for (int i = 0; i < a.size1(); ++i)
        for (int j = 0; j < b.size2(); ++j) {
            int acc = g.constant(0.0);
            for (int k = 0; k < a.size2(); ++k) acc = g.binary(Op::Add, acc, g.binary(Op::Mul, a.id(i, k), b.id(k, j)));
            r.id(i, j) = acc;
        }
    return r;
}

SX jacobian(const SX& expr, const SX& wrt) {
    SX r(expr.numel(), wrt.numel());
    for (int j = 0; j < wrt.numel(); ++j) {
        std::vector<int> col = graph().forward(expr.ids(), wrt.ids()[j]);
        for (int i = 0; i < expr.numel(); ++i) r.id(i, j) = col[i];
    }
    return r;
}

SX gradient(const SX& e, const SX& wrt) {
    if (!e.is_scalar()) throw std::invalid_argument("gradient: expression must be scalar");
    return SX::from_ids(wrt.numel(), 1, graph().reverse(e.ids()[0], wrt.ids()));
}

SX hessian(const SX& e, const SX& wrt) {
    SX g = gradient(e, wrt);
    const int n = wrt.numel();
    SX H(n, n);
    for (int j = 0; j < n; ++j) {  // forward-over-reverse, upper triangle then mirror
        std::vector<int> col = graph().forward(g.ids(), wrt.ids()[j]);
        for (int i = 0; i <= j; ++i) {
            H.id(i, j) = col[i];
            H.id(j, i) = col[i];
        }
    }
    return H;
}

SX substitute(const SX& expr, const SX& from, const SX& to) {
    if (from.numel() != to.numel()) throw std::invalid_argument("substitute: size mismatch");
    return SX::from_ids(expr.size1(), expr.size2(), graph().substitute(expr.ids(), from.ids(), to.ids()));
}

std::vector<double> evaluate(const SX& expr, const SX& vars, const std::vector<double>& values) {
    if (static_cast<int>(values.size()) != vars.numel()) throw std::invalid_argument("evaluate: value count mismatch");
    std::unordered_map<int, double> m;
    for (int i = 0; i < vars.numel(); ++i) m[vars.ids()[i]] = values[i];
    return graph().eval(expr.ids(), m);
}

// ---------------------------------------------------------------------------------------------
// printing / emission
// ---------------------------------------------------------------------------------------------
static std::string fmt_double(double v) {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.17g", v);
    std::string s(buf);
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";  // keep it a double literal
    return s;
}

std::string SX::str() const {
    const Graph& g = graph();
    std::ostringstream os;
    std::vector<int> order = g.live(m_ids);
    for (int id : order) {
        const Node& n = g.node(id);
        if (n.op == Op::Const || n.op == Op::Sym) continue;
        os << "v" << id << " = op" << static_cast<int>(n.op) << "(" << n.a << "," << n.b << "); ";
    }
    return os.str();
}

std::string emit_body(const std::vector<EmitInput>& ins, const std::vector<EmitOutput>& outs, const std::string& ind, bool cuda) {
    const Graph& g = graph();
    std::vector<int> all;
    for (const auto& o : outs) all.insert(all.end(), o.ids.begin(), o.ids.end());
    const std::vector<int> order = g.live(all);
    std::unordered_map<int, std::string> nm;
    std::ostringstream os;
    for (const auto& in : ins)
        for (int i = 0; i < in.syms.numel(); ++i) nm[in.syms.ids()[i]] = in.cname + "[" + std::to_string(i) + "]";
    // sin/cos of the same argument are fused into one sincos call
    std::unordered_map<int, int> sin_of, cos_of;
    for (int id : order) {
        const Node& n = g.node(id);
        if (n.op == Op::Sin) sin_of[n.a] = id;
        if (n.op == Op::Cos) cos_of[n.a] = id;
    }
    // CUDA only: sin/cos of a sum or difference come from the angle addition formulas when sin and cos of both terms are
    // computed anyway (4 multiply-adds instead of a ~60-instruction fp64 sincos; error a few ulp).  A term that has only
    // its sin or only its cos in this function gets the other one from the same sincos call, so that every function
    // emitted from one graph evaluates the shared sub-expressions identically.  The C emitters keep the plain calls.
    std::unordered_map<int, bool> want_both;
    if (cuda) {
        for (int id : order) {
            const Node& n = g.node(id);
            if (n.op != Op::Sin && n.op != Op::Cos) continue;
            const Node& arg = g.node(n.a);
            if (arg.op != Op::Add && arg.op != Op::Sub) continue;
            auto trig = [&](int t) { return sin_of.count(t) || cos_of.count(t); };
            if (trig(arg.a) && trig(arg.b)) want_both[arg.a] = want_both[arg.b] = true;
        }
    }
    // CUDA only: several quotients with the same denominator (the determinant of a mass matrix, typically) share ONE
    // reciprocal -- an fp64 division is a ~25-instruction expansion with a dependent latency of ~80 cycles, a multiplication is
    // one DMUL; the double pendulum's f, A, B, H divide 17 times by the same term.  Error <= 1.5 ulp per quotient.
    std::unordered_map<int, int> div_uses;
    std::unordered_map<int, std::string> recip_nm;
    if (cuda)
        for (int id : order) {
            const Node& n = g.node(id);
            if (n.op == Op::Div && g.node(n.b).op != Op::Const) ++div_uses[n.b];
        }
    std::unordered_map<int, std::string> sin_nm, cos_nm;  // argument id -> name of its emitted sin / cos
    auto ref = [&](int id) -> std::string {
        const Node& n = g.node(id);
        if (n.op == Op::Const) return n.val < 0 ? "(" + fmt_double(n.val) + ")" : fmt_double(n.val);
        auto it = nm.find(id);
        if (it == nm.end()) throw std::logic_error(n.op == Op::Sym ? "emit: unbound symbol " + n.name : "emit: node used before definition");
        return it->second;
    };
    std::unordered_map<int, bool> done;
    for (int id : order) {
        const Node& n = g.node(id);
        if (n.op == Op::Const || n.op == Op::Sym) continue;
        if (done.count(id)) continue;
        const std::string v = "v" + std::to_string(id);
        if (cuda && (n.op == Op::Sin || n.op == Op::Cos)) {
            const Node& arg = g.node(n.a);
            if ((arg.op == Op::Add || arg.op == Op::Sub) && sin_nm.count(arg.a) && cos_nm.count(arg.a) && sin_nm.count(arg.b) && cos_nm.count(arg.b)) {
                const std::string sp = sin_nm[arg.a], cp = cos_nm[arg.a], sq = sin_nm[arg.b], cq = cos_nm[arg.b];
                const bool add = arg.op == Op::Add;
                os << ind << "const double " << v << " = ";
                if (n.op == Op::Sin) os << "fma(" << sp << ", " << cq << ", " << (add ? "" : "-") << cp << " * " << sq << ")";
                else os << "fma(" << cp << ", " << cq << ", " << (add ? "-" : "") << sp << " * " << sq << ")";
                os << ";\n";
                nm[id] = v;
                (n.op == Op::Sin ? sin_nm : cos_nm)[n.a] = v;
                continue;
            }
        }
        if ((n.op == Op::Sin || n.op == Op::Cos) && ((sin_of.count(n.a) && cos_of.count(n.a)) || want_both.count(n.a))) {
            const std::string sname = sin_of.count(n.a) ? "v" + std::to_string(sin_of[n.a]) : "s" + std::to_string(n.a);
            const std::string cname = cos_of.count(n.a) ? "v" + std::to_string(cos_of[n.a]) : "c" + std::to_string(n.a);
            os << ind << "double " << sname << ", " << cname << "; " << (cuda ? "MMPC_SINCOS(" : "sincos(") << ref(n.a) << ", &" << sname << ", &" << cname << ");\n";
            if (sin_of.count(n.a)) { nm[sin_of[n.a]] = sname; done[sin_of[n.a]] = true; }
            if (cos_of.count(n.a)) { nm[cos_of[n.a]] = cname; done[cos_of[n.a]] = true; }
            sin_nm[n.a] = sname;
            cos_nm[n.a] = cname;
            continue;
        }
        if (cuda && n.op == Op::Div && div_uses.count(n.b) && div_uses[n.b] >= 2) {
            if (!recip_nm.count(n.b)) {
                recip_nm[n.b] = "r" + std::to_string(n.b);
                os << ind << "const double " << recip_nm[n.b] << " = 1.0 / " << ref(n.b) << ";\n";
            }
            os << ind << "const double " << v << " = " << ref(n.a) << " * " << recip_nm[n.b] << ";\n";
            nm[id] = v;
            continue;
        }
        os << ind << "const double " << v << " = ";
        switch (n.op) {
            case Op::Add: os << ref(n.a) << " + " << ref(n.b); break;
            case Op::Sub: os << ref(n.a) << " - " << ref(n.b); break;
            case Op::Mul: os << ref(n.a) << " * " << ref(n.b); break;
            case Op::Div: os << ref(n.a) << " / " << ref(n.b); break;
            case Op::Neg: os << "-" << ref(n.a); break;
            case Op::Sin: os << "sin(" << ref(n.a) << ")"; break;
            case Op::Cos: os << "cos(" << ref(n.a) << ")"; break;
            case Op::Tan: os << "tan(" << ref(n.a) << ")"; break;
            case Op::Exp: os << "exp(" << ref(n.a) << ")"; break;
            case Op::Log: os << "log(" << ref(n.a) << ")"; break;
            case Op::Sqrt: os << "sqrt(" << ref(n.a) << ")"; break;
            case Op::Tanh: os << "tanh(" << ref(n.a) << ")"; break;
            case Op::Atan: os << "atan(" << ref(n.a) << ")"; break;
            case Op::Fabs: os << "fabs(" << ref(n.a) << ")"; break;
            case Op::Sign: os << "((" << ref(n.a) << " > 0.0) - (" << ref(n.a) << " < 0.0))"; break;
            case Op::Pow: os << "pow(" << ref(n.a) << ", " << ref(n.b) << ")"; break;
            default: break;
        }
        os << ";\n";
        nm[id] = v;
    }
    for (const auto& o : outs)
        for (size_t i = 0; i < o.ids.size(); ++i) {
            if (o.skip_zero && g.is_zero(o.ids[i])) continue;
            os << ind << o.cname << "[" << i << "] = " << ref(o.ids[i]) << ";\n";
        }
    return os.str();
}

OpCount count_ops(const std::vector<int>& outs) {
    const Graph& g = graph();
    OpCount c;
    for (int id : g.live(outs)) {
        switch (g.node(id).op) {
            case Op::Const: case Op::Sym: break;
            case Op::Add: case Op::Sub: case Op::Mul: case Op::Div: case Op::Neg: case Op::Fabs: case Op::Sign: ++c.arith; break;
            default: ++c.trans;
        }
    }
    return c;
}

}  // namespace sym
}  // namespace mpc
}  // namespace mahi
