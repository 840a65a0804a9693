// casadi::external stand-in: loads a function with CasADi's generated-code ABI from a shared library
// (the drivers use it for the plant model, examples/model_control_example.cpp:46,81-82).
#include <casadi/casadi.hpp>
#include <dlfcn.h>

#include <stdexcept>

namespace casadi {

Function external(const std::string& name, const std::string& lib) {
    const std::string path = lib.find('/') == std::string::npos ? "./" + lib : lib;
    void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) throw std::runtime_error("casadi::external: cannot open '" + lib + "': " + dlerror());
    Function f;
    f.m_name = name;
    f.m_dl = h;
    f.m_fn = reinterpret_cast<Function::eval_t>(dlsym(h, name.c_str()));
    f.m_sp_out = reinterpret_cast<Function::sp_t>(dlsym(h, (name + "_sparsity_out").c_str()));
    if (!f.m_fn || !f.m_sp_out) throw std::runtime_error("casadi::external: '" + name + "' not found in '" + lib + "'");
    return f;
}

std::vector<DM> Function::operator()(const std::vector<DM>& args) const {
    if (is_symbolic()) {  // numeric evaluation of the expression graph (dense result, structural zeros are zeros)
        if (args.size() != m_in.size()) throw std::invalid_argument("casadi::Function '" + m_name + "': wrong number of arguments");
        std::vector<SX> parts;
        std::vector<double> values;
        for (size_t a = 0; a < m_in.size(); ++a) {
            if (static_cast<int>(args[a].nonzeros().size()) != m_in[a].numel())
                throw std::invalid_argument("casadi::Function '" + m_name + "': argument " + std::to_string(a) + " has the wrong size");
            for (int j = 0; j < m_in[a].size2(); ++j) for (int i = 0; i < m_in[a].size1(); ++i) parts.push_back(m_in[a](i, j));
            values.insert(values.end(), args[a].nonzeros().begin(), args[a].nonzeros().end());
        }
        const SX vars = SX::vertcat(parts);
        std::vector<DM> res;
        for (const SX& o : m_out) {
            const std::vector<double> v = mahi::mpc::sym::evaluate(o, vars, values);  // column-major like SX
            DM r(o.size1(), o.size2());
            for (int j = 0; j < o.size2(); ++j) for (int i = 0; i < o.size1(); ++i) r(i, j) = v[static_cast<size_t>(j) * o.size1() + i];
            res.push_back(r);
        }
        return res;
    }
    if (!m_fn) throw std::runtime_error("casadi::Function: null function called");
    std::vector<const double*> in;
    for (const DM& a : args) in.push_back(a.nonzeros().data());
    const long long* sp = m_sp_out(0);  // CCS: nrow, ncol, colind[ncol+1], row[nnz]
    const int nrow = static_cast<int>(sp[0]), ncol = static_cast<int>(sp[1]);
    const long long* colind = sp + 2;
    const long long* row = sp + 2 + ncol + 1;
    std::vector<double> nz(static_cast<size_t>(colind[ncol]), 0.0);
    double* out[1] = {nz.data()};
    if (m_fn(in.data(), out, nullptr, nullptr, 0) != 0) throw std::runtime_error("casadi::Function: evaluation of '" + m_name + "' failed");
    DM r(nrow, ncol);  // dense, zeros at structural zeros (what the reference's DM -> std::vector<double> yields)
    for (int j = 0; j < ncol; ++j)
        for (long long k = colind[j]; k < colind[j + 1]; ++k) r(static_cast<int>(row[k]), j) = nz[static_cast<size_t>(k)];
    return {r};
}

}  // namespace casadi
