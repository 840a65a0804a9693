// ORACLE/_ref: C wrapper around the reference's own pre-CasADi closed forms of the double
// pendulum (old/Models/DoublePendulumModel.hpp:15-132: get_A, get_B, get_Xdot), compiled from
// the reference tree in place.  Row-major outputs.
#include <cmath>
using std::cos;
using std::sin;
#include <Models/DoublePendulumModel.hpp>

LinearModel::LinearModel() {}
std::vector<double> LinearModel::flatten_matrix(Eigen::MatrixXd m) {
  std::vector<double> v;
  for (int j = 0; j < m.cols(); ++j)
    for (int i = 0; i < m.rows(); ++i) v.push_back(m(i, j));
  return v;
}

extern "C" void dp_old_ref_eval(int n, const double* x, const double* u, double* f, double* A, double* B) {
  DoublePendulumModel mdl;
  for (int k = 0; k < n; ++k) {
    std::vector<double> xv(x + 4 * k, x + 4 * k + 4), uv(u + 2 * k, u + 2 * k + 2);
    Eigen::MatrixXd a = mdl.get_A(xv, uv), b = mdl.get_B(xv, uv), xd = mdl.get_Xdot(xv, uv);
    for (int i = 0; i < 4; ++i) {
      f[4 * k + i] = xd(i, 0);
      for (int j = 0; j < 4; ++j) A[16 * k + 4 * i + j] = a(i, j);
      for (int j = 0; j < 2; ++j) B[8 * k + 2 * i + j] = b(i, j);
    }
  }
}
