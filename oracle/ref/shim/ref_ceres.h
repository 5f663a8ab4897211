// ORACLE / reference pin (test infrastructure only).
// Stand-ins for the few Ceres 2.1.0 types the reference's cost-functor text touches, so that
// include/utils/ceresfunc.h text compiles verbatim: Jet (dual numbers, jet.h), CostFunction,
// SizedCostFunction-style AutoDiffCostFunction (one Jet pass over all parameter blocks), HuberLoss
// (loss_function.cc: rho(s) = s for s <= a^2, 2 a sqrt(s) - a^2 above). Restated, not copied.
#ifndef MML_REF_CERES_H
#define MML_REF_CERES_H
#include <cmath>
#include <cstdint>
#include <algorithm>
#include <limits>
#include <utility>
#include <vector>

namespace ceres {

template <class T, int N> struct Jet {
  T a; T v[N];
  Jet() : a(T()) { for (int i = 0; i < N; i++) v[i] = T(); }
  Jet(const T& s) : a(s) { for (int i = 0; i < N; i++) v[i] = T(); }
  template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
  Jet(const S& s) : a(T(s)) { for (int i = 0; i < N; i++) v[i] = T(); }
  Jet(const T& s, int k) : a(s) { for (int i = 0; i < N; i++) v[i] = T(); v[k] = T(1); }
  Jet& operator+=(const Jet& o) { a += o.a; for (int i = 0; i < N; i++) v[i] += o.v[i]; return *this; }
  Jet& operator-=(const Jet& o) { a -= o.a; for (int i = 0; i < N; i++) v[i] -= o.v[i]; return *this; }
  Jet& operator*=(const Jet& o) { *this = *this * o; return *this; }
  Jet& operator/=(const Jet& o) { *this = *this / o; return *this; }
};
#define MML_JET template <class T, int N> inline
MML_JET Jet<T, N> operator+(const Jet<T, N>& f) { return f; }
MML_JET Jet<T, N> operator-(const Jet<T, N>& f) { Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; i++) h.v[i] = -f.v[i]; return h; }
MML_JET Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a + g.a; for (int i = 0; i < N; i++) h.v[i] = f.v[i] + g.v[i]; return h; }
MML_JET Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a - g.a; for (int i = 0; i < N; i++) h.v[i] = f.v[i] - g.v[i]; return h; }
MML_JET Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a * g.a; for (int i = 0; i < N; i++) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
MML_JET Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; const T gi = T(1) / g.a; const T fg = f.a * gi; h.a = fg;
  for (int i = 0; i < N; i++) h.v[i] = (f.v[i] - fg * g.v[i]) * gi;
  return h;
}
#define MML_JET_S(op) \
  MML_JET Jet<T, N> operator op(const Jet<T, N>& f, T s) { return f op Jet<T, N>(s); } \
  MML_JET Jet<T, N> operator op(T s, const Jet<T, N>& f) { return Jet<T, N>(s) op f; }
MML_JET_S(+) MML_JET_S(-) MML_JET_S(*) MML_JET_S(/)
#undef MML_JET_S
#define MML_JET_CMP(op) \
  MML_JET bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a op g.a; } \
  MML_JET bool operator op(const Jet<T, N>& f, T s) { return f.a op s; } \
  MML_JET bool operator op(T s, const Jet<T, N>& f) { return s op f.a; }
MML_JET_CMP(<) MML_JET_CMP(<=) MML_JET_CMP(>) MML_JET_CMP(>=) MML_JET_CMP(==) MML_JET_CMP(!=)
#undef MML_JET_CMP
MML_JET Jet<T, N> chain(const T& val, const T& d, const Jet<T, N>& f) { Jet<T, N> h; h.a = val; for (int i = 0; i < N; i++) h.v[i] = d * f.v[i]; return h; }
MML_JET Jet<T, N> sqrt(const Jet<T, N>& f) { T t = std::sqrt(f.a); return chain(t, T(1) / (T(2) * t), f); }
MML_JET Jet<T, N> abs(const Jet<T, N>& f) { return f.a < T(0) ? -f : f; }
MML_JET Jet<T, N> sin(const Jet<T, N>& f) { return chain(std::sin(f.a), std::cos(f.a), f); }
MML_JET Jet<T, N> cos(const Jet<T, N>& f) { return chain(std::cos(f.a), -std::sin(f.a), f); }
MML_JET Jet<T, N> atan(const Jet<T, N>& f) { return chain(std::atan(f.a), T(1) / (T(1) + f.a * f.a), f); }
MML_JET Jet<T, N> acos(const Jet<T, N>& f) { return chain(std::acos(f.a), -T(1) / std::sqrt(T(1) - f.a * f.a), f); }
MML_JET Jet<T, N> atan2(const Jet<T, N>& g, const Jet<T, N>& f) {
  Jet<T, N> h; const T t = T(1) / (f.a * f.a + g.a * g.a); h.a = std::atan2(g.a, f.a);
  for (int i = 0; i < N; i++) h.v[i] = t * (-g.a * f.v[i] + f.a * g.v[i]);
  return h;
}
#undef MML_JET
inline double sqrt(double x) { return std::sqrt(x); }
inline double abs(double x) { return std::fabs(x); }

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int32_t>& parameter_block_sizes() const { return sizes_; }
  int num_residuals() const { return nres_; }
 protected:
  std::vector<int32_t>* mutable_parameter_block_sizes() { return &sizes_; }
  void set_num_residuals(int n) { nres_ = n; }
 private:
  std::vector<int32_t> sizes_; int nres_ = 0;
};

namespace detail {
template <int... Ns> struct Sum;
template <> struct Sum<> { enum { value = 0 }; };
template <int N0, int... Ns> struct Sum<N0, Ns...> { enum { value = N0 + Sum<Ns...>::value }; };
template <class F, class J, int... Ns> struct Call {
  template <size_t... I> static bool runi(const F& f, J** p, J* r, std::index_sequence<I...>) { return f(p[I]..., r); }
  static bool run(const F& f, J** p, J* r) { return runi(f, p, r, std::make_index_sequence<sizeof...(Ns)>{}); }
};
}  // namespace detail

// Jacobians are row-major [residual][parameter of the block], as Ceres stores them.
template <class Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
  Functor* f_;
 public:
  explicit AutoDiffCostFunction(Functor* f) : f_(f) {
    set_num_residuals(kNumResiduals);
    for (int n : {Ns...}) mutable_parameter_block_sizes()->push_back(n);
  }
  ~AutoDiffCostFunction() override { delete f_; }
  const Functor& functor() const { return *f_; }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    constexpr int NP = detail::Sum<Ns...>::value;
    constexpr int NB = sizeof...(Ns);
    const int sizes[NB] = {Ns...};
    using J = Jet<double, NP>;
    if (!jacobians) {
      double* pp[NB]; std::vector<std::vector<double>> copy(NB);
      for (int b = 0; b < NB; b++) { copy[b].assign(parameters[b], parameters[b] + sizes[b]); pp[b] = copy[b].data(); }
      return detail::Call<Functor, double, Ns...>::run(*f_, pp, residuals);
    }
    std::vector<J> x(NP); J* pp[NB]; int off = 0;
    for (int b = 0; b < NB; b++) { pp[b] = x.data() + off; for (int i = 0; i < sizes[b]; i++) x[off + i] = J(parameters[b][i], off + i); off += sizes[b]; }
    J out[kNumResiduals];
    if (!detail::Call<Functor, J, Ns...>::run(*f_, pp, out)) return false;
    for (int r = 0; r < kNumResiduals; r++) residuals[r] = out[r].a;
    off = 0;
    for (int b = 0; b < NB; b++) {
      if (jacobians[b]) for (int r = 0; r < kNumResiduals; r++) for (int i = 0; i < sizes[b]; i++) jacobians[b][r * sizes[b] + i] = out[r].v[off + i];
      off += sizes[b];
    }
    return true;
  }
};

class LossFunction { public: virtual ~LossFunction() {} virtual void Evaluate(double sq_norm, double out[3]) const = 0; };
class HuberLoss : public LossFunction {
  const double a_, b_;
 public:
  explicit HuberLoss(double a) : a_(a), b_(a * a) {}
  void Evaluate(double s, double rho[3]) const override {
    if (s > b_) { const double r = std::sqrt(s); rho[0] = 2.0 * a_ * r - b_; rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r); rho[2] = -rho[1] / (2.0 * s); }
    else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  }
};

// ---- Problem / Solve: dense normal equations + the oracle's restatement of Ceres' trust-region
// loop (oracle/dogleg.h). Parameter blocks enter the state vector in AddParameterBlock order;
// blocks no residual uses are dropped, as Ceres' preprocessor does (RemoveFixedBlocks). The loss
// correction is Corrector::CorrectJacobian / CorrectResiduals (rho'' <= 0 branch and the alpha
// branch), cost = 1/2 sum rho(|r|^2).
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
class Problem {
 public:
  struct Options {};
  struct RB { CostFunction* cost; LossFunction* loss; std::vector<double*> params; };
  Problem() {}
  explicit Problem(const Options&) {}
  void AddParameterBlock(double* p, int size) { for (auto& b : blocks_) if (b.first == p) return; blocks_.push_back({p, size}); }
  void AddResidualBlock(CostFunction* c, LossFunction* l, const std::vector<double*>& ps) {
    const std::vector<int32_t>& sz = c->parameter_block_sizes();
    for (size_t i = 0; i < ps.size(); i++) AddParameterBlock(ps[i], sz[i]);
    rbs_.push_back({c, l, ps});
  }
  template <class... Ps> void AddResidualBlock(CostFunction* c, LossFunction* l, double* x0, Ps*... xs) {
    AddResidualBlock(c, l, std::vector<double*>{x0, xs...});
  }
  std::vector<std::pair<double*, int>> blocks_;
  std::vector<RB> rbs_;
};
struct Solver {
  struct Options {
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    TrustRegionStrategyType trust_region_strategy_type = LEVENBERG_MARQUARDT;
    int max_num_iterations = 50; bool minimizer_progress_to_stdout = false; int num_threads = 1;
  };
  struct Summary {
    double initial_cost = 0, final_cost = 0; int num_iterations = 0, num_successful_steps = 0, termination = 0;
    bool IsSolutionUsable() const { return termination >= 0; }
  };
};
}  // namespace ceres
#include "../../dogleg.h"
namespace ceres {
inline void Solve(const Solver::Options& opt, Problem* prob, Solver::Summary* sum) {
  // active blocks, in insertion order
  std::vector<int> off(prob->blocks_.size(), -1);
  std::vector<char> used(prob->blocks_.size(), 0);
  auto find = [&](double* p) { for (size_t i = 0; i < prob->blocks_.size(); i++) if (prob->blocks_[i].first == p) return (int)i; return -1; };
  for (auto& rb : prob->rbs_) for (double* p : rb.params) used[find(p)] = 1;
  int n = 0;
  for (size_t i = 0; i < prob->blocks_.size(); i++) if (used[i]) { off[i] = n; n += prob->blocks_[i].second; }
  if (n == 0) { *sum = Solver::Summary(); return; }
  std::vector<double> x(n);
  for (size_t i = 0; i < prob->blocks_.size(); i++) if (used[i]) for (int k = 0; k < prob->blocks_[i].second; k++) x[off[i] + k] = prob->blocks_[i].first[k];
  orc::EvalFn eval = [&](const double* xx, double* cost, double* H, double* g) {
    double c = 0;
    if (H) std::fill(H, H + (size_t)n * n, 0.0);
    if (g) std::fill(g, g + n, 0.0);
    std::vector<double> r, Jbuf;
    for (auto& rb : prob->rbs_) {
      const int nr = rb.cost->num_residuals();
      const std::vector<int32_t>& sz = rb.cost->parameter_block_sizes();
      const int nb = (int)sz.size();
      std::vector<const double*> pp(nb); std::vector<int> po(nb);
      int tot = 0;
      for (int b = 0; b < nb; b++) { po[b] = off[find(rb.params[b])]; pp[b] = xx + po[b]; tot += sz[b]; }
      r.assign(nr, 0.0); Jbuf.assign((size_t)nr * tot, 0.0);
      std::vector<double*> Jp(nb); int o = 0;
      for (int b = 0; b < nb; b++) { Jp[b] = Jbuf.data() + (size_t)nr * o; o += sz[b]; }
      if (!rb.cost->Evaluate(pp.data(), r.data(), Jp.data())) return false;
      double sq = 0; for (int i = 0; i < nr; i++) sq += r[i] * r[i];
      if (rb.loss) {
        double rho[3]; rb.loss->Evaluate(sq, rho);
        c += 0.5 * rho[0];
        const double sqrt_rho1 = std::sqrt(rho[1]);
        double residual_scaling, alpha_sq_norm;
        if (sq == 0.0 || rho[2] <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
        else { const double D = 1.0 + 2.0 * sq * rho[2] / rho[1]; const double alpha = 1.0 - std::sqrt(D); residual_scaling = sqrt_rho1 / (1 - alpha); alpha_sq_norm = alpha / sq; }
        for (int b = 0; b < nb; b++) {
          double* J = Jp[b];
          if (alpha_sq_norm != 0.0) {
            for (int k = 0; k < sz[b]; k++) { double rtj = 0; for (int i = 0; i < nr; i++) rtj += r[i] * J[i * sz[b] + k]; for (int i = 0; i < nr; i++) J[i * sz[b] + k] = sqrt_rho1 * (J[i * sz[b] + k] - alpha_sq_norm * r[i] * rtj); }
          } else for (int i = 0; i < nr * sz[b]; i++) J[i] *= sqrt_rho1;
        }
        for (int i = 0; i < nr; i++) r[i] *= residual_scaling;
      } else c += 0.5 * sq;
      if (H || g) for (int b = 0; b < nb; b++) for (int k = 0; k < sz[b]; k++) {
        const int gi = po[b] + k;
        if (g) { double s = 0; for (int i = 0; i < nr; i++) s += Jp[b][i * sz[b] + k] * r[i]; g[gi] += s; }
        if (H) for (int b2 = 0; b2 < nb; b2++) for (int k2 = 0; k2 < sz[b2]; k2++) {
          double s = 0; for (int i = 0; i < nr; i++) s += Jp[b][i * sz[b] + k] * Jp[b2][i * sz[b2] + k2];
          H[(size_t)gi * n + po[b2] + k2] += s;
        }
      }
    }
    *cost = c;
    return std::isfinite(c);
  };
  orc::DoglegSummary s = orc::dogleg_minimize(n, x.data(), eval, opt.max_num_iterations);
  for (size_t i = 0; i < prob->blocks_.size(); i++) if (used[i]) for (int k = 0; k < prob->blocks_[i].second; k++) prob->blocks_[i].first[k] = x[off[i] + k];
  sum->initial_cost = s.initial_cost; sum->final_cost = s.final_cost; sum->num_iterations = s.iterations;
  sum->num_successful_steps = s.successful; sum->termination = s.termination;
}

}  // namespace ceres
#endif
