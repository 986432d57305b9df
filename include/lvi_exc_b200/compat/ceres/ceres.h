// ceres/ceres.h — the Ceres-shaped surface of the calibration hot path (SURVEY §8b "Ceres-shaped surface": the complete list of ceres::
// members the reference tree touches), for hosts where the solve runs on the B200 through include/lvi_exc_b200.h instead of in Ceres.
//
// What it is: the TYPES and the bookkeeping.  kontiki::TrajectoryEstimator (kontiki/trajectory_estimator.h in this tree) keeps a
// ceres::Problem that records exactly what the reference's measurement wiring registers -- parameter blocks with their sizes and local
// parameterisations, constant blocks, bounds, one residual block per measurement with its loss -- so code that inspects or decorates the
// problem (estimator.problem().SetParameterBlockConstant(...), NumResidualBlocks(), ...) keeps working, and Solver::Options / Summary /
// IterationCallback carry the same fields.  What it is NOT: a CPU solver.  ceres::Solve() on a free-standing Problem throws: the only
// solver behind this header is the device LM of lvi_problem_solve, reached through TrajectoryEstimator::Solve (no CPU fallback).
// Call sites: K/kontiki/trajectory_estimator.h:22-68,88-94 ; K/entity/paramstore/paramstore.h:14-32 ; K/kontiki/sensors/sensors.h:94,161-162 ;
// K/kontiki/measurements/*.h (DynamicAutoDiffCostFunction, HuberLoss) ; L/include/utils/ceres_callbacks.h:31-66.
#ifndef LVI_EXC_B200_COMPAT_CERES_H
#define LVI_EXC_B200_COMPAT_CERES_H
#include <cmath>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace ceres {

enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };
enum MinimizerType { LINE_SEARCH, TRUST_REGION };
enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };
enum CallbackReturnType { SOLVER_CONTINUE, SOLVER_ABORT, SOLVER_TERMINATE_SUCCESSFULLY };

// scalar functions the residual functors call on double (the Jet overloads exist only inside Ceres' autodiff, which the device path
// replaces with analytic Jacobians, lvi_exc_b200/csrc/residuals.cuh)
using std::abs; using std::atan2; using std::cos; using std::exp; using std::pow; using std::sin; using std::sqrt;
template <class T> inline T DotProduct(const T x[3], const T y[3]) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; }
template <class T, int N> struct Jet { T a; T v[N]; Jet() : a(), v() {} explicit Jet(const T& s) : a(s), v() {} };   // data layout only

class LossFunction { public: virtual ~LossFunction() {} virtual void Evaluate(double s, double out[3]) const = 0; };
class HuberLoss : public LossFunction {
 public:
  explicit HuberLoss(double a) : a_(a), b_(a * a) {}
  void Evaluate(double s, double rho[3]) const override {
    if (s > b_) { const double r = std::sqrt(s); rho[0] = 2.0 * a_ * r - b_; rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r); rho[2] = -rho[1] / (2.0 * s); }
    else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  }
  double a() const { return a_; }
 private:
  double a_, b_;
};

class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};
// q+ = [sin|d|/|d| d, cos|d|] * q on storage x,y,z,w
class EigenQuaternionParameterization : public LocalParameterization {
 public:
  bool Plus(const double* x, const double* d, double* o) const override {
    const double n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (n > 0.0) {
      const double s = std::sin(n) / n, qx = s * d[0], qy = s * d[1], qz = s * d[2], qw = std::cos(n);
      o[0] = qw * x[0] + qx * x[3] + qy * x[2] - qz * x[1]; o[1] = qw * x[1] + qy * x[3] + qz * x[0] - qx * x[2];
      o[2] = qw * x[2] + qz * x[3] + qx * x[1] - qy * x[0]; o[3] = qw * x[3] - qx * x[0] - qy * x[1] - qz * x[2];
    } else { for (int i = 0; i < 4; ++i) o[i] = x[i]; }
    return true;
  }
  int GlobalSize() const override { return 4; }
  int LocalSize() const override { return 3; }
};

class CostFunction {
 public:
  virtual ~CostFunction() {}
  const std::vector<int>& parameter_block_sizes() const { return sizes_; }
  int num_residuals() const { return num_residuals_; }
 protected:
  std::vector<int> sizes_;
  int num_residuals_ = 0;
};
// Holds the functor and the block layout; never evaluated (the device evaluates the same residual analytically)
template <class Functor, int Stride = 4> class DynamicAutoDiffCostFunction : public CostFunction {
 public:
  explicit DynamicAutoDiffCostFunction(Functor* f) : functor_(f) {}
  void AddParameterBlock(int size) { sizes_.push_back(size); }
  void SetNumResiduals(int n) { num_residuals_ = n; }
 private:
  std::unique_ptr<Functor> functor_;
};

class Problem {
 public:
  struct Options {
    Ownership cost_function_ownership = TAKE_OWNERSHIP, loss_function_ownership = TAKE_OWNERSHIP, local_parameterization_ownership = TAKE_OWNERSHIP;
    bool enable_fast_removal = false, disable_all_safety_checks = false;
  };
  struct Block { int size = 0; LocalParameterization* parameterization = nullptr; bool constant = false; std::map<int, double> lower, upper; };
  struct Residual { std::shared_ptr<CostFunction> cost; LossFunction* loss = nullptr; std::vector<double*> blocks; };
  Problem() {}
  explicit Problem(const Options& o) : options_(o) {}
  void AddParameterBlock(double* values, int size) { Block& b = blocks_[values]; b.size = size; }
  void AddParameterBlock(double* values, int size, LocalParameterization* lp) { Block& b = blocks_[values]; b.size = size; b.parameterization = lp; }
  void SetParameterBlockConstant(double* values) { at(values).constant = true; }
  void SetParameterBlockVariable(double* values) { at(values).constant = false; }
  bool IsParameterBlockConstant(double* values) const { auto it = blocks_.find(values); return it != blocks_.end() && it->second.constant; }
  void SetParameterLowerBound(double* values, int index, double v) { at(values).lower[index] = v; }
  void SetParameterUpperBound(double* values, int index, double v) { at(values).upper[index] = v; }
  bool HasParameterBlock(const double* values) const { return blocks_.count(const_cast<double*>(values)) != 0; }
  int ParameterBlockSize(const double* values) const { return blocks_.at(const_cast<double*>(values)).size; }
  void AddResidualBlock(CostFunction* cost, LossFunction* loss, const std::vector<double*>& parameter_blocks) {
    Residual r; r.cost.reset(cost); r.loss = loss; r.blocks = parameter_blocks;
    for (size_t k = 0; k < parameter_blocks.size(); ++k)
      if (!blocks_.count(parameter_blocks[k])) blocks_[parameter_blocks[k]].size = k < cost->parameter_block_sizes().size() ? cost->parameter_block_sizes()[k] : 0;
    num_residuals_ += cost->num_residuals();
    residuals_.push_back(std::move(r));
  }
  int NumParameterBlocks() const { return static_cast<int>(blocks_.size()); }
  int NumParameters() const { int n = 0; for (auto& kv : blocks_) n += kv.second.size; return n; }
  int NumResidualBlocks() const { return static_cast<int>(residuals_.size()) + counted_blocks_; }
  int NumResiduals() const { return num_residuals_ + counted_residuals_; }
  // the estimator registers its residual blocks by count (one per measurement; the measurement tables are the data)
  void CountResidualBlocks(int blocks, int residuals) { counted_blocks_ += blocks; counted_residuals_ += residuals; }
  const std::map<double*, Block>& parameter_blocks() const { return blocks_; }
  const Options& options() const { return options_; }
 private:
  Block& at(double* v) { auto it = blocks_.find(v); if (it == blocks_.end()) throw std::logic_error("ceres::Problem: parameter block not found"); return it->second; }
  Options options_;
  std::map<double*, Block> blocks_;
  std::vector<Residual> residuals_;
  int num_residuals_ = 0, counted_blocks_ = 0, counted_residuals_ = 0;
};

struct IterationSummary {
  int iteration = 0;
  bool step_is_valid = true, step_is_nonmonotonic = false, step_is_successful = true;
  double cost = 0, cost_change = 0, gradient_max_norm = 0, gradient_norm = 0, step_norm = 0, relative_decrease = 0, trust_region_radius = 0, eta = 0, step_size = 0;
  int line_search_function_evaluations = 0, linear_solver_iterations = 0;
  double iteration_time_in_seconds = 0, step_solver_time_in_seconds = 0, cumulative_time_in_seconds = 0;
};
class IterationCallback {
 public:
  virtual ~IterationCallback() {}
  virtual CallbackReturnType operator()(const IterationSummary& summary) = 0;
};

struct Solver {
  struct Options {
    MinimizerType minimizer_type = TRUST_REGION;
    TrustRegionStrategyType trust_region_strategy_type = LEVENBERG_MARQUARDT;
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    bool minimizer_progress_to_stdout = false, update_state_every_iteration = false, jacobi_scaling = true, use_nonmonotonic_steps = false;
    int num_threads = 1, num_linear_solver_threads = 1, max_num_iterations = 50, max_num_consecutive_invalid_steps = 5;
    double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32, min_relative_decrease = 1e-3;
    double min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32, function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    std::vector<IterationCallback*> callbacks;
  };
  struct Summary {
    TerminationType termination_type = NO_CONVERGENCE;
    std::string message;
    double initial_cost = 0, final_cost = 0, fixed_cost = 0, total_time_in_seconds = 0, jacobian_evaluation_time_in_seconds = 0, linear_solver_time_in_seconds = 0;
    int num_successful_steps = 0, num_unsuccessful_steps = 0, num_parameter_blocks = 0, num_parameters = 0, num_effective_parameters = 0, num_residual_blocks = 0,
        num_residuals = 0, num_threads_used = 0;
    std::vector<IterationSummary> iterations;
    bool IsSolutionUsable() const { return termination_type == CONVERGENCE || termination_type == NO_CONVERGENCE || termination_type == USER_SUCCESS; }
    std::string BriefReport() const {
      static const char* term[] = {"CONVERGENCE", "NO_CONVERGENCE", "FAILURE", "USER_SUCCESS", "USER_FAILURE"};
      std::ostringstream s;
      s.setf(std::ios::scientific); s.precision(6);
      s << "Ceres Solver Report: Iterations: " << (iterations.empty() ? 0 : static_cast<int>(iterations.size()) - 1) << ", Initial cost: " << initial_cost
        << ", Final cost: " << final_cost << ", Termination: " << term[termination_type];
      return s.str();
    }
    std::string FullReport() const {
      std::ostringstream s;
      s << BriefReport() << "\nSolver: TRUST_REGION / LEVENBERG_MARQUARDT / exact Schur + band Cholesky on the device (lvi_exc_b200)\nResidual blocks: "
        << num_residual_blocks << "  Residuals: " << num_residuals << "  Effective parameters: " << num_effective_parameters << "\nSuccessful steps: "
        << num_successful_steps << "  Unsuccessful steps: " << num_unsuccessful_steps << "\nTime (s): total " << total_time_in_seconds << "  Jacobian "
        << jacobian_evaluation_time_in_seconds << "  linear solver " << linear_solver_time_in_seconds << "\n";
      return s.str();
    }
  };
};

// A free-standing ceres::Problem cannot be solved here: there is no CPU solver in this library (by design), and the device solver
// takes the typed measurement tables of kontiki::TrajectoryEstimator, not arbitrary functors.
inline void Solve(const Solver::Options&, Problem*, Solver::Summary*) {
  throw std::logic_error("lvi_exc_b200: ceres::Solve on a free-standing Problem is not available (no CPU solver); use kontiki::TrajectoryEstimator::Solve");
}

}  // namespace ceres
#endif
