// partitionsolvers/score.hpp -- the reference's score-context surface (src/cpp/include/score.hpp:31-221) as far as
// callers outside the solver use it: construct a context over (a, b), init(), then read range scores with
// get_score(i, j) (python_dpsolver.cpp:74-111 reads (0, n)).  The reference fills an (n+1)^2 table in init();
// here nothing is tabulated: a score is computed on request by the engine (ps_range_score_*), with the same
// left-to-right sums and rounding recipe.
#ifndef PARTITIONSOLVERS_SCORE_HPP
#define PARTITIONSOLVERS_SCORE_HPP

#include <cstddef>
#include <string>
#include <vector>

#include "DP.hpp"

namespace Objectives {

template <typename DataType>
class ParametricContext {
 public:
  ParametricContext(const std::vector<DataType>& a, const std::vector<DataType>& b, std::size_t n,
                    bool risk_partitioning_objective, bool use_rational_optimization, std::string name, int dist)
      : a_{a}, b_{b}, n_{n}, risk_partitioning_objective_{risk_partitioning_objective},
        use_rational_optimization_{use_rational_optimization}, name_{name}, dist_{dist} {}
  virtual ~ParametricContext() = default;

  void init() {}  // the reference tabulates every range score here; the engine evaluates on demand

  DataType get_score(int i, int j) const {
    if (i < 0 || j <= i || j > static_cast<int>(n_)) return DataType(0);
    DataType out = 0;
    ps_ctx* ctx = partitionsolvers::ThreadContext::get();
    ps_status st = range_score(ctx, j - i, a_.data() + i, b_.data() + i, &out);
    if (st != PS_OK) throw partitionsolvers::engine_error(ps_last_error(ctx));
    return out;
  }

  std::string getName() const { return name_; }
  bool getRiskPartitioningObjective() const { return risk_partitioning_objective_; }
  bool getUseRationalOptimization() const { return use_rational_optimization_; }

 protected:
  std::vector<DataType> a_;
  std::vector<DataType> b_;
  std::size_t n_;
  bool risk_partitioning_objective_;
  bool use_rational_optimization_;
  std::string name_;
  int dist_;

 private:
  ps_status range_score(ps_ctx* c, int n, const float* a, const float* b, float* out) const {
    return ps_range_score_f32(c, n, a, b, dist_, risk_partitioning_objective_ ? 1 : 0, out);
  }
  ps_status range_score(ps_ctx* c, int n, const double* a, const double* b, double* out) const {
    return ps_range_score_f64(c, n, a, b, dist_, risk_partitioning_objective_ ? 1 : 0, out);
  }
};

#define PS_DEFINE_CONTEXT(NAME, LABEL, DIST)                                                                     \
  template <typename DataType>                                                                                   \
  class NAME : public ParametricContext<DataType> {                                                              \
   public:                                                                                                       \
    NAME(const std::vector<DataType>& a, const std::vector<DataType>& b, std::size_t n,                          \
         bool risk_partitioning_objective, bool use_rational_optimization)                                       \
        : ParametricContext<DataType>(a, b, n, risk_partitioning_objective, use_rational_optimization, LABEL,    \
                                      DIST) {}                                                                   \
  }

PS_DEFINE_CONTEXT(GaussianContext, "Gaussian", PS_DIST_GAUSSIAN);
PS_DEFINE_CONTEXT(PoissonContext, "Poisson", PS_DIST_POISSON);
PS_DEFINE_CONTEXT(RationalScoreContext, "RationalScore", PS_DIST_RATIONAL);
#undef PS_DEFINE_CONTEXT

}  // namespace Objectives

#endif  // PARTITIONSOLVERS_SCORE_HPP
