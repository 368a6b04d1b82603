// partitionsolvers/LTSS.hpp -- the reference's LTSSSolver surface (src/cpp/include/LTSS.hpp:27-71) over the
// B200 engine: stable priority sort, then the best single subset among the prefixes and suffixes of the sorted
// order under the multiple-clustering score (LTSS_impl.hpp:73-111).  One call into the C ABI (ps_ltss_f32);
// the engine implements DataType = float, which is the only instantiation the reference's bindings use
// (python_ltsssolver.cpp:3-29).
#ifndef PARTITIONSOLVERS_LTSS_HPP
#define PARTITIONSOLVERS_LTSS_HPP

#include <type_traits>
#include <vector>

#include "DP.hpp"

template <typename DataType>
class LTSSSolver {
  static_assert(std::is_same<DataType, float>::value, "the B200 engine implements LTSSSolver<float>");

 public:
  LTSSSolver(std::vector<DataType> a, std::vector<DataType> b, objective_fn parametric_dist = objective_fn::Gaussian)
      : n_{static_cast<int>(a.size())}, a_{a}, b_{b}, parametric_dist_{parametric_dist} {
    _init();
  }
  LTSSSolver(int n, std::vector<DataType> a, std::vector<DataType> b,
             objective_fn parametric_dist = objective_fn::Gaussian)
      : n_{n}, a_{a}, b_{b}, parametric_dist_{parametric_dist} {
    _init();
  }

  std::vector<int> priority_sortind_;
  std::vector<int> get_optimal_subset_extern() const { return subset_; }
  float get_optimal_score_extern() const { return optimal_score_; }

 private:
  int n_;
  std::vector<DataType> a_;
  std::vector<DataType> b_;
  float optimal_score_ = 0.f;
  std::vector<int> subset_;
  objective_fn parametric_dist_;

  void _init() {
    using namespace partitionsolvers;
    const int dist = static_cast<int>(parametric_dist_);
    if (dist < 0 || dist > 2) throw distributionException();  // LTSS_impl.hpp:64-66
    if (n_ > static_cast<int>(a_.size())) n_ = static_cast<int>(a_.size());
    if (n_ < 1) throw engine_error("LTSSSolver: n must be >= 1");
    priority_sortind_.assign(n_, 0);
    int lo = 0, hi = 0;
    ps_ctx* ctx = ThreadContext::get();
    ps_status st = ps_ltss_f32(ctx, n_, a_.data(), b_.data(), dist, priority_sortind_.data(), &lo, &hi,
                               &optimal_score_);
    if (st == PS_ERR_DIST) throw distributionException();
    if (st != PS_OK) throw engine_error(std::string("ps_ltss_f32: ") + ps_last_error(ctx));
    subset_.assign(priority_sortind_.begin() + lo, priority_sortind_.begin() + hi);
  }
};

#endif  // PARTITIONSOLVERS_LTSS_HPP
