// python_ltsssolver.cpp -- reference: src/cpp/python_ltsssolver.cpp:3-29; each call is one LTSSSolver<float>.
#include "python_ltsssolver.hpp"

std::vector<int> find_optimal_partition__LTSS(int n, std::vector<float> a, std::vector<float> b, int parametric_dist) {
  return LTSSSolver<float>(n, a, b, static_cast<objective_fn>(parametric_dist)).get_optimal_subset_extern();
}

float find_optimal_score__LTSS(int n, std::vector<float> a, std::vector<float> b, int parametric_dist) {
  return LTSSSolver<float>(n, a, b, static_cast<objective_fn>(parametric_dist)).get_optimal_score_extern();
}

std::pair<std::vector<int>, float> optimize_one__LTSS(int n, std::vector<float> a, std::vector<float> b,
                                                      int parametric_dist) {
  LTSSSolver<float> s(n, a, b, static_cast<objective_fn>(parametric_dist));
  return std::make_pair(s.get_optimal_subset_extern(), s.get_optimal_score_extern());
}
