// python_dpsolver.cpp -- bodies of the flat API (reference: src/cpp/python_dpsolver.cpp:21-316), each a
// thin shell around one or more DPSolver<float> constructions, which is where the GPU work happens.
#include "python_dpsolver.hpp"

#include <algorithm>
#include <limits>

namespace {

DPSolver<float> make_solver(int n, int T, const std::vector<float>& a, const std::vector<float>& b, int dist,
                            bool rp, bool rational, float gamma, int reg_power, bool sweep_down = false,
                            bool find_t = false) {
  return DPSolver<float>(n, T, a, b, static_cast<objective_fn>(dist), rp, rational, gamma, reg_power, sweep_down,
                         find_t);
}

}  // namespace

// python_dpsolver.cpp:21-47: one pass to level T, every S backtraced, OLS choice
std::pair<std::vector<ps_scored_partition>, int> compute_optimal_num_clusters_OLS_all(
    int n, int T, std::vector<float> a, std::vector<float> b, int parametric_dist,
    bool risk_partitioning_objective, bool use_rational_optimization, float gamma, int reg_power) {
  auto dp = make_solver(n, T, a, b, parametric_dist, risk_partitioning_objective, use_rational_optimization, gamma,
                        reg_power, false, true);
  return std::make_pair(dp.get_all_subsets_and_scores_extern(), dp.get_optimal_num_clusters_OLS_extern());
}

// python_dpsolver.cpp:49-73
int compute_optimal_num_clusters_OLS(int n, int T, std::vector<float> a, std::vector<float> b, int parametric_dist,
                                     bool risk_partitioning_objective, bool use_rational_optimization, float gamma,
                                     int reg_power) {
  return make_solver(n, T, a, b, parametric_dist, risk_partitioning_objective, use_rational_optimization, gamma,
                     reg_power, false, true)
      .get_optimal_num_clusters_OLS_extern();
}

// python_dpsolver.cpp:74-111: S(0,n) of the inputs as given
float compute_score(std::vector<float> a, std::vector<float> b, int parametric_dist,
                    bool risk_partitioning_objective, bool use_rational_optimization) {
  (void)use_rational_optimization;
  if (parametric_dist < 0 || parametric_dist > 2) throw distributionException();
  float out = 0.f;
  ps_ctx* ctx = partitionsolvers::ThreadContext::get();
  ps_status st = ps_range_score_f32(ctx, static_cast<int>(a.size()), a.data(), b.data(), parametric_dist,
                                    risk_partitioning_objective ? 1 : 0, &out);
  if (st == PS_ERR_DIST) throw distributionException();
  if (st != PS_OK) throw partitionsolvers::engine_error(ps_last_error(ctx));
  return out;
}

// python_dpsolver.cpp:113-132
ps_subsets find_optimal_partition__DP(int n, int T, std::vector<float> a, std::vector<float> b, int parametric_dist,
                                      bool risk_partitioning_objective, bool use_rational_optimization, float gamma,
                                      int reg_power) {
  return make_solver(n, T, a, b, parametric_dist, risk_partitioning_objective, use_rational_optimization, gamma,
                     reg_power)
      .get_optimal_subsets_extern();
}

// python_dpsolver.cpp:134-153
float find_optimal_score__DP(int n, int T, std::vector<float> a, std::vector<float> b, int parametric_dist,
                             bool risk_partitioning_objective, bool use_rational_optimization, float gamma,
                             int reg_power) {
  return make_solver(n, T, a, b, parametric_dist, risk_partitioning_objective, use_rational_optimization, gamma,
                     reg_power)
      .get_optimal_score_extern();
}

// python_dpsolver.cpp:155-176
std::vector<ps_scored_partition> optimize_all__DP(int n, int T, std::vector<float> a, std::vector<float> b,
                                                  int parametric_dist, bool risk_partitioning_objective,
                                                  bool use_rational_optimization, float gamma, int reg_power) {
  return make_solver(n, T, a, b, parametric_dist, risk_partitioning_objective, use_rational_optimization, gamma,
                     reg_power, true)
      .get_all_subsets_and_scores_extern();
}

// python_dpsolver.cpp:179-201
ps_scored_partition optimize_one__DP(int n, int T, std::vector<float> a, std::vector<float> b, int parametric_dist,
                                     bool risk_partitioning_objective, bool use_rational_optimization, float gamma,
                                     int reg_power) {
  auto dp = make_solver(n, T, a, b, parametric_dist, risk_partitioning_objective, use_rational_optimization, gamma,
                        reg_power);
  return std::make_pair(dp.get_optimal_subsets_extern(), dp.get_optimal_score_extern());
}

// python_dpsolver.cpp:203-228
ps_scored_partition sweep_best_OLS__DP(int n, int T, std::vector<float> a, std::vector<float> b, int parametric_dist,
                                       bool risk_partitioning_objective, bool use_rational_optimization,
                                       float gamma, int reg_power) {
  auto dp = make_solver(n, T, a, b, parametric_dist, risk_partitioning_objective, use_rational_optimization, gamma,
                        reg_power, false, true);
  return std::make_pair(dp.get_optimal_subsets_extern(), dp.get_optimal_score_extern());
}

// python_dpsolver.cpp:230-261: the reference constructs T-1 solvers (t = T..2) and keeps the strictly best score.  The
// levels of one T-solve are the levels of every t-solve (DPSolver::sweep_extern_scores), so ONE pass serves them all.
ps_scored_partition sweep_best__DP(int n, int T, std::vector<float> a, std::vector<float> b, int parametric_dist,
                                   bool risk_partitioning_objective, bool use_rational_optimization, float gamma,
                                   int reg_power) {
  float best = std::numeric_limits<float>::lowest();
  ps_subsets subsets;
  if (T < 2) return std::make_pair(subsets, best);  // the reference's loop body never runs
  auto dp = make_solver(n, T, a, b, parametric_dist, risk_partitioning_objective, use_rational_optimization, gamma,
                        reg_power, true);
  const auto all = dp.get_all_subsets_and_scores_extern();
  const auto& ext = dp.sweep_extern_scores();
  for (int t = static_cast<int>(all.size()) - 1; t > 1; --t) {  // t > T_ (T capped at n) repeats t = T_
    if (ext[t] > best) {
      best = ext[t];
      subsets = all[t].first;
    }
  }
  return std::make_pair(subsets, best);
}

// python_dpsolver.cpp:263-316: the reference runs T..1 separate solves on its thread pool and files each result under the
// number of subsets it returned.  Here: one pass to level T, T backtraces (same results, see sweep_best__DP).
std::vector<ps_scored_partition> sweep_parallel__DP(int n, int T, std::vector<float> a, std::vector<float> b,
                                                    int parametric_dist, bool risk_partitioning_objective,
                                                    bool use_rational_optimization, float gamma, int reg_power) {
  if (parametric_dist < 0 || parametric_dist > 2) throw distributionException();
  std::vector<ps_scored_partition> results(static_cast<size_t>(std::max(T, 0) + 1));
  if (T < 1) return results;
  auto dp = make_solver(n, T, a, b, parametric_dist, risk_partitioning_objective, use_rational_optimization, gamma,
                        reg_power, true);
  const auto all = dp.get_all_subsets_and_scores_extern();
  const auto& ext = dp.sweep_extern_scores();
  for (size_t t = 1; t < all.size() && t < results.size(); ++t) results[t] = std::make_pair(all[t].first, ext[t]);
  return results;
}
