// python_dpsolver.hpp -- the flat free-function API the bindings export, with the reference's exact
// signatures (reference: src/cpp/include/python_dpsolver.hpp:16-109).  Bodies live in python_dpsolver.cpp
// and construct the engine-backed DPSolver<float> from <partitionsolvers/DP.hpp>.
#ifndef PS_B200_PYTHON_DPSOLVER_HPP
#define PS_B200_PYTHON_DPSOLVER_HPP

#include <utility>
#include <vector>

#include <partitionsolvers/DP.hpp>

using ps_subsets = std::vector<std::vector<int> >;
using ps_scored_partition = std::pair<ps_subsets, float>;

std::pair<std::vector<ps_scored_partition>, int> compute_optimal_num_clusters_OLS_all(
    int n, int T, std::vector<float> a, std::vector<float> b, int parametric_dist,
    bool risk_partitioning_objective, bool use_rational_optimization, float gamma = 0., int reg_power = 1.);

int compute_optimal_num_clusters_OLS(int n, int T, std::vector<float> a, std::vector<float> b, int parametric_dist,
                                     bool risk_partitioning_objective, bool use_rational_optimization,
                                     float gamma = 0., int reg_power = 1.);

float compute_score(std::vector<float> a, std::vector<float> b, int parametric_dist,
                    bool risk_partitioning_objective, bool use_rational_optimization);

ps_subsets find_optimal_partition__DP(int n, int T, std::vector<float> a, std::vector<float> b, int parametric_dist,
                                      bool risk_partitioning_objective, bool use_rational_optimization,
                                      float gamma = 0., int reg_power = 1);

float find_optimal_score__DP(int n, int T, std::vector<float> a, std::vector<float> b, int parametric_dist,
                             bool risk_partitioning_objective, bool use_rational_optimization, float gamma = 0.,
                             int reg_power = 1);

ps_scored_partition optimize_one__DP(int n, int T, std::vector<float> a, std::vector<float> b, int parametric_dist,
                                     bool risk_partitioning_objective, bool use_rational_optimization,
                                     float gamma = 0., int reg_power = 1);

std::vector<ps_scored_partition> optimize_all__DP(int n, int T, std::vector<float> a, std::vector<float> b,
                                                  int parametric_dist, bool risk_partitioning_objective,
                                                  bool use_rational_optimization, float gamma = 0.,
                                                  int reg_power = 1.);

ps_scored_partition sweep_best_OLS__DP(int n, int T, std::vector<float> a, std::vector<float> b, int parametric_dist,
                                       bool risk_partitioning_objective, bool use_rational_optimization,
                                       float gamma = 0., int reg_power = 1.);

ps_scored_partition sweep_best__DP(int n, int T, std::vector<float> a, std::vector<float> b, int parametric_dist,
                                   bool risk_partitioning_objective, bool use_rational_optimization,
                                   float gamma = 0., int reg_power = 1);

std::vector<ps_scored_partition> sweep_parallel__DP(int n, int T, std::vector<float> a, std::vector<float> b,
                                                    int parametric_dist, bool risk_partitioning_objective,
                                                    bool use_rational_optimization, float gamma = 0.,
                                                    int reg_power = 1);

#endif
