// oracle/ref_shim.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" shim around the UNMODIFIED reference solver headers that live under
// /root/reference/src/cpp/include.  It is compiled by oracle/Makefile with the reference's own
// flags (-O3 -mavx2, C++11 mode so the absent Eigen is not needed) into
// oracle/_ref/libps_ref.so.  Nothing from the reference is copied into this repository: the
// headers are included from where they lie.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load the resulting library.
//
// What it exposes (all plain pointers, caller-allocated):
//   ref_dp_solve_{f32,f64}   run DPSolver<D>(n,T,a,b,...) (DP.hpp:43-68) and dump the private state
//                            priority_sortind_, maxScore_, nextStart_ (DP.hpp:162-166), the subsets,
//                            score_by_subset_, optimal_score_ and the getter value
//                            get_optimal_score_extern() (DP_impl.hpp:267-276), plus the sweep table.
//   ref_dp_time_{f32,f64}    wall time of the constructor only (what solver_timer.cpp:71-73 times).
//   ref_dp_time_pool_f32     R independent solvers dispatched on the reference's own ThreadPool
//                            (threadpool.hpp:105-113), the call pattern of sweep_parallel__DP
//                            (python_dpsolver.cpp:263-316) -- the "native-threaded" baseline.
//   ref_score_{f32,f64}      S(0,n) of unsorted inputs == compute_score (python_dpsolver.cpp:74-111)
//   ref_ltss_f32             LTSSSolver<float> (LTSS_impl.hpp:73-111)
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <exception>
#include <functional>
#include <future>
#include <iostream>
#include <iterator>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <condition_variable>
#include <numeric>
#include <queue>
#include <sstream>
#include <iomanip>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>
#include <immintrin.h>
#include <string.h>
#include <math.h>

// The reference keeps the DP tables private; the parity tests need them row by row.
#define private public
#define protected public
#include "DP.hpp"
#include "LTSS.hpp"
#include "threadpool.hpp"
#undef private
#undef protected

namespace {

template <typename D>
int solve_impl(int n, int T, const D* a, const D* b, int dist, int risk_part, D gamma, int reg_power,
               int sweep_down, int* perm, D* max_score, int* next_start, int* n_subsets,
               int* subset_sizes, int* subset_elems, D* score_by_subset, D* optimal_score_getter,
               D* optimal_score_member, int* all_sizes, int* all_elems, D* all_scores) {
  try {
    std::vector<D> av(a, a + n), bv(b, b + n);
    DPSolver<D> dp(n, T, av, bv, static_cast<objective_fn>(dist), risk_part != 0, true, gamma,
                   reg_power, sweep_down != 0, false);
    const int Tc = dp.T_;
    const int nc = dp.n_;
    if (perm) std::copy(dp.priority_sortind_.begin(), dp.priority_sortind_.end(), perm);
    if (max_score)
      for (int j = 0; j <= Tc; ++j)
        std::copy(dp.maxScore_[j].begin(), dp.maxScore_[j].end(), max_score + (size_t)j * nc);
    if (next_start)
      for (int j = 0; j <= Tc; ++j)
        std::copy(dp.nextStart_[j].begin(), dp.nextStart_[j].end(), next_start + (size_t)j * nc);
    auto subsets = dp.get_optimal_subsets_extern();
    if (n_subsets) *n_subsets = (int)subsets.size();
    if (subset_sizes && subset_elems) {
      int off = 0;
      for (size_t s = 0; s < subsets.size(); ++s) {
        subset_sizes[s] = (int)subsets[s].size();
        for (int v : subsets[s]) subset_elems[off++] = v;
      }
    }
    if (score_by_subset) {
      auto sbs = dp.get_score_by_subset_extern();
      std::copy(sbs.begin(), sbs.end(), score_by_subset);
    }
    if (optimal_score_getter) *optimal_score_getter = dp.get_optimal_score_extern();
    if (optimal_score_member) *optimal_score_member = dp.optimal_score_;
    if (sweep_down && all_sizes && all_elems && all_scores) {
      auto all = dp.get_all_subsets_and_scores_extern();  // T+1 entries, entry 0 empty
      for (int S = 0; S <= Tc; ++S) {
        all_scores[S] = all[S].second;
        int off = 0;
        for (int s = 0; s < Tc; ++s) all_sizes[(size_t)S * Tc + s] = 0;
        for (size_t s = 0; s < all[S].first.size(); ++s) {
          all_sizes[(size_t)S * Tc + s] = (int)all[S].first[s].size();
          for (int v : all[S].first[s]) all_elems[(size_t)S * nc + off++] = v;
        }
      }
    }
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "ref_shim: " << e.what() << std::endl;
    return 1;
  }
}

template <typename D>
double time_impl(int n, int T, const D* a, const D* b, int dist, int risk_part, D* score_out) {
  std::vector<D> av(a, a + n), bv(b, b + n);
  auto t0 = std::chrono::steady_clock::now();
  DPSolver<D> dp(n, T, av, bv, static_cast<objective_fn>(dist), risk_part != 0, true);
  auto t1 = std::chrono::steady_clock::now();
  if (score_out) *score_out = dp.get_optimal_score_extern();
  return std::chrono::duration<double>(t1 - t0).count();
}

// S(i,j) table probe: returns ParametricContext::get_score(i,j) for a list of (i,j) pairs on the
// inputs AS GIVEN (no sort) -- python_dpsolver.cpp:74-111 uses (0,n).
template <typename D>
int score_pairs_impl(int n, const D* a, const D* b, int dist, int risk_part, int npairs,
                            const int* ii, const int* jj, D* out) {
  try {
    std::vector<D> av(a, a + n), bv(b, b + n);
    std::unique_ptr<ParametricContext<D>> ctx;
    objective_fn f = static_cast<objective_fn>(dist);
    if (f == objective_fn::Gaussian)
      ctx.reset(new GaussianContext<D>(av, bv, n, risk_part != 0, true));
    else if (f == objective_fn::Poisson)
      ctx.reset(new PoissonContext<D>(av, bv, n, risk_part != 0, true));
    else if (f == objective_fn::RationalScore)
      ctx.reset(new RationalScoreContext<D>(av, bv, n, risk_part != 0, true));
    else
      return 2;
    ctx->init();
    for (int p = 0; p < npairs; ++p) out[p] = ctx->get_score(ii[p], jj[p]);
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "ref_shim: " << e.what() << std::endl;
    return 1;
  }
}

}  // namespace

extern "C" {

int ref_dp_solve_f32(int n, int T, const float* a, const float* b, int dist, int risk_part,
                     float gamma, int reg_power, int sweep_down, int* perm, float* max_score,
                     int* next_start, int* n_subsets, int* subset_sizes, int* subset_elems,
                     float* score_by_subset, float* optimal_score_getter,
                     float* optimal_score_member, int* all_sizes, int* all_elems,
                     float* all_scores) {
  return solve_impl<float>(n, T, a, b, dist, risk_part, gamma, reg_power, sweep_down, perm,
                           max_score, next_start, n_subsets, subset_sizes, subset_elems,
                           score_by_subset, optimal_score_getter, optimal_score_member, all_sizes,
                           all_elems, all_scores);
}

int ref_dp_solve_f64(int n, int T, const double* a, const double* b, int dist, int risk_part,
                     double gamma, int reg_power, int sweep_down, int* perm, double* max_score,
                     int* next_start, int* n_subsets, int* subset_sizes, int* subset_elems,
                     double* score_by_subset, double* optimal_score_getter,
                     double* optimal_score_member, int* all_sizes, int* all_elems,
                     double* all_scores) {
  return solve_impl<double>(n, T, a, b, dist, risk_part, gamma, reg_power, sweep_down, perm,
                            max_score, next_start, n_subsets, subset_sizes, subset_elems,
                            score_by_subset, optimal_score_getter, optimal_score_member, all_sizes,
                            all_elems, all_scores);
}

double ref_dp_time_f32(int n, int T, const float* a, const float* b, int dist, int risk_part,
                       float* score_out) {
  return time_impl<float>(n, T, a, b, dist, risk_part, score_out);
}

double ref_dp_time_f64(int n, int T, const double* a, const double* b, int dist, int risk_part,
                       double* score_out) {
  return time_impl<double>(n, T, a, b, dist, risk_part, score_out);
}

// R independent instances (row-major a[R][n], b[R][n]) pushed through the reference's own
// process-global pool; returns wall seconds, fills scores[R].  Worker count is whatever the
// reference picks: hardware_concurrency()-1 (threadpool.hpp:105-113).
double ref_dp_time_pool_f32(int R, int n, int T, const float* a, const float* b, int dist,
                            int risk_part, float* scores, int* n_workers) {
  if (n_workers) *n_workers = (int)std::max(std::thread::hardware_concurrency(), 2u) - 1;
  std::vector<ThreadPool::TaskFuture<void>> futs;
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < R; ++r) {
    const float* ar = a + (size_t)r * n;
    const float* br = b + (size_t)r * n;
    float* out = scores ? scores + r : nullptr;
    auto task = [=]() {
      std::vector<float> av(ar, ar + n), bv(br, br + n);
      DPSolver<float> dp(n, T, av, bv, static_cast<objective_fn>(dist), risk_part != 0, true);
      if (out) *out = dp.get_optimal_score_extern();
    };
    futs.push_back(DefaultThreadPool::submitJob(task));
  }
  for (auto& f : futs) f.get();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

int ref_score_pairs_f32(int n, const float* a, const float* b, int dist, int risk_part, int npairs,
                        const int* ii, const int* jj, float* out) {
  return score_pairs_impl<float>(n, a, b, dist, risk_part, npairs, ii, jj, out);
}
int ref_score_pairs_f64(int n, const double* a, const double* b, int dist, int risk_part,
                        int npairs, const int* ii, const int* jj, double* out) {
  return score_pairs_impl<double>(n, a, b, dist, risk_part, npairs, ii, jj, out);
}

int ref_ltss_f32(int n, const float* a, const float* b, int dist, int* perm, int* subset_size,
                 int* subset, float* score) {
  try {
    std::vector<float> av(a, a + n), bv(b, b + n);
    LTSSSolver<float> s(n, av, bv, static_cast<objective_fn>(dist));
    auto sub = s.get_optimal_subset_extern();
    if (perm) std::copy(s.priority_sortind_.begin(), s.priority_sortind_.end(), perm);
    if (subset_size) *subset_size = (int)sub.size();
    if (subset) std::copy(sub.begin(), sub.end(), subset);
    if (score) *score = s.get_optimal_score_extern();
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "ref_shim: " << e.what() << std::endl;
    return 1;
  }
}

int ref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

// The constructor called with n < a.size() (DP_impl.hpp:10-27: the sort spans the whole vectors, the first n sorted elements
// are kept): subsets and the getter's score.
int ref_dp_solve_trunc_f32(int n, int len, int T, const float* a, const float* b, int dist, int risk_part,
                           int* n_subsets, int* subset_sizes, int* subset_elems, float* score) {
  try {
    std::vector<float> av(a, a + len), bv(b, b + len);
    DPSolver<float> dp(n, T, av, bv, static_cast<objective_fn>(dist), risk_part != 0, true);
    auto subsets = dp.get_optimal_subsets_extern();
    *n_subsets = (int)subsets.size();
    int off = 0;
    for (size_t s = 0; s < subsets.size(); ++s) {
      subset_sizes[s] = (int)subsets[s].size();
      for (int v : subsets[s]) subset_elems[off++] = v;
    }
    *score = dp.get_optimal_score_extern();
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "ref_shim: " << e.what() << std::endl;
    return 1;
  }
}

}  // extern "C"
