// partitionsolvers/replicas.hpp -- R independent DPSolver instances of one shape in one call, over every visible device.
//
// The reference has no batched solver: its Monte-Carlo replicas are one DPSolver per instance, spread over a
// multiprocessing pool (src/python/SimulatedExperiments.py:178-222) or a ThreadPool (python_dpsolver.cpp:275-301).
// Here the R instances go through the engine's batch dimension (ps_problem.batch, include/ps_b200.h), and when more than
// one device is visible the batch is dealt out in contiguous slices, one per device, each slice solved by its own host
// thread on its own engine context (contexts are created once per device and kept for the life of the process).  No
// data-path collective: the instances are independent.
//
// Per-replica getters return exactly what DPSolver<DataType>(n, T, a_r, b_r, dist, risk_partitioning) returns for that
// replica (DP_impl.hpp:122-153 subsets, :232-259 reorder, :267-276 score getter); tests/cpp/test_facade.cpp pins that.
#ifndef PARTITIONSOLVERS_REPLICAS_HPP
#define PARTITIONSOLVERS_REPLICAS_HPP

#include <memory>
#include <mutex>
#include <thread>

#include "DP.hpp"

namespace partitionsolvers {

// One engine context per device, created on first use, shared by the host threads of successive ReplicaSolver calls (a
// context serialises its callers: the per-device mutex is held for the duration of a slice).
class DeviceContexts {
 public:
  struct Slot {
    std::mutex mu;
    ps_ctx* ctx = nullptr;
  };
  static DeviceContexts& instance() {
    static DeviceContexts d;
    return d;
  }
  int count() const { return static_cast<int>(slots_.size()); }
  Slot& slot(int device) { return *slots_[device]; }
  ~DeviceContexts() {
    for (auto& s : slots_)
      if (s->ctx) ps_destroy(s->ctx);
  }

 private:
  DeviceContexts() {
    const int n = ps_device_count();
    for (int d = 0; d < n; ++d) slots_.emplace_back(new Slot());
  }
  std::vector<std::unique_ptr<Slot> > slots_;
};

template <typename DataType>
class ReplicaSolver {
 public:
  // a, b: [R][n] row-major host arrays.  max_devices = 0: every visible device (never more devices than replicas).
  ReplicaSolver(int R, int n, int T, const DataType* a, const DataType* b,
                objective_fn parametric_dist = objective_fn::Gaussian, bool risk_partitioning_objective = false,
                int max_devices = 0)
      : R_{R}, n_{n}, T_{T <= n ? T : n}, risk_partitioning_objective_{risk_partitioning_objective} {
    const int dist = static_cast<int>(parametric_dist);
    if (dist < 0 || dist > 2) throw distributionException();
    if (R < 1 || n < 1 || T < 1) throw engine_error("ReplicaSolver: R, n and T must be >= 1");
    DeviceContexts& pool = DeviceContexts::instance();
    if (pool.count() < 1) throw engine_error("ReplicaSolver: no CUDA device");
    int D = (max_devices > 0 && max_devices < pool.count()) ? max_devices : pool.count();
    if (D > R) D = R;
    devices_used_ = D;
    perm_.assign(static_cast<size_t>(R) * n_, 0);
    bounds_.assign(static_cast<size_t>(R) * (T_ + 1), 0);
    scores_.assign(static_cast<size_t>(R) * T_, DataType(0));
    slice_ms_.assign(D, 0.f);

    std::vector<std::string> errors(D);
    auto work = [&](int d) {
      // contiguous slice of the replicas: [lo, hi)
      const long long lo = static_cast<long long>(R) * d / D, hi = static_cast<long long>(R) * (d + 1) / D;
      DeviceContexts::Slot& s = pool.slot(d);
      std::lock_guard<std::mutex> lock(s.mu);
      if (!s.ctx) {
        ps_status st = ps_create(&s.ctx, d);
        if (st != PS_OK) {
          errors[d] = std::string("ps_create(device ") + std::to_string(d) + "): " + ps_status_string(st);
          return;
        }
      }
      ps_problem p;
      std::memset(&p, 0, sizeof(p));
      p.n = n_;
      p.T = T_;
      p.dist = dist;
      p.risk_part = risk_partitioning_objective_ ? 1 : 0;
      p.sum_mode = env_flag("PARTITIONSOLVERS_PREFIX_SUMS") ? PS_SUM_PREFIX : PS_SUM_RUNNING;
      p.batch = static_cast<int>(hi - lo);
      ps_result r;
      std::memset(&r, 0, sizeof(r));
      r.perm = perm_.data() + lo * n_;
      r.bounds = bounds_.data() + lo * (T_ + 1);
      r.subset_scores = scores_.data() + lo * T_;
      ps_status st = solve(s.ctx, &p, a + lo * n_, b + lo * n_, &r);
      if (st != PS_OK) {
        errors[d] = std::string("ps_solve(device ") + std::to_string(d) + "): " + ps_last_error(s.ctx);
        return;
      }
      ps_timings t;
      ps_get_timings(s.ctx, &t);
      slice_ms_[d] = t.total_ms;
    };
    if (D == 1) {
      work(0);
    } else {
      std::vector<std::thread> threads;
      for (int d = 0; d < D; ++d) threads.emplace_back(work, d);
      for (auto& t : threads) t.join();
    }
    for (const std::string& e : errors)
      if (!e.empty()) throw engine_error(e);
  }

  int replicas() const { return R_; }
  int devices_used() const { return devices_used_; }
  // device time of each device's slice (CUDA events on the context's stream), in device order
  const std::vector<float>& slice_ms() const { return slice_ms_; }

  // priority_sortind_ of replica r: sorted position -> original index
  const int* priority_sortind(int r) const { return perm_.data() + static_cast<size_t>(r) * n_; }

  std::vector<std::vector<int> > get_optimal_subsets_extern(int r) const {
    std::vector<std::vector<int> > subsets;
    std::vector<DataType> sc;
    partition(r, subsets, sc);
    return subsets;
  }
  std::vector<DataType> get_score_by_subset_extern(int r) const {
    std::vector<std::vector<int> > subsets;
    std::vector<DataType> sc;
    partition(r, subsets, sc);
    return sc;
  }
  DataType get_optimal_score_extern(int r) const {  // DP_impl.hpp:267-276
    const DataType* s = scores_.data() + static_cast<size_t>(r) * T_;
    if (risk_partitioning_objective_) {  // optimal_score_: accumulated left to right in DataType (DP_impl.hpp:137)
      DataType optimal = 0.;
      for (int t = 0; t < T_; ++t) optimal += s[t];
      return optimal;
    }
    std::vector<DataType> sc(s, s + T_);
    std::stable_sort(sc.begin(), sc.end());
    double acc = 0.;
    for (int q = 1; q < T_; ++q) acc += sc[q];
    return static_cast<DataType>(acc);
  }

 private:
  int R_, n_, T_;
  bool risk_partitioning_objective_;
  int devices_used_ = 0;
  std::vector<int> perm_, bounds_;
  std::vector<DataType> scores_;
  std::vector<float> slice_ms_;

  void partition(int r, std::vector<std::vector<int> >& subsets, std::vector<DataType>& sc) const {
    const int* perm = priority_sortind(r);
    const int* bounds = bounds_.data() + static_cast<size_t>(r) * (T_ + 1);
    const DataType* s = scores_.data() + static_cast<size_t>(r) * T_;
    subsets.assign(T_, std::vector<int>());
    sc.assign(s, s + T_);
    for (int t = 0; t < T_; ++t) {
      const int lo = bounds[t], hi = bounds[t + 1];
      if (lo >= 0 && hi >= lo && hi <= n_) subsets[t].assign(perm + lo, perm + hi);
    }
    if (!risk_partitioning_objective_) {  // reorder_subsets, DP_impl.hpp:232-259
      std::vector<int> ind(T_);
      std::iota(ind.begin(), ind.end(), 0);
      std::stable_sort(ind.begin(), ind.end(), [&sc](int i, int j) { return sc[i] < sc[j]; });
      std::vector<std::vector<int> > ss(T_);
      std::vector<DataType> sv(T_);
      for (int q = 0; q < T_; ++q) {
        ss[q].swap(subsets[ind[q]]);
        sv[q] = sc[ind[q]];
      }
      subsets.swap(ss);
      sc.swap(sv);
    }
  }
};

}  // namespace partitionsolvers

#endif  // PARTITIONSOLVERS_REPLICAS_HPP
