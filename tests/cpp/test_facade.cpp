// tests/cpp/test_facade.cpp -- the reference's gtest cases (src/cpp/test/gtest_all.cpp) re-expressed
// against the engine-backed facade <partitionsolvers/DP.hpp>, without gtest (not installed here).
// Each block names the reference test it follows.  Needs a GPU; run by tests/test_gpu_facade.py.
#include <partitionsolvers/DP.hpp>
#include <partitionsolvers/replicas.hpp>

#include <cstdio>
#include <random>
#include <thread>

#include "../../partitionsolvers_b200/cxx/python_dpsolver.hpp"

static int g_fail = 0;
#define CHECK(cond)                                                         \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);           \
      ++g_fail;                                                             \
    }                                                                       \
  } while (0)

static void gen(int n, unsigned seed, bool positive, std::vector<float>& a, std::vector<float>& b) {
  std::mt19937 eng(seed);
  std::uniform_real_distribution<float> da(positive ? 0.1f : -10.f, 10.f), db(0.1f, 10.f);
  a.resize(n);
  b.resize(n);
  for (int i = 0; i < n; ++i) {
    a[i] = da(eng);
    b[i] = db(eng);
  }
}

static float rational_obj(const std::vector<float>& a, const std::vector<float>& b, int s, int e) {
  if (s == e) return 0.f;
  float num = 0.f, den = 0.f;
  for (int i = s; i < e; ++i) {
    num += a[i];
    den += b[i];
  }
  return num * num / den;
}

int main() {
  const objective_fn objs[] = {objective_fn::Gaussian, objective_fn::Poisson, objective_fn::RationalScore};

  {  // TestBaselines, gtest_all.cpp:373-419
    std::vector<float> a{0.0212651,   -0.20654906, -0.20654906, -0.20654906, -0.20654906, 0.0212651,  -0.20654906,
                         0.0212651,   -0.20654906, 0.0212651,   -0.20654906, 0.0212651,   -0.20654906, -0.06581402,
                         0.0212651,   0.03953075,  -0.20654906, 0.16200014,  0.0212651,   -0.20654906, 0.20296943,
                         -0.18828341, -0.20654906, -0.20654906, -0.06581402, -0.20654906, 0.16200014,  0.03953075,
                         -0.20654906, -0.20654906, 0.03953075,  0.20296943,  -0.20654906, 0.0212651,   0.20296943,
                         -0.20654906, 0.0212651,   0.03953075,  -0.20654906, 0.03953075};
    std::vector<float> b{0.22771114, 0.21809504, 0.21809504, 0.21809504, 0.21809504, 0.22771114, 0.21809504, 0.22771114,
                         0.21809504, 0.22771114, 0.21809504, 0.22771114, 0.21809504, 0.22682739, 0.22771114, 0.22745816,
                         0.21809504, 0.2218354,  0.22771114, 0.21809504, 0.218429,   0.219738,   0.21809504, 0.21809504,
                         0.22682739, 0.21809504, 0.2218354,  0.22745816, 0.21809504, 0.21809504, 0.22745816, 0.218429,
                         0.21809504, 0.22771114, 0.218429,   0.21809504, 0.22771114, 0.22745816, 0.21809504, 0.22745816};
    std::vector<std::vector<int> > expected = {
        {1, 2, 3, 4, 6, 8, 10, 12, 16, 19, 21, 22, 23, 25, 28, 29, 32, 35, 38},
        {13, 24},
        {0, 5, 7, 9, 11, 14, 15, 18, 27, 30, 33, 36, 37, 39},
        {17, 26},
        {20, 31, 34}};
    auto dp = DPSolver<float>(40, 5, a, b, objective_fn::Gaussian, true, true);
    auto opt = dp.get_optimal_subsets_extern();
    CHECK(opt.size() == expected.size());
    for (size_t i = 0; i < expected.size() && i < opt.size(); ++i) {
      std::sort(opt[i].begin(), opt[i].end());
      CHECK(opt[i] == expected[i]);
    }
    std::vector<float> a1{2.26851454, 2.86139335, 5.51314769, 6.84829739, 6.96469186, 7.1946897, 9.80764198, 4.2310646};
    std::vector<float> b1{3.43178016, 3.92117518, 7.29049707, 7.37995406, 4.80931901, 4.38572245, 3.98044255, 0.59677897};
    auto dp1 = DPSolver<float>(8, 3, a1, b1, objective_fn::Poisson, false, true);
    auto o1 = dp1.get_optimal_subsets_extern();
    CHECK((o1 == std::vector<std::vector<int> >{{0, 1, 2, 3}, {4, 5, 6}, {7}}));
  }

  for (auto obj : objs) {
    const bool pos = obj == objective_fn::Poisson;
    {  // TestOrderedProperty, gtest_all.cpp:421-459: pre-sorted input -> runs of consecutive indices
      std::vector<float> a, b;
      gen(100, 11, pos, a, b);
      std::vector<int> ind(100);
      std::iota(ind.begin(), ind.end(), 0);
      std::stable_sort(ind.begin(), ind.end(), [&](int i, int j) { return a[i] / b[i] < a[j] / b[j]; });
      std::vector<float> as(100), bs(100);
      for (int i = 0; i < 100; ++i) {
        as[i] = a[ind[i]];
        bs[i] = b[ind[i]];
      }
      auto dp = DPSolver<float>(100, 8, as, bs, obj, false, true);
      for (auto& s : dp.get_optimal_subsets_extern()) {
        std::vector<int> v(s);
        std::sort(v.begin(), v.end());
        for (size_t q = 1; q < v.size(); ++q) CHECK(v[q] == v[q - 1] + 1);
      }
    }
    {  // TestSinglePartitionSpecification, gtest_all.cpp:461-498
      std::vector<float> a, b;
      gen(50, 12, pos, a, b);
      auto dp = DPSolver<float>(50, 1, a, b, obj, true, true);
      auto s = dp.get_optimal_subsets_extern();
      CHECK(s.size() == 1 && s[0].size() == 50);
    }
    {  // TestSweepResultsMatchSingleTResults, gtest_all.cpp:603-662
      std::vector<float> a, b;
      gen(150, 13, pos, a, b);
      const int T = 10;
      auto sw = DPSolver<float>(150, T, a, b, obj, true, true, 0., 1, true, false).get_all_subsets_and_scores_extern();
      CHECK((int)sw.size() == T + 1);
      for (int S = 1; S <= T; ++S) {
        auto one = DPSolver<float>(150, S, a, b, obj, true, true);
        CHECK(one.get_optimal_score_extern() == sw[S].second);
        CHECK(one.get_optimal_subsets_extern() == sw[S].first);
      }
    }
    {  // TestSweepWithOptimalTMatchesManualSweep, gtest_all.cpp:735-856: in-solver sweep == threaded solves
      std::vector<float> a, b;
      gen(200, 14, pos, a, b);
      const int T = 8;
      auto sw = optimize_all__DP(200, T, a, b, (int)obj, true, true, 0.f, 1);
      auto par = sweep_parallel__DP(200, T, a, b, (int)obj, true, true, 0.f, 1);
      CHECK(par.size() == sw.size());
      for (int S = 1; S <= T; ++S) {
        CHECK(par[S].second == sw[S].second);
        CHECK(par[S].first == sw[S].first);
      }
    }
    {  // TestConsistencyofOLSReturnedPartitionSize, gtest_all.cpp:664-733
      std::vector<float> a, b;
      gen(2000, 15, pos, a, b);
      auto dp = DPSolver<float>(2000, 10, a, b, obj, true, true, 0., 1, false, true);
      auto all = dp.get_all_subsets_and_scores_extern();
      const int t = dp.get_optimal_num_clusters_OLS_extern();
      CHECK(t >= 1 && t <= 10);
      CHECK((int)dp.get_optimal_subsets_extern().size() == t);
      CHECK(all.size() == 11);
      for (size_t i = 0; i < all.size(); ++i) CHECK(all[i].first.size() == i);
    }
  }

  {  // TestOptimalityWithRandomPartitions, gtest_all.cpp:960-1029 (Rational objective)
    std::vector<float> a, b;
    gen(100, 16, true, a, b);
    auto dp = DPSolver<float>(100, 3, a, b, objective_fn::RationalScore, true, true);
    auto perm = dp.priority_sortind();
    std::vector<float> as(100), bs(100);
    for (int i = 0; i < 100; ++i) {
      as[i] = a[perm[i]];
      bs[i] = b[perm[i]];
    }
    const float best = dp.get_optimal_score_extern();
    std::mt19937 eng(5);
    std::uniform_int_distribution<int> cut(1, 99);
    for (int trial = 0; trial < 500; ++trial) {
      int c1 = cut(eng), c2 = cut(eng);
      if (c1 == c2) continue;
      if (c1 > c2) std::swap(c1, c2);
      float tot = rational_obj(as, bs, 0, c1) + rational_obj(as, bs, c1, c2) + rational_obj(as, bs, c2, 100);
      CHECK(tot <= best * (1.f + 1e-5f));
    }
  }

  {  // TestDisreteValuedPriorityFunction, gtest_all.cpp:924-958: massive ties, OLS t >= 1, no crash
    std::mt19937 eng(17);
    std::uniform_int_distribution<int> d(1, 2);
    std::vector<float> a(500), b(500, 1.f);
    for (auto& v : a) v = (float)d(eng);
    CHECK(compute_optimal_num_clusters_OLS(500, 6, a, b, (int)objective_fn::RationalScore, true, true) >= 1);
  }

  {  // error behaviour: bad enum -> distributionException (DP_impl.hpp:62-64)
    std::vector<float> a(10, 1.f), b(10, 1.f);
    bool thrown = false;
    try {
      DPSolver<float>(10, 2, a, b, static_cast<objective_fn>(9), true, true);
    } catch (const distributionException&) {
      thrown = true;
    }
    CHECK(thrown);
  }

  {  // double instantiation + move-assign (benchmarks.cpp:120-145 usage) + concurrent construction
    std::vector<double> a(300), b(300);
    std::mt19937 eng(18);
    std::uniform_real_distribution<double> d(0.1, 10.);
    for (int i = 0; i < 300; ++i) {
      a[i] = d(eng);
      b[i] = d(eng);
    }
    DPSolver<double> dp;
    dp = DPSolver<double>(300, 5, a, b, objective_fn::Poisson, false, true);
    const double ref = dp.get_optimal_score_extern();
    double got[4];
    std::vector<std::thread> th;
    for (int q = 0; q < 4; ++q)
      th.emplace_back([&, q]() { got[q] = DPSolver<double>(300, 5, a, b, objective_fn::Poisson, false, true)
                                               .get_optimal_score_extern(); });
    for (auto& t : th) t.join();
    for (int q = 0; q < 4; ++q) CHECK(got[q] == ref);
  }

  {  // ReplicaSolver: R instances in one call (over every visible device) == R separate DPSolver constructions.  Count
     // data (tied priorities: both take the engine's stable order), both objectives, R not a multiple of the device count.
    const int R = 37, n = 400, T = 5;
    std::vector<float> a((size_t)R * n), b((size_t)R * n);
    std::mt19937 eng(77);
    std::poisson_distribution<int> pb(100);
    for (size_t i = 0; i < a.size(); ++i) {
      b[i] = (float)std::max(1, pb(eng));
      std::poisson_distribution<int> pa((double)b[i]);
      a[i] = (float)std::max(1, pa(eng));
    }
    for (int rp = 0; rp < 2; ++rp) {
      partitionsolvers::ReplicaSolver<float> rs(R, n, T, a.data(), b.data(), objective_fn::Poisson, rp != 0);
      CHECK(rs.replicas() == R && rs.devices_used() >= 1 && rs.devices_used() <= ps_device_count());
      for (int r = 0; r < R; ++r) {
        std::vector<float> ar(a.begin() + (size_t)r * n, a.begin() + (size_t)(r + 1) * n);
        std::vector<float> br(b.begin() + (size_t)r * n, b.begin() + (size_t)(r + 1) * n);
        DPSolver<float> one(n, T, ar, br, objective_fn::Poisson, rp != 0, true);
        CHECK(rs.get_optimal_score_extern(r) == one.get_optimal_score_extern());
        CHECK(rs.get_optimal_subsets_extern(r) == one.get_optimal_subsets_extern());
        CHECK(rs.get_score_by_subset_extern(r) == one.get_score_by_subset_extern());
      }
      // one device asked for: same results
      partitionsolvers::ReplicaSolver<float> r1(R, n, T, a.data(), b.data(), objective_fn::Poisson, rp != 0, 1);
      CHECK(r1.devices_used() == 1);
      for (int r = 0; r < R; ++r) CHECK(r1.get_optimal_score_extern(r) == rs.get_optimal_score_extern(r));
    }
    std::printf("ReplicaSolver: %d device(s) visible\n", ps_device_count());
  }

  if (g_fail) {
    std::printf("%d check(s) failed\n", g_fail);
    return 1;
  }
  std::printf("facade: all checks passed\n");
  return 0;
}
