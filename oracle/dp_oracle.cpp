// oracle/dp_oracle.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement ("streaming oracle") of the reference's consecutive-partitions dynamic program,
// written from the algorithm's definition with O(n) memory per level instead of the reference's three
// (n+1)x(n+1) tables, so that it can run at sizes where the reference cannot allocate (n >= 40k).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load it; the product (partitionsolvers_b200/csrc) never links or calls it.
//
// Parity status: PINNED.  tests/test_oracle.py checks this file bit-for-bit (sort index, every
// maxScore_/nextStart_ row, subsets, per-subset scores, objective) against
//   * the reference's own golden vector TestBaselines (src/cpp/test/gtest_all.cpp:373-419), and
//   * fixtures produced by the unmodified reference headers (oracle/_ref, built by oracle/Makefile;
//     generator tools/make_golden.py) committed under tests/golden/.
// The OLS partition-size step is the exception: the reference runs it through Eigen
// (DP_impl.hpp:171-181), which is not available here -- "parity unpinned" for the regression
// coefficients, the integer result is pinned by the reference's property tests only.
//
// Each function cites the reference lines it restates (paths relative to /root/reference).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

namespace {

enum Dist { GAUSSIAN = 0, POISSON = 1, RATIONAL = 2 };  // score.hpp:20-22

// ---- range scores ---------------------------------------------------------------------------------
// Rounding recipes.  The reference is a template over DataType; with DataType=float
// std::log is logf and everything Poisson stays in float, whereas std::pow(float,int) yields double,
// so the Gaussian mult-clust and Rational scores are evaluated in double from float sums and rounded
// once; Gaussian risk-part squares in float then divides in double.  No fused multiply-add anywhere
// (reference flags: -O3 -mavx2 only) -- this file is compiled with -ffp-contract=off.
template <typename D>
struct Score;

template <>
struct Score<float> {
  static inline float eval(int dist, bool rp, float A, float B) {
    switch (dist) {
      case POISSON:
        if (rp) return A * std::log(A / B);                          // score_impl.hpp:374-380
        return (A > B) ? A * std::log(A / B) + B - A : 0.f;         // score_impl.hpp:382-390
      case GAUSSIAN:
        if (rp) {                                                    // score_impl.hpp:435-441
          float sq = A * A;
          return (float)((double)sq / 2. / (double)B);
        }
        if (A > B) {                                                 // score_impl.hpp:426-433
          double a = A, b = B;
          return (float)(.5 * (a * a / b + b) - a);
        }
        return 0.f;
      default: {                                                     // score_impl.hpp:461-474
        double a = A, b = B;
        return (float)(a * a / b);
      }
    }
  }
};

template <>
struct Score<double> {
  static inline double eval(int dist, bool rp, double A, double B) {
    switch (dist) {
      case POISSON:
        if (rp) return A * std::log(A / B);
        return (A > B) ? A * std::log(A / B) + B - A : 0.;
      case GAUSSIAN:
        if (rp) return A * A / 2. / B;
        return (A > B) ? .5 * (A * A / B + B) - A : 0.;
      default:
        return A * A / B;
    }
  }
};

template <typename F>
void parallel_rows(int nrows, int nthreads, F&& body) {
  if (nrows <= 0) return;
  const int grain = 16;
  if (nthreads <= 1 || nrows < 4 * grain) {
    for (int i = 0; i < nrows; ++i) body(i);
    return;
  }
  std::atomic<int> next(0);
  auto worker = [&]() {
    for (;;) {
      int s = next.fetch_add(grain);
      if (s >= nrows) break;
      int e = std::min(nrows, s + grain);
      for (int i = s; i < e; ++i) body(i);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
}

template <typename D>
struct Solver {
  int n, T, dist;
  bool rp;
  int nthreads;
  std::vector<D> a, b;                   // priority-sorted
  std::vector<int> perm;                 // priority_sortind_
  std::vector<std::vector<D>> M;         // maxScore_  (T+1) x n
  std::vector<std::vector<int>> N;       // nextStart_ (T+1) x n

  // S(i,j): sums accumulated left to right starting at i (score_impl.hpp:248-256).
  D range_score(int i, int j) const {
    D A = 0, B = 0;
    for (int k = i; k < j; ++k) {
      A = A + a[k];
      B = B + b[k];
    }
    return Score<D>::eval(dist, rp, A, B);
  }

  // DP_impl.hpp:7-28.  std::sort with the reference's comparator: the quotient is recomputed in D
  // for every comparison; ties land wherever libstdc++'s introsort leaves them, which is why the
  // same library call is used here rather than a hand-rolled sort.
  void sort_by_priority(const D* a_in, const D* b_in, const int* perm_in) {
    perm.resize(n);
    if (perm_in) {
      std::copy(perm_in, perm_in + n, perm.begin());
    } else {
      std::iota(perm.begin(), perm.end(), 0);
      std::sort(perm.begin(), perm.end(), [a_in, b_in](int i, int j) {
        return (a_in[i] / b_in[i]) < (a_in[j] / b_in[j]);
      });
    }
    a.resize(n);
    b.resize(n);
    for (int i = 0; i < n; ++i) {
      a[i] = a_in[perm[i]];
      b[i] = b_in[perm[i]];
    }
  }

  // DP_impl.hpp:83-118.
  void fill() {
    const D low = std::numeric_limits<D>::lowest();
    M.assign(T + 1, std::vector<D>(n, low));
    N.assign(T + 1, std::vector<int>(n, -1));
    for (int i = 0; i < n; ++i) {
      M[0][i] = 0;
      N[0][i] = -1;
    }
    // level 1: S(i,n) for every i (DP_impl.hpp:90-95)
    {
      D* m1 = M[1].data();
      int* n1 = N[1].data();
      parallel_rows(n, nthreads, [&](int r) {
        int i = r;
        m1[i] = range_score(i, n);
        n1[i] = n;
      });
    }
    for (int j = 2; j <= T; ++j) {
      const D* prev = M[j - 1].data();
      D* cur = M[j].data();
      int* arg = N[j].data();
      const int kmax = n - (j - 1);
      const int nrows = (j == T) ? 1 : (n - j + 1);  // DP_impl.hpp:114-116
      const D* ap = a.data();
      const D* bp = b.data();
      const int dist_ = dist;
      const bool rp_ = rp;
      parallel_rows(nrows, nthreads, [=](int i) {
        D best = std::numeric_limits<D>::lowest();
        int bestk = -1;
        D A = 0, B = 0;
        for (int k = i + 1; k <= kmax; ++k) {
          A = A + ap[k - 1];
          B = B + bp[k - 1];
          D v = Score<D>::eval(dist_, rp_, A, B) + prev[k];
          if (v > best) {  // strict: first maximum wins, NaN never wins
            best = v;
            bestk = k;
          }
        }
        cur[i] = best;
        arg[i] = bestk;
      });
    }
  }

  struct Part {
    std::vector<std::vector<int>> subsets;
    std::vector<D> scores;  // per subset, after the mult-clust reorder
    D objective;
  };

  // DP_impl.hpp:122-153 (+ reorder_subsets 232-259).
  Part backtrace(int S, D gamma, int reg_power) const {
    Part p;
    p.subsets.assign(S, std::vector<int>());
    p.scores.assign(S, 0);
    D obj = 0;
    int cur = 0;
    for (int t = S; t > 0; --t) {
      int nxt = N[t][cur];
      for (int i = cur; i < nxt; ++i) p.subsets[S - t].push_back(perm[i]);
      p.scores[S - t] = range_score(cur, nxt);
      obj += p.scores[S - t];
      cur = nxt;
    }
    if (!rp) {
      std::vector<int> ord(S);
      std::iota(ord.begin(), ord.end(), 0);
      const std::vector<D>& sc = p.scores;
      std::stable_sort(ord.begin(), ord.end(), [&sc](int x, int y) { return sc[x] < sc[y]; });
      std::vector<std::vector<int>> ss(S);
      std::vector<D> sv(S);
      for (int s = 0; s < S; ++s) {
        ss[s] = p.subsets[ord[s]];
        sv[s] = p.scores[ord[s]];
      }
      p.subsets.swap(ss);
      p.scores.swap(sv);
    }
    obj -= gamma * std::pow(S, reg_power);  // double arithmetic, stored back into D
    p.objective = obj;
    return p;
  }
};

// DP_impl.hpp:155-198 restated without Eigen: ordinary least squares of y=log(score[S]-score[S-1])
// on x=log(S+1), S=2..T, in float like the reference's MatrixXf; returns argmin residual + 1.
template <typename D>
int ols_num_clusters(const std::vector<D>& scores /* size T+1, scores[0]=0 */) {
  const int m = (int)scores.size() - 2;
  if (m <= 0) return 1;
  std::vector<float> x(m), y(m);
  for (int i = 0; i < m; ++i) {
    D xv = (D)(i + 3);
    D dv = scores[i + 2] - scores[i + 1];
    x[i] = (float)std::log(xv);
    y[i] = (float)std::log(dv);
  }
  double sx = 0, sy = 0, sxx = 0, sxy = 0;
  for (int i = 0; i < m; ++i) {
    sx += x[i];
    sy += y[i];
    sxx += (double)x[i] * x[i];
    sxy += (double)x[i] * y[i];
  }
  double den = m * sxx - sx * sx;
  float b0, b1;
  if (den != 0) {
    b0 = (float)((m * sxy - sx * sy) / den);
    b1 = (float)((sy - (double)b0 * sx) / m);
  } else {  // single sample: minimum-norm solution
    double nrm = (double)x[0] * x[0] + 1.0;
    b0 = (float)(x[0] * y[0] / nrm);
    b1 = (float)(y[0] / nrm);
  }
  int best = 0;
  D minres = std::numeric_limits<D>::max();
  for (int i = 0; i < m; ++i) {
    float r = y[i] - b0 * x[i] - b1 * 1.f;
    if (r < minres) {
      best = i;
      minres = r;
    }
  }
  return best + 1;
}

template <typename D>
int solve_impl(int n, int T, const D* a, const D* b, int dist, int risk_part, D gamma, int reg_power,
               int sweep_down, int nthreads, const int* perm_in, int* perm, D* max_score,
               int* next_start, int* n_subsets, int* subset_sizes, int* subset_elems,
               D* score_by_subset, D* optimal_score_getter, D* optimal_score_member, int* all_sizes,
               int* all_elems, D* all_scores, int* ols_t) {
  if (n < 1 || T < 1 || dist < 0 || dist > 2) return 2;
  Solver<D> s;
  s.n = n;
  s.T = std::min(T, n);  // DP_impl.hpp:35-38
  s.dist = dist;
  s.rp = risk_part != 0;
  s.nthreads = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
  s.sort_by_priority(a, b, perm_in);
  s.fill();
  const int Tc = s.T;
  if (perm) std::copy(s.perm.begin(), s.perm.end(), perm);
  if (max_score)
    for (int j = 0; j <= Tc; ++j) std::copy(s.M[j].begin(), s.M[j].end(), max_score + (size_t)j * n);
  if (next_start)
    for (int j = 0; j <= Tc; ++j)
      std::copy(s.N[j].begin(), s.N[j].end(), next_start + (size_t)j * n);

  // optimize(): DP_impl.hpp:200-230
  std::vector<std::vector<int>> subsets(Tc);
  std::vector<D> sbs(Tc, 0);
  D optimal_score = 0;
  if (sweep_down || ols_t) {
    std::vector<typename Solver<D>::Part> all(Tc + 1);
    std::vector<D> sc(Tc + 1, 0);
    for (int S = Tc; S >= 1; --S) {
      all[S] = s.backtrace(S, gamma, reg_power);
      sc[S] = all[S].objective;
      if (S == Tc) sbs = all[S].scores;
    }
    if (all_sizes && all_elems && all_scores) {
      for (int S = 0; S <= Tc; ++S) {
        all_scores[S] = sc[S];
        for (int q = 0; q < Tc; ++q) all_sizes[(size_t)S * Tc + q] = 0;
        int off = 0;
        for (size_t q = 0; q < all[S].subsets.size(); ++q) {
          all_sizes[(size_t)S * Tc + q] = (int)all[S].subsets[q].size();
          for (int v : all[S].subsets[q]) all_elems[(size_t)S * n + off++] = v;
        }
      }
    }
    if (ols_t) {
      int t = ols_num_clusters<D>(sc);
      *ols_t = t;
      subsets = all[t].subsets;
      optimal_score = all[t].objective;
    }
  } else {
    auto p = s.backtrace(Tc, gamma, reg_power);
    subsets = p.subsets;
    sbs = p.scores;
    optimal_score = p.objective;
  }
  if (n_subsets) *n_subsets = (int)subsets.size();
  if (subset_sizes && subset_elems) {
    int off = 0;
    for (size_t q = 0; q < subsets.size(); ++q) {
      subset_sizes[q] = (int)subsets[q].size();
      for (int v : subsets[q]) subset_elems[off++] = v;
    }
  }
  if (score_by_subset) std::copy(sbs.begin(), sbs.end(), score_by_subset);
  if (optimal_score_member) *optimal_score_member = optimal_score;
  if (optimal_score_getter) {  // DP_impl.hpp:267-276
    if (s.rp) {
      *optimal_score_getter = optimal_score;
    } else {
      double acc = 0.;
      for (int q = 1; q < Tc; ++q) acc += sbs[q];
      *optimal_score_getter = (D)(acc - gamma * std::pow(Tc, reg_power));
    }
  }
  return 0;
}

// Timing entry for the CPU baseline: one full solve (sort + levels + backtrace), all threads.
template <typename D>
double time_impl(int n, int T, const D* a, const D* b, int dist, int risk_part, int nthreads,
                 D* score_out) {
  auto t0 = std::chrono::steady_clock::now();
  D getter = 0;
  solve_impl<D>(n, T, a, b, dist, risk_part, (D)0, 1, 0, nthreads, nullptr, nullptr, nullptr, nullptr,
                nullptr, nullptr, nullptr, nullptr, &getter, nullptr, nullptr, nullptr, nullptr,
                nullptr);
  auto t1 = std::chrono::steady_clock::now();
  if (score_out) *score_out = getter;
  return std::chrono::duration<double>(t1 - t0).count();
}

// One DP row in isolation (DP_impl.hpp:103-113): best/argmax over k=i+1..kmax of S(i,k)+prev[k] on
// ALREADY SORTED inputs.  Lets the GPU tests check sampled rows at sizes where a full CPU solve would
// take minutes: feed the GPU's own level j-1 row as `prev`.
template <typename D>
void rows_impl(int n, int kmax, int dist, int rp, const D* a, const D* b, const D* prev, int nrows,
                      const int* rows, D* best_out, int* arg_out) {
  (void)n;
  for (int r = 0; r < nrows; ++r) {
    const int i = rows[r];
    D best = std::numeric_limits<D>::lowest();
    int bestk = -1;
    D A = 0, B = 0;
    for (int k = i + 1; k <= kmax; ++k) {
      A = A + a[k - 1];
      B = B + b[k - 1];
      D v = Score<D>::eval(dist, rp != 0, A, B) + prev[k];
      if (v > best) {
        best = v;
        bestk = k;
      }
    }
    best_out[r] = best;
    arg_out[r] = bestk;
  }
}
}  // namespace

extern "C" {

int oracle_dp_solve_f32(int n, int T, const float* a, const float* b, int dist, int risk_part,
                        float gamma, int reg_power, int sweep_down, int nthreads,
                        const int* perm_in, int* perm, float* max_score, int* next_start,
                        int* n_subsets, int* subset_sizes, int* subset_elems,
                        float* score_by_subset, float* optimal_score_getter,
                        float* optimal_score_member, int* all_sizes, int* all_elems,
                        float* all_scores, int* ols_t) {
  return solve_impl<float>(n, T, a, b, dist, risk_part, gamma, reg_power, sweep_down, nthreads,
                           perm_in, perm, max_score, next_start, n_subsets, subset_sizes,
                           subset_elems, score_by_subset, optimal_score_getter,
                           optimal_score_member, all_sizes, all_elems, all_scores, ols_t);
}

int oracle_dp_solve_f64(int n, int T, const double* a, const double* b, int dist, int risk_part,
                        double gamma, int reg_power, int sweep_down, int nthreads,
                        const int* perm_in, int* perm, double* max_score, int* next_start,
                        int* n_subsets, int* subset_sizes, int* subset_elems,
                        double* score_by_subset, double* optimal_score_getter,
                        double* optimal_score_member, int* all_sizes, int* all_elems,
                        double* all_scores, int* ols_t) {
  return solve_impl<double>(n, T, a, b, dist, risk_part, gamma, reg_power, sweep_down, nthreads,
                            perm_in, perm, max_score, next_start, n_subsets, subset_sizes,
                            subset_elems, score_by_subset, optimal_score_getter,
                            optimal_score_member, all_sizes, all_elems, all_scores, ols_t);
}

double oracle_dp_time_f32(int n, int T, const float* a, const float* b, int dist, int risk_part,
                          int nthreads, float* score_out) {
  return time_impl<float>(n, T, a, b, dist, risk_part, nthreads, score_out);
}
double oracle_dp_time_f64(int n, int T, const double* a, const double* b, int dist, int risk_part,
                          int nthreads, double* score_out) {
  return time_impl<double>(n, T, a, b, dist, risk_part, nthreads, score_out);
}

// Range score of the inputs as given (no sort), sums accumulated from i: the value the reference's
// table holds at partialSums_[i][j] (score_impl.hpp:248-256).
float oracle_range_score_f32(int n, const float* a, const float* b, int dist, int risk_part, int i,
                             int j) {
  (void)n;
  float A = 0, B = 0;
  for (int k = i; k < j; ++k) {
    A = A + a[k];
    B = B + b[k];
  }
  return Score<float>::eval(dist, risk_part != 0, A, B);
}
double oracle_range_score_f64(int n, const double* a, const double* b, int dist, int risk_part,
                              int i, int j) {
  (void)n;
  double A = 0, B = 0;
  for (int k = i; k < j; ++k) {
    A = A + a[k];
    B = B + b[k];
  }
  return Score<double>::eval(dist, risk_part != 0, A, B);
}

void oracle_dp_rows_f32(int n, int kmax, int dist, int rp, const float* a, const float* b, const float* prev,
                        int nrows, const int* rows, float* best_out, int* arg_out) {
  rows_impl<float>(n, kmax, dist, rp, a, b, prev, nrows, rows, best_out, arg_out);
}
void oracle_dp_rows_f64(int n, int kmax, int dist, int rp, const double* a, const double* b,
                        const double* prev, int nrows, const int* rows, double* best_out, int* arg_out) {
  rows_impl<double>(n, kmax, dist, rp, a, b, prev, nrows, rows, best_out, arg_out);
}

int oracle_ols_f32(int m, const float* scores) {
  return ols_num_clusters<float>(std::vector<float>(scores, scores + m));
}
int oracle_ols_f64(int m, const double* scores) {
  return ols_num_clusters<double>(std::vector<double>(scores, scores + m));
}

// LTSS_impl.hpp:14-35,73-111: stable sort by priority, best prefix S(0,i) then best suffix S(i,n)
// under the mult-clust score, float result regardless of D (LTSS.hpp:53,59).
int oracle_ltss_f32(int n, const float* a_in, const float* b_in, int dist, int* perm_out,
                    int* subset_size, int* subset, float* score) {
  std::vector<int> perm(n);
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [a_in, b_in](int i, int j) {
    return (a_in[i] / b_in[i]) < (a_in[j] / b_in[j]);
  });
  std::vector<float> a(n), b(n);
  for (int i = 0; i < n; ++i) {
    a[i] = a_in[perm[i]];
    b[i] = b_in[perm[i]];
  }
  float best = -std::numeric_limits<float>::max();
  int lo = 0, hi = 0;
  {
    float A = 0, B = 0;
    for (int i = 1; i <= n; ++i) {
      A = A + a[i - 1];
      B = B + b[i - 1];
      float s = Score<float>::eval(dist, false, A, B);
      if (s > best) {
        best = s;
        lo = 0;
        hi = i;
      }
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    float A = 0, B = 0;
    for (int k = i; k < n; ++k) {
      A = A + a[k];
      B = B + b[k];
    }
    float s = Score<float>::eval(dist, false, A, B);
    if (s > best) {
      best = s;
      lo = i;
      hi = n;
    }
  }
  if (perm_out) std::copy(perm.begin(), perm.end(), perm_out);
  if (subset_size) *subset_size = hi - lo;
  if (subset)
    for (int i = lo; i < hi; ++i) subset[i - lo] = perm[i];
  if (score) *score = best;
  return 0;
}

// The host libm's log / logf on arrays: what the reference's std::log calls resolve to on this machine.
void oracle_log_f64(long long n, const double* x, double* y) {
  for (long long i = 0; i < n; ++i) y[i] = std::log(x[i]);
}
void oracle_log_f32(long long n, const float* x, float* y) {
  for (long long i = 0; i < n; ++i) y[i] = std::log(x[i]);
}

int oracle_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
