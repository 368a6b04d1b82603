// proto_pybind.cpp -- pybind11 twin of the reference's SWIG module `proto` (src/cpp/proto.i): same module
// name, function names, argument order and defaults.  SWIG's std_vector.i/std_pair.i typemaps return
// tuples (of tuples), so results are converted to tuples here; solver_tests.py:79-87 relies on that.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "python_dpsolver.hpp"
#include "python_ltsssolver.hpp"

namespace py = pybind11;

namespace {

py::tuple subsets_to_tuple(const ps_subsets& s) {
  py::tuple out(s.size());
  for (size_t i = 0; i < s.size(); ++i) {
    py::tuple inner(s[i].size());
    for (size_t j = 0; j < s[i].size(); ++j) inner[j] = py::int_(s[i][j]);
    out[i] = inner;
  }
  return out;
}
py::tuple scored_to_tuple(const ps_scored_partition& p) {
  return py::make_tuple(subsets_to_tuple(p.first), py::float_(p.second));
}
py::tuple all_to_tuple(const std::vector<ps_scored_partition>& v) {
  py::tuple out(v.size());
  for (size_t i = 0; i < v.size(); ++i) out[i] = scored_to_tuple(v[i]);
  return out;
}

using fvec = std::vector<float>;

}  // namespace

PYBIND11_MODULE(proto, m) {
  m.doc() = "B200-backed drop-in for PartitionSolvers' SWIG module `proto` (DP entry points)";
  py::register_exception<distributionException>(m, "distributionException");
  py::register_exception<partitionsolvers::engine_error>(m, "EngineError");

#define PS_ARGS                                                                                             \
  py::arg("n"), py::arg("T"), py::arg("a"), py::arg("b"), py::arg("parametric_dist"),                       \
      py::arg("risk_partitioning_objective"), py::arg("use_rational_optimization"), py::arg("gamma") = 0.f, \
      py::arg("reg_power") = 1

  m.def("optimize_one__DP", [](int n, int T, fvec a, fvec b, int d, bool rp, bool ro, float g, int r) {
    return scored_to_tuple(optimize_one__DP(n, T, std::move(a), std::move(b), d, rp, ro, g, r));
  }, PS_ARGS);
  m.def("optimize_all__DP", [](int n, int T, fvec a, fvec b, int d, bool rp, bool ro, float g, int r) {
    return all_to_tuple(optimize_all__DP(n, T, std::move(a), std::move(b), d, rp, ro, g, r));
  }, PS_ARGS);
  m.def("find_optimal_partition__DP", [](int n, int T, fvec a, fvec b, int d, bool rp, bool ro, float g, int r) {
    return subsets_to_tuple(find_optimal_partition__DP(n, T, std::move(a), std::move(b), d, rp, ro, g, r));
  }, PS_ARGS);
  m.def("find_optimal_score__DP", [](int n, int T, fvec a, fvec b, int d, bool rp, bool ro, float g, int r) {
    return find_optimal_score__DP(n, T, std::move(a), std::move(b), d, rp, ro, g, r);
  }, PS_ARGS);
  m.def("sweep_best__DP", [](int n, int T, fvec a, fvec b, int d, bool rp, bool ro, float g, int r) {
    return scored_to_tuple(sweep_best__DP(n, T, std::move(a), std::move(b), d, rp, ro, g, r));
  }, PS_ARGS);
  m.def("sweep_best_OLS__DP", [](int n, int T, fvec a, fvec b, int d, bool rp, bool ro, float g, int r) {
    return scored_to_tuple(sweep_best_OLS__DP(n, T, std::move(a), std::move(b), d, rp, ro, g, r));
  }, PS_ARGS);
  m.def("sweep_parallel__DP", [](int n, int T, fvec a, fvec b, int d, bool rp, bool ro, float g, int r) {
    std::vector<ps_scored_partition> res;
    {
      py::gil_scoped_release nogil;
      res = sweep_parallel__DP(n, T, std::move(a), std::move(b), d, rp, ro, g, r);
    }
    return all_to_tuple(res);
  }, PS_ARGS);
  m.def("compute_optimal_num_clusters_OLS", [](int n, int T, fvec a, fvec b, int d, bool rp, bool ro, float g, int r) {
    return compute_optimal_num_clusters_OLS(n, T, std::move(a), std::move(b), d, rp, ro, g, r);
  }, PS_ARGS);
  m.def("compute_optimal_num_clusters_OLS_all", [](int n, int T, fvec a, fvec b, int d, bool rp, bool ro, float g, int r) {
    auto res = compute_optimal_num_clusters_OLS_all(n, T, std::move(a), std::move(b), d, rp, ro, g, r);
    return py::make_tuple(all_to_tuple(res.first), res.second);
  }, PS_ARGS);
  m.def("compute_score", [](fvec a, fvec b, int d, bool rp, bool ro) {
    return compute_score(std::move(a), std::move(b), d, rp, ro);
  }, py::arg("a"), py::arg("b"), py::arg("parametric_dist"), py::arg("risk_partitioning_objective"),
        py::arg("use_rational_optimization"));
#undef PS_ARGS
  // LTSS entry points (reference python_ltsssolver.hpp:10-23)
  auto ints_to_tuple = [](const std::vector<int>& v) {
    py::tuple out(v.size());
    for (size_t i = 0; i < v.size(); ++i) out[i] = py::int_(v[i]);
    return out;
  };
  m.def("find_optimal_partition__LTSS", [ints_to_tuple](int n, fvec a, fvec b, int d) {
    return ints_to_tuple(find_optimal_partition__LTSS(n, std::move(a), std::move(b), d));
  }, py::arg("n"), py::arg("a"), py::arg("b"), py::arg("parametric_dist"));
  m.def("find_optimal_score__LTSS", [](int n, fvec a, fvec b, int d) {
    return find_optimal_score__LTSS(n, std::move(a), std::move(b), d);
  }, py::arg("n"), py::arg("a"), py::arg("b"), py::arg("parametric_dist"));
  m.def("optimize_one__LTSS", [ints_to_tuple](int n, fvec a, fvec b, int d) {
    auto r = optimize_one__LTSS(n, std::move(a), std::move(b), d);
    return py::make_tuple(ints_to_tuple(r.first), py::float_(r.second));
  }, py::arg("n"), py::arg("a"), py::arg("b"), py::arg("parametric_dist"));
  // SWIG template names used by callers as constructors (solverSWIG_DP.py:31-32)
  m.def("FArray", []() { return py::list(); });
  m.def("IArray", []() { return py::list(); });
}
