"""Python wrapper classes with the interface of the reference's src/python/solverSWIG_DP.py:9-97 and
solverSWIG_LTSS.py:8-21 (task objects whose __call__ runs the solve through the `proto` module), so that
scripts written against the reference's wrappers run against the B200-backed module unchanged.

    from partitionsolvers_b200.solverSWIG import OptimizerSWIG, LTSSOptimizerSWIG, Distribution
    subsets, score = OptimizerSWIG(4, g, h, Distribution.POISSON, True, True)()
"""
import os
import sys

_LIBDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib")


def _proto():
    if _LIBDIR not in sys.path:
        sys.path.insert(0, _LIBDIR)
    import proto  # the pybind11 twin of the reference's SWIG module, built by partitionsolvers_b200.build
    return proto


class Distribution:
    GAUSSIAN = 0
    POISSON = 1
    RATIONALSCORE = 2


class OptimizerSWIG(object):
    """Mode switch of the reference wrapper: one solve / sweep_all / sweep_best (OLS) / parallel_sweep."""

    def __init__(self, num_partitions, g, h, objective_fn=Distribution.GAUSSIAN, risk_partitioning_objective=False,
                 use_rational_optimization=False, gamma=0., reg_power=1., parallel_sweep=False, sweep_all=False,
                 sweep_best=False):
        assert not (parallel_sweep and sweep_all)
        assert not (sweep_all and sweep_best)
        self.N = len(g)
        self.num_partitions = num_partitions
        self.objective_fn = objective_fn
        self.risk_partitioning_objective = risk_partitioning_objective
        self.use_rational_optimization = use_rational_optimization
        self.gamma = gamma
        self.reg_power = reg_power
        self.g_c = g
        self.h_c = h
        self.parallel_sweep = parallel_sweep
        self.sweep_all = sweep_all
        self.sweep_best = sweep_best

    def _args(self):
        return (self.N, self.num_partitions, self.g_c, self.h_c, self.objective_fn, self.risk_partitioning_objective,
                self.use_rational_optimization, self.gamma, int(self.reg_power))

    def compute_score(self):
        return _proto().compute_score(self.g_c, self.h_c, self.objective_fn, self.risk_partitioning_objective,
                                      self.use_rational_optimization)

    def compute_partial_score(self, i, j):
        return _proto().compute_score(self.g_c[i:j], self.h_c[i:j], self.objective_fn,
                                      self.risk_partitioning_objective, self.use_rational_optimization)

    def __call__(self):
        p = _proto()
        if self.parallel_sweep:
            return p.sweep_parallel__DP(*self._args())
        if self.sweep_best:
            return p.compute_optimal_num_clusters_OLS_all(*self._args())
        if self.sweep_all:
            return p.optimize_all__DP(*self._args())
        return p.optimize_one__DP(*self._args())


class LTSSOptimizerSWIG(object):
    """solverSWIG_LTSS.OptimizerSWIG: best single subset."""

    def __init__(self, g, h, objective_fn=Distribution.GAUSSIAN):
        self.N = len(g)
        self.g_c = g
        self.h_c = h
        self.objective_fn = objective_fn

    def __call__(self):
        return _proto().optimize_one__LTSS(self.N, self.g_c, self.h_c, self.objective_fn)
