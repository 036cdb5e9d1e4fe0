"""BO loop layer (reference distBO/bayes_opt/)."""
from .helpers import UtilityFunction, acq_max  # noqa: F401

__all__ = ["BayesianOptimization", "UtilityFunction", "acq_max"]


def __getattr__(name):
    if name == "BayesianOptimization":
        from .bayesian_optimization import BayesianOptimization
        return BayesianOptimization
    raise AttributeError(name)
