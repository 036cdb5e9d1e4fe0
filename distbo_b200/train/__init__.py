"""Surrogate layer (reference distBO/train/): gp_regressor facade over the sm_100a engine."""
from .gp_regressor import gp_regressor  # noqa: F401
