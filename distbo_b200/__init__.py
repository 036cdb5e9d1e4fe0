"""distbo_b200 - B200-native (sm_100a) implementation of the distBO joint-GP hot path.

Public surface mirrors the reference (hcllaw/distBO):
  distbo_b200.train.gp_regressor.gp_regressor      (reference distBO/train/gp_regressor.py)
  distbo_b200.bayes_opt.BayesianOptimization       (reference distBO/bayes_opt/)
  distbo_b200.model_embed.bayes_opt_wrap           (reference distBO/model_embed/bayes_optimise.py)
The arithmetic lives in libdistbo_b200.so (C ABI, include/distbo_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
