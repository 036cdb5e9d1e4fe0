"""BO wrapper and objectives (reference distBO/model_embed/)."""
from .bayes_optimise import bayes_opt_wrap  # noqa: F401
from .models import set_model  # noqa: F401
from .meta_fea import meta  # noqa: F401
