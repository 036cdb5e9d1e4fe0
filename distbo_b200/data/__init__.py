"""Dataset containers and the toy generators needed to drive the hot path end to end
(reference distBO/data/)."""
from .data import data  # noqa: F401
