"""Stand-in for NVIDIA apex: only `amp` is touched by the reference (train_generator.py:161-169,318,356)."""
from . import amp  # noqa: F401
