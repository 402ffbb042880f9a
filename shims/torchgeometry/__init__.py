"""Stand-in for the `torchgeometry` package: only `image.GaussianBlur` is used by the reference (train_generator.py:181)."""
from . import image  # noqa: F401
