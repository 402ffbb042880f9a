"""Import-only stand-in (eval_models / LPIPS is out of scope)."""
from . import color, metrics, transform  # noqa: F401
