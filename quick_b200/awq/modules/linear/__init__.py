from .quick import WQLinear_QUICK  # noqa: F401
