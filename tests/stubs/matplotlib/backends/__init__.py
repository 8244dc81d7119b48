from . import backend_agg      # noqa: F401
