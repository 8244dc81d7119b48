"""Inert stand-in for matplotlib (test infrastructure): figures accept the calls the reference's logging makes and draw nothing."""
from . import pyplot          # noqa: F401


def use(*a, **k):
    pass
