"""Stub: plotting is out of scope for the hot path (see oracle/stubs/matplotlib)."""


def __getattr__(name):
    raise RuntimeError("matplotlib stub: plotting is not available (%s)" % name)
