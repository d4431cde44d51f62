"""Minimal stand-in so that the reference package imports on boxes without
matplotlib (test infrastructure only; the plotting helpers are out of scope)."""
