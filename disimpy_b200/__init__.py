"""disimpy_b200: B200-native drop-in for the random-walk hot path of disimpy.

``from disimpy_b200 import gradients, simulations, substrates, utils`` mirrors the
reference package's public modules (docs/source/reference.rst:5-41) for everything on
the path ``simulations.simulation()`` -> per-walker walk -> signal.
"""

from . import gradients, simulations, substrates, utils  # noqa: F401

__version__ = "0.1.0"
