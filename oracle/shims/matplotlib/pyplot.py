"""Shim for matplotlib.pyplot used ONLY by oracle/ref_oracle.py (test infrastructure).

The reference imports pyplot at module top (utils.py:36, utils_class.py:5), evaluates `plt.Axes`
annotations at def time (utils.py:881,938) and calls `plt.rcParams.update` (utils_class.py:12-15).
Nothing on the numeric hot path plots, so every other attribute raises if actually called.
"""


class Axes:  # annotation target only
    pass


rcParams = {}


def __getattr__(name):
    def _no_plot(*a, **k):
        raise RuntimeError("matplotlib shim: plotting (%s) is out of scope for the oracle" % name)
    return _no_plot
