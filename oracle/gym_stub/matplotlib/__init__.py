"""Import-time stand-in for matplotlib -- TEST INFRASTRUCTURE ONLY.

The reference's ``evaluation`` package imports ``matplotlib.pyplot`` at module top
(evaluation/seed_variance.py:10, evaluation/failure_statistics.py:10), so importing
``evaluation.evaluator`` drags it in.  Plotting is not on the hot path; every plot call
on this stub raises, so nothing can silently pretend to have plotted.
"""
__version__ = "0.stub"


def use(*args, **kwargs):
    return None
