"""``matplotlib.pyplot`` stand-in: importable, but any attribute access raises."""


def __getattr__(name):
    raise RuntimeError(
        f"matplotlib.pyplot.{name}: matplotlib is not installed; this is the oracle's "
        "import-time stub (plots are outside the hot path)"
    )
