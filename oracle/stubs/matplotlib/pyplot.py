axes = None


def __getattr__(name):
    def _missing(*a, **k):
        raise RuntimeError(f"matplotlib stub: pyplot.{name} is not available")
    return _missing
