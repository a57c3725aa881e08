def __getattr__(name):
    def _unavailable(*a, **k):
        raise RuntimeError("matplotlib stub: plotting is not available in the oracle harness")
    return _unavailable
