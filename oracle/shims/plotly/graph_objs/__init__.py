class Scatter:  # pragma: no cover - never instantiated on the hot path
    def __init__(self, *a, **k):
        raise RuntimeError("plotly shim")
