"""Import shim (test infrastructure only): src/utils.py:6-9 imports plotly for
plotting helpers that the latent-dynamics path never calls."""


class _Offline:
    @staticmethod
    def plot(*_a, **_k):
        raise RuntimeError("plotly shim: plotting is not available")


offline = _Offline()
