"""Import shim (test infrastructure only): src/env.py:7 imports gym at module
scope; no environment is ever constructed by the oracle harness."""


class Env:  # pragma: no cover
    pass


class Wrapper:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("gym shim")


class ObservationWrapper(Wrapper):  # pragma: no cover
    pass


def make(*_a, **_k):  # pragma: no cover
    raise RuntimeError("gym shim: no environments available")
