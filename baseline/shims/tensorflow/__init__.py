"""Substrate stub: ``distkeras/workers.py`` imports tensorflow at module level but only touches it
when ``keras.backend.backend() == 'tensorflow'`` (never true with the torch-backed Keras shim)."""
__version__ = "0.0.0+stub"


class ConfigProto(object):  # pragma: no cover - unreachable with the torch backend
    def __init__(self, *a, **k):
        class _G(object):
            allow_growth = False
        self.gpu_options = _G()


class Session(object):  # pragma: no cover
    def __init__(self, *a, **k):
        pass
