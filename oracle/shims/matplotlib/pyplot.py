"""No-op pyplot surface used by the reference (subplots / savefig / show)."""


class _Axes:
    def __getattr__(self, name):
        return lambda *a, **k: None


def subplots(*a, **k):
    return None, _Axes()


def savefig(*a, **k):
    return None


def show(*a, **k):
    return None
