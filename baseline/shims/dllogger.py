"""No-op stand-in for NVIDIA dllogger (reference arm only)."""


class Verbosity:
    OFF, DEFAULT, VERBOSE = -1, 0, 1


class JSONStreamBackend:
    def __init__(self, *a, **k): pass


class StdOutBackend:
    def __init__(self, *a, **k): pass


def init(backends=None): pass
def log(step=None, data=None, verbosity=0): pass
def flush(): pass
