"""TEST INFRASTRUCTURE -- see matplotlib/__init__.py in this directory.  Every attribute is a recorder."""
import sys

CALLS = []


class _Cm:
    def __getattr__(self, name):
        return lambda *a, **k: (0.0, 0.0, 0.0, 1.0)


cm = _Cm()


def _recorder(name):
    def f(*args, **kwargs):
        CALLS.append((name, args, kwargs))
    return f


def __getattr__(name):   # PEP 562: plt.figure, plt.plot, plt.boxplot, plt.savefig, ...
    if name.startswith('__'):
        raise AttributeError(name)
    return _recorder(name)


sys.modules[__name__].__dict__.setdefault('CALLS', CALLS)
