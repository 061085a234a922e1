"""Every attribute is a function that does nothing (``plt.figure()``, ``plt.plot(...)``, ``plt.savefig(...)`` ...)."""


def __getattr__(name):
    def _noop(*args, **kwargs):
        return None
    _noop.__name__ = name
    return _noop
