"""No-op pyplot: the reference imports it at load (iterative_ilqr/utils/base.py:10)."""


class _Obj:
    def __getattr__(self, name):
        return lambda *a, **k: [_Obj()]

    def __getitem__(self, i):
        return _Obj()


def subplots(*a, **k):
    return _Obj(), _Obj()


def plot(*a, **k):
    return [_Obj()]


def legend(*a, **k):
    pass


def show(*a, **k):
    pass
