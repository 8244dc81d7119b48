class _Any:
    def __init__(self, name='obj'):
        self._name = name
        self.calls = []

    def __getattr__(self, k):
        def f(*a, **kw):
            self.calls.append((k, a, kw))
            return _Any(k)
        return f


class Figure(_Any):
    pass


def figure(*a, **k):
    return Figure('figure')


def close(*a, **k):
    pass


def show(*a, **k):
    pass


def subplots(*a, **k):
    return Figure('figure'), _Any('axes')
