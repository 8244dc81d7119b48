"""Inert pyplot: figures accept the calls the reference's logging makes (add_subplot / imshow / colorbar / tick labels ...) and
render as a blank 500 x 300 canvas (tests/stubs/matplotlib/backends/backend_agg.py)."""


class _Any:
    """accepts any attribute access / call chain"""
    def __init__(self, name='obj'):
        self.__dict__['_name'] = name

    def __getattr__(self, k):
        return _Any(k)

    def __call__(self, *a, **kw):
        return _Any(self._name + '()')


class _Canvas:
    def get_width_height(self):
        return 500, 300


class Figure(_Any):
    def __init__(self):
        super().__init__('figure')
        self.__dict__['canvas'] = _Canvas()


def figure(*a, **k):
    return Figure()


def close(*a, **k):
    pass


def show(*a, **k):
    pass


def subplots(*a, **k):
    return Figure(), _Any('axes')
