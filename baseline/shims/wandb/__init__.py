class Table:
    def __init__(self, *a, **k):
        pass


class Histogram:
    def __init__(self, *a, **k):
        pass


def init(*a, **k):
    return None


def log(*a, **k):
    pass
