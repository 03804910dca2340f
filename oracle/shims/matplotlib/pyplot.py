"""No-op pyplot: the reference only draws one PNG per population export (search_organisms.py:720-761)."""


def plot(*args, **kwargs):
    return []


def legend(*args, **kwargs):
    return None


def savefig(path, *args, **kwargs):
    open(path, "wb").close()


def close(*args, **kwargs):
    return None
