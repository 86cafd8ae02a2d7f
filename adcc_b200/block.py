"""Shorthand space strings: b.oovv -> "o1o1v1v1" (o -> o1, v -> v1, c -> o2), the helper the
reference defines in adcc/block.py:30-38."""

_MAP = {"o": "o1", "v": "v1", "c": "o2"}


def __getattr__(attr):
    if attr.startswith("__"):
        raise AttributeError(attr)
    try:
        return "".join(_MAP[c] for c in attr)
    except KeyError:
        raise AttributeError(attr)


# also usable as attributes of the module object for static tools
o, v, c = "o1", "v1", "o2"
