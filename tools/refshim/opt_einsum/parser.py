"""opt_einsum.parser: the one helper adcc imports."""
import string

_letters = string.ascii_lowercase + string.ascii_uppercase


def gen_unused_symbols(used, n):
    count = 0
    for c in _letters:
        if count >= n:
            return
        if c not in used:
            yield c
            count += 1
