"""opt_einsum.backends.dispatch: the registries adcc/opt_einsum_integration.py:116-132 fills in."""
EVAL_CONSTS_BACKENDS = {}
_cached_funcs = {}
_has_einsum = {}
