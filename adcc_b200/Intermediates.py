"""Evaluate-once cache of named ADC intermediates (adcc/Intermediates.py:27-75)."""
from .timings import Timer


class Intermediates:
    generators = {}

    def __init__(self, ground_state):
        self.ground_state = ground_state
        self.reference_state = ground_state.reference_state
        self.timer = Timer()
        self.cached_tensors = {}

    def __getattr__(self, attr):
        if attr.startswith("_"):
            raise AttributeError(attr)
        if attr in self.cached_tensors:
            return self.cached_tensors[attr]
        from .adc_pp import INTERMEDIATES
        gens = dict(INTERMEDIATES)
        gens.update(self.generators)
        if attr not in gens:
            raise AttributeError(f"{self.__class__.__name__} object has no attribute '{attr}'")
        with self.timer.record(attr):
            t = gens[attr](self.reference_state, self.ground_state, self)
            t = t.evaluate()
            t._contig()
            t.set_immutable()
        self.cached_tensors[attr] = t
        return t

    def clear(self):
        self.cached_tensors = {}


def register_as_intermediate(function):
    Intermediates.generators[function.__name__] = function
    return function
