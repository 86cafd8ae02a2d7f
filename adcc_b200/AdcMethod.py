"""ADC method names: "adc0".."adc3", "adc2x" and their "cvs-" variants.
(Reduced mirror of adcc/AdcMethod.py for the PP-ADC methods on the hot path.)"""


class AdcMethod:
    available_methods = ["adc0", "adc1", "adc2", "adc2x", "adc3",
                         "cvs-adc0", "cvs-adc1", "cvs-adc2", "cvs-adc2x", "cvs-adc3"]

    def __init__(self, method):
        if isinstance(method, AdcMethod):
            method = method.name
        if method not in self.available_methods:
            raise ValueError("Invalid method " + str(method) + ". Only available are: "
                             + ", ".join(self.available_methods))
        split = method.split("-")
        self.__base_method = split[-1]
        split = split[:-1]
        self.is_core_valence_separated = "cvs" in split
        self.level = int(self.__base_method[3])
        self.adc_type = "pp"

    def at_level(self, newlevel):
        prefix = "cvs-" if self.is_core_valence_separated else ""
        return AdcMethod(prefix + "adc" + str(newlevel))

    @property
    def name(self):
        return ("cvs-" if self.is_core_valence_separated else "") + self.__base_method

    @property
    def property_method(self):
        """Level at which ISR properties are evaluated (adc3 -> adc2)."""
        if self.level < 3:
            return self.name.replace("2x", "2")
        return self.at_level(2).name

    @property
    def base_method(self):
        return AdcMethod(self.__base_method)

    def __eq__(self, other):
        return isinstance(other, AdcMethod) and self.name == other.name

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return f"Method(name={self.name})"
