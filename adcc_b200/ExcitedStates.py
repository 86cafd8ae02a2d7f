"""Container of converged excited states (reduced mirror of adcc/ExcitedStates.py and
adcc/ElectronicStates.py: energies, vectors, solver bookkeeping, ``describe``; the transition
properties of SURVEY 8(f1) live in adcc_b200/properties.py)."""
import numpy as np

from .AdcMethod import AdcMethod

from .timings import Timer


class ExcitedStates:
    def __init__(self, data, method=None, property_method=None):
        """``data``: solver state or another ExcitedStates; ``property_method``: level of the ISR
        used for properties ("isr0" / "isr1" / "isr2" or an ADC method name; default: the
        method's own level, ADC(3) -> second order) (adcc/ElectronicStates.py:39-96)."""
        self.matrix = data.matrix
        self.reference_state = self.matrix.reference_state
        self.ground_state = self.matrix.ground_state
        self.method = self.matrix.method if method is None else AdcMethod(method)
        if property_method is None:
            property_method = self.method.property_method
        elif str(property_method).startswith("isr"):
            prefix = "cvs-" if self.method.is_core_valence_separated else ""
            property_method = prefix + "adc" + str(property_method)[3:]
        self.property_method = AdcMethod(property_method).name
        data = getattr(data, "solver_state", data)
        self.timer = Timer()
        if hasattr(data, "timer"):
            self.timer.attach(data.timer, "solver")
        self.converged = data.converged
        self.n_iter = data.n_iter
        self.n_applies = data.n_applies
        self.residual_norm = np.array(data.residual_norms) if data.residual_norms is not None else None
        self._excitation_energy_uncorrected = np.array(data.eigenvalues)
        self.excitation_vector = data.eigenvectors
        self.kind = getattr(data, "kind", "any")
        self.spin_change = None
        self.solver_state = data

    @property
    def size(self):
        return self._excitation_energy_uncorrected.size

    @property
    def excitation_energy_uncorrected(self):
        return self._excitation_energy_uncorrected

    @property
    def excitation_energy(self):
        return self._excitation_energy_uncorrected

    @property
    def total_energy(self):
        level = self.method.level
        if self.method.is_core_valence_separated:
            level = min(level, 2)
        return self.ground_state.energy(level) + self.excitation_energy

    # -- transition properties
    @property
    def transition_dm(self):
        from .properties import transition_dm
        return [transition_dm(self.property_method, self.ground_state, v) for v in self.excitation_vector]

    @property
    def transition_dipole_moment(self):
        from .properties import transition_dipole_moments
        return transition_dipole_moments(self)

    @property
    def oscillator_strength(self):
        tdm = self.transition_dipole_moment
        return 2.0 / 3.0 * np.array([np.linalg.norm(t) ** 2 * e
                                     for t, e in zip(tdm, self.excitation_energy)])

    def describe(self, oscillator_strengths=False):
        eV = 27.211386245988
        kind = {"singlet": "singlet", "triplet": "triplet", "any": "any", "spin_flip": "spin_flip"}.get(self.kind, "")
        conv = "converged" if self.converged else "NOT CONVERGED"
        lines = ["+" + "-" * 52 + "+",
                 "| {0:<30s} {1:>19s} |".format(self.method.name + "  " + kind, conv),
                 "+" + "-" * 52 + "+",
                 "|  #        excitation energy     " + ("osc str " if oscillator_strengths else "        ") + "           |",
                 "|          (au)           (eV)                       |"]
        osc = self.oscillator_strength if oscillator_strengths else None
        for i, e in enumerate(self.excitation_energy):
            extra = f"{osc[i]:8.4f}" if osc is not None else " " * 8
            lines.append("| {0:2d} {1:13.7g} {2:13.7g}   {3}           |".format(i, e, e * eV, extra))
        lines.append("+" + "-" * 52 + "+")
        return "\n".join(lines)
