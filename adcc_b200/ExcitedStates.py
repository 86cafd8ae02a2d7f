"""Container of converged excited states (reduced mirror of adcc/ExcitedStates.py and
adcc/ElectronicStates.py: energies, vectors, solver bookkeeping, ``describe``; the transition
properties of SURVEY 8(f1) live in adcc_b200/properties.py)."""
import numpy as np

from .AdcMethod import AdcMethod

from .timings import Timer


def _column(header, values, unit="", width=None):
    """One column of the describe() table: every cell centred in a common width."""
    need = max(len(header), len(unit), *(len(v) for v in values)) if values else max(len(header), len(unit))
    if width is None or width < need:
        width = need + 1
    return {"width": width, "header": f"{header:^{width}s}", "unit": f"{unit:^{width}s}",
            "values": [f"{v:^{width}s}" for v in values]}


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
        outer = data
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
        self.kind = getattr(outer, "kind", getattr(data, "kind", "any"))
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

    # -- state densities (adcc/ElectronicStates.py:205-264)
    @property
    def state_diffdm(self):
        """Difference densities (state minus ground state) as {block: Tensor}, upper blocks of a
        Hermitian block matrix."""
        from .properties import state_diffdm
        return [state_diffdm(self.property_method, self.ground_state, v) for v in self.excitation_vector]

    @property
    def state_dipole_moment(self):
        from .properties import state_dipole_moments
        return state_dipole_moments(self)

    def describe(self, oscillator_strengths=True, rotatory_strengths=False, state_dipole_moments=False,
                 transition_dipole_moments=False, block_norms=True, ssq=False):
        """Human-readable table of the states (adcc/ExcitedStates.py:50-166): one column group per
        requested quantity, each column as wide as its widest entry plus one, header line
        ``method (property method)    kind, converged``.  Rotatory strengths and <S^2> belong to
        subsystems outside this path and are not shown."""
        eV = 27.211386245988
        has_dipole = hasattr(self.reference_state.operator_integral_provider, "electric_dipole")
        columns = [_column("#", [str(i) for i in range(self.size)])]
        columns.append(_column("excitation energy", [f"{e:^13.7g} {e * eV:^13.7g}" for e in self.excitation_energy],
                               unit="(au)         (eV)"))
        if transition_dipole_moments and has_dipole:
            vals = [f"{t[0]:^8.4f} {t[1]:^8.4f} {t[2]:^8.4f}{np.linalg.norm(t):^8.4f}"
                    for t in self.transition_dipole_moment]
            columns.append(_column("transition dipole moment", vals, unit="x(au)    y(au)    z(au)    abs(au)"))
        if oscillator_strengths and has_dipole:
            columns.append(_column("osc str", [f"{f:^8.4f}" for f in self.oscillator_strength], unit="(au)"))
        if block_norms:
            for block, label in (("ph", "|v1|^2"), ("pphh", "|v2|^2")):
                if block in self.matrix.axis_blocks:
                    columns.append(_column(label, [f"{v[block].dot(v[block]):^9.4f}"
                                                   for v in self.excitation_vector]))
        if state_dipole_moments and has_dipole:
            vals = [f"{d[0]:^8.4f} {d[1]:^8.4f} {d[2]:^8.4f}{np.linalg.norm(d):^8.4f}"
                    for d in self.state_dipole_moment]
            columns.append(_column("state dipole moment", vals, unit="x(au)    y(au)    z(au)    abs(au)"))
        width = sum(c["width"] for c in columns)
        # "adc2 (isr2)", "cvs-adc2x (cvs-isr2)": the property method is an IsrMethod in the reference and
        # never compares equal to the AdcMethod, so it is always shown (adcc/ElectronicStates.py:490-492)
        method = f"{self.method.name} ({self.property_method.replace('adc', 'isr')})"
        info = []
        if self.kind:
            info.append(self.kind)
        if self.spin_change is not None and self.spin_change != 0:
            info.append(f"(\u0394MS={self.spin_change:+2d})")
        if info:
            info[-1] += ","
        info.append("converged" if self.converged else "NOT CONVERGED")
        info = " ".join(info)
        header = f"{method}    {info:>{max(width - len(method) - 4, 0)}s}"
        if len(header) > width:                                   # few narrow columns: pad the last one
            columns[-1] = _column(columns[-1]["header"].strip(), [v.strip() for v in columns[-1]["values"]],
                                  unit=columns[-1]["unit"].strip(),
                                  width=columns[-1]["width"] + len(header) - width)
            width = len(header)
        hline = "+" + "-" * (width + 2) + "+"
        lines = [hline, f"| {header} |", hline,
                 "| " + "".join(c["header"] for c in columns) + " |",
                 "| " + "".join(c["unit"] for c in columns) + " |"]
        for row in zip(*[c["values"] for c in columns]):
            lines.append("| " + "".join(row) + " |")
        lines.append(hline)
        return "\n".join(lines)


class State2States:
    """Transition properties from the excited state ``initial`` to all higher-lying states of an
    ExcitedStates object (adcc/State2States.py:33-140): excitation energies, state-to-state
    transition densities, transition dipole moments and oscillator strengths."""

    def __init__(self, data, method=None, property_method=None, initial=0):
        self.parent = data if isinstance(data, ExcitedStates) else ExcitedStates(data, method, property_method)
        if not 0 <= initial < self.parent.size:
            raise ValueError(f"initial = {initial} is not one of the {self.parent.size} states")
        self.initial = int(initial)
        self.method = self.parent.method
        self.property_method = self.parent.property_method
        self.reference_state = self.parent.reference_state
        self.ground_state = self.parent.ground_state
        if self.method.is_core_valence_separated:
            raise NotImplementedError("State2States is not implemented for CVS methods.")

    @property
    def size(self):
        return self.parent.size - self.initial - 1

    @property
    def excitation_energy(self):
        e = self.parent.excitation_energy
        return np.array([e[f] - e[self.initial] for f in range(self.initial + 1, self.parent.size)])

    excitation_energy_uncorrected = excitation_energy

    @property
    def excitation_vector(self):
        return self.parent.excitation_vector[self.initial + 1:]

    @property
    def transition_dm(self):
        from .properties import state2state_transition_dm
        start = self.parent.excitation_vector[self.initial]
        return [state2state_transition_dm(self.property_method, self.ground_state, start, final)
                for final in self.excitation_vector]

    @property
    def transition_dipole_moment(self):
        from .properties import _dipole_of_density
        cache = {}
        return np.array([_dipole_of_density(self.reference_state, dm, cache, hermitian=False)
                         for dm in self.transition_dm]).reshape(self.size, 3)

    @property
    def oscillator_strength(self):
        return 2.0 / 3.0 * np.array([np.linalg.norm(t) ** 2 * abs(e)
                                     for t, e in zip(self.transition_dipole_moment, self.excitation_energy)])
