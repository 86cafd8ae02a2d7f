"""Result container of the eigensolvers (adcc/solver/SolverStateBase.py:26-78)."""
from ..timings import Timer


class EigenSolverStateBase:
    def __init__(self, matrix):
        self.matrix = matrix
        self.eigenvalues = None
        self.eigenvectors = None
        self.residual_norms = None
        self.converged = False
        self.n_iter = 0
        self.n_applies = 0
        self.reortho_triggers = []
        self.timer = Timer()

    def describe(self):
        conv = "converged" if self.converged else "NOT CONVERGED"
        algorithm = getattr(self, "algorithm", "")
        text = "+" + 60 * "-" + "+\n"
        text += "| {0:<41s}  {1:>15s} |\n".format(algorithm, conv)
        text += "| {0:30s} n_iter={1:<3d}  n_applies={2:<5d} |\n".format(
            str(self.matrix)[:30], self.n_iter, self.n_applies)
        text += "| n_reortho={0:<7d}  max_overlap_before_reortho={1:<10s}   |\n".format(
            len(self.reortho_triggers),
            "{:<10.4E}".format(max(self.reortho_triggers)) if self.reortho_triggers else "N/A")
        text += "+" + 60 * "-" + "+\n"
        text += "|  #     eigenvalue  res. norm                               |\n"
        for i, _ in enumerate(self.eigenvectors or []):
            text += "| {0:2d} {1:14.7g}  {2:9.4g}                               |\n".format(
                i, self.eigenvalues[i], self.residual_norms[i])
        text += "+" + 60 * "-" + "+"
        return text
