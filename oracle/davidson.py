"""Oracle Jacobi-Davidson (test infrastructure, see oracle/__init__.py).

Restates the control flow of adcc/solver/davidson.py:78-350 (davidson_iterations) and
:353-468 (eigsh argument handling / defaults), adcc/solver/preconditioner.py:40-87
(Jacobi preconditioner r/(D - sigma)), adcc/solver/explicit_symmetrisation.py:38-98 and
libadcc_src/amplitude_vector_enforce_spin_kind.cc:33-170 (singlet purification), and
adcc/solver/common.py:26-46 (select_eigenpairs) on dict-of-ndarray vectors.

One documented deviation: the Rayleigh-Ritz step uses a dense ``numpy.linalg.eigh`` and
keeps the ``which`` n_block pairs instead of ARPACK (davidson.py:185-193); the selected
eigenpairs are mathematically identical and the oracle becomes deterministic.
"""
import warnings
import numpy as np


def vdot(a, b):
    return sum(float(np.vdot(a[k], b[k])) for k in a)


def lincomb(coeffs, vecs):
    out = {k: np.zeros_like(v) for k, v in vecs[0].items()}
    for c, v in zip(coeffs, vecs):
        for k in out:
            out[k] += c * v[k]
    return out


def select_eigenpairs(evals, n_ep, which):
    if which == "SA":
        return np.arange(len(evals))[:n_ep]
    if which == "LA":
        return np.arange(len(evals))[-n_ep:]
    if which == "LM":
        return np.sort(np.argsort(np.abs(evals))[-n_ep:])
    if which == "SM":
        return np.sort(np.argsort(np.abs(evals))[:n_ep])
    raise ValueError(which)


def enforce_spin_kind(u2, n_alpha, kind):
    """amplitude_vector_enforce_spin_kind.cc:33-170 on a dense (i,j,a,b) array, in place.

    singlet: u[ia ja aa ba] <- u[ia jb aa bb] + u[ia jb ab ba]; bbbb follows the +1 map.
    triplet: no-op (:58-65)."""
    if kind == "triplet":
        return u2
    if kind != "singlet":
        raise NotImplementedError(kind)
    ni, nj, na, nb = n_alpha
    A = (slice(0, ni), slice(0, nj), slice(0, na), slice(0, nb))
    abab = u2[0:ni, nj:, 0:na, nb:]
    abba = u2[0:ni, nj:, na:, 0:nb]
    # element (i,j,a,b) of the aaaa block from spatially identical indices
    new = abab + abba
    u2[A] = new
    u2[ni:, nj:, na:, nb:] = new
    return u2


def _flip(t, n_alpha):
    """alpha<->beta relabelling of every axis (restricted reference: equal halves)."""
    for ax, na in enumerate(n_alpha):
        t = np.concatenate((np.take(t, range(na, t.shape[ax]), axis=ax),
                            np.take(t, range(0, na), axis=ax)), axis=ax)
    return t


def enforce_spin_maps(v, n_alpha, fac):
    """Spin-block symmetry of restricted singlet/triplet trial vectors
    (adcc/guess/guess_zero.py:121-161: aa<->bb, aaaa<->bbbb, abab<->baba, abba<->baab with
    factor fac; spin-changing blocks forbidden).  libtensor keeps this structurally (mapped
    blocks are not stored, as_lt_symmetry.cc:100-172); on dense arrays it is the orthogonal
    projector T <- (T + fac * flip(T)) / 2 followed by zeroing the forbidden blocks."""
    ni, nj, na, nb = n_alpha
    if "ph" in v:
        t = v["ph"]
        t = 0.5 * (t + fac * _flip(t, (nj, na)))
        t[:nj, na:] = 0.0
        t[nj:, :na] = 0.0
        v["ph"] = t
    if "pphh" in v:
        t = v["pphh"]
        t = 0.5 * (t + fac * _flip(t, n_alpha))
        s = [np.array([0] * n + [1] * (t.shape[k] - n)) for k, n in enumerate(n_alpha)]
        # spin change = (#beta holes) - (#beta particles) must vanish
        keep = (s[0][:, None, None, None] + s[1][None, :, None, None]
                == s[2][None, None, :, None] + s[3][None, None, None, :])
        v["pphh"] = t * keep
    return v


class Symmetrisation:
    """IndexSymmetrisation / IndexSpinSymmetrisation (explicit_symmetrisation.py:38-98)."""

    def __init__(self, matrix, kind="any"):
        self.matrix = matrix
        self.kind = kind
        hf = matrix.hf
        occ = "o2" if matrix.cvs else "o1"
        self.n_alpha = [hf.n_alpha_of[s] for s in ("o1", occ, "v1", "v1")]

    def symmetrise(self, vecs):
        for v in vecs:
            if self.kind in ("singlet", "triplet"):
                enforce_spin_maps(v, self.n_alpha, 1.0 if self.kind == "singlet" else -1.0)
            if "pphh" in v:
                v["pphh"] = self.matrix.symmetrise_doubles(v["pphh"])
                if self.kind in ("singlet", "triplet"):
                    enforce_spin_kind(v["pphh"], self.n_alpha, self.kind)
        return vecs


class DavidsonState:
    def __init__(self, matrix, guesses):
        self.matrix = matrix
        self.eigenvalues = None
        self.eigenvectors = None
        self.residual_norms = None
        self.converged = False
        self.n_iter = 0
        self.n_applies = 0
        self.reortho_triggers = []
        self.subspace_vectors = list(guesses)
        self.residuals = None


def davidson_eigsh(matrix, guesses, n_ep=None, n_block=None, max_subspace=None, conv_tol=1e-9,
                   which="SA", max_iter=70, preconditioner=True, explicit_symmetrisation=None,
                   residual_min_norm=None, callback=None):
    if n_ep is None:
        n_ep = len(guesses)
    elif n_ep > len(guesses):
        raise ValueError("n_ep cannot exceed the number of guess vectors")
    if n_block is None:
        n_block = n_ep
    elif n_block < n_ep or n_block > len(guesses):
        raise ValueError("invalid n_block")
    if not max_subspace:
        max_subspace = max(6 * n_ep, 20, 5 * len(guesses))
    elif max_subspace < 2 * n_block or max_subspace < len(guesses):
        raise ValueError("invalid max_subspace")

    state = DavidsonState(matrix, guesses)
    n_problem = matrix.shape[1]
    eps = np.finfo(float).eps
    if residual_min_norm is None:
        residual_min_norm = 2 * n_problem * eps
    diagonal = matrix.diagonal()

    SS = state.subspace_vectors
    n_ss = len(SS)
    Ass_cont = np.empty((max_subspace, max_subspace))
    Ax = [matrix.matvec(v) for v in SS]
    state.n_applies += n_ss
    Ass = Ass_cont[:n_ss, :n_ss]
    for i in range(n_ss):
        for j in range(i, n_ss):
            Ass[i, j] = Ass[j, i] = vdot(SS[i], Ax[j])

    while state.n_iter < max_iter:
        state.n_iter += 1
        evals, evecs = np.linalg.eigh(Ass)
        if Ass.shape == (n_block, n_block):
            rvals, rvecs = evals, evecs
        else:
            pick = {"SA": np.arange(n_block), "LA": np.arange(len(evals) - n_block, len(evals))}
            sel = pick[which] if which in pick else select_eigenpairs(evals, n_block, which)
            rvals, rvecs = evals[sel], evecs[:, sel]

        residuals = [lincomb(np.hstack((rvecs[:, k], -rvals[k] * rvecs[:, k])), Ax + SS)
                     for k in range(n_block)]
        mask = select_eigenpairs(rvals, n_ep, which)
        state.eigenvalues = rvals[mask]
        state.residuals = [residuals[i] for i in mask]
        state.residual_norms = np.array([np.sqrt(vdot(r, r)) for r in state.residuals])
        if callback:
            callback(state, "next_iter")
        state.converged = bool(np.all(state.residual_norms < conv_tol))
        if state.converged or state.n_iter == max_iter:
            if not state.converged:
                warnings.warn("Maximum number of iterations reached in davidson procedure.")
            state.eigenvectors = [lincomb(rvecs[:, i], SS) for i in range(n_block) if i in mask]
            return state

        if n_ss + n_block > max_subspace:
            SS = [lincomb(rvecs[:, k], SS) for k in range(n_block)]
            state.subspace_vectors = SS
            Ax = [lincomb(rvecs[:, k], Ax) for k in range(n_block)]
            n_ss = len(SS)
            Ass = Ass_cont[:n_ss, :n_ss]
            for i in range(n_ss):
                for j in range(i, n_ss):
                    Ass[i, j] = Ass[j, i] = vdot(SS[i], Ax[j])

        if preconditioner:
            shifts = rvals - 1e-6
            preconds = [{k: r[k] / (diagonal[k] - shifts[i]) for k in r}
                        for i, r in enumerate(residuals)]
        else:
            preconds = residuals
        if explicit_symmetrisation is not None:
            explicit_symmetrisation.symmetrise(preconds)

        n_added = 0
        for i in range(n_block):
            pvec = preconds[i]
            coeff = np.hstack(([1.0], -np.array([vdot(pvec, s) for s in SS])))
            pvec = lincomb(coeff, [pvec] + SS)
            pnorm = np.sqrt(vdot(pvec, pvec))
            if pnorm < residual_min_norm:
                continue
            overlap = np.array([vdot(pvec, s) for s in SS])
            loss = np.max(np.abs(overlap)) / pnorm
            if loss > n_problem * eps:
                pvec = lincomb(np.hstack(([1.0], -overlap)), [pvec] + SS)
                pnorm = np.sqrt(vdot(pvec, pvec))
                state.reortho_triggers.append(loss)
            if pnorm >= residual_min_norm:
                SS.append({k: t / pnorm for k, t in pvec.items()})
                n_added += 1
                n_ss = len(SS)
        if n_added == 0:
            state.converged = False
            state.eigenvectors = [lincomb(rvecs[:, i], SS) for i in range(n_block) if i in mask]
            warnings.warn("Davidson procedure could not generate any further vectors.")
            return state

        Ax.extend(matrix.matvec(v) for v in SS[-n_added:])
        state.n_applies += n_added
        Ass = Ass_cont[:n_ss, :n_ss]
        for i in range(n_ss - n_added, n_ss):
            for j in range(i + 1):
                Ass[i, j] = Ass[j, i] = vdot(SS[i], Ax[j])
    return state


def jacobi_davidson(*args, **kwargs):
    return davidson_eigsh(*args, preconditioner=True, **kwargs)
