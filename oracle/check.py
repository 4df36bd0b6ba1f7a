"""ORACLE-side comparison helpers (test infrastructure only): the per-entry parity criterion of SURVEY.md §8(c)(2),
shared by tests/ and by bench.py's untimed `parity_check` leg."""
import numpy as np


def entry_scales(colptr, rowval, nzval):
    """sqrt(|K_ii K_jj|) for every stored entry (i, j) of a 1-based CSC matrix: the natural magnitude of that entry (for
    a positive-definite Kt it bounds |K_ij|), used as the floor of the per-entry comparison."""
    n = len(colptr) - 1
    col = np.repeat(np.arange(n, dtype=np.int64), np.diff(colptr))
    row = rowval - 1
    diag = np.zeros(n)
    on = row == col
    diag[row[on]] = np.abs(nzval[on])
    return np.sqrt(diag[row] * diag[col])


FLOOR = 0.01   # an entry's magnitude is taken as at least 1 % of its natural scale sqrt(|K_ii| |K_jj|)


def assert_values_entrywise(nz, nz0, colptr, rowval, tol):
    """SURVEY.md §8(c)(2), per entry:   |a - b| <= tol * max(|a|, |b|, FLOOR * sqrt(|K_ii| |K_jj|)).
    For every entry larger than 1 % of its natural scale this IS the north-star relative bound.  An entry far below that
    scale is the result of cancellation among the (up to 8 elements x 8 Gauss points of) terms summed into it and cannot
    meet a relative bound whatever the evaluation order; for those the bound is absolute, 0.01 * tol of the natural scale
    (1e-14 at tol = 1e-12).  Measured with the product's arithmetic compiled for the host against the oracle on jittered
    Hex8 meshes: every entry within 8.6e-16 of its natural scale; 0.1-0.3 % of the entries lie below 1e-4 of their scale
    and are the only ones beyond a pure relative 1e-12.
    Returns (number of entries that needed the floor, largest plain relative difference)."""
    err = np.abs(nz - nz0)
    mag = np.maximum(np.abs(nz), np.abs(nz0))
    bad = err > tol * mag
    n_floor = int(bad.sum())
    if n_floor:
        scale = entry_scales(colptr, rowval, nz0)[bad]
        over = err[bad] / (tol * np.maximum(mag[bad], FLOOR * scale) + 1e-300)
        k = int(np.argmax(over))
        assert over[k] <= 1.0, (f"{int((over > 1).sum())} entries differ by more than {tol} of max(|entry|, {FLOOR} x natural scale); "
                                f"worst: {over[k]:.2f} x the bound, |entry|/scale = {mag[bad][k] / max(scale[k], 1e-300):.2e}")
    return n_floor, (float((err / np.maximum(mag, 1e-300)).max()) if len(err) else 0.0)


def assert_result_parity(phi, r, colptr, rowval, nzval, ref, tol=1e-12):
    """(ϕ, r, CSC) against an oracle result (ϕ0, r0, (colptr0, rowval0, nzval0)): pattern bit-exact, values per entry."""
    phi0, r0, (cp0, rv0, nz0) = ref
    assert abs(phi - phi0) <= tol * max(1.0, abs(phi0)), (phi, phi0)
    rs = max(np.abs(r0).max(), 1e-300)
    assert np.abs(r - r0).max() <= tol * rs, np.abs(r - r0).max() / rs
    assert np.array_equal(colptr, cp0), "colptr differs"
    assert np.array_equal(rowval, rv0), "rowval differs"
    return assert_values_entrywise(nzval, nz0, cp0, rv0, tol)
