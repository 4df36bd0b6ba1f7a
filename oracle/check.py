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


def assert_values_entrywise(nz, nz0, colptr, rowval, tol):
    """Every stored value |a - b| <= tol * max(|a|, |b|).  An entry that is itself the result of cancellation (|K_ij| far
    below the magnitude of the terms summed into it) cannot meet a relative bound whatever the evaluation order -- 64
    terms (8 elements x 8 Gauss points) of magnitude M carry ~64 eps M ~ 7e-15 M of rounding -- so for those the bound
    is the FLOOR 0.1 * tol * sqrt(|K_ii| |K_jj|): a tenth of the tolerance relative to the entry's natural scale, still
    ten times tighter than a matrix-scale comparison; at most 0.1 % of the entries (or 8) may need it.
    Returns (number of entries that needed the floor, largest relative difference)."""
    err = np.abs(nz - nz0)
    mag = np.maximum(np.abs(nz), np.abs(nz0))
    bad = err > tol * mag
    n_floor = int(bad.sum())
    if n_floor:
        floor = 0.1 * tol * entry_scales(colptr, rowval, nz0)
        worst = (err[bad] / np.maximum(floor[bad], 1e-300)).max()
        assert worst <= 1.0, f"{n_floor} entries beyond relative {tol}; worst is {worst:.2f} x the cancellation floor"
        assert n_floor <= max(8, len(err) // 1000), f"{n_floor} of {len(err)} entries needed the cancellation floor"
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
