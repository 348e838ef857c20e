"""Synthetic per-cell initial states: y0[c][i] = var_i * (1 + amp*u), u ~ U(-1,1)
from a counter-based generator keyed by (seed, cell, component), so any process can regenerate the
same inputs without files.  Cell 0 keeps the example driver's unperturbed state (the reference's
golden files then double as an in-run check).  Zeros stay zero."""
import numpy as np

SEED = 20130430
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return z ^ (z >> np.uint64(31))


def perturbed_states(base: np.ndarray, ncells: int, first_cell: int = 0, amp: float = 0.1, seed: int = SEED,
                     out: np.ndarray | None = None) -> np.ndarray:
    """[ncells, N] array for global cell indices first_cell .. first_cell+ncells-1."""
    base = np.asarray(base, dtype=np.float64)
    n = base.shape[0]
    with np.errstate(over="ignore"):
        cells = (np.arange(first_cell, first_cell + ncells, dtype=np.uint64)[:, None])
        comps = np.arange(n, dtype=np.uint64)[None, :]
        key = _splitmix64(np.uint64(seed) + _splitmix64(cells * np.uint64(1000003) + comps))
    u = (key >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)   # [0,1)
    u = 2.0 * u - 1.0
    y = base[None, :] * (1.0 + amp * u)
    if first_cell == 0 and ncells > 0:
        y[0, :] = base
    if out is not None:
        out[...] = y
        return out
    return y


def swe_members(nmembers: int, first_member: int = 0, amp: float = 0.1, seed: int = SEED, amplitude: float = 30.0) -> np.ndarray:
    """[nmembers, 4800] shallow-water ensemble: the driver's Gaussian bump with amplitude
    a * (1 + amp * u) per member, u from the same counter-based generator; member 0 keeps the driver's a = 30."""
    import fatode_b200 as F
    with np.errstate(over="ignore"):
        mem = np.arange(first_member, first_member + nmembers, dtype=np.uint64)
        key = _splitmix64(np.uint64(seed) + _splitmix64(mem * np.uint64(1000003) + np.uint64(4800)))
    u = 2.0 * ((key >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)) - 1.0
    a = amplitude * (1.0 + amp * u)
    if first_member == 0 and nmembers > 0:
        a[0] = amplitude
    return np.stack([F.swe_initial_state(float(x)) for x in a]) if nmembers > 0 else np.zeros((0, 4800))
