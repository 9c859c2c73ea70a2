"""Synthetic N x K categorical tables for configs C3 / C4.

The generator is COUNTER-BASED: cell (row, col) is a pure function of (seed, row, col) through a
splitmix64 hash, so any row range can be produced independently -- on the host with NumPy or on the
GPU with torch integer ops, bit-identically -- which is what lets every rank of the sharded run
materialise exactly its rows of the same global table.  (SURVEY.md proposes a PCG64 stream; a
stream cannot be sharded without generating the whole table, hence this deviation.)

Column i has domain size d_i = [2,4,8,16,32,64,48,24][i mod 8]; base columns draw from
p(v) ~ (v+1)^-1.2; dependent columns: i mod 8 = 4 -> c_i = perm_i[c_{i+1}] mod 32 (determinant
c_{i+1}, d = 64), i mod 8 = 7 -> c_i = perm_i[c_{i-1}] mod 24 (determinant c_{i-1}, d = 48).
NULLs: every cell of the listed columns independently with probability `null_ratio`.
FD violations (C4): rows whose determinant is one of its 3 least frequent values get their
dependent replaced by (v + 1) mod d with probability `violation_ratio`.
"""
import numpy as np

DOMAIN_CYCLE = [2, 4, 8, 16, 32, 64, 48, 24]
SEED_TABLE, SEED_NULL, SEED_VIOL = 20240922, 20240923, 20240924
_M64 = (1 << 64) - 1
_GOLD = 0x9E3779B97F4A7C15
_C1 = 0xBF58476D1CE4E5B9
_C2 = 0x94D049BB133111EB


def domain_size(i):
    return DOMAIN_CYCLE[i % 8]


def determinant_of(i, n_cols):
    """Column that determines column i, or None for base columns."""
    if i % 8 == 4 and i + 1 < n_cols:
        return i + 1
    if i % 8 == 7:
        return i - 1
    return None


def fd_constraints(n_cols):
    """``X->Y`` statements for the (i mod 8 = 4) dependencies: "c05->c04;c13->c12;..."."""
    return ";".join("c%02d->c%02d" % (i + 1, i) for i in range(n_cols) if i % 8 == 4 and i + 1 < n_cols)


def column_names(n_cols):
    return ["c%02d" % i for i in range(n_cols)]


def _thresholds(d):
    """Integer CDF thresholds on a 53-bit uniform: code = #(thresholds <= u)."""
    w = np.arange(1, d + 1, dtype=np.float64) ** -1.2
    cdf = np.cumsum(w / w.sum())
    thr = np.floor(cdf[:-1] * float(1 << 53)).astype(np.int64)
    return thr


def _perm(seed, i, d):
    return np.random.default_rng([seed, i]).permutation(d).astype(np.int64)


# ---- splitmix64, NumPy (uint64) ------------------------------------------------------------------
def _mix_np(seed, col, rows):
    with np.errstate(over="ignore"):
        z = rows.astype(np.uint64) * np.uint64(_GOLD)
        z = z + np.uint64((seed * 0x2545F4914F6CDD1D + (col + 1) * 0xD6E8FEB86659FD93) & _M64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_C1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_C2)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.int64)  # 53-bit uniform


# ---- splitmix64, torch (int64 two's complement, logical shifts emulated) ------------------------
def _s64(x):
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(z, s):
    return (z >> s) & ((1 << (64 - s)) - 1)


def _mix_torch(seed, col, rows):
    z = rows * _s64(_GOLD)
    z = z + _s64((seed * 0x2545F4914F6CDD1D + (col + 1) * 0xD6E8FEB86659FD93) & _M64)
    z = (z ^ _lsr(z, 30)) * _s64(_C1)
    z = (z ^ _lsr(z, 27)) * _s64(_C2)
    z = z ^ _lsr(z, 31)
    return _lsr(z, 11)


class SynthSpec:
    def __init__(self, n_rows, n_cols, null_ratio=0.01, null_cols=None, violation_ratio=0.0, seed=0):
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.null_ratio = float(null_ratio)
        self.null_cols = list(range(n_cols)) if null_cols is None else list(null_cols)
        self.violation_ratio = float(violation_ratio)
        self.seed = int(seed)
        self.dom = [domain_size(i) for i in range(n_cols)]

    @classmethod
    def c3(cls, n_rows=10_000_000, n_cols=16, seed=0):
        return cls(n_rows, n_cols, 0.01, None, 0.0, seed)

    @classmethod
    def c4(cls, n_rows=100_000_000, n_cols=32, seed=0):
        fd_cols = set()
        for i in range(n_cols):
            if i % 8 == 4 and i + 1 < n_cols:
                fd_cols |= {i, i + 1}
        return cls(n_rows, n_cols, 0.01, [i for i in range(n_cols) if i not in fd_cols], 0.02, seed)


def _generate(spec, lo, hi, xp, mix, searchsorted, take, where, rows):
    cols = [None] * spec.n_cols
    order = [i for i in range(spec.n_cols) if determinant_of(i, spec.n_cols) is None] + \
            [i for i in range(spec.n_cols) if determinant_of(i, spec.n_cols) is not None]
    null_thr = int(spec.null_ratio * float(1 << 53))
    viol_thr = int(spec.violation_ratio * float(1 << 53))
    for i in order:
        d = spec.dom[i]
        det = determinant_of(i, spec.n_cols)
        if det is None:
            u = mix(SEED_TABLE + spec.seed, i, rows)
            c = searchsorted(_thresholds(d), u)
        else:
            dd = spec.dom[det]
            c = take(_perm(SEED_TABLE + spec.seed, i, dd) % d, cols[det])
            if viol_thr > 0 and i % 8 == 4:
                uv = mix(SEED_VIOL + spec.seed, i, rows)
                dirty = (cols[det] >= dd - 3) & (uv < viol_thr)
                c = where(dirty, (c + 1) % d, c)
        cols[i] = c
    out = []
    for i in range(spec.n_cols):
        c = cols[i]
        if i in spec.null_cols and null_thr > 0:
            un = mix(SEED_NULL + spec.seed, i, rows)
            c = where(un < null_thr, -1, c)
        out.append(c)
    return out


def generate_numpy(spec, lo=0, hi=None):
    """-> list of n_cols int32 arrays for global rows [lo, hi)."""
    hi = spec.n_rows if hi is None else hi
    rows = np.arange(lo, hi, dtype=np.int64)
    cols = _generate(
        spec, lo, hi, np, _mix_np,
        lambda thr, u: np.searchsorted(thr, u, side="right").astype(np.int64),
        lambda lut, idx: lut[idx],
        lambda m, a, b: np.where(m, a, b), rows)
    return [c.astype(np.int32) for c in cols]


def generate_torch(spec, device, lo=0, hi=None, out=None, chunk=1 << 24):
    """-> int32 tensor [n_cols, n_pad] on `device` holding global rows [lo, hi) (padding = -1)."""
    import torch
    hi = spec.n_rows if hi is None else hi
    n = hi - lo
    n_pad = (n + 127) // 128 * 128 or 128
    if out is None:
        out = torch.full((spec.n_cols, n_pad), -1, dtype=torch.int32, device=device)

    def searchsorted(thr, u):
        return torch.searchsorted(torch.from_numpy(thr).to(device), u, right=True)

    def take(lut, idx):
        return torch.from_numpy(lut).to(device)[idx]

    def where(m, a, b):
        a = torch.full_like(b, a) if not torch.is_tensor(a) else a
        return torch.where(m, a, b)

    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        rows = torch.arange(lo + s, lo + e, dtype=torch.int64, device=device)
        cols = _generate(spec, lo + s, lo + e, torch, _mix_torch, searchsorted, take, where, rows)
        for i, c in enumerate(cols):
            out[i, s:e] = c.to(torch.int32)
    return out
