"""Deterministic synthetic inputs for tests and benchmarks.

ViennaRNA is not available in this image, so base-pair probabilities cannot come from
``RNAfold -p``.  This module writes LocARNA PP 2.0 files (format: reference
``src/LocARNA/rna_data.cc:984-1103``, example ``Data/Examples/example.pp``) whose dot plots are
"helix structured": every maximal run of >= 3 stacked canonical/GU pairs of the sequence is a
candidate stem, stems get Boltzmann-like weights from a toy stacking energy, and pair
probabilities are normalised per position so that the paired mass of a position stays < 1.
The inverse temperature is bisected per sequence so that about ``density * n`` pairs have
p >= 0.001 (the survey's probe saw ~2.2 n for RNAfold dot plots of random sequences).

The same files feed the oracle and the GPU path, so this generator only shapes the workload;
it is not part of the parity argument.
"""
from __future__ import annotations

import os
import numpy as np

NT = "ACGU"
# canonical + wobble pairs, indexed by nucleotide codes A=0 C=1 G=2 U=3
_PAIR_E = np.zeros((4, 4))
_PAIR_E[0, 3] = _PAIR_E[3, 0] = 2.0
_PAIR_E[1, 2] = _PAIR_E[2, 1] = 3.0
_PAIR_E[2, 3] = _PAIR_E[3, 2] = 1.0


def random_sequence(n: int, seed: int) -> str:
    rng = np.random.RandomState(seed)
    return "".join(NT[c] for c in rng.randint(0, 4, size=n))


def mutate(seq: str, identity: float, seed: int, indel_rate: float = 0.03) -> str:
    """A BRAliBase-like relative of ``seq``: substitutions to reach about ``identity`` plus a few indels."""
    rng = np.random.RandomState(seed)
    out = []
    for ch in seq:
        r = rng.rand()
        if r < indel_rate / 2:
            continue  # deletion
        if r < indel_rate:
            out.append(NT[rng.randint(0, 4)])  # insertion before
        if rng.rand() > identity:
            choices = [c for c in NT if c != ch]
            out.append(choices[rng.randint(0, 3)])
        else:
            out.append(ch)
    return "".join(out)


def _stems(codes: np.ndarray, min_len: int = 3, min_loop: int = 3):
    n = len(codes)
    can = _PAIR_E[codes[:, None], codes[None, :]] > 0
    run = np.zeros((n + 2, n + 2), dtype=np.int32)  # run[i+1, j+1] = stacked run length going inward from (i, j)
    for d in range(min_loop + 1, n):
        i = np.arange(0, n - d)
        j = i + d
        inner = run[i + 2, j] if d - 2 >= min_loop + 1 else 0
        run[i + 1, j + 1] = np.where(can[i, j], 1 + inner, 0)
    stems = []
    ii, jj = np.nonzero(run[1:-1, 1:-1] >= min_len)
    for i, j in zip(ii.tolist(), jj.tolist()):
        if i > 0 and j < n - 1 and can[i - 1, j + 1]:
            continue  # not the outermost pair of a maximal run
        stems.append((i, j, int(run[i + 1, j + 1])))
    return stems


def dotplot(seq: str, seed: int = 0, density: float = 2.2, cutoff: float = 0.0005):
    """Return a sorted list of (i, j, p), 1-based, p > cutoff."""
    codes = np.array([NT.index(c) if c in NT else 0 for c in seq.upper().replace("T", "U")], dtype=np.int64)
    n = len(codes)
    if n < 8:
        return []
    stems = _stems(codes)
    if not stems:
        return []
    rng = np.random.RandomState(seed * 7919 + n)
    energies = []
    for (i, j, L) in stems:
        e = sum(_PAIR_E[codes[i + k], codes[j - k]] for k in range(L))
        e += -4.0 - 0.004 * (j - i) + 0.75 * rng.randn()
        energies.append(e)
    energies = np.array(energies)
    energies -= energies.max()

    def probs(beta):
        w = np.exp(beta * energies)
        zpos = np.zeros(n)
        for (i, j, L), ws in zip(stems, w):
            zpos[i:i + L] += ws
            zpos[j - L + 1:j + 1] += ws
        out = []
        for (i, j, L), ws in zip(stems, w):
            for k in range(L):
                a, b = i + k, j - k
                # fraying: outer and inner ends of a stem are a little less probable
                fr = 1.0 if 0 < k < L - 1 else 0.8
                out.append((a + 1, b + 1, fr * ws / (0.05 + max(zpos[a], zpos[b]))))
        return out

    target = density * n
    lo, hi = 0.01, 8.0
    for _ in range(18):
        mid = 0.5 * (lo + hi)
        cnt = sum(1 for (_, _, p) in probs(mid) if p >= 0.001)
        if cnt > target:
            lo = mid  # sharper distribution -> fewer pairs above threshold
        else:
            hi = mid
    res = [(a, b, float("%.6g" % p)) for (a, b, p) in probs(0.5 * (lo + hi))]
    res = [(a, b, min(p, 0.999999)) for (a, b, p) in res if p > cutoff]
    res.sort()
    return res


def write_pp(path: str, name: str, seq: str, pairs, cutoff: float = 0.0005, stacking: bool = False) -> None:
    """PP 2.0 file. With ``stacking`` the base pair lines carry the joint probability of (i, j) and (i+1, j-1) as a fourth column
    (rna_data.cc:1058-1095, keyword #STACK): here 0.85 x the smaller of the two pair probabilities where the inner pair is listed."""
    prob = {(i, j): p for (i, j, p) in pairs}
    with open(path, "w") as f:
        f.write("#PP 2.0\n\n")
        f.write("%s %s\n" % (name, seq))
        f.write("\n#END\n\n#SECTION BASEPAIRS\n\n#BPCUT %g\n" % cutoff)
        if stacking:
            f.write("#STACK\n")
        f.write("\n")
        for (i, j, p) in pairs:
            q = prob.get((i + 1, j - 1))
            if stacking and q is not None and 0.85 * min(p, q) > cutoff:
                f.write("%d %d %.6g %.6g\n" % (i, j, p, 0.85 * min(p, q)))
            else:
                f.write("%d %d %.6g\n" % (i, j, p))
        f.write("\n#END\n")


def make_pp(path: str, name: str, seq: str, seed: int = 0, density: float = 2.2, stacking: bool = False) -> None:
    write_pp(path, name, seq, dotplot(seq, seed=seed, density=density), stacking=stacking)


def _make_one(job):
    path, name, seq, seed = job
    if not os.path.exists(path):
        tmp = "%s.tmp%d" % (path, os.getpid())
        make_pp(tmp, name, seq, seed=seed)
        os.replace(tmp, path)
    return path


def make_family(outdir: str, config: int, count: int, length, related: bool = False, workers: int = 1):
    """Write ``count`` PP files named s<k>.pp; returns the list of paths.

    seed = 1000*config + index.  ``length`` is an int or a callable(rng)->int.
    With ``related`` the odd members are 70 %-identity relatives of their even predecessor.
    ``workers`` > 1 writes the files with a process pool (every file depends only on its own seed); existing files are kept.
    """
    os.makedirs(outdir, exist_ok=True)
    jobs = []
    prev = None
    for k in range(count):
        seed = 1000 * config + k
        n = length if isinstance(length, int) else int(length(np.random.RandomState(seed)))
        if related and k % 2 == 1 and prev is not None:
            seq = mutate(prev, 0.7, seed)
        else:
            seq = random_sequence(n, seed)
        prev = seq
        jobs.append((os.path.join(outdir, "s%d.pp" % k), "s%d" % k, seq, seed))
    if workers > 1 and count > 1:
        import concurrent.futures as cf
        with cf.ProcessPoolExecutor(max_workers=workers) as ex:
            return list(ex.map(_make_one, jobs, chunksize=16))
    return [_make_one(j) for j in jobs]
