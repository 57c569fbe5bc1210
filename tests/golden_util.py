import hashlib
import json
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cases():
    return json.load(open(os.path.join(GOLD, "reference_outputs.json")))["cases"]


def digest(obj) -> str:
    return hashlib.sha256(json.dumps(obj, separators=(",", ":")).encode()).hexdigest()


def full_edges(edges, n, m):
    """alignment_edges(only_local=false) (alignment.cc:120-167): raw trace edges (-1 = gap) -> edges incl. locality gaps (-3)."""
    out, la, lb = [], 1, 1
    for i, j in edges:
        if i > 0:
            while la < i:
                out.append([la, -3]); la += 1
        if j > 0:
            while lb < j:
                out.append([-3, lb]); lb += 1
        if i > 0:
            la += 1
        if j > 0:
            lb += 1
        out.append([i, j])
    while la <= n:
        out.append([la, -3]); la += 1
    while lb <= m:
        out.append([-3, lb]); lb += 1
    return out


# ---- command line tests: the cases of a module are started together, a few processes at a time ------------------------------------
# Every CLI case is its own process (CUDA context set-up dominates its run time); running them one after the other makes the GPU suite
# take many minutes. prefetch() starts the listed commands concurrently and keeps their results, run() hands a result out (or runs the
# command on the spot if it was not prefetched). Output files must therefore have names that do not depend on pytest's tmp_path: out_dir().
import atexit
import shutil
import subprocess
import tempfile
from concurrent.futures import ThreadPoolExecutor

_RESULTS = {}
_OUT = None


def out_dir() -> str:
    """One scratch directory per test session for the files the prefetched commands write."""
    global _OUT
    if _OUT is None:
        _OUT = tempfile.mkdtemp(prefix="lb200_cli_")
        atexit.register(shutil.rmtree, _OUT, True)
    return _OUT


def _key(argv, cwd):
    return (tuple(argv), cwd)


def _run(argv, cwd):
    return subprocess.run(list(argv), capture_output=True, text=True, cwd=cwd)


def prefetch(jobs, workers=None):
    """jobs: iterable of (argv, cwd). Runs those not yet known, `workers` at a time (default: half the host cores, at most 8)."""
    todo = {}
    for argv, cwd in jobs:
        k = _key(argv, cwd)
        if k not in _RESULTS:
            todo[k] = (k, list(argv), cwd)
    todo = list(todo.values())
    if not todo:
        return
    if workers is None:
        workers = max(1, min(8, (os.cpu_count() or 2) // 2))
    with ThreadPoolExecutor(max_workers=workers) as pool:
        for (k, _, _), r in zip(todo, pool.map(lambda t: _run(t[1], t[2]), todo)):
            _RESULTS[k] = r


def run(argv, cwd=None):
    """subprocess.run(argv, capture_output=True, text=True, cwd=cwd), served from the prefetched results when there is one."""
    r = _RESULTS.pop(_key(argv, cwd), None)
    return r if r is not None else _run(argv, cwd)
