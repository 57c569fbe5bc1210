"""Python access to the oracle. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under locarna_b200/ does.

Two checkers live here:
  * ``port_align``  - the CPU restatement (oracle/port/locarna_port.cc -> oracle/_build/liboracle.so)
  * ``ref_align``   - the reference's own sources compiled by oracle/Makefile (oracle/_ref/ref_harness);
                      present wherever `make -C oracle ref` was run (the build container; the binaries
                      travel to the GPU box with the snapshot).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
REF_HARNESS = os.path.join(HERE, "_ref", "ref_harness")
REF_LOCARNA = os.path.join(HERE, "_ref", "locarna")
NEG_INF = -4611686018427387904


class Params(C.Structure):
    _fields_ = [
        ("min_prob", C.c_double), ("max_diff_am", C.c_int), ("max_diff_at_am", C.c_int), ("max_diff", C.c_int),
        ("min_trace_probability", C.c_double), ("noLP", C.c_int), ("struct_local", C.c_int), ("sequ_local", C.c_int),
        ("free_endgaps", C.c_char * 8), ("struct_weight", C.c_int), ("indel", C.c_int), ("indel_opening", C.c_int),
        ("tau", C.c_int), ("exclusion", C.c_int), ("match", C.c_int), ("mismatch", C.c_int), ("use_ribosum", C.c_int),
        ("temperature_alipf", C.c_int), ("unpaired_penalty", C.c_int), ("pf_double", C.c_int), ("do_trace", C.c_int),
        ("setup_only", C.c_int),
    ]


class Result(C.Structure):
    _fields_ = [
        ("lenA", C.c_int), ("lenB", C.c_int), ("seqA", C.c_char_p), ("seqB", C.c_char_p),
        ("n_arcsA", C.c_int), ("n_arcsB", C.c_int), ("arcsA", C.POINTER(C.c_int)), ("arcsB", C.POINTER(C.c_int)),
        ("weightsA", C.POINTER(C.c_long)), ("weightsB", C.POINTER(C.c_long)),
        ("min_col", C.POINTER(C.c_long)), ("max_col", C.POINTER(C.c_long)),
        ("n_am", C.c_long), ("am", C.POINTER(C.c_int)), ("am_score", C.POINTER(C.c_long)), ("D", C.POINTER(C.c_long)),
        ("score", C.c_long), ("score_is_neg_inf", C.c_int), ("max_i", C.c_int), ("max_j", C.c_int),
        ("cells", C.c_uint64), ("terms", C.c_uint64), ("tasks", C.c_uint64),
        ("n_edges", C.c_int), ("edgesA", C.POINTER(C.c_int)), ("edgesB", C.POINTER(C.c_int)),
        ("strA", C.c_char_p), ("strB", C.c_char_p),
        ("rowA", C.c_char_p), ("rowB", C.c_char_p), ("structA", C.c_char_p), ("structB", C.c_char_p),
    ]


_lib = None


def build_port() -> None:
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "port")], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build_port()
        _lib = C.CDLL(LIB)
        _lib.locarna_port_default_params.argtypes = [C.POINTER(Params)]
        _lib.locarna_port_align.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(Params), C.POINTER(Result), C.c_char_p, C.c_int]
        _lib.locarna_port_align.restype = C.c_int
        _lib.locarna_port_free.argtypes = [C.POINTER(Result)]
        _lib.locarna_port_inside_p.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(Params), C.c_double, C.POINTER(C.c_double),
                                               C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_long), C.c_char_p, C.c_int]
        _lib.locarna_port_inside_p.restype = C.c_int
        dpp = C.POINTER(C.POINTER(C.c_double))
        _lib.locarna_port_probs_p.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(Params), C.c_double, C.c_double, C.POINTER(C.c_double), dpp, dpp, dpp,
                                              C.POINTER(C.c_long), dpp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]
        _lib.locarna_port_probs_p.restype = C.c_int
    return _lib


# flag name (as on the reference CLI, without leading dashes) -> Params field
_FLAG_FIELDS = {
    "min-prob": "min_prob", "max-diff-am": "max_diff_am", "max-diff-at-am": "max_diff_at_am", "max-diff": "max_diff",
    "min-trace-probability": "min_trace_probability", "noLP": "noLP", "struct-local": "struct_local",
    "sequ-local": "sequ_local", "free-endgaps": "free_endgaps", "struct-weight": "struct_weight", "indel": "indel",
    "indel-opening": "indel_opening", "tau": "tau", "exclusion": "exclusion", "match": "match", "mismatch": "mismatch",
    "temperature-alipf": "temperature_alipf", "unpaired-penalty": "unpaired_penalty",
}


def make_params(flags: dict | None = None, do_trace: bool = True, setup_only: bool = False) -> Params:
    p = Params()
    lib().locarna_port_default_params(C.byref(p))
    for k, v in (flags or {}).items():
        if k == "no-ribosum":
            p.use_ribosum = 0 if v else 1
        elif k == "pf-double":
            p.pf_double = 1 if v else 0
        elif k == "free-endgaps":
            p.free_endgaps = v.encode()
        else:
            setattr(p, _FLAG_FIELDS[k], v)
    p.do_trace = 1 if do_trace else 0
    p.setup_only = 1 if setup_only else 0
    return p


def _res_to_dict(r: Result, setup_only: bool, do_trace: bool) -> dict:
    nA, nB, K = r.n_arcsA, r.n_arcsB, r.n_am
    d = {
        "lenA": r.lenA, "lenB": r.lenB, "seqA": r.seqA.decode(), "seqB": r.seqB.decode(),
        "arcsA": [(r.arcsA[2 * k], r.arcsA[2 * k + 1]) for k in range(nA)],
        "arcsB": [(r.arcsB[2 * k], r.arcsB[2 * k + 1]) for k in range(nB)],
        "weightsA": [r.weightsA[k] for k in range(nA)], "weightsB": [r.weightsB[k] for k in range(nB)],
        "min_col": [r.min_col[i] for i in range(r.lenA + 1)], "max_col": [r.max_col[i] for i in range(r.lenA + 1)],
        "am": [tuple(r.am[5 * k + t] for t in range(5)) for k in range(K)],
        "am_score": [r.am_score[k] for k in range(K)],
    }
    if setup_only:
        return d
    d.update({
        "score": None if r.score_is_neg_inf else r.score,
        "D": [None if r.D[k] == NEG_INF else r.D[k] for k in range(K)],
        "max_i": r.max_i, "max_j": r.max_j, "cells": r.cells, "terms": r.terms, "tasks": r.tasks,
    })
    if do_trace:
        d.update({
            "edges": [(r.edgesA[k], r.edgesB[k]) for k in range(r.n_edges)],
            "strA": r.strA.decode(), "strB": r.strB.decode(), "rowA": r.rowA.decode(), "rowB": r.rowB.decode(),
            "structA": r.structA.decode(), "structB": r.structB.decode(),
        })
    return d


def port_align(ppA: str, ppB: str, flags: dict | None = None, do_trace: bool = True, setup_only: bool = False) -> dict:
    """Run the CPU restatement on two PP files; returns a dict of intermediates and results."""
    p = make_params(flags, do_trace, setup_only)
    r = Result()
    err = C.create_string_buffer(512)
    rc = lib().locarna_port_align(ppA.encode(), ppB.encode(), C.byref(p), C.byref(r), err, 512)
    if rc != 0:
        raise RuntimeError("oracle port: " + err.value.decode())
    try:
        return _res_to_dict(r, setup_only, do_trace)
    finally:
        lib().locarna_port_free(C.byref(r))


def port_inside_p(ppA: str, ppB: str, flags: dict | None = None, pf_scale: float = 1.0) -> dict:
    """LocARNA-P inside by the CPU restatement (T = double): partition function Z and the inside value of every arc match
    (reference index order). ``flags`` as for port_align; locarna_p's own defaults differ from locarna's in
    min-trace-probability (1e-5) and the envelope precision (pf-double) - pass them explicitly."""
    p = make_params(flags, False, False)
    Z = C.c_double()
    D = C.POINTER(C.c_double)()
    n = C.c_long()
    err = C.create_string_buffer(512)
    rc = lib().locarna_port_inside_p(ppA.encode(), ppB.encode(), C.byref(p), pf_scale, C.byref(Z), C.byref(D), C.byref(n), err, 512)
    if rc != 0:
        raise RuntimeError("oracle port: " + err.value.decode())
    try:
        return {"Z": Z.value, "D": [D[k] for k in range(n.value)]}
    finally:
        C.CDLL(None).free(D)


def port_probs_p(ppA: str, ppB: str, flags: dict | None = None, pf_scale: float = 1.0, min_am_prob: float = 0.001) -> dict:
    """LocARNA-P complete by the CPU restatement (T = double): Z, inside table D, outside table Dprime and the arc-match
    probabilities (all per arc match, reference index order), and the base-match probability matrix ``bm[i][j]`` (1-based)."""
    p = make_params(flags, False, False)
    Z = C.c_double()
    D, Dp, amp, bmp = (C.POINTER(C.c_double)() for _ in range(4))
    n, la, lb = C.c_long(), C.c_int(), C.c_int()
    err = C.create_string_buffer(512)
    rc = lib().locarna_port_probs_p(ppA.encode(), ppB.encode(), C.byref(p), pf_scale, min_am_prob, C.byref(Z), C.byref(D), C.byref(Dp), C.byref(amp),
                                    C.byref(n), C.byref(bmp), C.byref(la), C.byref(lb), err, 512)
    if rc != 0:
        raise RuntimeError("oracle port: " + err.value.decode())
    try:
        K, W = n.value, lb.value + 1
        return {"Z": Z.value, "D": [D[k] for k in range(K)], "Dprime": [Dp[k] for k in range(K)], "am_prob": [amp[k] for k in range(K)],
                "bm": [[bmp[i * W + j] for j in range(W)] for i in range(la.value + 1)], "lenA": la.value, "lenB": lb.value}
    finally:
        libc = C.CDLL(None)
        for ptr in (D, Dp, amp, bmp):
            libc.free(ptr)


def flags_to_argv(flags: dict | None) -> list:
    argv = []
    for k, v in (flags or {}).items():
        if isinstance(v, bool):
            if v:
                argv.append("--" + k)
        else:
            argv += ["--" + k, str(v)]
    return argv


def have_ref() -> bool:
    return os.path.exists(REF_HARNESS)


def parse_harness(text: str) -> list:
    """Parse ref_harness output into a list of dicts shaped like ``port_align``'s."""
    out, cur = [], None
    lines = text.splitlines()
    i = 0

    def sc(tok):
        return None if tok == "-inf" else int(tok)

    while i < len(lines):
        t = lines[i].split()
        i += 1
        if not t:
            continue
        k = t[0]
        if k == "PAIR":
            cur = {"files": (t[2], t[3])}
        elif k == "LEN":
            cur["lenA"], cur["lenB"] = int(t[1]), int(t[2])
        elif k in ("SEQA", "SEQB"):
            cur["seq" + k[-1]] = t[1] if len(t) > 1 else ""
        elif k == "NARCS":
            cur["n_arcsA"], cur["n_arcsB"] = int(t[1]), int(t[2])
        elif k == "NAM":
            cur["n_am"] = int(t[1])
        elif k in ("ARCSA", "ARCSB"):
            n = int(t[1])
            rows = [lines[i + r].split() for r in range(n)]
            i += n
            cur["arcs" + k[-1]] = [(int(r[0]), int(r[1])) for r in rows]
            cur["probs" + k[-1]] = [float(r[2]) for r in rows]
            cur["weights" + k[-1]] = [int(r[3]) for r in rows]
        elif k == "BANDMIN":
            cur["min_col"] = [int(x) for x in t[1:]]
        elif k == "BANDMAX":
            cur["max_col"] = [int(x) for x in t[1:]]
        elif k == "SIGMA":
            cur["sigma"] = [int(x) for x in t[1:]]
        elif k == "GAPA":
            cur["gapA"] = [int(x) for x in t[1:]]
        elif k == "GAPB":
            cur["gapB"] = [int(x) for x in t[1:]]
        elif k == "AM":
            n = int(t[1])
            rows = [lines[i + r].split() for r in range(n)]
            i += n
            cur["am"] = [(int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[5])) for r in rows]
            cur["am_score"] = [int(r[4]) for r in rows]
            if rows and len(rows[0]) > 6:
                cur["D"] = [sc(r[6]) for r in rows]
            elif not rows:
                cur["D"] = []
        elif k == "SCORE":
            cur["score"] = sc(t[1])
        elif k == "EDGES":
            n = int(t[1])
            cur["edges_full"] = [tuple(int(x) for x in lines[i + r].split()) for r in range(n)]
            i += n
        elif k in ("STRA", "STRB"):
            cur["struct" + k[-1]] = t[1] if len(t) > 1 else ""
        elif k in ("ROWA", "ROWB"):
            cur["row" + k[-1]] = t[1] if len(t) > 1 else ""
        elif k == "PF":
            cur["Z"] = float(t[1])
        elif k == "PFD":
            n = int(t[1])
            cur["pfD"] = [float(lines[i + r]) for r in range(n)]
            i += n
        elif k == "AMPROBS":
            n = int(t[1])
            cur["am_probs"] = [tuple(int(x) for x in lines[i + r].split()[:4]) + (float(lines[i + r].split()[4]),) for r in range(n)]
            i += n
        elif k == "BMPROBS":
            n = int(t[1])
            cur["bm_probs"] = [(int(lines[i + r].split()[0]), int(lines[i + r].split()[1]), float(lines[i + r].split()[2])) for r in range(n)]
            i += n
        elif k in ("TIMEPF_INSIDE", "TIMEPF_OUTSIDE"):
            cur.setdefault("time_pf_ms", {})[k[7:].lower()] = float(t[1])
        elif k == "TIME":
            cur["time_ms"] = {t[j]: float(t[j + 1]) for j in range(1, len(t), 2)}
        elif k == "END":
            out.append(cur)
            cur = None
    return out


def ref_align(ppA: str, ppB: str, flags: dict | None = None, dump: str = "arcs,band,am,D,aln", do_trace: bool = True, timing: bool = False) -> dict:
    """Run the compiled reference (ref_harness) on two PP files."""
    argv = [REF_HARNESS] + flags_to_argv(flags) + ["--dump", dump]
    if not do_trace:
        argv.append("--no-trace")
    if timing:
        argv.append("--time")
    argv += [ppA, ppB]
    r = subprocess.run(argv, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("ref_harness failed: " + r.stderr)
    return parse_harness(r.stdout)[0]


def ref_inside_p(ppA: str, ppB: str, flags: dict | None = None, pf_scale: float = 1.0, probs: bool = False, timing: bool = False) -> dict:
    """LocARNA-P by the compiled reference (AlignerP<double> through ref_harness --pf): Z, inside table, optionally the
    arc-match / base-match probabilities (outside + probability passes)."""
    argv = [REF_HARNESS] + flags_to_argv(flags) + ["--dump", "am,D", "--no-trace", "--pf-probs" if probs else "--pf", "--pf-scale", repr(pf_scale)]
    if timing:
        argv.append("--time")
    argv += [ppA, ppB]
    r = subprocess.run(argv, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("ref_harness failed: " + r.stderr)
    return parse_harness(r.stdout)[0]


def ref_batch(pairs, flags: dict | None = None, dump: str = "", do_trace: bool = True, timing: bool = True, list_path: str | None = None) -> list:
    """Run ref_harness sequentially over a list of (ppA, ppB) pairs in one process."""
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".pairs", delete=False) as f:
        for a, b in pairs:
            f.write("%s %s\n" % (a, b))
        lp = f.name
    argv = [REF_HARNESS] + flags_to_argv(flags) + (["--dump", dump] if dump else [])
    if not do_trace:
        argv.append("--no-trace")
    if timing:
        argv.append("--time")
    argv += ["--pairs", lp]
    r = subprocess.run(argv, capture_output=True, text=True)
    os.unlink(lp)
    if r.returncode != 0:
        raise RuntimeError("ref_harness failed: " + r.stderr)
    return parse_harness(r.stdout)
