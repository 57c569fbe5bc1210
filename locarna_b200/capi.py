"""ctypes binding of the C ABI in include/locarna_b200.h (the product's only entry point).

This module deliberately contains no algorithmic code: everything is done by
``liblocarna_b200.so`` (host C++ + sm_100a CUDA kernels). If the library is missing or no CUDA
device is present the calls fail loudly; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LB200_LIB", os.path.join(_HERE, "liblocarna_b200.so"))

OK = 0
DEVICE_NONE = -1
RUN_SCORE_ONLY, RUN_TRACE, RUN_KEEP_D = 0, 1, 2
SCORE_NEG_INF = -(2 ** 63)


class Params(C.Structure):
    _fields_ = [
        ("min_prob", C.c_double), ("max_diff_am", C.c_int), ("max_diff_at_am", C.c_int), ("max_diff", C.c_int),
        ("min_trace_probability", C.c_double),
        ("struct_weight", C.c_int), ("indel", C.c_int), ("indel_opening", C.c_int), ("tau", C.c_int), ("exclusion", C.c_int),
        ("match", C.c_int), ("mismatch", C.c_int), ("use_ribosum", C.c_int), ("unpaired_penalty", C.c_int),
        ("temperature_alipf", C.c_int),
        ("no_lonely_pairs", C.c_int), ("struct_local", C.c_int), ("sequ_local", C.c_int),
        ("free_endgaps", C.c_char * 8), ("pf_double", C.c_int), ("exp_prob", C.c_double), ("max_bps_length_ratio", C.c_double), ("max_bp_span", C.c_int),
        ("stacking", C.c_int), ("new_stacking", C.c_int),
    ]


class PairInfo(C.Structure):
    _fields_ = [("lenA", C.c_int), ("lenB", C.c_int), ("n_arcsA", C.c_int), ("n_arcsB", C.c_int),
                ("n_arcmatches", C.c_int64), ("n_tasks", C.c_int64), ("cells", C.c_int64), ("terms", C.c_int64), ("n_edges", C.c_int64)]


EXPORTS = [
    "lb200_default_params", "lb200_ctx_create", "lb200_ctx_destroy", "lb200_last_error", "lb200_set_params",
    "lb200_seq_add_pp", "lb200_seq_add", "lb200_seq_length", "lb200_seq_get", "lb200_pair_add", "lb200_num_pairs", "lb200_clear_pairs",
    "lb200_prepare", "lb200_upload", "lb200_run", "lb200_last_kernel_ms", "lb200_last_h2d_bytes", "lb200_last_d2h_bytes", "lb200_last_dfill_ms", "lb200_last_dfill_launches", "lb200_last_launches", "lb200_envelope_stats", "lb200_pair_score", "lb200_get_scores",
    "lb200_pair_get_info", "lb200_pair_band", "lb200_pair_arcmatches", "lb200_pair_alignment", "lb200_upgma_newick",
    "lb200_run_pf", "lb200_pair_partition_function", "lb200_pair_arcmatch_pf", "lb200_run_pf_probs", "lb200_pair_arcmatch_probs",
    "lb200_pair_basematch_probs", "lb200_pairs_add", "lb200_all_vs_all", "lb200_pair_cost", "lb200_shard_pairs", "lb200_seq_num_arcs", "lb200_seqs_add_pp", "lb200_last_dfill_kind", "lb200_rows_fallbacks", "lb200_release_device_cache", "lb200_run_normalized", "lb200_run_penalized", "lb200_shard_job", "lb200_pair_set_restriction", "lb200_run_pair_toplevel", "lb200_pair_add_restricted", "lb200_band_from_alignment", "lb200_seq_anchors", "lb200_seqs_copy", "lb200_set_ribosum_file", "lb200_seq_pairs",
    "lb200_seq_num_rows", "lb200_seq_get_row",
]

_lib = None


def load():
    """Load liblocarna_b200.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("liblocarna_b200.so is missing (%s); build it with `make` - there is no fallback path" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, ip, dp, i64p = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int64)
    lib.lb200_default_params.argtypes = [C.POINTER(Params)]
    lib.lb200_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.lb200_ctx_destroy.argtypes = [vp]
    lib.lb200_last_error.argtypes = [vp]
    lib.lb200_last_error.restype = C.c_char_p
    lib.lb200_set_params.argtypes = [vp, C.POINTER(Params)]
    lib.lb200_seq_add_pp.argtypes = [vp, C.c_char_p]
    lib.lb200_seqs_add_pp.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p)]
    lib.lb200_seq_add.argtypes = [vp, C.c_char_p, C.c_char_p, ip, ip, dp, C.c_int]
    lib.lb200_seq_length.argtypes = [vp, C.c_int]
    lib.lb200_seq_get.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    lib.lb200_seq_num_arcs.argtypes = [vp, C.c_int]
    lib.lb200_seq_num_rows.argtypes = [vp, C.c_int]
    lib.lb200_seq_get_row.argtypes = [vp, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    lib.lb200_pairs_add.argtypes = [vp, C.c_int, ip, ip]
    lib.lb200_all_vs_all.argtypes = [C.c_int, ip, ip]
    lib.lb200_all_vs_all.restype = C.c_int64
    lib.lb200_pair_cost.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    lib.lb200_pair_cost.restype = C.c_double
    lib.lb200_shard_pairs.argtypes = [C.c_int64, dp, C.c_int, ip, i64p, i64p]
    lib.lb200_pair_add.argtypes = [vp, C.c_int, C.c_int, ip, ip]
    lib.lb200_num_pairs.argtypes = [vp]
    lib.lb200_clear_pairs.argtypes = [vp]
    lib.lb200_prepare.argtypes = [vp]
    lib.lb200_upload.argtypes = [vp]
    lib.lb200_last_dfill_ms.argtypes = [vp]
    lib.lb200_last_dfill_ms.restype = C.c_double
    lib.lb200_last_dfill_launches.argtypes = [vp]
    lib.lb200_last_dfill_launches.restype = C.c_int64
    lib.lb200_last_h2d_bytes.argtypes = [vp]
    lib.lb200_last_h2d_bytes.restype = C.c_int64
    lib.lb200_last_d2h_bytes.argtypes = [vp]
    lib.lb200_last_d2h_bytes.restype = C.c_int64
    lib.lb200_run.argtypes = [vp, C.c_int]
    lib.lb200_last_kernel_ms.argtypes = [vp]
    lib.lb200_last_kernel_ms.restype = C.c_double
    lib.lb200_shard_job.argtypes = [vp, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lb200_seq_pairs.argtypes = [vp, C.c_int, ip, ip, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), ip]
    lib.lb200_seq_pairs.restype = C.c_int64
    lib.lb200_seqs_copy.argtypes = [vp, vp]
    lib.lb200_set_ribosum_file.argtypes = [vp, C.c_char_p]
    lib.lb200_seq_anchors.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
    lib.lb200_pair_add_restricted.argtypes = [vp, C.c_int, C.c_int, ip, ip]
    lib.lb200_band_from_alignment.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, ip, ip]
    lib.lb200_pair_set_restriction.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.lb200_run_pair_toplevel.argtypes = [vp, C.c_int, C.c_int, C.c_int64, C.c_int]
    lib.lb200_run_normalized.argtypes = [vp, C.c_int64]
    lib.lb200_run_penalized.argtypes = [vp, C.c_int64]
    lib.lb200_last_dfill_kind.argtypes = [vp]
    lib.lb200_rows_fallbacks.argtypes = [vp]
    lib.lb200_rows_fallbacks.restype = C.c_int64
    lib.lb200_last_launches.argtypes = [vp]
    lib.lb200_last_launches.restype = C.c_int64
    lib.lb200_envelope_stats.argtypes = [vp, i64p, i64p]
    lib.lb200_pair_score.argtypes = [vp, C.c_int, i64p]
    lib.lb200_get_scores.argtypes = [vp, i64p, C.c_int]
    lib.lb200_pair_get_info.argtypes = [vp, C.c_int, C.POINTER(PairInfo)]
    lib.lb200_pair_band.argtypes = [vp, C.c_int, ip, ip]
    lib.lb200_pair_arcmatches.argtypes = [vp, C.c_int, ip, ip, ip, ip, ip, i64p]
    lib.lb200_pair_alignment.argtypes = [vp, C.c_int, ip, ip, C.c_char_p, C.c_char_p]
    lib.lb200_run_pf.argtypes = [vp, C.c_double]
    lib.lb200_pair_partition_function.argtypes = [vp, C.c_int, dp]
    lib.lb200_pair_arcmatch_pf.argtypes = [vp, C.c_int, dp]
    lib.lb200_run_pf_probs.argtypes = [vp, C.c_double, C.c_double]
    lib.lb200_pair_arcmatch_probs.argtypes = [vp, C.c_int, dp]
    lib.lb200_pair_basematch_probs.argtypes = [vp, C.c_int, dp]
    lib.lb200_upgma_newick.argtypes = [C.c_int, C.POINTER(C.c_char_p), i64p, C.c_char_p, C.c_size_t]
    _lib = lib
    return lib


# reference CLI flag name -> lb200_params field
FLAG_FIELDS = {
    "min-prob": "min_prob", "max-diff-am": "max_diff_am", "max-diff-at-am": "max_diff_at_am", "max-diff": "max_diff",
    "min-trace-probability": "min_trace_probability", "noLP": "no_lonely_pairs", "struct-local": "struct_local",
    "sequ-local": "sequ_local", "struct-weight": "struct_weight", "indel": "indel", "indel-opening": "indel_opening",
    "tau": "tau", "exclusion": "exclusion", "match": "match", "mismatch": "mismatch",
    "temperature-alipf": "temperature_alipf", "unpaired-penalty": "unpaired_penalty", "pf-double": "pf_double", "exp-prob": "exp_prob", "maxBPspan": "max_bp_span", "max-bps-length-ratio": "max_bps_length_ratio", "stacking": "stacking", "new-stacking": "new_stacking",
}


def make_params(flags: dict | None = None) -> Params:
    p = Params()
    load().lb200_default_params(C.byref(p))
    for k, v in (flags or {}).items():
        if k == "free-endgaps":
            p.free_endgaps = v.encode()
        elif k == "no-ribosum":
            p.use_ribosum = 0 if v else 1
        else:
            setattr(p, FLAG_FIELDS[k], int(v) if isinstance(v, bool) else v)
    return p


def band_from_alignment(lenA: int, lenB: int, aliA: str, aliB: str, delta: int, relaxed: bool = False):
    """min_col / max_col of TraceController(seqA, seqB, reference alignment, delta) (lb200_band_from_alignment)."""
    lo, hi = (C.c_int * (lenA + 1))(), (C.c_int * (lenA + 1))()
    rc = load().lb200_band_from_alignment(lenA, lenB, aliA.encode(), aliB.encode(), delta, int(relaxed), lo, hi)
    if rc != OK:
        raise Error("lb200_band_from_alignment failed with code %d" % rc)
    return list(lo), list(hi)


class Error(RuntimeError):
    pass


class Context:
    """One GPU context: sequences are uploaded once, pairs are batched, `run` aligns them all."""

    def __init__(self, device: int = 0, flags: dict | None = None):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.lb200_ctx_create(device, C.byref(h))
        if rc != OK:
            raise Error("lb200_ctx_create failed with code %d (no CUDA device? there is no CPU fallback)" % rc)
        self.h = h
        if flags is not None:
            self.set_params(flags)

    def close(self):
        if getattr(self, "h", None):
            self.lib.lb200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise Error("locarna_b200 error %d: %s" % (rc, self.lib.lb200_last_error(self.h).decode()))
        return rc

    def set_params(self, flags: dict):
        p = make_params(flags)
        self._chk(self.lib.lb200_set_params(self.h, C.byref(p)))

    def add_pp(self, path: str) -> int:
        return self._chk(self.lib.lb200_seq_add_pp(self.h, path.encode()))

    def add_pps(self, paths) -> int:
        """Add PP 2.0 files with one call (parsed in parallel on the host); returns the id of the first, ids are consecutive."""
        arr = (C.c_char_p * max(len(paths), 1))(*[p.encode() for p in paths])
        return self._chk(self.lib.lb200_seqs_add_pp(self.h, len(paths), arr))

    def add_seq(self, name: str, seq: str, pairs) -> int:
        n = len(pairs)
        ai = (C.c_int * max(n, 1))(*[p[0] for p in pairs])
        aj = (C.c_int * max(n, 1))(*[p[1] for p in pairs])
        ap = (C.c_double * max(n, 1))(*[p[2] for p in pairs])
        return self._chk(self.lib.lb200_seq_add(self.h, name.encode(), seq.encode(), ai, aj, ap, n))

    def seq_length(self, seq: int) -> int:
        return self._chk(self.lib.lb200_seq_length(self.h, seq))

    def set_ribosum_file(self, path):
        """--ribosum-file: score tables from a matrix file (None / "RIBOSUM85_60": the built-in matrix)."""
        self._chk(self.lib.lb200_set_ribosum_file(self.h, None if path is None else path.encode()))

    def copy_seqs_from(self, other: "Context") -> int:
        """Append the parsed sequences of another context (lb200_seqs_copy); returns the index of the first one."""
        return self._chk(self.lib.lb200_seqs_copy(self.h, other.h))

    def seq_anchors(self, seq: int) -> str:
        n = self._chk(self.lib.lb200_seq_anchors(self.h, seq, None, 0))
        buf = C.create_string_buffer(n + 1)
        self._chk(self.lib.lb200_seq_anchors(self.h, seq, buf, n + 1))
        return buf.value.decode()

    def seq_rows(self, seq: int):
        """(name, aligned string) of every row of a sequence (one row for a single sequence, several for a profile input)."""
        n = self._chk(self.lib.lb200_seq_num_rows(self.h, seq))
        length = self._chk(self.lib.lb200_seq_length(self.h, seq))
        out = []
        for k in range(n):
            name, s = C.create_string_buffer(256), C.create_string_buffer(length + 1)
            self._chk(self.lib.lb200_seq_get_row(self.h, seq, k, name, 256, s, length + 1))
            out.append((name.value.decode(), s.value.decode()))
        return out

    def seq_num_arcs(self, seq: int) -> int:
        return self._chk(self.lib.lb200_seq_num_arcs(self.h, seq))

    def add_pairs(self, pairs) -> int:
        """Add a list of (seqA, seqB) pairs with one call (bands derived like the reference does)."""
        n = len(pairs)
        a = (C.c_int * max(n, 1))(*[p[0] for p in pairs])
        b = (C.c_int * max(n, 1))(*[p[1] for p in pairs])
        return self._chk(self.lib.lb200_pairs_add(self.h, n, a, b))

    def add_pairs_np(self, a, b) -> int:
        """add_pairs for two int32 numpy arrays (no per-pair Python work)."""
        import numpy as np
        a = np.ascontiguousarray(a, dtype=np.int32); b = np.ascontiguousarray(b, dtype=np.int32)
        if a.shape != b.shape or a.ndim != 1:
            raise Error("add_pairs_np: two 1-d arrays of equal length expected")
        return self._chk(self.lib.lb200_pairs_add(self.h, len(a), a.ctypes.data_as(C.POINTER(C.c_int)), b.ctypes.data_as(C.POINTER(C.c_int))))

    def shard_job(self, a, b, world: int):
        """Cost-balanced split (lb200_shard_job) of the pairs (a[k], b[k]) of this context's sequences over `world` shares:
        list of int64 numpy arrays of pair indices, each by descending cost."""
        import numpy as np
        a = np.ascontiguousarray(a, dtype=np.int32); b = np.ascontiguousarray(b, dtype=np.int32)
        n = len(a)
        order = np.empty(max(n, 1), dtype=np.int64); begin = np.zeros(world + 1, dtype=np.int64)
        self._chk(self.lib.lb200_shard_job(self.h, n, a.ctypes.data, b.ctypes.data, world, None, order.ctypes.data, begin.ctypes.data))
        return [order[begin[r]:begin[r + 1]] for r in range(world)]

    def scores_np(self):
        """Scores as an int64 numpy array; SCORE_NEG_INF marks -inf."""
        import numpy as np
        n = self.num_pairs()
        out = np.empty(max(n, 1), dtype=np.int64)
        self._chk(self.lib.lb200_get_scores(self.h, out.ctypes.data_as(C.POINTER(C.c_int64)), n))
        return out[:n]

    def add_pair(self, a: int, b: int, band=None) -> int:
        if band is None:
            return self._chk(self.lib.lb200_pair_add(self.h, a, b, None, None))
        lo, hi = band
        n = self.seq_length(a) + 1
        if len(lo) != n or len(hi) != n:
            raise Error("band arrays must hold lenA + 1 = %d entries (got %d / %d)" % (n, len(lo), len(hi)))
        return self._chk(self.lib.lb200_pair_add(self.h, a, b, (C.c_int * len(lo))(*lo), (C.c_int * len(hi))(*hi)))

    def add_pair_restricted(self, a: int, b: int, lo, hi) -> int:
        """A pair whose band (before the probability envelope) is given: reference-alignment or anchor bands."""
        n = self.seq_length(a) + 1
        if len(lo) != n or len(hi) != n:
            raise Error("band arrays must hold lenA + 1 = %d entries (got %d / %d)" % (n, len(lo), len(hi)))
        return self._chk(self.lib.lb200_pair_add_restricted(self.h, a, b, (C.c_int * n)(*lo), (C.c_int * n)(*hi)))

    def clear_pairs(self):
        self._chk(self.lib.lb200_clear_pairs(self.h))

    def prepare(self):
        self._chk(self.lib.lb200_prepare(self.h))

    def upload(self):
        self._chk(self.lib.lb200_upload(self.h))

    @property
    def h2d_bytes(self) -> int:
        return self.lib.lb200_last_h2d_bytes(self.h)

    @property
    def d2h_bytes(self) -> int:
        return self.lib.lb200_last_d2h_bytes(self.h)

    def run(self, flags: int = RUN_SCORE_ONLY):
        self._chk(self.lib.lb200_run(self.h, flags))

    def run_normalized(self, L: int):
        """Normalized local alignment (locarna --normalized L); scores() gives the normalized scores, alignment(k) the alignments."""
        self._chk(self.lib.lb200_run_normalized(self.h, L))

    def set_restriction(self, pair: int, startA: int, startB: int, endA: int, endB: int):
        self._chk(self.lib.lb200_pair_set_restriction(self.h, pair, startA, startB, endA, endB))

    def run_pair_toplevel(self, pair: int, mode: int = 0, arg: int = 0, flags: int = RUN_TRACE):
        """Top level (+ trace) of one pair on the resident D table under its restriction (k-best building block)."""
        self._chk(self.lib.lb200_run_pair_toplevel(self.h, pair, mode, arg, flags))

    def run_penalized(self, position_penalty: int):
        self._chk(self.lib.lb200_run_penalized(self.h, position_penalty))

    @property
    def kernel_ms(self) -> float:
        return self.lib.lb200_last_kernel_ms(self.h)

    @property
    def dfill_ms(self) -> float:
        return self.lib.lb200_last_dfill_ms(self.h)

    @property
    def dfill_launches(self) -> int:
        return self.lib.lb200_last_dfill_launches(self.h)

    @property
    def dfill_kind(self) -> int:
        """0: one launch per level group, 1: dependency-driven boxes, 2: row-grouped kernel"""
        return self.lib.lb200_last_dfill_kind(self.h)

    @property
    def rows_fallbacks(self) -> int:
        return self.lib.lb200_rows_fallbacks(self.h)

    @property
    def launches(self) -> int:
        return self.lib.lb200_last_launches(self.h)

    def envelope_stats(self):
        """(pairs whose band was decided on the GPU, pairs recomputed on the host) of the last prepare/upload."""
        a, b = C.c_int64(), C.c_int64()
        self._chk(self.lib.lb200_envelope_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def num_pairs(self) -> int:
        return self._chk(self.lib.lb200_num_pairs(self.h))

    def scores(self):
        n = self.num_pairs()
        out = (C.c_int64 * max(n, 1))()
        self._chk(self.lib.lb200_get_scores(self.h, out, n))
        return [None if out[k] == SCORE_NEG_INF else out[k] for k in range(n)]

    def info(self, pair: int) -> PairInfo:
        inf = PairInfo()
        self._chk(self.lib.lb200_pair_get_info(self.h, pair, C.byref(inf)))
        return inf

    def band(self, pair: int):
        n = self.info(pair).lenA + 1
        lo, hi = (C.c_int * n)(), (C.c_int * n)()
        self._chk(self.lib.lb200_pair_band(self.h, pair, lo, hi))
        return list(lo), list(hi)

    def arcmatches(self, pair: int, with_D: bool = False):
        K = self.info(pair).n_arcmatches
        arrs = [(C.c_int * max(K, 1))() for _ in range(5)]
        D = (C.c_int64 * max(K, 1))() if with_D else None
        self._chk(self.lib.lb200_pair_arcmatches(self.h, pair, *arrs, D))
        am = [tuple(a[k] for a in arrs[:4]) for k in range(K)]
        score = [arrs[4][k] for k in range(K)]
        if with_D:
            return am, score, [None if D[k] == SCORE_NEG_INF else D[k] for k in range(K)]
        return am, score

    def run_pf(self, pf_scale: float = 1.0):
        """LocARNA-P inside pass (FP64 on the GPU) for all pairs."""
        self._chk(self.lib.lb200_run_pf(self.h, pf_scale))

    def partition_function(self, pair: int) -> float:
        z = C.c_double()
        self._chk(self.lib.lb200_pair_partition_function(self.h, pair, C.byref(z)))
        return z.value

    def arcmatch_pf(self, pair: int):
        """Inside values D(a,b) in the reference's arc-match index order."""
        K = self.info(pair).n_arcmatches
        D = (C.c_double * max(K, 1))()
        self._chk(self.lib.lb200_pair_arcmatch_pf(self.h, pair, D))
        return [D[k] for k in range(K)]

    def run_pf_probs(self, pf_scale: float = 1.0, min_am_prob: float = 0.001):
        """LocARNA-P complete: inside, outside, arc-match and base-match probabilities (FP64 on the GPU)."""
        self._chk(self.lib.lb200_run_pf_probs(self.h, pf_scale, min_am_prob))

    def arcmatch_probs(self, pair: int):
        K = self.info(pair).n_arcmatches
        D = (C.c_double * max(K, 1))()
        self._chk(self.lib.lb200_pair_arcmatch_probs(self.h, pair, D))
        return [D[k] for k in range(K)]

    def basematch_probs(self, pair: int):
        inf = self.info(pair)
        W = inf.lenB + 1
        buf = (C.c_double * ((inf.lenA + 1) * W))()
        self._chk(self.lib.lb200_pair_basematch_probs(self.h, pair, buf))
        return [[buf[i * W + j] for j in range(W)] for i in range(inf.lenA + 1)]

    def alignment(self, pair: int):
        inf = self.info(pair)
        n = inf.n_edges
        ea, eb = (C.c_int * max(n, 1))(), (C.c_int * max(n, 1))()
        sa, sb = C.create_string_buffer(inf.lenA + 2), C.create_string_buffer(inf.lenB + 2)
        self._chk(self.lib.lb200_pair_alignment(self.h, pair, ea, eb, sa, sb))
        return [(ea[k], eb[k]) for k in range(n)], sa.value.decode(), sb.value.decode()


def upgma_newick(names, matrix) -> str:
    """UPGMA guide tree (newick, without ';') of a symmetric score matrix, as mlocarna builds it (host code, no GPU needed)."""
    n = len(names)
    arr = (C.c_char_p * n)(*[x.encode() for x in names])
    flat = (C.c_int64 * (n * n))(*[int(matrix[i][j]) for i in range(n) for j in range(n)])
    out = C.create_string_buffer(64 * n + 1024)
    rc = load().lb200_upgma_newick(n, arr, flat, out, len(out))
    if rc != OK:
        raise Error("lb200_upgma_newick failed with code %d" % rc)
    return out.value.decode()
