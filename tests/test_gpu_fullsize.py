"""BASELINE config 5 at its real sizes (300-nt sequences, one bench step of ~1000 pairs, mlocarna's tree-stage flags): properties
that do not need the oracle on every pair - the result must not depend on the D-fill schedule, the entry-stream format, the
chunking of the batch or the number of boxes in flight - plus an oracle spot check and the reference's cell count."""
import hashlib

import pytest

from locarna_b200 import capi, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
FLAGS = {"noLP": True, "max-diff-am": 30, "struct-weight": 200, "min-prob": 0.001}
N_SEQ = 46  # 1035 pairs: one step of bench.py


@pytest.fixture(scope="module")
def family(tmp_path_factory):
    return synth.make_family(str(tmp_path_factory.mktemp("cfg5")), 5, N_SEQ, 300)


def run(family, run_flags=capi.RUN_SCORE_ONLY):
    ctx = capi.Context(device=0, flags=FLAGS)
    ids = [ctx.add_pp(p) for p in family]
    pairs = [(a, b) for a in range(N_SEQ) for b in range(a)]       # mlocarna's pair order
    for a, b in pairs:
        ctx.add_pair(ids[a], ids[b])
    ctx.run(run_flags)
    scores = ctx.scores()
    cells = [ctx.info(k).cells for k in range(len(pairs))]
    kind = (ctx.dfill_kind, ctx.rows_fallbacks)
    ctx.close()
    run.kind = kind
    return pairs, scores, cells


def digest(xs):
    return hashlib.sha256(",".join(map(str, xs)).encode()).hexdigest()


def test_config5_step_is_schedule_format_and_chunk_independent(family, monkeypatch):
    pairs, base, cells = run(family)
    assert len(pairs) == 1035 and all(s is not None for s in base)
    assert run.kind == (2, 0)                                    # the row-grouped kernel ran, no chunk fell back
    want = digest(base)
    for env in ({"LB200_DFILL": "levels"}, {"LB200_DFILL": "dep", "LB200_PACK": "0"}, {"LB200_CHUNK_PAIRS": "300"},
                {"LB200_DFILL": "dep", "LB200_CTAS_PER_SM": "7"}, {"LB200_DFILL": "rows", "LB200_ROWS_FORCE_NC": "2"},
                {"LB200_DFILL": "rows", "LB200_ROWS_CTAS_PER_SM": "5"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        _, scores, cells2 = run(family)
        assert digest(scores) == want, env
        assert cells2 == cells, env
        assert run.kind == ({"levels": 0, "dep": 1, "rows": 2}.get(env.get("LB200_DFILL"), 2), 0), env
        for k in env:
            monkeypatch.delenv(k)
    # oracle spot check (the oracle needs about a second per 300-nt pair) incl. the reference's count of align_noex calls
    for k in (0, 517, 1034):
        a, b = pairs[k]
        ref = O.port_align(family[a], family[b], FLAGS, do_trace=False)
        assert base[k] == ref["score"], k
        assert cells[k] == ref["cells"], k


def test_locarna_p_probabilities_are_probabilities_at_300nt(family):
    """LocARNA-P (inside, outside, probabilities) on one 300-nt pair of the config-5 family: size-independent sanity of the result -
    finite positive partition function, every arc-match / base-match probability inside [0, 1], and every position matched with total
    probability <= 1 (row and column sums of the base-match matrix). Slack 1e-3: the reference's read of its -1 debugging fill
     perturbs a few probabilities in the 4th digit, and the device reproduces that."""
    import math
    ctx = capi.Context(0, {"pf-double": True, "min-trace-probability": 1e-5})
    a, b = ctx.add_pp(family[0]), ctx.add_pp(family[1])
    ctx.add_pair(a, b)
    ctx.run_pf_probs(1.0, 0.001)
    z = ctx.partition_function(0)
    assert math.isfinite(z) and z > 0
    am = ctx.arcmatch_probs(0)
    assert am and all(math.isfinite(p) and -1e-3 <= p <= 1 + 1e-3 for p in am)
    bm = ctx.basematch_probs(0)
    assert all(math.isfinite(p) and -1e-3 <= p <= 1 + 1e-3 for row in bm for p in row)
    assert max(sum(row) for row in bm) <= 1 + 1e-3
    assert max(sum(row[j] for row in bm) for j in range(len(bm[0]))) <= 1 + 1e-3
    assert max(sum(row) for row in bm) > 0.5                          # the alignment is not empty
    ctx.close()
