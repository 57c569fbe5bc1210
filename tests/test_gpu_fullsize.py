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
    ctx.close()
    return pairs, scores, cells


def digest(xs):
    return hashlib.sha256(",".join(map(str, xs)).encode()).hexdigest()


def test_config5_step_is_schedule_format_and_chunk_independent(family, monkeypatch):
    pairs, base, cells = run(family)
    assert len(pairs) == 1035 and all(s is not None for s in base)
    want = digest(base)
    for env in ({"LB200_DFILL": "levels"}, {"LB200_DFILL": "dep", "LB200_PACK": "0"}, {"LB200_CHUNK_PAIRS": "300"},
                {"LB200_DFILL": "dep", "LB200_CTAS_PER_SM": "7"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        _, scores, cells2 = run(family)
        assert digest(scores) == want, env
        assert cells2 == cells, env
        for k in env:
            monkeypatch.delenv(k)
    # oracle spot check (the oracle needs about a second per 300-nt pair) incl. the reference's count of align_noex calls
    for k in (0, 517, 1034):
        a, b = pairs[k]
        ref = O.port_align(family[a], family[b], FLAGS, do_trace=False)
        assert base[k] == ref["score"], k
        assert cells[k] == ref["cells"], k
