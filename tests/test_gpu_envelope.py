"""Probability envelope on the GPU (FP64 screening + exact host re-check): bands identical to the oracle's 80-bit bands."""
import pytest

from locarna_b200 import capi
from oracle import oracle as O

pytestmark = pytest.mark.gpu

FLAGSETS = [{}, {"noLP": True, "max-diff-am": 30}, {"sequ-local": True}, {"free-endgaps": "++++"}, {"free-endgaps": "+--+"},
            {"max-diff": 25}, {"min-trace-probability": 1e-6}, {"min-trace-probability": 0.01, "temperature-alipf": 150},
            {"no-ribosum": True, "struct-weight": 100}, {"pf-double": True}]


@pytest.mark.parametrize("flags", FLAGSETS + [dict(_v1=True), dict({"sequ-local": True}, _v1=True), dict({"free-endgaps": "+--+"}, _v1=True)],
                         ids=lambda f: "_".join("%s=%s" % kv for kv in f.items()) or "default")
def test_device_bands_equal_oracle(synth_dir, flags, monkeypatch):
    flags = dict(flags)
    if flags.pop("_v1", False):   # the global-memory kernel that serves sequences too long for the shared-memory sweep
        monkeypatch.setenv("LB200_ENVELOPE_V1", "1")
    files = synth_dir["cfg3"][:6] + synth_dir["short"][:4] + synth_dir["cfg2"][:2]
    pairs = [(files[a], files[b]) for a in range(len(files)) for b in range(a)]
    ctx = capi.Context(0, flags)
    ids = {f: ctx.add_pp(f) for f in files}
    for a, b in pairs:
        ctx.add_pair(ids[a], ids[b])
    ctx.prepare()
    dev, host = ctx.envelope_stats()
    assert dev + host == len(pairs)
    if "no-ribosum" not in flags:  # match/mismatch scores of 50/0 overflow FP64 for longer pairs: those fall back to the host
        assert dev >= len(pairs) - 2, "the FP64 screening should decide (almost) every pair: device %d host %d" % (dev, host)
    for k, (a, b) in enumerate(pairs):
        ref = O.port_align(a, b, flags, setup_only=True)
        lo, hi = ctx.band(k)
        assert lo == ref["min_col"] and hi == ref["max_col"], (flags, a, b)
    ctx.close()


def test_host_mode_env(synth_dir, monkeypatch):
    monkeypatch.setenv("LB200_ENVELOPE", "host")
    a, b = synth_dir["cfg3"][:2]
    ctx = capi.Context(0, {})
    ctx.add_pair(ctx.add_pp(a), ctx.add_pp(b))
    ctx.prepare()
    assert ctx.envelope_stats() == (0, 1)
    ref = O.port_align(a, b, {}, setup_only=True)
    assert ctx.band(0) == (ref["min_col"], ref["max_col"])
