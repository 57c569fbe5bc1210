"""bench.py end to end on the GPU with a small job: the JSON line carries every contract field, the resident and the e2e leg
agree on every score, the CPU sample agrees with the GPU, and memory does not grow with the step count (the round-1 driver run
failed with out-of-memory at --steps 20 --warmup 5)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_bench_small_job_many_steps():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "20", "--warmup", "5", "--seqs", "48", "--job-pairs", "256",
           "--sub-batch", "64", "--cpu-sample", "4"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config",
                "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "score_checksum"):
        assert key in line, key
    assert line["steps"] == 20 and line["warmup"] == 5 and line["scaling"] == "strong"
    assert line["value"] > 0 and line["e2e"]["value"] > 0 and line["gpu_launches"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["resident_vs_e2e_score_mismatches"] == 0
    assert line["cpu_baseline"]["scores_match_gpu"] is True
    assert 0 < line["roofline"]["frac"] < 1
