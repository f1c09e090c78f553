"""CPU-side checks of bench.py: the quantisation tables it builds the workload with (the host mirror's
CompressionLevel) equal the oracle's (encode.swift:286-333), and the reference arm prints the contract's JSON line."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O


def _bench():
    sys.path.insert(0, ROOT)
    import bench
    return bench


@pytest.mark.parametrize("level", [0.0, 0.125, 0.25, 0.5, 1.0, 2.0, 4.0, 8.0])
def test_bench_quanta_equal_the_oracles(level):
    b = _bench()
    for chroma in (0, 1):
        assert np.array_equal(b.quanta(level, chroma), O.quanta(level, chroma)), (level, chroma)


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference: the reference's CPU path (restated oracle) on a bounded sample; one JSON line."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--batch", "8"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    b = _bench()
    assert d["impl"] == "reference" and d["metric"] == b.METRIC and d["unit"] == "Mpixels/s" and d["higher_is_better"] is True
    assert d["config"] == b.base_config(8) and d["value"] > 0 and d["n_gpus"] == 1  # the same object our arm prints
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
