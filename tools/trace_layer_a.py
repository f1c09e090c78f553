#!/usr/bin/env python3
"""Layer A (staged whole-file API) on the two 4K files of bench_configs.layer_a with JPEG_SM100_TRACE=1: per-scan wall times on
stderr, the summary on stdout.  usage: [JPEG_SM100_ACR=0|1] trace_layer_a.py"""
import json
import os
import sys

os.environ.setdefault("JPEG_SM100_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bench_configs  # noqa: E402

q = [bench.quanta(bench.LEVEL, 0), bench.quanta(bench.LEVEL, 1), bench.quanta(bench.LEVEL, 1)]
print(json.dumps(bench_configs.layer_a(0, q), indent=1))
