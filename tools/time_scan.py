#!/usr/bin/env python3
"""Times layer-B entropy decode (K3) of synthetic 4K 4:2:0 frames for a given restart interval (0 = none, the reference
encoder's own output) -- device-resident, CUDA events.  usage: time_scan.py <n_frames> <interval_rows>"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import FACTORS, H, LEVEL, W, quanta  # noqa: E402
from jpeg_b200 import batch, lib, synth  # noqa: E402

n, rows = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream()
ctx = lib.Context(0, stream=stream.cuda_stream)
geo = batch.Geometry((W, H), FACTORS)
q = np.stack([quanta(LEVEL, 0), quanta(LEVEL, 1), quanta(LEVEL, 1)])
ecs_all, tabs_all, ref = [], [], []
for base in range(0, n, 8):
    frames = torch.stack([synth.frame(i, W, H, dev) for i in range(base, min(n, base + 8))])
    ecs, tabs, enc = batch.encode_frames(ctx, frames, geo, q, rows * geo.blocks[0])
    ecs_all += [e.copy() for e in ecs]
    tabs_all += list(tabs)
    ref.append([c.clone() for c in enc.coef])
inputs = batch.DecodeInputs(ecs_all, tabs_all)
tables = (lib.HuffTable * (8 * n))(*tabs_all)
desc = batch.sequential_scan(geo)
buf = batch.DeviceBuffers(geo, n, dev)
d_ecs = torch.from_numpy(inputs.ecs).to(dev)
d_off = torch.from_numpy(inputs.offsets.view(np.int64)).to(dev)
d_st = torch.zeros(n, dtype=torch.int32, device=dev)
interval = rows * geo.blocks[0] if rows else lib.INTERVAL_NONE


def run():
    ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(desc), d_ecs.data_ptr(), d_off.data_ptr(), inputs.n_ecs, interval,
                                               lib.SCAN_FRESH, tables, 0, C.byref(buf.sp), d_st.data_ptr()))


for _ in range(2):
    run()
torch.cuda.synchronize()
assert d_st.cpu().abs().sum().item() == 0
for p in range(3):
    want = torch.cat([r[p] for r in ref])
    assert torch.equal(buf.coef[p], want), p
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(5):
    run()
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"frames {n} interval_rows {rows} n_ecs {inputs.n_ecs}: K3 {ms:.3f} ms  ({n * W * H / ms / 1e3:.0f} Mpixels/s)")
