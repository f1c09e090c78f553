#!/usr/bin/env python3
"""bench.py -- Mpixels/s decode of 4K (3840x2160) baseline 4:2:0 JPEG, batch 64 synthetic frames per GPU.

    python bench.py --gpus N --steps K --warmup W            our CUDA path (libjpeg_sm100.so)
    python bench.py --impl reference ...                      the reference's CPU path (restated oracle), host cores
    python bench.py --quick [--sweep T:WARM,...]              device-resident stage times only (kernel A/B runs; not a bench line)

A "step" is one pass of the hot path over one batch: scan lexing (N1: unstuff + RSTn split) -> entropy decode (K3, which also
clears the fresh coefficient planes and resolves the DC predictions) -> dequantise + IDCT (K1) -> upsample + YCbCr->RGB +
pack (K2).  `value` times it with every input (the raw scan bytes of the files) already resident in HBM (CUDA events on the
launching stream, max over ranks); `e2e` -- the headline -- times the same work through the host-buffer C-ABI call
jpeg_sm100_decode_batch_raw_rgb8 with pinned host buffers, H2D and D2H copies inside the timed region, over all K steps, and
reports next to it what the PCIe links sustain device -> host alone AND with all ranks copying at once (the e2e bound: 3 bytes
of RGB per pixel leave the device).
Inputs are produced by our own GPU encoder (K4-K7) from deterministic synthetic frames, DRI = one MCU row.
Multi-GPU: independent images are sharded across ranks, no collective on the data path ("weak": 64 frames per GPU).
The other configurations of BASELINE.json (#3 4K encode, #4 1080p progressive decode, #5 12-Mpixel 4:4:4 decode -> re-encode,
64 frames per GPU) and the staged whole-file API (layer A) are timed in the same run and reported under `configs` / `layer_a`.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, BATCH = 3840, 2160, 64
FACTORS = [(2, 2), (1, 1), (1, 1)]
LEVEL = 0.25
WORKLOAD = "3840x2160 baseline 4:2:0 decode, batch 64 synthetic frames per GPU, DRI = 240 MCUs (1 MCU row), level 0.25"
METRIC = "Mpixels/s decode (4K 4:2:0 baseline)"


def base_config(frames):
    """the `config` object, identical in both arms (the driver compares them)"""
    return {"workload": WORKLOAD, "frames_per_step_per_gpu": int(frames),
            "l2": "inputs larger than L2 (coefficients 1.6 GB, RGB 1.6 GB per step)"}


def k1_traffic():
    """DRAM bytes per (average) K1 launch from the committed ncu --set full capture of this same command"""
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            return int(json.load(f)["dram_bytes_per_launch_avg"])
    except Exception:
        return None


def k3_traffic():
    """DRAM bytes of one k_decode_par launch (64 frames) from the committed ncu --set full capture"""
    try:
        with open(os.path.join(ROOT, "profiles", "k3_traffic.json")) as f:
            return int(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def quanta(level, chroma):
    """JPEG.CompressionLevel.luminance(level).quanta / .chrominance(level).quanta (encode.swift:286-333) from the product's host
    mirror (the reference arm takes the oracle's)"""
    from jpeg_b200.host import CompressionLevel
    return (CompressionLevel.chrominance(level) if chroma else CompressionLevel.luminance(level)).quanta


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (restated C oracle; no Swift toolchain exists here)
# ----------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor

    import torch

    from jpeg_b200 import synth
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    frames = args.batch  # the whole batch of the workload per step, one frame per task, all host threads
    q = [O.quanta(LEVEL, 0), O.quanta(LEVEL, 1), O.quanta(LEVEL, 1)]

    # inputs: the same synthetic frames, encoded by the oracle's reference-equivalent encoder (DRI = 240 MCUs)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import jpegfile as J

    def make2(i):
        rgb = synth.frame(i, W, H, "cpu").numpy()
        planes = O.decompose(O.pack_rgb(rgb), FACTORS)
        s = O.Spectral.create((W, H), FACTORS)
        for p in range(3):
            s.coefficients(p)[...] = O.fdct_plane(planes[p], q[p])
        ecs, dct, act = s.encode_scan((0, 64), (0, None), [0, 1, 2], [0, 1, 1], [0, 1, 1], 240)
        return J.unstuff_split(ecs), dct, act

    torch.set_num_threads(1)
    with ThreadPoolExecutor(cores) as pool:
        inputs = list(pool.map(make2, range(frames)))

        def decode(inp):  # Spectral.decode(ecss:) -> idct -> interleaved -> unpack(as: RGB) on one core
            parts, dct, act = inp
            s = O.Spectral.create((W, H), FACTORS)
            for p in range(3):
                s.set_quanta(p, q[p])
            s.decode_scan((0, 64), (0, None), [0, 1, 2], [0, 1, 1], [0, 1, 1], dct, act, parts, interval=240)
            return O.unpack_rgb(s.to_rectangular())[0, 0, 0]

        for _ in range(args.warmup):
            list(pool.map(decode, inputs))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            list(pool.map(decode, inputs))
        dt = (time.perf_counter() - t0) / args.steps
    value = frames * W * H / dt / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": "Mpixels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config(frames),
            "details": {"note": "the reference's CPU path (C restatement of the Swift code; no Swift toolchain here): every step decodes the "
                                "whole batch from pre-lexed entropy-coded segments (the lexer, which our arm includes, is left out), one frame per task"},
            "cpu_baseline": {"value": round(value, 2), "unit": "Mpixels/s", "cores": cores, "kind": "port",
                             "sample": f"{frames} frames of the workload per step, one frame per task, {cores} threads"},
            "e2e": {"value": round(value, 2), "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
class StdoutGuard:
    """Rank 0 prints exactly ONE line on stdout.  NCCL's init report (NCCL_DEBUG=INFO: nranks, transports -- the record that the
    job really ran on N ranks) and any other library chatter go to file descriptor 1 as well, so for the duration of the run
    fd 1 is pointed at stderr and the JSON line is written to the saved descriptor."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())


def run_ours(args):
    import torch
    import torch.distributed as dist

    from jpeg_b200 import batch, lib, synth
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: jpeg_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    placement = "unchanged (single rank)"
    out = StdoutGuard()
    if world > 1:
        # one process per GPU: keep the rank's threads and pinned staging buffers on the GPU's own NUMA node (best effort)
        if os.environ.get("JPEG_SM100_NUMA", "1") != "0":
            from jpeg_b200 import affinity
            placement = affinity.bind_to_gpu(local)
        os.environ.setdefault("NCCL_DEBUG", "INFO")  # the init report (on stderr, see StdoutGuard)
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    ctx = lib.Context(local, stream=stream.cuda_stream)
    geo = batch.Geometry((W, H), FACTORS)
    q = np.stack([quanta(LEVEL, 0), quanta(LEVEL, 1), quanta(LEVEL, 1)])
    n = args.batch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()  # (the first collective: NCCL finishes its init here)

    # ---- inputs: synthetic frames -> our GPU encoder -> (host) lexer ------------------------------------------------
    ecs_all, tables_all = [], []
    chunk = 8
    for base in range(0, n, chunk):
        frames = torch.stack([synth.frame(rank * n + i, W, H, dev) for i in range(base, min(n, base + chunk))])
        ecs, tabs, _ = batch.encode_frames(ctx, frames, geo, q, geo.blocks[0])
        ecs_all += [e.copy() for e in ecs]
        tables_all += list(tabs)
        del frames, _
    torch.cuda.empty_cache()
    inputs = batch.DecodeInputs(ecs_all, tables_all, n_ecs_expected=geo.blocks[1])
    tables = (lib.HuffTable * (8 * n))(*tables_all)
    desc = batch.sequential_scan(geo)
    buf = batch.DeviceBuffers(geo, n, dev)
    # raw scan bytes (stuffed, RSTn-delimited) as they sit in the files; the GPU lexer (N1) is the first stage of a step
    raw_len = np.array([len(e) for e in ecs_all], dtype=np.uint64)
    raw_off = np.concatenate([[0], np.cumsum(raw_len)[:-1]]).astype(np.uint64)
    raw_cat = np.concatenate(ecs_all + [np.zeros(64, np.uint8)])
    d_raw = torch.from_numpy(raw_cat).to(dev)
    d_ecs = torch.zeros(raw_cat.size + 64, dtype=torch.uint8, device=dev)
    d_off = torch.zeros(n * inputs.n_ecs + 1, dtype=torch.int64, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    d_lex_status = torch.zeros(n, dtype=torch.int32, device=dev)
    qz = np.ascontiguousarray(q, dtype=np.uint16)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    # the back half of a step: the staged kernels K1 -> sample planes -> K2 (default: faster, see fused.cu) or K1+K2 in one kernel
    # (JPEG_SM100_FUSE=1: 6 instead of 9 bytes of HBM traffic per pixel, but issue-bound and slower)
    fused = os.environ.get("JPEG_SM100_FUSE", "0") not in ("", "0")
    STAGES = ("lexer", "huffman", "idct_color") if fused else ("lexer", "huffman", "idct", "color")
    stage_ms = {k: [] for k in STAGES}

    def step(record):
        e = []

        def mark():
            if record:
                e.append(ev())
                e[-1].record(stream)

        mark()
        ctx.check(ctx.L.jpeg_sm100_dev_lex_scan(ctx.h, d_raw.data_ptr(), raw_off.ctypes.data, raw_len.ctypes.data, n, inputs.n_ecs,
                                                d_ecs.data_ptr(), d_off.data_ptr(), d_lex_status.data_ptr()))
        mark()
        # SCAN_FRESH: the coefficient planes are new Spectral planes; K3 clears the rows it decodes into (no separate memset)
        ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(desc), d_ecs.data_ptr(), d_off.data_ptr(), inputs.n_ecs,
                                                   geo.blocks[0], lib.SCAN_FRESH, tables, 0, C.byref(buf.sp), d_status.data_ptr()))
        mark()
        if fused:
            ctx.check(ctx.L.jpeg_sm100_dev_spectral_to_rgb8(ctx.h, C.byref(buf.sp), qz.ctypes.data, W, H, 0, buf.rgb.data_ptr()))
        else:
            ctx.check(ctx.L.jpeg_sm100_dev_idct(ctx.h, C.byref(buf.sp), qz.ctypes.data, 8, C.byref(buf.pl)))
            mark()
            ctx.check(ctx.L.jpeg_sm100_dev_planar_to_rgb8(ctx.h, C.byref(buf.pl), W, H, 0, buf.rgb.data_ptr()))
        mark()
        return e

    for _ in range(args.warmup):
        step(False)
    barrier()
    assert d_status.cpu().abs().sum().item() == 0, "decode reported an error"
    assert d_lex_status.cpu().abs().sum().item() == 0, "lexer reported an error"
    assert np.array_equal(d_off.cpu().numpy().view(np.uint64), inputs.offsets), "GPU lexer offsets differ from the host lexer's"
    checksum = int(buf.rgb[0, ::97, ::89].to(torch.int64).sum().item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    t0, t1 = ev(), ev()
    barrier()
    t0.record(stream)
    recs = [step(True) for _ in range(args.steps)]
    t1.record(stream)
    barrier()
    launches = ctx.launches - launches0
    total_ms = t0.elapsed_time(t1)
    for e in recs:
        for k, name in enumerate(STAGES):
            stage_ms[name].append(e[k].elapsed_time(e[k + 1]))
    ms_per_step = total_ms / args.steps

    # the K1+K2 fused kernel (opt-in, fused.cu) on the same coefficients, next to the staged pair of the timed step: same bytes?
    fused_side = None
    if not fused and not args.quick:
        try:
            os.environ["JPEG_SM100_FUSE"] = "1"
            rgb2 = torch.zeros_like(buf.rgb)
            call = lambda: ctx.check(ctx.L.jpeg_sm100_dev_spectral_to_rgb8(ctx.h, C.byref(buf.sp), qz.ctypes.data, W, H, 0, rgb2.data_ptr()))
            call()
            f0, f1 = ev(), ev()
            f0.record(stream)
            for _ in range(5):
                call()
            f1.record(stream)
            torch.cuda.synchronize()
            f_ms = f0.elapsed_time(f1) / 5
            f_bytes = (128.0 * geo.total_blocks + 3.0 * W * H) * n
            fused_side = {"kernel": "k_idct_rgb420", "ms": round(f_ms, 4), "algorithmic_bytes": int(f_bytes),
                          "GBps": round(f_bytes / (f_ms * 1e-3) / 1e9, 1), "equals_staged": bool(torch.equal(rgb2, buf.rgb)),
                          "note": "opt-in (JPEG_SM100_FUSE=1): 6 B/px of HBM traffic instead of 9, but bound by instruction issue like K1 and K2 "
                                  "themselves; compare ms with stages.ms.idct + stages.ms.color"}
            del rgb2
        except Exception as e:
            fused_side = {"error": f"{type(e).__name__}: {e}"[:200]}
        finally:
            os.environ.pop("JPEG_SM100_FUSE", None)

    if args.quick:  # kernel A/B runs: device-resident stages only; --sweep "T:WARM,T:WARM,..." re-times K3p settings
        if rank == 0:
            sampler.stop()

        def report(ms, stages):
            if rank == 0:
                out.emit(json.dumps({"quick": True, "tag": os.environ.get("JPEG_SM100_LIB", ""), "ms_per_step": round(ms, 4),
                                  "value": round(n * W * H * world / (ms * 1e-3) / 1e6, 1),
                                  "stages_ms": {k: round(statistics.mean(v), 4) for k, v in stages.items()},
                                  "env": {k: v for k, v in os.environ.items() if k.startswith("JPEG_SM100_")}}))

        report(ms_per_step, stage_ms)
        for item in [x for x in args.sweep.split(",") if x]:
            ts, warm = item.split(":")
            os.environ["JPEG_SM100_PAR_T"], os.environ["JPEG_SM100_PAR_WARM"] = ts, warm
            for _ in range(2):
                step(False)
            barrier()
            assert d_status.cpu().abs().sum().item() == 0
            a0, a1 = ev(), ev()
            a0.record(stream)
            rr = [step(True) for _ in range(args.steps)]
            a1.record(stream)
            barrier()
            st2 = {k: [] for k in STAGES}
            for e in rr:
                for k, name in enumerate(STAGES):
                    st2[name].append(e[k].elapsed_time(e[k + 1]))
            report(a0.elapsed_time(a1) / args.steps, st2)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region -------------------------------------------
    # Images are independent, so the batch may be sharded over a few contexts (one stream each) driven by as many host threads:
    # while one share is copying its RGB back over PCIe another is uploading / entropy-decoding.  One call for the whole batch
    # reaches ~95 % of this by itself (the entry point pipelines groups of images internally), which is what ranks of a
    # multi-GPU job use: fewer host threads per rank fighting for the host's cores (JPEG_BENCH_STREAMS overrides).
    # Every step of every share uploads its raw scan bytes from pinned memory and reads its RGB back (Bi + Bo per step).
    rgb_bytes = n * W * H * 3
    default_streams = 4 if world == 1 else 1
    n_streams = max(1, min(n, int(os.environ.get("JPEG_BENCH_STREAMS", str(default_streams)))))
    halves = []
    for k in range(n_streams):
        i0, i1 = k * n // n_streams, (k + 1) * n // n_streams
        b0 = int(raw_off[i0])
        b1 = int(raw_off[i1 - 1] + raw_len[i1 - 1])
        halves.append({
            "n": i1 - i0, "i0": i0,
            "ctx": lib.Context(local),  # own stream + own device staging
            "raw": torch.from_numpy(np.concatenate([raw_cat[b0:b1], np.zeros(64, np.uint8)])).pin_memory(),
            "raw_off": (raw_off[i0:i1] - np.uint64(b0)).astype(np.uint64), "raw_len": raw_len[i0:i1].copy(),
            "rgb": torch.empty((i1 - i0) * W * H * 3, dtype=torch.uint8).pin_memory(),
            "status": np.zeros(i1 - i0, dtype=np.int32),
            "tables": (lib.HuffTable * (8 * (i1 - i0)))(*tables_all[8 * i0:8 * i1]),
        })

    def e2e_worker(h, steps):
        c = h["ctx"]
        for _ in range(steps):
            c.check(c.L.jpeg_sm100_decode_batch_raw_rgb8(c.h, C.byref(desc), h["n"], h["raw"].data_ptr(), h["raw_off"].ctypes.data,
                                                         h["raw_len"].ctypes.data, inputs.n_ecs, geo.blocks[0], h["tables"], 0,
                                                         qz.ctypes.data, W, H, 0, h["rgb"].data_ptr(), h["status"].ctypes.data))

    def e2e_run(steps):
        if len(halves) == 1:
            return e2e_worker(halves[0], steps)
        ths = [threading.Thread(target=e2e_worker, args=(h, steps)) for h in halves]
        for t_ in ths:
            t_.start()
        for t_ in ths:
            t_.join()

    e2e_run(max(1, min(args.warmup, 2)))
    e2e_steps = args.steps
    barrier()
    w0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_s = (time.perf_counter() - w0) / e2e_steps
    clocks = sampler.stop() if rank == 0 else None
    assert all(int(np.abs(h["status"]).sum()) == 0 for h in halves)
    got = halves[0]["rgb"].view(halves[0]["n"], H, W, 3)[0, ::97, ::89].to(torch.int64).sum().item()
    assert int(got) == checksum, "e2e result differs from the device-resident result"
    e2e_launches = sum(h["ctx"].launches for h in halves)

    # The e2e bound: every frame's 3 bytes per pixel cross PCIe once.  What the link sustains device -> pinned host, same buffer
    # as a share of the batch, (a) this rank alone while the others idle (ranks take turns), (b) all ranks at the same time --
    # the ceiling the N-rank e2e figure has to be read against.
    def d2h_rate(reps=3):
        src = buf.rgb.view(-1)[:halves[0]["rgb"].numel()]
        halves[0]["rgb"].copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        a0, a1 = ev(), ev()
        a0.record(stream)
        for _ in range(reps):
            halves[0]["rgb"].copy_(src, non_blocking=True)
        a1.record(stream)
        torch.cuda.synchronize()
        return reps * src.numel() / (a0.elapsed_time(a1) * 1e-3) / 1e9

    d2h_alone = d2h_conc = None
    try:
        for r in range(world):
            barrier()
            if r == rank:
                d2h_alone = d2h_rate()
        barrier()
        d2h_conc = d2h_rate()
        barrier()
    except Exception:
        pass

    # ---- the other configurations of BASELINE.json + the staged whole-file API, same run (bounded: a few seconds each) ----
    configs, layer_a = {}, None
    del d_raw, d_ecs
    for h in halves:
        h["ctx"].close()
    halves_rgb_numel = halves[0]["rgb"].numel()
    del halves
    torch.cuda.empty_cache()
    if not args.no_configs:
        import bench_configs
        for name, fn in (("3", bench_configs.config3), ("4", bench_configs.config4), ("5", bench_configs.config5)):
            try:
                barrier()
                configs[name] = fn(ctx, dev, q, rank, n)
            except Exception as e:  # reported, never fatal: the headline numbers are already measured
                configs[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()
        if rank == 0:
            try:
                layer_a = bench_configs.layer_a(local, q)
            except Exception as e:
                layer_a = {"error": f"{type(e).__name__}: {e}"[:300]}

    per_rank = {"placement": placement, "d2h_alone": d2h_alone, "d2h_concurrent": d2h_conc, "e2e_s": e2e_s, "ms_per_step": ms_per_step,
                "configs": {k: {kk: vv for kk, vv in v.items() if kk in ("ms", "error", "parity_checked")} for k, v in configs.items()}}
    if world > 1:
        t = torch.tensor([ms_per_step, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step, e2e_s = t[0].item(), t[1].item()
        gathered = [None] * world
        dist.all_gather_object(gathered, per_rank)
    else:
        gathered = [per_rank]

    if rank == 0:
        peak, peak_src = measured_peaks()
        px_per_step = n * W * H * world
        if fused:  # one launch: 128 B of coefficients per block in, 3 B of RGB per pixel out
            idct_ms = statistics.mean(stage_ms["idct_color"])
            idct_bytes = (128.0 * geo.total_blocks + 3.0 * W * H) * n
        else:
            idct_ms = statistics.mean(stage_ms["idct"]) / 3.0  # three launches (Y, Cb, Cr) per step
            idct_bytes = 192.0 * geo.total_blocks * n / 3.0    # algorithmic bytes of the average launch
        achieved = idct_bytes / (idct_ms * 1e-3) / 1e9
        shares = {k: round(statistics.mean(v) / (ms_per_step if world == 1 else sum(statistics.mean(x) for x in stage_ms.values())), 4)
                  for k, v in stage_ms.items()}
        huff_ms = statistics.mean(stage_ms["huffman"])
        huff_bytes = inputs.ecs_bytes + 2.0 * 64 * geo.total_blocks * n
        color_ms = statistics.mean(stage_ms["idct_color" if fused else "color"])
        color_bytes = (64.0 * geo.total_blocks + 3.0 * W * H) * n
        # CPU baseline on a bounded sample (single thread), same run
        cpu = cpu_baseline_sample(ecs_all[:3], tables_all[:24], q)
        conc = [g["d2h_concurrent"] for g in gathered if g["d2h_concurrent"]]
        alone = [g["d2h_alone"] for g in gathered if g["d2h_alone"]]
        e2e_gbps_per_rank = rgb_bytes / e2e_s / 1e9
        # configs: the job's time is the slowest rank's
        cfg_out = {}
        for k, v in configs.items():
            o = dict(v)
            ms_all = [g["configs"].get(k, {}).get("ms") for g in gathered]
            if all(m is not None for m in ms_all) and "pixels_per_gpu" in o:
                o["ms"] = round(max(ms_all), 3)
                o["Mpixels_per_s"] = round(o["pixels_per_gpu"] * world / (o["ms"] * 1e-3) / 1e6, 1)
                o["parity_checked"] = all(g["configs"][k].get("parity_checked") for g in gathered)
                o["n_gpus"] = world
            errs = [g["configs"].get(k, {}).get("error") for g in gathered if g["configs"].get(k, {}).get("error")]
            if errs:
                o["error"] = errs[0]
            cfg_out[k] = o
        line = {
            "metric": METRIC, "value": round(px_per_step / (ms_per_step * 1e-3) / 1e6, 1), "unit": "Mpixels/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config(n),
            "details": {"value_is": "device-resident (inputs in HBM when the timed region starts); the headline against the reference arm is e2e",
                        "ecs_bytes_per_frame": inputs.ecs_bytes // n, "parallelism": f"images sharded over {world} GPU(s), no collective",
                        "cpu_affinity": [g["placement"] for g in gathered],
                        "wait_mode": os.environ.get("JPEG_SM100_WAIT", "yield")},
            "e2e": {"value": round(px_per_step / e2e_s / 1e6, 1), "unit": "Mpixels/s",
                    "h2d_bytes_per_step": int(raw_len.sum() + raw_len.nbytes * 2), "d2h_bytes_per_step": int(rgb_bytes + 4 * n),
                    "pcie_d2h_GBps_measured": None if not alone else round(min(alone), 1),
                    "pcie_d2h_GBps_measured_alone_per_rank": [round(x, 1) for x in alone],
                    "pcie_d2h_GBps_measured_concurrent": None if not conc else {"min": round(min(conc), 1), "mean": round(statistics.mean(conc), 1),
                                                                                 "per_rank": [round(x, 1) for x in conc]},
                    "pcie_d2h_GBps_achieved": round(e2e_gbps_per_rank, 1),
                    "pcie_both_directions_GBps_achieved": round((rgb_bytes + float(raw_len.sum())) / e2e_s / 1e9, 1),
                    "frac_of_concurrent_ceiling": None if not conc else round(e2e_gbps_per_rank / min(conc), 3),
                    "streams": n_streams, "call": "jpeg_sm100_decode_batch_raw_rgb8 (raw scan bytes in pinned host memory -> GPU lexer -> RGB8 in pinned host memory)",
                    "steps": e2e_steps, "per_rank_s": [round(g["e2e_s"], 4) for g in gathered]},
            "gpu_launches": int(launches), "gpu_launches_e2e": int(e2e_launches),
            "roofline": {"kernel": "k_idct_rgb420 (K1+K2 in one kernel)" if fused else "k_idct_tma<u8> (fused de-zigzag+dequant+IDCT+clamp, K1)", "bound": "hbm",
                         "achieved": round(achieved, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": k1_traffic() if (n == BATCH and not fused) else None,
                         "bytes_per_launch": int(idct_bytes), "ms_per_launch": round(idct_ms, 5),
                         "note": "the HBM-bound kernel BASELINE.json's north_star sets the >= 70 % target on; the kernel that takes most of the step "
                                 "(K3, bound by dependent-instruction latency and instruction issue, not HBM) is under roofline_dominant"},
            "roofline_dominant": {"kernel": "K3 stage = k_build_luts + k_decode_par (self-synchronising subsequence-parallel Huffman decode: 16 threads per restart "
                                            "interval, 8-9 CTAs x 4 warps per SM, blocks assembled in shared memory and flushed by the whole warp as 128-byte lines, "
                                            "DC predictions resolved before the flush) + k_zero_flagged + k_decode_fast(flagged only) + k_reduce_status; latency- and issue-bound "
                                            "(ncu: profiles/), not HBM-bound",
                                  "bound": "hbm", "achieved": round(huff_bytes / (huff_ms * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                                  "frac": round(huff_bytes / (huff_ms * 1e-3) / 1e9 / peak, 4), "traffic": k3_traffic() if n == BATCH else None,
                                  "bytes_per_launch": int(huff_bytes), "ms_per_launch": round(huff_ms, 4),
                                  "share_of_step": shares["huffman"]},
            "stages": {"ms": {k: round(statistics.mean(v), 4) for k, v in stage_ms.items()}, "share_of_step": shares,
                       "lexer_GBps": round(3.0 * float(raw_len.sum()) / (statistics.mean(stage_ms["lexer"]) * 1e-3) / 1e9, 1),
                       "huffman_GBps": round(huff_bytes / (huff_ms * 1e-3) / 1e9, 1),
                       "color_GBps": round(color_bytes / (color_ms * 1e-3) / 1e9, 1)},
            "k1k2_fused": fused_side,
            "configs": cfg_out,
            "layer_a": layer_a,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        out.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_sample(ecs_list, tables, q):
    """the oracle (CPU restatement of the reference, 'port') on a bounded sample: 3 frames of the workload, one thread, ~10-20 s"""
    try:
        from jpeg_b200 import batch
        from oracle import oracle as O
        specs = []
        for i, e in enumerate(ecs_list):
            data, lens = batch.unstuff_split(e)
            offs = np.concatenate([[0], np.cumsum(lens)])
            parts = [data[offs[k]:offs[k + 1]].tobytes() for k in range(len(lens))]
            dct = [O.HuffSpec.make(bytes(t.counts), bytes(t.values)) if t.present else O.HuffSpec() for t in tables[8 * i:8 * i + 4]]
            act = [O.HuffSpec.make(bytes(t.counts), bytes(t.values)) if t.present else O.HuffSpec() for t in tables[8 * i + 4:8 * i + 8]]
            specs.append((parts, dct, act))
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 12.0 and reps < 20:
            for parts, dct, act in specs:
                s = O.Spectral.create((W, H), FACTORS)
                for p in range(3):
                    s.set_quanta(p, q[p])
                s.decode_scan((0, 64), (0, None), [0, 1, 2], [0, 1, 1], [0, 1, 1], dct, act, parts, interval=240)
                O.unpack_rgb(s.to_rectangular())
            reps += 1
        dt = time.perf_counter() - t0
        return {"value": round(reps * len(specs) * W * H / dt / 1e6, 2), "unit": "Mpixels/s", "cores": 1, "kind": "port",
                "sample": f"{len(specs)} frames of the workload x {reps} repetitions, single thread ({os.cpu_count()} cores present)"}
    except Exception as e:  # the baseline is reported, never required
        return {"value": None, "unit": "Mpixels/s", "cores": 1, "kind": "port", "sample": f"failed: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--sweep", default="", help="with --quick: K3p settings to re-time, log2(threads per interval):warm-up bits, ...")
    ap.add_argument("--quick", action="store_true", help="device-resident stage times only (kernel A/B runs); not a bench line")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE.json's configurations #3-#5 and the layer-A timings")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
