"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch of images, decode their shards with the oracle
(standing in for the per-rank GPU context), and the results come back in input order; timing is the max over ranks."""
import hashlib
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, files, out):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from jpeg_b200 import shard
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.shard_indices(len(files), rank, world)
    assert all(shard.owner_of(i, world) == rank for i in mine)
    local = {}
    for i in mine:
        rgb, _, _ = O.decode_rgb(open(os.path.join(GOLDEN, files[i]), "rb").read())
        local[i] = hashlib.sha256(rgb.tobytes()).hexdigest()
    dist.barrier()
    ordered = shard.gather_in_input_order(local, len(files))
    seen = []

    def work(i, rel):
        seen.append(i)
        return (rank, local[i])
    mapped = shard.map_sharded(files, work)
    assert seen == mine and [m[1] for m in mapped] == ordered and [m[0] for m in mapped] == [i % world for i in range(len(files))]
    slowest = shard.max_over_ranks(1.0 + rank)
    if rank == 0:
        out.put((ordered, slowest))
    dist.destroy_process_group()


def test_two_rank_sharding(manifest):
    files = [v["jpeg"] for v in manifest["decode"][:5]]
    want = [v["rgb_sha256"] for v in manifest["decode"][:5]]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, files, out)) for r in range(2)]
    for p in procs:
        p.start()
    ordered, slowest = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ordered == want
    assert slowest == 2.0


def test_shard_partition_properties():
    from jpeg_b200 import shard
    for n in (0, 1, 7, 64, 512):
        for world in (1, 2, 4, 8):
            parts = [shard.shard_indices(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_numa_placement_plan(tmp_path):
    """jpeg_b200/affinity.py: a rank binds to the CPUs of its GPU's NUMA node, and only when that is a proper, non-empty subset
    of what the process may use; a missing or negative numa_node means no change."""
    from jpeg_b200 import affinity
    assert affinity.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert affinity.parse_cpulist("") == set()
    assert affinity.pci_address(0, 0x1b, 0) == "0000:1b:00.0"
    sysfs = tmp_path
    for pci, node in (("0000:1b:00.0", "0"), ("0000:9d:00.0", "1"), ("0000:aa:00.0", "-1")):
        d = sysfs / "bus" / "pci" / "devices" / pci
        d.mkdir(parents=True)
        (d / "numa_node").write_text(node + "\n")
    for node, cpus in ((0, "0-11,24-35"), (1, "12-23,36-47")):
        d = sysfs / "devices" / "system" / "node" / f"node{node}"
        d.mkdir(parents=True)
        (d / "cpulist").write_text(cpus + "\n")
    everything = set(range(48))
    assert affinity.plan("0000:1b:00.0", everything, str(sysfs)) == (0, set(range(12)) | set(range(24, 36)))
    assert affinity.plan("0000:9d:00.0", everything, str(sysfs)) == (1, set(range(12, 24)) | set(range(36, 48)))
    assert affinity.plan("0000:aa:00.0", everything, str(sysfs)) == (None, None)      # the kernel does not know
    assert affinity.plan("0000:ff:00.0", everything, str(sysfs)) == (None, None)      # no such device
    assert affinity.plan("0000:9d:00.0", set(range(12)), str(sysfs)) == (1, None)      # cpuset excludes the node: unchanged
    assert affinity.plan("0000:1b:00.0", set(range(12)), str(sysfs)) == (0, None)      # already confined to it
    assert affinity.bind_to_gpu(0, str(sysfs)).startswith("unchanged")                   # no CUDA device here: a hint never raises
