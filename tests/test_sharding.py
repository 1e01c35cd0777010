"""Multi-GPU path on CPU: two `gloo` ranks shard streams the way bench.py does (no data-path
collective), each decodes its shard through the emulated test build, and the union of the shards is
checked against a single-process decode of all streams."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_file

sys.path.insert(0, ROOT)
import bench  # noqa: E402

NAMES = ["1test", "2test"]   # small files: the emulator runs every CUDA thread as an OS thread
TOTAL = 6


def _decode_shard(lib_path, first, count):
    from vorbispizza_b200 import Context, decode_files
    files = [load_file(n) for n in NAMES]
    with Context(0, lib_path=lib_path) as ctx:
        ctx.set("host_threads", 2)
        ctx.set("bulk_group", 2)     # several pipeline groups even for a tiny shard
        datas = [files[(first + i) % len(files)] for i in range(count)]
        pcm, counts = decode_files(ctx, datas, clip=True)
    out, off = [], 0
    for i, c in enumerate(counts):
        n = int(c)  # mono files
        out.append((first + i, n, hashlib.sha256(pcm[off:off + n].tobytes()).hexdigest()))
        off += n
    return out, int(pcm.size)


def _worker(rank, world, port, lib_path, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = bench.stream_assignment(TOTAL, rank, world, "strong")
    res, samples = _decode_shard(lib_path, first, count)
    # the only cross-rank traffic: the barrier and the reductions of the timing / totals
    t = torch.tensor([float(samples)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
    q.put((rank, res, float(t.item()), float(ms.item())))
    dist.destroy_process_group()


def test_assignment_is_a_partition():
    for world in (1, 2, 3, 4, 8):
        for total in (1, 7, 8, 4096):
            seen = []
            for r in range(world):
                f, c = bench.stream_assignment(total, r, world, "strong")
                seen += list(range(f, f + c))
            assert seen == list(range(total))
        f, c = bench.stream_assignment(4096, 3, world, "weak") if world > 3 else (3 * 4096, 4096)
        assert (f, c) == (3 * 4096, 4096)
    assert bench.stream_file_and_start(9, 4, [25, 310, 366, 606]) == (1, 14)


def test_two_ranks_equal_one(emu_lib_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, emu_lib_path, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    whole, total = _decode_shard(emu_lib_path, 0, TOTAL)
    shards = sorted(sum((g[1] for g in got), []))
    assert shards == sorted(whole)
    for g in got:
        assert g[2] == float(total)     # SUM over ranks of the samples = the whole job
        assert g[3] == 11.0             # MAX over ranks of the step time


# ---- BASELINE config 5 across ranks: the excerpt list shards like the streams do ---------------------------
def _excerpt_shard(lib_path, rank, world, n):
    from vorbispizza_b200 import Context, decode_excerpts
    file_of, start, count, _ = bench.config5_excerpts(rank, world, "strong", n=n)
    count = np.minimum(count, 600).astype(np.int32)   # short reads: the emulator runs every CUDA thread as an OS thread
    # config 5 draws from {2test, 3test, issue6test}; the emulated run maps them onto the two small files
    files = [load_file(n_) for n_ in NAMES]
    totals = [17318, 315790]
    file_of = (file_of % 2).astype(np.uint32)
    start = np.array([int(s_) % (totals[int(f)] - 700) for s_, f in zip(start, file_of)], np.int64)
    with Context(0, lib_path=lib_path) as ctx:
        ctx.set("host_threads", 2)
        pcm, offsets, got = decode_excerpts(ctx, files, file_of, start, count, clip=True)
    out = []
    for i in range(file_of.size):
        n_fl = int(got[i])  # mono files
        out.append((int(file_of[i]), int(start[i]), n_fl, hashlib.sha256(pcm[int(offsets[i]):int(offsets[i]) + n_fl].tobytes()).hexdigest()))
    return out


def _excerpt_worker(rank, world, port, lib_path, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = _excerpt_shard(lib_path, rank, world, n)
    t = torch.tensor([float(sum(r[2] for r in res))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.barrier()
    q.put((rank, res, float(t.item())))
    dist.destroy_process_group()


def test_excerpt_list_is_partitioned():
    whole = bench.config5_excerpts(0, 1, "strong", n=101)
    for world in (2, 3, 8):
        parts = [bench.config5_excerpts(r, world, "strong", n=101) for r in range(world)]
        for k in range(3):
            assert np.array_equal(np.concatenate([p[k] for p in parts]), whole[k])
    a, b = bench.config5_excerpts(0, 2, "weak", n=16), bench.config5_excerpts(1, 2, "weak", n=16)
    assert a[0].size == b[0].size == 16 and not np.array_equal(a[1], b[1])


def test_two_ranks_excerpts_equal_one(emu_lib_path):
    n = 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_excerpt_worker, args=(r, 2, port, emu_lib_path, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=900) for _ in procs])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    whole = _excerpt_shard(emu_lib_path, 0, 1, n)
    assert got[0][1] + got[1][1] == whole          # rank order = list order: the shards concatenate to the job
    assert got[0][2] == got[1][2] == float(sum(r[2] for r in whole))
