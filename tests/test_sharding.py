"""Multi-GPU path on CPU: the batch shards by image index across ranks with no data-path collective. Two gloo
ranks each take their shard, 'decode' it with the oracle (the host logic under test is the sharding and the
result gathering, not the kernels), and the gathered results must cover the batch exactly once, in order."""
import hashlib
import os
import socket
import sys

import pytest

from conftest import ROOT, GOLDEN


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import json
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import libwebp_b200 as W
    from oracle import portwebp as P
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names = [e["file"] for e in json.load(open(os.path.join(GOLDEN, "manifest.json")))]
    datas = [open(os.path.join(GOLDEN, n), "rb").read() for n in names]
    idx = W.shard_indices(len(datas), rank, world)
    mine = []
    for i in idx:
        st, px = P.decode(datas[i], P.RGBA, 0)
        mine.append((i, st, hashlib.sha256(px.tobytes()).hexdigest() if st == 0 else None))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)     # results only (a status array + digests), never pixels
    dist.barrier()
    if rank == 0:
        q.put(gathered)
    dist.destroy_process_group()


def test_two_ranks_cover_the_batch_once(manifest):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    flat = sorted(x for part in gathered for x in part)
    assert [i for i, _, _ in flat] == list(range(len(manifest)))          # every image exactly once
    assert [i for i, _, _ in gathered[0]] == list(range(0, len(manifest), 2))
    assert [i for i, _, _ in gathered[1]] == list(range(1, len(manifest), 2))
    for (i, st, digest), e in zip(flat, manifest):
        assert st == 0 and digest == e["sha256"]["1:0"], e["file"]


def test_shard_indices_partition():
    import libwebp_b200 as W
    for n in (0, 1, 7, 8, 4096):
        for world in (1, 2, 4, 8):
            parts = [W.shard_indices(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        W.shard_indices(4, 4, 4)
