"""CPU tests (gloo, world_size 2) of the frame-sharded exchange: block sharding, the single all-gather of
per-frame feature records, and the predecessor-frame hand-off across the block boundary."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    from pilotguru_b200.dist import shard_range
    for n, w in [(10000, 8), (10, 3), (7, 8), (1, 1), (0, 4)]:
        blocks = [shard_range(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
            assert a1 == b0 and a0 <= a1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pilotguru_b200.dist import FeatureExchange, KP_BYTES
        B, cap = 3, 16
        x = FeatureExchange(world, rank, B, cap, device=torch.device("cpu"))
        rng = np.random.default_rng(100 + rank)
        # fill slots 1..B with recognisable records (what the extract kernels would write)
        counts = x.counts_view(); kps = x.kps_view(); desc = x.desc_view()
        for s in range(1, B + 1):
            counts[s] = 5 + rank + s
            kps[s] = torch.from_numpy(rng.normal(size=(cap, 7)).astype(np.float32))
            desc[s] = torch.from_numpy(rng.integers(0, 256, (cap, 32), dtype=np.uint8))
        counts[0] = 99                                            # rank 0 keeps its own predecessor
        mine_last = (int(counts[B]), kps[B].clone(), desc[B].clone())
        x.exchange()
        # every rank now holds every rank's block
        for r in range(world):
            cnt, kb, db = x.frame_record(r, B)
            assert cnt == 5 + r + B
        if rank == 0:
            assert int(counts[0]) == 99
        else:
            lc, lk, ld = x.frame_record(rank - 1, B)
            assert int(counts[0]) == lc == 5 + (rank - 1) + B
            assert torch.equal(x.kps_view()[0].view(torch.uint8).reshape(-1)[:lc * KP_BYTES], lk)
            assert torch.equal(x.desc_view()[0][:lc], ld)
        # own slots untouched by the exchange
        assert int(counts[B]) == mine_last[0] and torch.equal(kps[B], mine_last[1]) and torch.equal(desc[B], mine_last[2])
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_exchange_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
