"""world_size-2 gloo test of the N>1 path's host logic (CPU): contiguous read sharding and
read-order reassembly, with the oracle standing in for the per-GPU engine.  The reassembled
text must be byte-identical to the reference rb_align's stdout on the whole file."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, read_fastx
from rowbowt_b200.shard import run_sharded, shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 2, 5, 97, 1000):
        for w in (1, 2, 3, 8):
            blocks = [shard_bounds(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    prefix = os.path.join(GOLDEN, "tiny", "tiny")
    names, seqs = read_fastx(os.path.join(GOLDEN, "tiny", "noisy.fq"))
    ix = O.OracleIndex.open(prefix, sa=True, markers=True)      # every rank holds a full replica
    items = list(zip(names, seqs))

    def engine(block):
        return [ix.report([n], [s], sa=True, markers=True) for n, s in block]

    out = run_sharded(items, engine, rank, world)
    dist.barrier()
    if rank == 0:
        q.put("".join(out))
    else:
        assert out is None
    dist.destroy_process_group()


def test_two_rank_sharded_output_is_in_read_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    text = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp = open(os.path.join(GOLDEN, "expected", "tiny.noisy.fq.sm.txt")).read()
    assert text == exp
