"""world_size-2 gloo tests (CPU) of the multi-process host logic: weight broadcast from rank 0,
batch sharding, and that the shards of a CPU sampling reproduce the single-process result."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.gen_golden_cfg import TINY_ADM


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from azula_b200 import parallel
        from azula_b200.plugins import adm
        from azula_b200.sample import DDIMSampler

        torch.manual_seed(100 + rank)  # different initial weights on every rank
        den = adm.make_model(**TINY_ADM).eval()
        adm.seed_parameters(den.backbone, seed=7 + rank)
        sent = parallel.broadcast_parameters(den.backbone, src=0)
        digest = torch.stack([p.double().sum() for p in den.backbone.parameters()]).sum().item()

        g = torch.Generator().manual_seed(3)
        x1 = torch.randn(4, 3, 16, 16, generator=g)  # the same global batch on every rank
        mine = parallel.shard_of(x1, rank, world)
        with torch.no_grad():
            x0 = DDIMSampler(den, steps=3, silent=True)(mine)  # eta = 0: deterministic
        gathered = [torch.empty_like(x0) for _ in range(world)]
        dist.all_gather(gathered, x0)
        if rank == 0:
            with torch.no_grad():
                full = DDIMSampler(den, steps=3, silent=True)(x1)
            out.put({"sent": sent, "equal": bool(torch.equal(torch.cat(gathered), full)), "digest": digest})
        else:
            out.put({"digest": digest})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_broadcast_and_sharded_sampling_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    digests = {round(r["digest"], 6) for r in results}
    assert len(digests) == 1, "ranks hold different weights after the broadcast"
    root = next(r for r in results if "equal" in r)
    assert root["sent"] > 0
    assert root["equal"], "concatenated shards differ from the single-process sampling"


def test_shard_range():
    from azula_b200 import parallel

    assert list(parallel.shard_range(8, 1, 4)) == [2, 3]
    with pytest.raises(ValueError):
        parallel.shard_range(10, 0, 4)
