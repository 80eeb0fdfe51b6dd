"""CPU, world_size 2 over gloo: the multi-GPU path's host logic - contiguous mesh shards per rank, rank-local
tokenization (the oracle stands in for the device pipeline here), histogram all-reduce - gives the same
codes and the same codebook-usage histogram as one process over the whole collection."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import SMALL_CFG, make_oracle
import synth
from meshgpt_pytorch_b200 import sharding
from oracle import meshtok_oracle as O

N_MESHES = 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _collection():
    return synth.grid_batch("tri", 4, 3, list(range(N_MESHES)))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    o = make_oracle(**SMALL_CFG)
    v, f = _collection()
    lo, hi = sharding.shard_range(N_MESHES, rank, world)
    codes = o.tokenize(v[lo:hi], f[lo:hi])
    hist = O.codebook_usage_histogram(codes, o.num_quantizers, o.codebook_size)
    sharding.allreduce_histogram(hist)
    torch.save(dict(lo=lo, hi=hi, codes=codes, hist=hist), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_tokenization_and_histogram_allreduce(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    o = make_oracle(**SMALL_CFG)
    v, f = _collection()
    full = o.tokenize(v, f)
    assert [(p["lo"], p["hi"]) for p in parts] == [(0, 3), (3, 5)]
    assert torch.equal(torch.cat([p["codes"] for p in parts]), full)          # meshes are independent
    want = O.codebook_usage_histogram(full, o.num_quantizers, o.codebook_size)
    assert torch.equal(parts[0]["hist"], want) and torch.equal(parts[1]["hist"], want)
    assert int(want.sum()) == full.numel()
