"""Worker of tests/test_gpu_dist.py (launched by torch.distributed.run, one process per GPU, NCCL): tokenizes this rank's contiguous
shard of a ragged collection with the PRODUCT (libmeshtok through meshgpt_pytorch_b200), accumulates the codebook-usage histogram on
the device and all-reduces it over NCCL.  Writes its codes + the reduced histogram to <out_dir>/rank<r>.pt."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def collection():
    """16 meshes, 100 .. 800 faces: the whole collection in the padded layout + per-mesh face counts."""
    import synth
    g = torch.Generator().manual_seed(7)
    nf = torch.randint(100, 801, (16,), generator=g).tolist()
    v, f = synth.ragged_batch("tri", 20, 20, nf, range(16))
    return v, f, nf


def main():
    out_dir = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from helpers import make_pair
    from meshgpt_pytorch_b200 import data as Dt
    from meshgpt_pytorch_b200 import sharding
    o, m = make_pair(use_residual_lfq=False)                      # full-size RVQ model, same seed on every rank
    v, f, nf = collection()
    lo, hi = sharding.balanced_contiguous_shards(nf, world)[rank]
    hist = torch.zeros(m.num_quantizers, m.codebook_size, dtype=torch.int64, device="cuda")
    # the shard in the ragged layout (what a sharded loader hands over), tokens back to back
    r = Dt.padded_to_ragged(v[lo:hi], f[lo:hi])
    flat = m.tokenize_ragged(r["vertices"].cuda(), r["faces"].cuda(), r["face_offsets"], r["vertex_offsets"], histogram=hist)
    m.check_last_status()
    local_hist = hist.clone()
    sharding.allreduce_histogram(hist)                             # NCCL over NVLink: the path's one collective
    torch.save(dict(lo=lo, hi=hi, face_offsets=r["face_offsets"], codes=flat.cpu(), hist=hist.cpu(), local_hist=local_hist.cpu()),
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
