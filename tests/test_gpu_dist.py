"""GPU, world size 2 over NCCL (skipped with fewer than 2 GPUs): the PRODUCT per rank - contiguous Sigma-faces-balanced shards of a
ragged collection, rank-local tokenization in the ragged layout, device-side histogram, ONE NCCL all-reduce - gives the same tokens
and the same codebook-usage histogram as one process over the whole collection, and those tokens agree with the oracle.
(tests/test_dist_gloo.py covers the host logic of the same path on CPU with the oracle standing in for the device.)"""
import os
import socket
import subprocess
import sys

import pytest
import torch

from helpers import assert_tokens_match_or_attributed, make_pair
from oracle import meshtok_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_ranks_nccl_product_path(tmp_path):
    from dist_worker import collection
    from meshgpt_pytorch_b200 import data as Dt
    from meshgpt_pytorch_b200 import sharding
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    parts = [torch.load(tmp_path / f"rank{k}.pt") for k in range(world)]
    v, f, nf = collection()
    cuts = sharding.balanced_contiguous_shards(nf, world)
    assert [(p["lo"], p["hi"]) for p in parts] == cuts and cuts[0][1] == cuts[1][0] and cuts[1][1] == len(nf)
    faces_per_rank = [sum(nf[lo:hi]) for lo, hi in cuts]
    assert max(faces_per_rank) <= 1.25 * min(faces_per_rank), faces_per_rank     # balanced by faces, not by mesh count
    # one process over the whole collection
    o, m = make_pair(use_residual_lfq=False)
    hist = torch.zeros(m.num_quantizers, m.codebook_size, dtype=torch.int64, device="cuda")
    full = m.forward(vertices=v.cuda(), faces=f.cuda(), return_codes=True, histogram=hist)
    for p in parts:
        padded = Dt.ragged_to_padded_codes(p["codes"], p["face_offsets"], 3, m.pad_id)
        want = full[p["lo"]:p["hi"]].cpu()
        assert torch.equal(padded, want[:, : padded.shape[1]]) and bool((want[:, padded.shape[1]:] == m.pad_id).all())
        assert torch.equal(p["hist"], hist.cpu())                                  # every rank holds the job's histogram
    assert torch.equal(parts[0]["local_hist"] + parts[1]["local_hist"], hist.cpu())
    assert torch.equal(hist.cpu(), O.codebook_usage_histogram(full.cpu(), m.num_quantizers, m.codebook_size))
    assert_tokens_match_or_attributed(o, m, v, f, full, label="2-rank collection")
