"""N>1 host logic on CPU with the gloo backend (world_size 2): static track partition, partition-invariant
synthetic data, gather of per-track results, MAX-over-ranks timing.  The compute engine stand-in is the CPU
restatement (the CUDA path needs a GPU); what is tested is the multi-process plumbing bench.py uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from roft_b200.partition import max_over_ranks_ms, owner_of, track_range


def test_track_partition_is_a_disjoint_cover():
    for total in (1, 7, 256, 2048):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = track_range(r, world, total)
                assert 0 <= lo <= hi <= total
                seen += list(range(lo, hi))
                for t in (lo, hi - 1):
                    if lo < hi:
                        assert owner_of(t, world, total) == r
            assert seen == list(range(total))
            sizes = [track_range(r, world, total)[1] - track_range(r, world, total)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert track_range(3, 8, 2048) == (768, 1024)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, frames, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cpu_ref
    from helpers import frame_inputs, sequence, small_cfg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = small_cfg(96, 64, subsampling_radius=2.0, segm_delay=2, pose_delay=2)
    lo, hi = track_range(rank, world, total)
    seq = sequence(cfg, hi - lo, frames, first_track_id=lo)
    res = torch.zeros((total, 19), dtype=torch.float64)
    for t in range(hi - lo):
        x0 = np.zeros(13); x0[6:] = seq.pose[0, t].numpy()
        f = cpu_ref.CFilter(cfg, x0)
        for k in range(frames):
            fr = frame_inputs(seq, cfg, k, t)
            f.step(fr.depth, fr.flow, fr.mask, fr.pose, fr.dt)
        pm, _, vm, _, _ = f.state()
        res[lo + t, :13] = torch.from_numpy(pm); res[lo + t, 13:] = torch.from_numpy(vm)
    dist.barrier()
    dist.all_reduce(res)  # disjoint rows: sum == gather (test-only; the data path itself has no collective)
    ms = max_over_ranks_ms(10.0 + rank)
    if rank == 0:
        torch.save({"res": res, "ms": ms}, out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_partition_matches_single_process(tmp_path):
    total, frames = 3, 6
    outs = []
    for world in (1, 2):
        out = str(tmp_path / f"w{world}.pt")
        mp.spawn(_worker, args=(world, _free_port(), total, frames, out), nprocs=world, join=True)
        outs.append(torch.load(out))
    assert torch.equal(outs[0]["res"], outs[1]["res"])  # bit-identical: per-track data and results do not depend on the partition
    assert outs[0]["ms"] == 10.0 and outs[1]["ms"] == 11.0  # MAX over ranks
    assert torch.isfinite(outs[0]["res"]).all() and outs[0]["res"][:, 13:].abs().sum() > 0
