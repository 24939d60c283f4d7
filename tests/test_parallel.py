"""CPU: the N>1 host logic (sharding + final gather) with world_size-2 and -3 gloo groups."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmmot_b200.parallel import gather_pairs, shard_range


def test_shard_ranges_cover_batch_exactly():
    for total in (1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [h - l for l, h in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    # stand-in for this rank's per-pair assignment indices: a deterministic function of the pair id
    local = (torch.arange(lo, hi).reshape(-1, 1) * 10 + torch.arange(4).reshape(1, -1)).to(torch.int32)
    full = gather_pairs(local, total)
    torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 8), (2, 7), (3, 10)])
def test_gather_equals_single_process_result(tmp_path, world, total):
    port = 29500 + world * 10 + total
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    expect = (torch.arange(total).reshape(-1, 1) * 10 + torch.arange(4).reshape(1, -1)).to(torch.int32)
    for r in range(world):
        assert torch.equal(torch.load(os.path.join(tmp_path, f"r{r}.pt")), expect)
