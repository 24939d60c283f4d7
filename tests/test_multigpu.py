"""GPU, >= 2 devices: sharded frame-pairs + final gather must equal the 1-GPU result bit for bit
(pairs are independent and the GroupNorm statistics are reduced in a fixed order)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import mmmot_b200
    from mmmot_b200.parallel import gather_pairs, shard_range
    from mmmot_b200.synthetic import synthetic_batch, synthetic_state_dict
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    B, n = 5, 16
    net = mmmot_b200.TrackingNet(2, appear_skippool=True, score_arch="branch_cls", score_fusion_arch="C",
                                 affinity_op="minus_abs", softmax_mode="dual_add", neg_threshold=0.2, test_mode=2, dropblock=0)
    net.load_state_dict(synthetic_state_dict("C", 3))
    net.cuda(rank).eval()
    crops, pts, split = synthetic_batch(B, n, pts=32, hw=32, seed=90)
    L = 2 * n
    lo, hi = shard_range(B, rank, world)
    s = split[lo * L:hi * L + 1]
    o = net.predict_batch(crops[lo * L:hi * L].cuda(rank), pts[int(s[0]):int(s[-1])].cuda(rank), s - s[0], n)
    match = gather_pairs(o["match"], B)
    link = gather_pairs(o["link"][:, 2].contiguous(), B)
    if rank == 0:
        full = net.predict_batch(crops.cuda(0), pts.cuda(0), split, n)
        torch.save({"same_match": torch.equal(match, full["match"]), "same_link": torch.equal(link, full["link"][:, 2])},
                   os.path.join(out_dir, "res.pt"))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_equals_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, 29611, str(tmp_path)), nprocs=2, join=True)
    r = torch.load(os.path.join(tmp_path, "res.pt"))
    assert r["same_match"] and r["same_link"]
