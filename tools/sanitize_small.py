"""Small forward + LP for compute-sanitizer runs (memcheck / racecheck): N=M=8, 32x32 crops, tcgen05 engine forced so the
tensor-core kernels (incl. the first layer's in-kernel operand producers) are the ones checked.
  compute-sanitizer --tool memcheck  python tools/sanitize_small.py
  compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmmot_b200
from mmmot_b200.synthetic import synthetic_batch, synthetic_state_dict

mmmot_b200.set_engine("tcgen05")
net = mmmot_b200.TrackingNet(2, appear_skippool=True, score_arch="branch_cls", score_fusion_arch="C", affinity_op="minus_abs",
                             softmax_mode="dual_add", neg_threshold=0.2, test_mode=2, dropblock=0)
net.load_state_dict(synthetic_state_dict("C", seed=3))
net.cuda().eval()
crops, pts, split = synthetic_batch(2, 8, pts=24, hw=32, seed=5)
out = net.predict_batch(crops.cuda(), pts.cuda(), split, 8)
torch.cuda.synchronize()
print("match", out["match"].cpu().tolist(), "status", int(out["status"]))
