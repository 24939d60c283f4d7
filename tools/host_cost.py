"""Host-side cost of enqueueing one bench step (cfg4): wall time of predict_batch with an EMPTY stream and no
synchronisation inside, against the device time of the same step.  If the two are close the step is launch-bound."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmmot_b200
from mmmot_b200.synthetic import synthetic_state_dict

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n, pts, hw = 128, 512, 64
L = 2 * n
dev = torch.device("cuda", 0)
net = mmmot_b200.TrackingNet(2, appear_skippool=True, score_arch="branch_cls", score_fusion_arch="C", affinity_op="minus_abs",
                             softmax_mode="dual_add", neg_threshold=0.2, test_mode=2, dropblock=0)
net.load_state_dict(synthetic_state_dict("C", seed=0))
net.cuda(dev).eval()
g = torch.Generator(device=dev).manual_seed(1)
crops = torch.randn(pairs * L, 3, hw, hw, device=dev, generator=g)
points = torch.randn(pairs * L * pts, 3, device=dev, generator=g)
split = torch.arange(0, pairs * L * pts + 1, pts, dtype=torch.int32)
for _ in range(3):
    net.predict_batch(crops, points, split, n, check=False)
torch.cuda.synchronize()
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    net.predict_batch(crops, points, split, n, check=False)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"pairs={pairs} host enqueue {1e3 * (t1 - t0):7.2f} ms   device {e0.elapsed_time(e1):7.2f} ms")
