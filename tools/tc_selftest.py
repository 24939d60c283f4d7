"""tcgen05 engine self-test through the C ABI: single contractions vs fp64 matmul."""
import ctypes, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmmot_b200 import _lib
from mmmot_b200.weights import pack_tc

lib = _lib.load()
lib.mmmot_set_debug(int(os.environ.get('TC_DBG', '0')))
dev = torch.device("cuda")
vp = lambda t: ctypes.c_void_p(t.data_ptr())
g = torch.Generator().manual_seed(0)
SHAPES = ((128, 32, 256), (128, 64, 256), (256, 32, 256), (256, 96, 512), (512, 512, 4096), (64, 70, 300),
                  (1024, 128, 1000), (128, 4608, 2048), (512, 512, 148 * 256 * 2 + 77))
if os.environ.get('TC_BIG'):
    SHAPES = ((512, 512, 148 * 256 * 4),)
for (M, K, S) in SHAPES:
    Wt = torch.randn(K, M, generator=g)
    X = torch.randn(K, S, generator=g)
    b = torch.randn(M, generator=g)
    ref = (Wt.double().t() @ X.double()) + b.double()[:, None]
    Wt_d, X_d, b_d = Wt.to(dev), X.to(dev), b.to(dev)
    Wp, wps = pack_tc(Wt); Wp = Wp.to(dev)
    out = {}
    for eng in (1, 2):
        Y = torch.full((M, S), float("nan"), device=dev)
        rc = lib.mmmot_debug_linear(vp(Wt_d), vp(Wp), wps, vp(b_d), vp(X_d), vp(Y), M, K, S, eng, None)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(3):
            lib.mmmot_debug_linear(vp(Wt_d), vp(Wp), wps, vp(b_d), vp(X_d), vp(Y), M, K, S, eng, None)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / 3
        err = float((Y.double().cpu() - ref).abs().max() / ref.abs().max())
        out[eng] = (rc, err, 2.0 * M * K * S / dt / 1e12)
    print(f"M={M} K={K} S={S}: simt rc={out[1][0]} err={out[1][1]:.2e} {out[1][2]:.1f} TF/s | tc rc={out[2][0]} err={out[2][1]:.2e} {out[2][2]:.1f} TF/s", flush=True)
