"""ORACLE helper — imports the UNMODIFIED reference modules from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
oracle/make_goldens.py to generate tests/golden/ and by tests that are skipped when the
reference tree is absent.  Nothing is copied from the reference; it is imported in place.
"""
import contextlib
import io
import os
import sys

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "modules"))


def load_tracking_net(**kw):
    """Build reference modules.TrackingNet(**kw).eval() (reference: modules/tracking_net.py:17).

    Applies the compat shim from SURVEY F3: modern F.group_norm raises on one value per group
    where reference-era torch returned beta."""
    import torch.nn.functional as F
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    F._verify_batch_size = lambda size: None
    with contextlib.redirect_stdout(io.StringIO()):
        from modules import TrackingNet
        net = TrackingNet(**kw)
    return net.eval()
