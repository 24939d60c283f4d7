"""mmmot_b200 — B200-native (sm_100a) implementation of mmMOT's per-frame-pair association
forward behind the reference's own Python surface.  See DESIGN.md."""
from .config import build_model, model_kwargs  # noqa: F401
from .solvers import ortools_solve, solve_batch  # noqa: F401
from .tracking_net import TrackingNet  # noqa: F401
