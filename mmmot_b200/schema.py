"""Checkpoint schema of the association hot path.

The drop-in contract includes the 263 ``state_dict`` key names of the reference
``TrackingNet`` (reference: modules/tracking_net.py:17-104, utils/train_util.py:157-188
``load_state`` uses ``load_state_dict(strict=False)``).  This module is the single table
of those names and shapes; the parameter holder (``tracking_net.py``), the weight packer
(``weights.py``) and the synthetic weight generator (``synthetic.py``) are all driven by it.

Key order follows module registration order of the reference so that
``list(state_dict())`` matches too (fusion_module, appearance, point_net, w_link, w_det).
"""
from collections import OrderedDict

# VGG16 (configuration "D", reference: modules/vgg.py:87-90) split into the four stages
# that modules/appear_net.py:130-157 builds: the first max-pool does NOT close a stage.
# Entries: (index inside the stage's Sequential, cin, cout) for convs; BN sits at index+1.
VGG_STAGES = (
    ((0, 3, 64), (3, 64, 64), (7, 64, 128), (10, 128, 128)),   # pools at 6 and 13
    ((0, 128, 256), (3, 256, 256), (6, 256, 256)),             # pool at 9
    ((0, 256, 512), (3, 512, 512), (6, 512, 512)),
    ((0, 512, 512), (3, 512, 512), (6, 512, 512)),
)
# conv index after which a 2x2 max-pool follows, per stage
VGG_POOL_AFTER = ((3, 10), (6,), (6,), (6,))
SKIP_CHANNELS = (128, 256, 512, 512)          # channels of the four skip maps
SKIP_OUT = 128                                # each SkipPool head emits 128 -> concat 512
D = 512


def skip_mid(c):
    """Hidden width of a SkipPool head (reference: modules/appear_net.py:22)."""
    return max(c // 4, 64)


def _conv_bn(sd, prefix, idx, cin, cout, k):
    shape = (cout, cin, k, k) if k else (cout, cin, 1)
    sd[f"{prefix}.{idx}.weight"] = (shape, "conv")
    sd[f"{prefix}.{idx}.bias"] = ((cout,), "bias")
    sd[f"{prefix}.{idx + 1}.weight"] = ((cout,), "norm_w")
    sd[f"{prefix}.{idx + 1}.bias"] = ((cout,), "norm_b")
    sd[f"{prefix}.{idx + 1}.running_mean"] = ((cout,), "run_mean")
    sd[f"{prefix}.{idx + 1}.running_var"] = ((cout,), "run_var")
    sd[f"{prefix}.{idx + 1}.num_batches_tracked"] = ((), "nbt")


def _lin_gn(sd, conv, norm, cin, cout, wshape):
    sd[f"{conv}.weight"] = (wshape, "conv")
    sd[f"{conv}.bias"] = ((cout,), "bias")
    if norm is not None:
        sd[f"{norm}.weight"] = ((cout,), "norm_w")
        sd[f"{norm}.bias"] = ((cout,), "norm_b")


def _stn(sd, p, cin, k):
    sd[f"{p}.idt"] = ((k, k), "eye")
    _lin_gn(sd, f"{p}.conv1", f"{p}.bn1", cin, 64, (64, cin, 1))
    _lin_gn(sd, f"{p}.conv2", f"{p}.bn2", 64, 128, (128, 64, 1))
    _lin_gn(sd, f"{p}.conv3", f"{p}.bn3", 128, 1024, (1024, 128, 1))
    _lin_gn(sd, f"{p}.fc1", f"{p}.fc_bn1", 1024, 512, (512, 1024))
    _lin_gn(sd, f"{p}.fc2", f"{p}.fc_bn2", 512, 256, (256, 512))
    sd[f"{p}.output.weight"] = ((k * k, 256), "stn_out")
    sd[f"{p}.output.bias"] = ((k * k,), "stn_out")


def state_schema(fusion="C", point_in=3):
    """OrderedDict key -> (shape, kind).  ``kind`` drives init / synthetic generation:
    conv, bias, norm_w, norm_b, run_mean, run_var, nbt, eye, stn_out."""
    sd = OrderedDict()
    # --- fusion_module (modules/fusion_net.py) ---
    f = "fusion_module"
    if fusion == "C":
        for g in ("gate_p", "gate_i"):
            _lin_gn(sd, f"{f}.{g}.0", None, D, D, (D, D, 1))
        for g in ("input_p", "input_i"):
            _lin_gn(sd, f"{f}.{g}.0", f"{f}.{g}.1", D, D, (D, D, 1))
    elif fusion == "B":
        for g in ("input_p", "input_i"):
            _lin_gn(sd, f"{f}.{g}.0", f"{f}.{g}.1", D, D, (D, D, 1))
    elif fusion == "A":
        _lin_gn(sd, f"{f}.input_w.0", f"{f}.input_w.1", 2 * D, D, (D, 2 * D, 1))
    else:
        raise ValueError(f"unknown fusion arch {fusion!r}")
    # --- appearance (modules/appear_net.py, modules/vgg.py) ---
    for s, stage in enumerate(VGG_STAGES):
        for idx, cin, cout in stage:
            _conv_bn(sd, f"appearance.layers.{s}", idx, cin, cout, 3)
    for s, c in enumerate(SKIP_CHANNELS):
        p = f"appearance.global_pool.{s}.fc"
        m = skip_mid(c)
        sd[f"{p}.0.weight"] = ((c,), "norm_w")
        sd[f"{p}.0.bias"] = ((c,), "norm_b")
        _lin_gn(sd, f"{p}.1", f"{p}.2", c, m, (m, c, 1, 1))
        _lin_gn(sd, f"{p}.4", f"{p}.5", m, SKIP_OUT, (SKIP_OUT, m, 1, 1))
    # --- point_net (modules/point_net.py) ---
    p = "point_net.feat"
    _stn(sd, f"{p}.stn1", point_in, point_in)
    _lin_gn(sd, f"{p}.conv1", f"{p}.bn1", point_in, 64, (64, point_in, 1))
    _lin_gn(sd, f"{p}.conv2", f"{p}.bn2", 64, 64, (64, 64, 1))
    _stn(sd, f"{p}.stn2", 64, 64)
    _lin_gn(sd, f"{p}.conv3", f"{p}.bn3", 64, 64, (64, 64, 1))
    _lin_gn(sd, f"{p}.conv4", f"{p}.bn4", 64, 128, (128, 64, 1))
    _lin_gn(sd, f"{p}.conv5", f"{p}.bn5", 128, 1024, (1024, 128, 1))
    sd["point_net.conv1.weight"] = ((512, 1088, 1), "conv")
    sd["point_net.conv1.bias"] = ((512,), "bias")
    sd["point_net.conv2.weight"] = ((512, 512, 1), "conv")
    sd["point_net.conv2.bias"] = ((512,), "bias")
    for n in ("bn1", "bn2", "avg_bn"):
        sd[f"point_net.{n}.weight"] = ((512,), "norm_w")
        sd[f"point_net.{n}.bias"] = ((512,), "norm_b")
    # --- w_link (modules/gcn.py, modules/new_end.py) ---
    ne = "w_link.w_new_end"
    _lin_gn(sd, f"{ne}.conv0.0", f"{ne}.conv0.1", D, D, (D, D, 1, 1))
    _lin_gn(sd, f"{ne}.conv1.0", f"{ne}.conv1.1", D, D, (D, D, 1))
    _lin_gn(sd, f"{ne}.conv1.3", f"{ne}.conv1.4", D, 128, (128, D, 1))
    _lin_gn(sd, f"{ne}.conv1.6", None, 128, 1, (1, 128, 1))
    _lin_gn(sd, "w_link.conv1.0", "w_link.conv1.1", D, D, (D, D, 1, 1))
    _lin_gn(sd, "w_link.conv1.3", "w_link.conv1.4", D, D, (D, D, 1, 1))
    _lin_gn(sd, "w_link.conv1.6", "w_link.conv1.7", D, 128, (128, D, 1, 1))
    _lin_gn(sd, "w_link.conv1.9", None, 128, 1, (1, 128, 1, 1))
    # --- w_det (modules/tracking_net.py:91-100) ---
    _conv_bn(sd, "w_det", 0, D, D, 0)
    _conv_bn(sd, "w_det", 3, D, D // 2, 0)
    sd["w_det.6.weight"] = ((1, D // 2, 1), "conv")
    sd["w_det.6.bias"] = ((1,), "bias")
    return sd


BUFFER_KINDS = ("run_mean", "run_var", "nbt")
