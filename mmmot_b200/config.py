"""experiments/*/config.yaml -> TrackingNet, mirroring reference utils/build_util.py:62-83
(``build_model``) without the easydict dependency."""
import yaml

from .tracking_net import TrackingNet


def model_kwargs(common):
    """`common` is the dict under the YAML's top-level ``common:`` key."""
    mdl = common["model"]
    return dict(
        seq_len=common["sample_max_len"],
        score_arch=mdl["score_arch"],
        appear_arch=mdl["appear_arch"],
        appear_len=mdl["appear_len"],
        appear_skippool=mdl["appear_skippool"],
        appear_fpn=mdl["appear_fpn"],
        point_arch=mdl["point_arch"],
        point_len=mdl["point_len"],
        without_reflectivity=common["without_reflectivity"],
        softmax_mode=mdl["softmax_mode"],
        affinity_op=mdl["affinity_op"],
        end_arch=mdl["end_arch"],
        end_mode=mdl["end_mode"],
        test_mode=mdl["test_mode"],
        score_fusion_arch=mdl["score_fusion_arch"],
        neg_threshold=mdl["neg_threshold"],
        dropblock=common["dropblock"],
        use_dropout=common["use_dropout"],
    )


def build_model(config):
    """Accepts a path to a config.yaml, the parsed YAML dict, or its ``common`` sub-dict."""
    if isinstance(config, str):
        with open(config) as f:
            config = yaml.safe_load(f)
    common = config.get("common", config)
    return TrackingNet(**model_kwargs(common))
