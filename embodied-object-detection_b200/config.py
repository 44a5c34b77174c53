"""The MODEL.* memory keys of detic/config.py:56-74 with the reference's names and defaults, as a plain
nested namespace (yacs/detectron2 are not required; a detectron2 CfgNode with the same keys is accepted by
``from_cfg`` helpers because only attribute access is used)."""
from __future__ import annotations

from types import SimpleNamespace

from .fpn_fusion import MemoryFusion
from .memory import SpatialFeatureMemory

_DEFAULTS = dict(
    MAP_MERGE_TYPE="",            # config.py:57
    MAP_FEAT_FUSION="",           # config.py:58  sum | mem_only | image_only
    MEMORY_FEATURE_WEIGHT=100,    # config.py:62  (stored, unused by forward)
    TEST_SAVE_SEMMAP=False,       # config.py:65
    SEMMAP_PATH="",               # config.py:66
    MEMORY_TYPE="",               # config.py:67  image_only | implicit_memory
    MEMORY_CLS_SCORE_THRESH=0.3,  # config.py:68
    MEMORY_OBS_SCORE_THRESH=0.4,  # config.py:69
    MAP_FEATURE_WEIGHT=500,       # config.py:70
    TEST_DATA_PATH="", TRAIN_DATA_PATH="", MEMORY_PATH="",   # config.py:71-73
    TEST_TYPE="default",          # config.py:74  default | episodic | longterm
    DEVICE="cuda",
)


def add_detic_memory_config(cfg=None):
    """Add the memory keys (reference defaults) to ``cfg.MODEL``; creates a namespace when cfg is None."""
    if cfg is None:
        cfg = SimpleNamespace(MODEL=SimpleNamespace())
    for k, v in _DEFAULTS.items():
        if not hasattr(cfg.MODEL, k):
            setattr(cfg.MODEL, k, v)
    return cfg


def merge_from_list(cfg, opts):
    """``KEY VALUE`` overrides like the reference CLI (train_mp3d.py:813-822), e.g.
    ['MODEL.MEMORY_TYPE', 'implicit_memory', 'MODEL.MAP_FEAT_FUSION', 'sum', 'MODEL.MAP_FEATURE_WEIGHT', '5']."""
    for key, val in zip(opts[0::2], opts[1::2]):
        node = cfg
        *path, leaf = key.split(".")
        for p in path:
            node = getattr(node, p)
        old = getattr(node, leaf, None)
        if isinstance(old, bool):
            val = str(val).lower() in ("1", "true", "yes")
        elif isinstance(old, (int, float)) and not isinstance(val, (int, float)):
            val = float(val) if ("." in str(val) or isinstance(old, float)) else int(val)
        setattr(node, leaf, val)
    return cfg


def build_memory_fusion(cfg) -> MemoryFusion:
    m = cfg.MODEL
    if m.MEMORY_TYPE == "implicit_memory" and m.MAP_FEAT_FUSION not in ("sum", "mem_only", "image_only"):
        raise ValueError(f"MODEL.MAP_FEAT_FUSION={m.MAP_FEAT_FUSION!r}: the reference defines sum | mem_only | image_only")
    return MemoryFusion(memory_type=m.MEMORY_TYPE, fusion=m.MAP_FEAT_FUSION, map_feature_weight=float(m.MAP_FEATURE_WEIGHT),
                        memory_feature_weight=float(m.MEMORY_FEATURE_WEIGHT), merge_type=m.MAP_MERGE_TYPE)


def build_spatial_memory(cfg, **kw) -> SpatialFeatureMemory:
    return SpatialFeatureMemory(device=cfg.MODEL.DEVICE, test_type=cfg.MODEL.TEST_TYPE, **kw)
