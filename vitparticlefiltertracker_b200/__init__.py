"""B200-native ViT particle-filter tracking hot path (sm_100a CUDA behind a C ABI).

Public surface: ``Tracker.init(frame, bbox)`` / ``Tracker.step(frame) -> State``.
"""
from .tracker import State, Tracker  # noqa: F401
from .multistream import MultiStreamTracker  # noqa: F401
from .ingest import FrameIngest, track  # noqa: F401
from . import sources  # noqa: F401
from .vit_weights import (VIT_CONFIGS, flops_per_crop, init_state_dict, load_checkpoint,  # noqa: F401
                          pack_for_device, pack_ln_fold)

__all__ = ["Tracker", "MultiStreamTracker", "FrameIngest", "track", "State", "VIT_CONFIGS", "flops_per_crop", "init_state_dict",
           "load_checkpoint", "pack_for_device", "pack_ln_fold"]
