"""speech2lip_b200 — B200-native (sm_100a) implementation of the Speech2Lip per-frame rendering hot path:
AudioNet + the canonical-space implicit MLP renderer of src/face_simple, behind the reference's own
TalkingFace interface.  See DESIGN.md / INTEGRATION.md."""
from ._cabi import LIB_PATH, PARAM_NAMES  # noqa: F401
from .renderer import (LipRenderer, PackedWeights, audio_encode, audio_windows, frames_to_bgr8, density2outputs, get_rays, mlp_points,  # noqa: F401
                       post_fusion_compose, rgb_forward_rows)
from .staging import NpyPrefetcher  # noqa: F401
from .talking_face import TalkingFace  # noqa: F401

__all__ = ["TalkingFace", "LipRenderer", "PackedWeights", "audio_encode", "rgb_forward_rows", "mlp_points",
           "density2outputs", "get_rays", "post_fusion_compose", "audio_windows", "frames_to_bgr8", "NpyPrefetcher", "LIB_PATH", "PARAM_NAMES"]
