"""Drop-in twin of the reference's `roi` module (roi.py:5-30): where each region of interest sits relative
to a detection.  Names and values are the reference's (drawer.py reads roi.SELECTED_ROI_CONFIGS at import).

An ROIConfig says: take the landmarks `landmark_indices` of the largest `model_type` detection, average them
into an anchor point, and span the box anchor + relative_bbox * (detection bbox size), relative_bbox being
(left, top, right, bottom) fractions.  A Location is the 6-vector (x, y, x0, y0, x1, y1) — anchor plus box
corners — of ints, or of NaNs when there is no detection (signal_processor.py:142-154).
"""
import dataclasses

import numpy as np

import model

type Location = tuple[int, int, int, int, int, int] | tuple[float, float, float, float, float, float] | np.ndarray[float]


@dataclasses.dataclass(eq=False)
class ROIConfig:
    model_type: model.ModelType
    landmark_indices: list[int]
    relative_bbox: tuple[float, float, float, float]


# landmark indices of the MediaPipe topologies the reference anchors on (roi.py:16-22), exported as module constants
_LANDMARK = {
    'FACE_DETECTION': {'NOSE': 2},
    'FACE_LANDMARKS': {'NOSE': 4, 'FOREHEAD': 151, 'CHEEK': 330, 'EYEBROW': 337},
    'HAND_LANDMARKS': {'WRIST': 0, 'MIDDLE': 9},
}
globals().update({f'{topo}_{part}_INDEX': idx for topo, parts in _LANDMARK.items() for part, idx in parts.items()})


def _config(detector: model.ModelType, topo: str, parts: list[str], box: tuple[float, float, float, float]) -> ROIConfig:
    return ROIConfig(detector, [_LANDMARK[topo][p] for p in parts], box)


_FACE, _HAND = model.ModelType.FACE_LANDMARKER, model.ModelType.HAND_LANDMARKER
# (left, top, right, bottom) fractions of the detection bbox around the anchor (roi.py:24-28)
FACE_CHEEK_CONFIG = _config(_FACE, 'FACE_LANDMARKS', ['CHEEK'], (-0.05, -0.05, 0.15, 0.05))
FACE_EYEBROW_CONFIG = _config(_FACE, 'FACE_LANDMARKS', ['EYEBROW'], (-0.10, -0.15, 0.25, 0.00))
FACE_FOREHEAD_CONFIG = _config(_FACE, 'FACE_LANDMARKS', ['FOREHEAD'], (-0.00, -0.10, 0.20, 0.05))
HAND_WRIST_CONFIG = _config(_HAND, 'HAND_LANDMARKS', ['WRIST'], (-0.10, -0.10, 0.10, 0.10))
HAND_PALM_CONFIG = _config(_HAND, 'HAND_LANDMARKS', ['WRIST', 'MIDDLE'], (-0.10, -0.10, 0.10, 0.10))

SELECTED_ROI_CONFIGS = [FACE_FOREHEAD_CONFIG, HAND_PALM_CONFIG]   # roi.py:30
