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


# landmark indices of the MediaPipe topologies the reference anchors on (roi.py:16-22)
FACE_DETECTION_NOSE_INDEX = 2
FACE_LANDMARKS_NOSE_INDEX = 4
FACE_LANDMARKS_FOREHEAD_INDEX = 151
FACE_LANDMARKS_CHEEK_INDEX = 330
FACE_LANDMARKS_EYEBROW_INDEX = 337
HAND_LANDMARKS_WRIST_INDEX = 0
HAND_LANDMARKS_MIDDLE_INDEX = 9

_FACE, _HAND = model.ModelType.FACE_LANDMARKER, model.ModelType.HAND_LANDMARKER
# (left, top, right, bottom) fractions of the detection bbox around the anchor (roi.py:24-28)
FACE_CHEEK_CONFIG = ROIConfig(_FACE, [FACE_LANDMARKS_CHEEK_INDEX], (-0.05, -0.05, 0.15, 0.05))
FACE_EYEBROW_CONFIG = ROIConfig(_FACE, [FACE_LANDMARKS_EYEBROW_INDEX], (-0.10, -0.15, 0.25, 0.00))
FACE_FOREHEAD_CONFIG = ROIConfig(_FACE, [FACE_LANDMARKS_FOREHEAD_INDEX], (-0.00, -0.10, 0.20, 0.05))
HAND_WRIST_CONFIG = ROIConfig(_HAND, [HAND_LANDMARKS_WRIST_INDEX], (-0.10, -0.10, 0.10, 0.10))
HAND_PALM_CONFIG = ROIConfig(_HAND, [HAND_LANDMARKS_WRIST_INDEX, HAND_LANDMARKS_MIDDLE_INDEX], (-0.10, -0.10, 0.10, 0.10))

SELECTED_ROI_CONFIGS = [FACE_FOREHEAD_CONFIG, HAND_PALM_CONFIG]   # roi.py:30
