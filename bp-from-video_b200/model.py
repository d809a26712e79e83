"""Drop-in twin of the reference's `model` module (model.py:4-8): the detector kinds whose outputs the
signal path can anchor an ROI on.  Values equal the lower-cased member names, as enum.auto() yields for a
StrEnum, so pickles and comparisons interoperate with the reference's own enum values."""
import enum


class ModelType(enum.StrEnum):
    FACE_DETECTOR = 'face_detector'
    FACE_LANDMARKER = 'face_landmarker'
    HAND_LANDMARKER = 'hand_landmarker'
    PERSON_SEGMENTER = 'person_segmenter'
