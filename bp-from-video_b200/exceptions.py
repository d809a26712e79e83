"""Drop-in twin of the reference's `exceptions` module (exceptions.py:1-2)."""


class CaptureError(RuntimeError):
    """Raised by a frame source that cannot deliver frames (the drivers bp.py:29 / pbp.py stop on it)."""
