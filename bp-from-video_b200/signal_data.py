"""Drop-in twin of the reference's `signal_data` module (signal_data.py:12-117): NaN-aware bounded time
series (`Signal`) and a list-of-signals container (`SignalGroup`).

Host-side bookkeeping only — the windows that feed the CUDA kernels live in device ring buffers
(bpv/engine.py); these classes are what `SignalStore` hands to the drawer / pbp pipeline, so they keep the
reference's attribute surface (x, y, v, w, range_x, range_y, signals, num_signals) and the exact semantics
of get_fs / get_mean / get_peak.  Storage is a numpy array window instead of two deques: `x` and `y` are
ndarrays (convertible by np.array, iterable, picklable), sized like the reference's deques.
"""
from __future__ import annotations

import collections.abc
import math
import warnings

import numpy as np

import roi

type XType = int | float
type YType = int | float | roi.Location


def _initial(value, maxlen):
    """Contents of deque(value if sequence else [value]*maxlen (or []), maxlen) as an ndarray (signal_data.py:18-19)."""
    if isinstance(value, (list, np.ndarray)):
        data = np.asarray(value, dtype=float)
        if data.ndim == 0:
            data = data.reshape(1)
        return data[len(data) - maxlen:] if maxlen is not None and len(data) > maxlen else data.copy()
    if maxlen is None:
        return np.empty(0)
    row = np.asarray(value, dtype=float)
    return np.repeat(row[None, ...], maxlen, axis=0) if row.ndim else np.full(maxlen, float(value))


def _span(values, mask):
    """(nanmin, nanmax) when at least two entries are usable, else (nan, nan) (signal_data.py:47-49).
    fmin / fmax reductions ignore NaN like nanmin / nanmax (all-NaN -> nan) without their warning machinery: this runs
    twice per Signal on every appended sample."""
    if np.count_nonzero(mask) < 2:
        return (np.nan, np.nan)
    return (np.fmin.reduce(values, axis=None), np.fmax.reduce(values, axis=None))


class Signal:
    """One bounded series of (x, y) samples; y may be scalar or a fixed-length vector (ROI Locations)."""

    def __init__(self, xi: XType | list[XType] = np.nan, yi: YType | list[YType] = np.nan, s_maxlen: int | None = None) -> None:
        self.maxlen = s_maxlen
        self.x = _initial(xi, s_maxlen)
        self.y = _initial(yi, s_maxlen)
        self.reset_mask()
        self.reset_range()

    def __repr__(self) -> str:
        with np.printoptions(legacy='1.25'):
            return f"{type(self).__name__}({', '.join(f'{k}={v}' for k, v in vars(self).items())})"

    # -- mutation ---------------------------------------------------------------------------------
    def _append(self, arr, value):
        value = np.asarray(value, dtype=float)
        if arr.size == 0:
            arr = np.empty((0, *value.shape))
        if self.maxlen is not None and len(arr) >= self.maxlen:
            if self.maxlen == 0:
                return arr
            return np.concatenate([arr[1:], value[None, ...]], axis=0)      # deque(maxlen).append: the oldest sample falls out
        return np.concatenate([arr, value[None, ...]], axis=0)

    def add_sample(self, xp: XType, yp: YType) -> None:
        self.x = self._append(self.x, xp)
        self.y = self._append(self.y, yp)
        self.reset_mask()
        self.reset_range()

    def set_data(self, data_x=None, data_y=None) -> None:
        if data_x is not None:
            self.x = _initial(list(data_x), self.maxlen)
        if data_y is not None:
            self.y = _initial(list(data_y), self.maxlen)
        self.reset_mask()
        self.reset_range()

    # -- masks and ranges -------------------------------------------------------------------------
    def reset_mask(self) -> None:
        self.v = np.isfinite(self.x)
        fy = np.isfinite(self.y)
        self.w = fy.all(axis=1) if fy.ndim == 2 else fy

    def reset_range(self) -> None:
        self.range_x = _span(self.x, self.v)
        self.range_y = _span(self.y, self.w)

    def set_range(self, range_x=None, range_y=None) -> None:
        if range_x is not None:
            self.range_x = range_x
        if range_y is not None:
            self.range_y = range_y

    # -- queries ----------------------------------------------------------------------------------
    def get_fs(self, only_valid: bool = False) -> XType:
        """Sampling rate 1 / mean(diff(x)) over the finite x (or finite y when only_valid)."""
        mask = self.w if only_valid else self.v
        if int(mask.sum()) < 2:
            return np.nan
        return 1 / np.nanmean(np.diff(np.asarray(self.x)[mask]))

    def get_mean(self, as_int: bool = False) -> YType:
        y = np.asarray(self.y)
        if not self.w.any():
            return y[-1]
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', RuntimeWarning)
            mean = np.squeeze(np.nanmean(y, axis=0))
        return mean.round().astype(int) if as_int else mean

    def get_peak(self, min_x: XType | None = None, max_x: XType | None = None) -> tuple[XType, YType]:
        """(x, y) at the largest finite y with min_x <= x <= max_x (defaults: range_x); first maximum wins."""
        x, y = np.asarray(self.x), np.asarray(self.y)
        lo = self.range_x[0] if min_x is None else min_x
        hi = self.range_x[1] if max_x is None else max_x
        with np.errstate(invalid='ignore'):
            pick = (lo <= x) & (x <= hi) & self.w
        if int(pick.sum()) >= 2:
            k = int(np.argmax(y[pick]))
            return (x[pick][k], y[pick][k])
        if y.ndim == 2:
            return ((np.nan,) * y.shape[-1], np.nan)
        return (np.nan, np.nan)


class SignalGroup:
    """A list of Signals that are sampled together, with the union of their ranges."""

    def __init__(self, num_signals: int | None = None, xi=np.nan, yi=np.nan, s_maxlen: int | None = None, *,
                 signals: list[Signal] | None = None) -> None:
        if signals is None:
            signals = [Signal(xi, yi, s_maxlen) for _ in range(num_signals)]
        self.signals = signals
        self.num_signals = len(signals)
        self.reset_ranges()

    def __repr__(self) -> str:
        return f"{type(self).__name__}({', '.join(f'{k}={v}' for k, v in vars(self).items())})"

    def __iter__(self) -> collections.abc.Iterator[Signal]:
        return iter(self.signals)

    def add_samples(self, xps, yps) -> None:
        if not isinstance(xps, (list, np.ndarray)):
            xps = [xps] * self.num_signals
        for sig, xp, yp in zip(self.signals, xps, yps):
            sig.add_sample(xp, yp)
        self.reset_ranges()

    @staticmethod
    def _union(pairs):
        """(nanmin of the lows, nanmax of the highs) when both sides have a finite entry, else (nan, nan)."""
        lows, highs = zip(*pairs) if pairs else ((), ())
        if not (any(math.isfinite(v) for v in lows) and any(math.isfinite(v) for v in highs)):
            return (np.nan, np.nan)
        return (np.float64(min(v for v in lows if v == v)), np.float64(max(v for v in highs if v == v)))

    def reset_ranges(self) -> None:
        # NB (reference behaviour, signal_data.py:100-102): wrapping signals in a group re-derives every
        # member's range from its data, discarding ranges set with Signal.set_range().
        for sig in self.signals:
            sig.reset_range()
        self.range_x = self._union([s.range_x for s in self.signals])
        self.range_y = self._union([s.range_y for s in self.signals])

    def set_ranges(self, range_x=None, range_y=None) -> None:
        for sig in self.signals:
            sig.set_range(range_x, range_y)
        if range_x is not None:
            self.range_x = range_x
        if range_y is not None:
            self.range_y = range_y

    def get_means(self, as_int: bool = False) -> list[YType]:
        return [sig.get_mean(as_int) for sig in self.signals]

    def get_peaks(self, min_x=None, max_x=None) -> list[tuple[XType, YType]]:
        return [sig.get_peak(min_x, max_x) for sig in self.signals]
