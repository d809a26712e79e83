"""Drop-in twin of the reference's `profiler` module (profiler.py:10-47).

The reference wraps every stage method in a cProfile-based decorator.  The hot path here runs on the GPU,
where host-side cProfile says nothing; kernels are profiled with ncu / CUDA events instead (profiles/).
This module keeps the import surface the drivers use — `profiler.enabled` (pbp.py:11), `timeit`, `printit` —
with a cheap wall-clock accumulator per decorated function.
"""
import functools
import time

PROFILER_ENABLED = True      # as the reference (profiler.py:7): bp.py prints the per-stage table at exit; pbp.py:11 switches it off


class Profiler:
    def __init__(self, enabled: bool = PROFILER_ENABLED) -> None:
        self.enabled = enabled
        self.funcs: list[str] = []
        self.stats: dict[str, list[float]] = {}

    def timeit(self, func):
        name = func.__name__
        if name not in self.funcs:
            self.funcs.append(name)

        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            if not self.enabled:
                return func(*args, **kwargs)
            t0 = time.perf_counter()
            try:
                return func(*args, **kwargs)
            finally:
                rec = self.stats.setdefault(name, [0, 0.0])
                rec[0] += 1
                rec[1] += time.perf_counter() - t0
        return wrapper

    def clear(self) -> None:
        self.stats.clear()

    def printit(self, clear_info: bool = False) -> None:
        if not self.enabled:
            return
        print(f'{"function":<28}{"calls":>8}{"total s":>12}{"ms/call":>12}')
        for name in sorted(self.stats):
            calls, total = self.stats[name]
            print(f'{name:<28}{calls:>8}{total:>12.4f}{1e3 * total / max(calls, 1):>12.4f}')
        if clear_info:
            self.clear()


profiler = Profiler()
timeit = profiler.timeit
printit = profiler.printit
