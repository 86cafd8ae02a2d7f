"""Named wall-clock timers with the interface of adcc/timings.py:84-253 (Timer.record /
restart / stop / total / current / describe, timed_member_call)."""
import time
from contextlib import contextmanager


class Timer:
    def __init__(self):
        self.time_construction = time.perf_counter()
        self.intervals = {}
        self.start_times = {}

    def restart(self, task):
        self.start_times[task] = time.perf_counter()

    def stop(self, task, now=None):
        if task in self.start_times:
            now = now or time.perf_counter()
            self.intervals.setdefault(task, []).append((self.start_times.pop(task), now))
            return now - self.intervals[task][-1][0]
        return 0.0

    def is_running(self, task):
        return task in self.start_times

    @contextmanager
    def record(self, task):
        self.restart(task)
        try:
            yield self
        finally:
            self.stop(task)

    @property
    def tasks(self):
        return sorted(set(self.intervals) | set(self.start_times))

    def total(self, task):
        now = time.perf_counter()
        tot = sum(e - s for s, e in self.intervals.get(task, []))
        if task in self.start_times:
            tot += now - self.start_times[task]
        for t in self.tasks:                      # nested keys "a/b" count towards "a"
            if t != task and t.startswith(task + "/"):
                tot += 0.0
        return tot

    def current(self, task):
        if task in self.start_times:
            return time.perf_counter() - self.start_times[task]
        if task in self.intervals:
            s, e = self.intervals[task][-1]
            return e - s
        raise ValueError(f"Unknown task {task}")

    @property
    def lifetime(self):
        return time.perf_counter() - self.time_construction

    def attach(self, other, subtree=""):
        for k, v in other.intervals.items():
            self.intervals.setdefault((subtree + "/" + k) if subtree else k, []).extend(v)

    def describe(self):
        rows = [f"{t:40s} {self.total(t):10.4f}s  x{len(self.intervals.get(t, []))}" for t in self.tasks]
        return "\n".join(rows)


def strtime(t):
    return f"{t:.3f}s" if t < 60 else f"{int(t // 60)}m {t % 60:.1f}s"


def strtime_short(t):
    if t < 1:
        return f"{1000 * t:.0f}ms"
    return f"{t:.1f}s" if t < 60 else f"{t / 60:.1f}m"


def timed_member_call(timer="timer"):
    def decorator(fn):
        def wrapped(self, *args, **kwargs):
            with getattr(self, timer).record(fn.__name__):
                return fn(self, *args, **kwargs)
        wrapped.__name__ = fn.__name__
        wrapped.__doc__ = fn.__doc__
        return wrapped
    return decorator
