"""Minimal stand-in for `plum.dispatch` (annotation based multiple dispatch).

Only needed to import the reference's EMLP policy code
(algos/emlp_torch/reps/representation.py:196-211,554-574).  Test infrastructure only.
"""
import inspect


class _Dispatcher:
    def __init__(self):
        self._table = {}

    def _register(self, fn, sig):
        self._table.setdefault(fn.__qualname__, []).append((sig, fn))
        table = self._table[fn.__qualname__]

        def call(*args):
            best, best_score = None, -1
            for s, f in table:
                if len(s) != len(args):
                    continue
                if all(isinstance(a, t) for a, t in zip(args, s)):
                    score = sum(len(t.__mro__) for t in s)
                    if score > best_score:
                        best, best_score = f, score
            if best is None:
                raise TypeError("no dispatch match for %r" % (tuple(type(a) for a in args),))
            return best(*args)

        call.__name__ = fn.__name__
        return call

    def __call__(self, fn):
        params = inspect.signature(fn).parameters.values()
        sig = tuple(p.annotation if p.annotation is not inspect._empty else object for p in params)
        return self._register(fn, sig)

    def multi(self, *sigs):
        def deco(fn):
            out = None
            for s in sigs:
                out = self._register(fn, tuple(s))
            return out
        return deco


dispatch = _Dispatcher()
