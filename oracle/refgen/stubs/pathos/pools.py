"""Fork-based stand-in for ``pathos.pools.ProcessPool``.

The reference only uses ``ProcessPool(nodes=n).map(func, iterable)``
(sympleints/Functions.py:45-52).  ``func`` is a closure and cannot be pickled,
so it is parked in a module global and inherited by forked workers.  Work items
are dispatched heaviest-first (largest total angular momentum) to shorten the
serial tail; results are returned in the original order.
"""
import multiprocessing as mp
import os

_FUNC = None


def _call(arg):
    return _FUNC(arg)


class ProcessPool:
    def __init__(self, nodes=None):
        env = os.environ.get("SYMPLEINTS_REF_WORKERS")
        self.nodes = int(env) if env else (nodes or 1)

    def map(self, func, iterable):
        global _FUNC
        items = list(iterable)
        if self.nodes <= 1 or len(items) <= 1:
            return [func(it) for it in items]
        order = sorted(range(len(items)), key=lambda i: -sum(items[i]))
        _FUNC = func
        ctx = mp.get_context("fork")
        with ctx.Pool(self.nodes) as pool:
            res = pool.map(_call, [items[i] for i in order], chunksize=1)
        out = [None] * len(items)
        for i, r in zip(order, res):
            out[i] = r
        return out
