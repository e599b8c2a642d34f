"""TEST INFRASTRUCTURE: recording proxy for the `mctqmc` surface (used to prove that w2dynamics_b200/solver_sequence.py issues the
calls of the unmodified reference plugin) + re-exports of the sequence for the tests."""
import hashlib

import numpy as np

from w2dynamics_b200.solver_sequence import MCConfig, PluginSequence, RESULT_READS, qmc_defaults  # noqa: F401


def _digest(x):
    if isinstance(x, np.ndarray):
        a = np.ascontiguousarray(x)
        return "nd%s%s:%s" % (a.dtype.str, a.shape, hashlib.sha1(a.tobytes()).hexdigest()[:12])
    if isinstance(x, (bool, np.bool_)):
        return "b%d" % int(x)
    if isinstance(x, (int, np.integer)):
        return "i%d" % int(x)
    if isinstance(x, (float, np.floating)):
        return "f%r" % float(x)
    if isinstance(x, (bytes, str)):
        s = x if isinstance(x, bytes) else x.encode()
        return "s%d:%s" % (len(s), hashlib.sha1(s).hexdigest()[:12])
    if x is None:
        return "None"
    return type(x).__name__


class Recorder:
    """Proxy around an mctqmc object: logs every subroutine call (name + argument digests) and every module-variable read / write."""

    def __init__(self, target):
        object.__setattr__(self, "_t", target)
        object.__setattr__(self, "log", [])

    def __getattr__(self, name):
        v = getattr(self._t, name)
        if callable(v):
            def call(*a, **kw):
                self.log.append(("call", name, tuple(_digest(x) for x in a), tuple(sorted((k, _digest(x)) for k, x in kw.items()))))
                return v(*a, **kw)
            return call
        self.log.append(("read", name))
        return v

    def __setattr__(self, name, value):
        self.log.append(("write", name, _digest(value)))
        setattr(self._t, name, value)


