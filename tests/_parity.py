"""Shared replay-parity checker (test infrastructure).

north_star: "per-move determinant and trace ratios match the reference within 1e-10 relative, integer state
(expansion order, operator flavours, quantum-number blocks) matches bit-exactly".  A ratio that is itself the
result of a cancellation (S = F - R.M.Q with |S| << |F|; a multiplet trace that vanishes by symmetry) carries an
absolute FP64 floor of a few ulp of the O(1) terms it is built from, so the comparison is
    |r_engine - r_oracle| <= 1e-10 * |r_oracle| + 1e-14
(the absolute term only matters for |r| < 1e-4; such proposals are rejected with probability 1 - O(r))."""
import numpy as np

# move, status, flavour, FPos(c^dag), FPos(c), det-ratio sign, n_ops, outer superstate, NPairs, config sign, window position
INTS = [0, 1, 4, 5, 6, 8, 12, 13, 14, 15, 16]
RTOL, AFLOOR = 1e-10, 1e-14


def ratio_close(le, lo):
    """log|ratio| columns of engine / oracle -> boolean mask of rows inside the tolerance."""
    with np.errstate(over="ignore", invalid="ignore"):
        re_, ro = np.exp(np.minimum(le, 700.0)), np.exp(np.minimum(lo, 700.0))
        return (np.abs(le - lo) <= RTOL) | (np.abs(re_ - ro) <= RTOL * np.abs(ro) + AFLOOR)


def floor_usage(log, ref):
    """How many replayed moves pass only through the absolute floor (their relative deviation exceeds RTOL): (det, trace, moves)."""
    out = []
    for col in (7, 9):
        le, lo = log[:, col], ref[:, col]
        with np.errstate(over="ignore", invalid="ignore"):
            re_, ro = np.exp(np.minimum(le, 700.0)), np.exp(np.minimum(lo, 700.0))
            rel_ok = (np.abs(le - lo) <= RTOL) | (np.abs(re_ - ro) <= RTOL * np.abs(ro))
        out.append(int((~rel_ok).sum()))
    return out[0], out[1], int(log.shape[0])


def assert_replay(log, ref):
    assert log.shape == ref.shape
    bad = np.nonzero((log[:, INTS] != ref[:, INTS]).any(1))[0]
    assert bad.size == 0, "integer state differs first at move %d:\n%s\n%s" % (bad[0], log[bad[0]], ref[bad[0]])
    okd, okt = ratio_close(log[:, 7], ref[:, 7]), ratio_close(log[:, 9], ref[:, 9])
    assert okd.all(), "determinant ratio: move %d  %r vs %r" % (np.nonzero(~okd)[0][0], log[~okd][0, 7], ref[~okd][0, 7])
    assert okt.all(), "trace ratio: move %d  %r vs %r" % (np.nonzero(~okt)[0][0], log[~okt][0, 9], ref[~okt][0, 9])
    # sign of the trace ratio: bit-exact wherever the ratio is not a numerical zero
    sig = np.exp(np.minimum(ref[:, 9], 0.0)) > 1e-12
    assert np.array_equal(log[sig, 10], ref[sig, 10]), "trace-ratio sign"
    # proposed times and the accept draw: same counter-based RNG stream, unfused arithmetic
    assert np.array_equal(log[:, [2, 3, 17]], ref[:, [2, 3, 17]]), "proposed times / accept draws"
    pr = np.abs(log[:, 11] - ref[:, 11]) <= 1e-9 * np.abs(ref[:, 11]) + 1e-13
    assert pr.all(), "acceptance probability"
