"""Per-mismatch near-tie verification for the query path (north_star: "identical except documented near-ties,
relative distance gap < 1e-5").  Every id / line that differs from the oracle is re-evaluated in float64 from the
codebooks and must be closer to the oracle's choice than REL_TIE; nothing is accepted as a fraction.

The query-time quantities cancel ||q||^2 (distances are ||q - y||^2 - ||q||^2, line scores are built from
||c||^2 - 2 q.c), so gaps are taken relative to the magnitudes that were cancelled: |value| + ||q||^2.
"""
import numpy as np

REL_TIE = 1e-5
REL_DIST = 1e-4


class Model64:
    """float64 evaluation of line scores and decode distances for one model (numpy, vectorised over pairs)"""

    def __init__(self, m):
        self.cent = np.asarray(m["cent"], np.float64)
        self.cn = (self.cent ** 2).sum(1)
        self.edge = np.asarray(m["edge"])
        self.ed2 = np.asarray(m["edge_d2"], np.float64)  # the stored fp32 values ARE the model (term5)
        self.lcb = np.asarray(m["lambda_cb"], np.float64)
        self.pq = np.asarray(m["pq"], np.float64)
        self.E = self.edge.shape[1]
        self.M = self.pq.shape[0]

    def coarse(self, q, c):
        """D[c] = ||c||^2 - 2 q.c   (gpu/impl/Distance.cu:287-290: no ||q||^2)"""
        return self.cn[c] - 2.0 * (self.cent[c] * q).sum(-1)

    def line_score(self, q, line):
        """BroadcastSum.cu:505-520: (v > 0) ? b : b - v^2 / (4 c2)"""
        c, e = line // self.E, line % self.E
        s = self.edge[c, e]
        a2, b2, c2 = self.coarse(q, s), self.coarse(q, c), self.ed2[c, e]
        v = a2 - b2 - c2
        return np.where(v > 0, b2, b2 - 0.25 * v * v / c2)

    def decode_dist(self, q, line, lamq, code):
        """||q - ((1-l)c + l s) - p(code)||^2 - ||q||^2"""
        c, e = line // self.E, line % self.E
        s = self.edge[c, e]
        lh = self.lcb[lamq][..., None]
        p = np.concatenate([self.pq[mm][code[..., mm]] for mm in range(self.M)], axis=-1)
        y = (1.0 - lh) * self.cent[c] + lh * self.cent[s] + p
        return ((q - y) ** 2).sum(-1) - (q ** 2).sum(-1)


def check_lines(lines_gpu, lines_ref, xq, m, rel=REL_TIE):
    """Top-W line lists [nq][W] (rank order).  Returns a bool mask of the queries whose line SETS agree.  Every
    rank-wise difference must be a near-tie in the float64 line score."""
    m64 = m if isinstance(m, Model64) else Model64(m)
    lg, lr = np.asarray(lines_gpu), np.asarray(lines_ref)
    assert lg.shape == lr.shape
    assert np.array_equal(lg < 0, lr < 0), "padding of the line lists differs"
    qi, r = np.nonzero(lg != lr)
    if len(qi):
        q = np.asarray(xq, np.float64)[qi]
        sg, sr = m64.line_score(q, lg[qi, r]), m64.line_score(q, lr[qi, r])
        scale = np.abs(sr) + (q ** 2).sum(1)
        bad = np.abs(sg - sr) > rel * scale
        assert not bad.any(), "unexplained line mismatches: %s" % [
            (int(qi[i]), int(r[i]), int(lg[qi[i], r[i]]), int(lr[qi[i], r[i]]), float(sg[i]), float(sr[i]))
            for i in np.nonzero(bad)[0][:5]]
    same_set = np.array([set(a.tolist()) == set(b.tolist()) for a, b in zip(lg, lr)])
    return same_set


def check_topk(D, I, Do, Io, xq, m, entry_line, entry_lamq, entry_codes, same_lines=None, rel=REL_TIE, rel_dist=REL_DIST):
    """GPU top-k (D, I) against the oracle's (Do, Io) on the same index; ids index the arrival-order arrays entry_*.
    * padding identical; GPU rows ascending;
    * every returned distance equals the float64 decode distance of its id within rel_dist;
    * every rank where the ids differ is a near-tie: the two ids' float64 distances differ by < rel (relative to
      |d| + ||q||^2).  Queries whose selected line sets differ (same_lines False; those line differences were themselves
      verified as near-ties by check_lines) are exempt from the rank-wise id check only.
    Returns (number of differing ranks, number of exempt queries)."""
    m64 = m if isinstance(m, Model64) else Model64(m)
    D, I, Do, Io = np.asarray(D), np.asarray(I), np.asarray(Do), np.asarray(Io)
    q64 = np.asarray(xq, np.float64)
    nq, k = I.shape
    if same_lines is None:
        same_lines = np.ones(nq, bool)
    fmax = np.finfo(np.float32).max
    assert np.all(np.diff(D, axis=1) >= 0), "GPU distances not ascending"
    strict = same_lines[:, None] & np.ones((1, k), bool)
    assert np.array_equal((I < 0) & strict, (Io < 0) & strict), "padding differs"
    assert np.all(D[I < 0] == fmax) and np.all(Do[Io < 0] == fmax)
    qn = (q64 ** 2).sum(1)
    # self-consistency of every GPU entry
    qi, r = np.nonzero(I >= 0)
    ids = I[qi, r]
    d64 = m64.decode_dist(q64[qi], entry_line[ids], entry_lamq[ids], entry_codes[ids])
    err = np.abs(D[qi, r] - d64)
    assert np.all(err <= rel_dist * (np.abs(d64) + qn[qi])), "GPU distance differs from the float64 decode distance"
    # rank-wise id differences must be near-ties
    diff = (I != Io) & (I >= 0) & (Io >= 0) & strict
    qi, r = np.nonzero(diff)
    if len(qi):
        ig, io = I[qi, r], Io[qi, r]
        dg = m64.decode_dist(q64[qi], entry_line[ig], entry_lamq[ig], entry_codes[ig])
        do = m64.decode_dist(q64[qi], entry_line[io], entry_lamq[io], entry_codes[io])
        bad = np.abs(dg - do) > rel * (np.abs(do) + qn[qi])
        assert not bad.any(), "unexplained id mismatches (query, rank, gpu id, oracle id, d64 gpu, d64 oracle): %s" % [
            (int(qi[i]), int(r[i]), int(ig[i]), int(io[i]), float(dg[i]), float(do[i])) for i in np.nonzero(bad)[0][:5]]
    # where ids agree the distances agree within rel_dist
    same = (I == Io) & (I >= 0)
    assert np.all(np.abs(D - Do)[same] <= rel_dist * (np.abs(Do) + qn[:, None])[same])
    return int(diff.sum()), int((~same_lines).sum())
