"""Synthetic DBN families, their Hugin .net text and ancestral samplers.

The benchmark configurations of BASELINE.json are "random-init CPTs of each
named architecture + sequences sampled from them".  This module produces, for
each family,
  * the `.net` text (what the unchanged reference host code parses), and
  * the `FlatModel` the reference's parser + join-tree builder would hand to the
    hot path for that text (tests/test_synth_vs_reference.py checks the two
    against each other while /root/reference is available), so that the GPU
    arm of bench.py does not need any reference code at run time;
  * a numpy ancestral sampler (the reference's generate_data, src/nip.c:2325-2478,
    runs several make_consistent per variable per slice and is unusable at
    benchmark sizes).

Host-side plumbing only: nothing here is on the timed path.
"""
from __future__ import annotations

import numpy as np

from .desc import IF_INCOMING, IF_OLD_OUTGOING, IF_OUTGOING, FlatModel


def _seq_normalise_blocks(a: np.ndarray, block: int) -> np.ndarray:
    """nip_normalise_cpd (src/nippotential.c:373-383) with the same left-to-right
    summation order; all-zero blocks stay as they are."""
    a = np.array(a, dtype=np.float64).reshape(-1, block)
    s = np.cumsum(a, axis=1)[:, -1:]
    s = np.where(s == 0, 1.0, s)
    return (a / s).reshape(-1)


def _fmt_rows(mat: np.ndarray, per_line: int = 6) -> str:
    """rows of `mat` as a nested Hugin data list, <= per_line numbers per text
    line (the reference's tokenizer reads lines of at most 10 000 chars,
    src/nipparsers.h:30)."""
    out = []
    for row in mat:
        nums = ["%.17g" % x for x in row]
        lines = [" ".join(nums[i:i + per_line]) for i in range(0, len(nums), per_line)]
        out.append("(" + "\n   ".join(lines) + ")")
    return "\n  ".join(out)


# ---------------------------------------------------------------------------
# HMM-style DBN  (configs C1/C2/C4/C5 of SURVEY §8):  P0 -> P1 -> M1
# ---------------------------------------------------------------------------
class HmmSpec:
    """Hidden chain with S states, M symbols.  Nodes are declared child first
    (M1, P1, P0) so that the parser's dimension-0 normalisation
    (src/huginnet.y:635-636) normalises over the child of every CPT."""

    def __init__(self, S: int, M: int, seed: int = 1, floor: float = 0.05):
        rng = np.random.default_rng(seed)
        self.S, self.M = S, M
        A = rng.random((S, S)) + floor   # A[p0, p1]
        E = rng.random((S, M)) + floor   # E[p1, m]
        pi = rng.random(S) + floor
        self.A = A / A.sum(1, keepdims=True)
        self.E = E / E.sum(1, keepdims=True)
        self.pi = pi / pi.sum()

    def net_text(self) -> str:
        S, M = self.S, self.M
        st = lambda n: " ".join('"%d"' % i for i in range(n))
        return "\n".join([
            "net { node_size = (80 40); }",
            "node M1 { label = \"obs\"; states = (%s); }" % st(M),
            "node P1 { label = \"state(t)\"; states = (%s); }" % st(S),
            "node P0 { label = \"state(t-1)\"; states = (%s); NIP_next = \"P1\"; }" % st(S),
            "potential (M1 | P1) { data = (\n  %s ); }" % _fmt_rows(self.E),
            "potential (P1 | P0) { data = (\n  %s ); }" % _fmt_rows(self.A),
            "potential (P0) { data = ( %s ); }" % _fmt_rows(self.pi[None, :])[1:-1],
            "",
        ])

    def flat(self) -> FlatModel:
        """What parse_model() builds for net_text(): variables M1=0, P1=1, P0=2;
        cliques {P1,P0} (P1 fastest) and {M1,P1} (M1 fastest); one sepset {P1};
        in_clique = out_clique = clique 0."""
        S, M = self.S, self.M
        # values as the parser sees them: the printed %.17g text round-trips exactly
        t0 = _seq_normalise_blocks(self.A.reshape(-1), S)   # [p0][p1], p1 fastest
        t1 = _seq_normalise_blocks(self.E.reshape(-1), M)   # [p1][m],  m fastest
        pr = np.array(self.pi, dtype=np.float64)
        s = np.cumsum(pr)[-1]
        pr = pr / s if s != 0 else pr
        return FlatModel(
            var_card=np.array([M, S, S]), var_flags=np.array([0, IF_INCOMING | IF_OUTGOING, IF_OLD_OUTGOING]),
            var_parent_off=np.array([0, 1, 2, 2]), var_parents=np.array([1, 2]),
            var_family=np.array([1, 0, 0]), var_prior_off=np.array([0, 0, 0, S]), var_prior=pr,
            clique_var_off=np.array([0, 2, 4]), clique_vars=np.array([1, 2, 0, 1]),
            clique_tab_off=np.array([0, S * S, S * S + S * M]), clique_tables=np.concatenate([t0, t1]),
            sepset_cliques=np.array([0, 1]), sepset_var_off=np.array([0, 1]), sepset_vars=np.array([1]),
            clique_adj_off=np.array([0, 1, 2]), clique_adj=np.array([0, 0]),
            outgoing=np.array([1]), prev_outgoing=np.array([2]), in_clique=0, out_clique=0,
            var_names=["M1", "P1", "P0"],
        ).normalise_dtypes()

    obs_vars = [0]      # data column: M1
    hidden_query = [1]  # P1

    def sample(self, n_series: int, T: int, seed: int = 2, missing: float = 0.0) -> np.ndarray:
        """[n_series, T, 1] int32 observations of M1 (-1 = missing)."""
        rng = np.random.default_rng(seed)
        cA = np.cumsum(self.A, axis=1)
        cE = np.cumsum(self.E, axis=1)
        cpi = np.cumsum(self.pi)
        out = np.empty((n_series, T, 1), dtype=np.int32)
        # slice 0: P0 ~ pi, P1 ~ A[P0]
        state = np.minimum(np.searchsorted(cpi, rng.random(n_series)), self.S - 1)
        for t in range(T):
            u = rng.random(n_series)
            state = np.minimum((cA[state] < u[:, None]).sum(1), self.S - 1)
            u = rng.random(n_series)
            out[:, t, 0] = np.minimum((cE[state] < u[:, None]).sum(1), self.M - 1)
        if missing > 0:
            out[rng.random(out.shape) < missing] = -1
        return out


# ---------------------------------------------------------------------------
# generic .net writer for small hand-specified DBNs (parity fixtures)
# ---------------------------------------------------------------------------
def net_text_generic(nodes, potentials) -> str:
    """nodes: list of (symbol, n_states, next_symbol_or_None) in declaration order.
    potentials: list of (child, [parents as written after '|'], ndarray) where the
    array is indexed [parents in written order..., child] (Hugin order: child
    fastest, last written parent next)."""
    lines = ["net { node_size = (80 40); }"]
    for sym, n, nxt in nodes:
        st = " ".join('"%d"' % i for i in range(n))
        extra = ' NIP_next = "%s";' % nxt if nxt else ""
        lines.append("node %s { states = (%s);%s }" % (sym, st, extra))
    for child, parents, arr in potentials:
        arr = np.asarray(arr, dtype=np.float64)
        head = "potential (%s%s)" % (child, (" | " + " ".join(parents)) if parents else "")
        flat = arr.reshape(-1, arr.shape[-1])
        lines.append("%s { data = (\n  %s ); }" % (head, _fmt_rows(flat)))
    return "\n".join(lines) + "\n"


# ---------------------------------------------------------------------------
# factorial DBN (config C3 of SURVEY §8): 4 ring-coupled chains
#   X^i_t | X^i_{t-1}, X^{i-1}_{t-1}   and   Y^i_t | X^i_t
# ---------------------------------------------------------------------------
class FactorialSpec:
    """K = 4 chains of `ns` states with one `ny`-symbol observation each.  Node order in the
    .net text: Y0..Y3, X0..X3, W0..W3 (W^i = X^i of the previous slice, NIP_next = X^i).

    flat() states the join tree the reference's parser + triangulation build for this text
    (verified for ns = 3 against the golden fixture factorial4x3.json, which was generated from
    the reference itself; the tree depends on the graph only, not on the cardinalities):
    four {Y^i, X^i} cliques and three 6-variable cliques chained by two 5-variable sepsets,
    interface |I| = ns^4."""

    K = 4
    _CLIQUES = [["Y0", "X0"], ["Y1", "X1"], ["Y3", "X3"], ["Y2", "X2"],
                ["X2", "X3", "W0", "W1", "W2", "W3"],
                ["X1", "X2", "X3", "W0", "W1", "W3"],
                ["X0", "X1", "X2", "X3", "W0", "W3"]]
    _SEPSETS = [(0, 6, ["X0"]), (1, 5, ["X1"]), (2, 6, ["X3"]), (3, 4, ["X2"]),
                (4, 5, ["X2", "X3", "W0", "W1", "W3"]), (5, 6, ["X1", "X2", "X3", "W0", "W3"])]
    _FAMILY = [0, 1, 3, 2, 6, 5, 4, 4, 4, 4, 4, 4]
    _ADJ = [[0], [1], [2], [3], [3, 4], [1, 5, 4], [0, 2, 5]]   # sepsets of each clique, reference order
    _IN_OUT = (4, 6)
    # the uncoupled variant (X^i_t | X^i_{t-1} only): four 5-variable cliques, 4-variable sepsets
    _U_CLIQUES = [["Y0", "X0"], ["Y1", "X1"], ["Y3", "X3"], ["Y2", "X2"],
                  ["X2", "W0", "W1", "W2", "W3"], ["X0", "X2", "W0", "W1", "W3"],
                  ["X0", "X1", "X2", "W1", "W3"], ["X0", "X1", "X2", "X3", "W3"]]
    _U_SEPSETS = [(0, 5, ["X0"]), (1, 7, ["X1"]), (2, 7, ["X3"]), (3, 7, ["X2"]),
                  (4, 5, ["X2", "W0", "W1", "W3"]), (5, 6, ["X0", "X2", "W1", "W3"]),
                  (6, 7, ["X0", "X1", "X2", "W3"])]
    _U_FAMILY = [0, 1, 3, 2, 5, 6, 4, 7, 4, 4, 4, 4]
    _U_ADJ = [[0], [1], [2], [3], [4], [0, 5, 4], [6, 5], [1, 3, 2, 6]]
    _U_IN_OUT = (4, 7)

    def __init__(self, ns: int = 16, ny: int = 4, seed: int = 1, coupled: bool = True):
        K = self.K
        rng = np.random.default_rng(seed)
        self.ns, self.ny, self.coupled = ns, ny, coupled
        if not coupled:
            self._CLIQUES, self._SEPSETS, self._FAMILY = self._U_CLIQUES, self._U_SEPSETS, self._U_FAMILY
            self._ADJ, self._IN_OUT = self._U_ADJ, self._U_IN_OUT
        # unnormalised, as written to the .net text: the parser normalises (and so does flat())
        self.E = [rng.random((ns, ny)) + 0.05 for _ in range(K)]            # [x][y]
        self.A = [rng.random((ns, ns, ns) if coupled else (ns, ns)) + 0.05 for _ in range(K)]  # [w_{i-1}][w_i][x] / [w_i][x]
        self.pi = [rng.random(ns) + 0.1 for _ in range(K)]
        self.names = ["Y%d" % i for i in range(K)] + ["X%d" % i for i in range(K)] + ["W%d" % i for i in range(K)]

    def net_text(self) -> str:
        K, ns, ny = self.K, self.ns, self.ny
        nodes = [("Y%d" % i, ny, None) for i in range(K)] + [("X%d" % i, ns, None) for i in range(K)] + \
                [("W%d" % i, ns, "X%d" % i) for i in range(K)]
        pots = [("Y%d" % i, ["X%d" % i], self.E[i]) for i in range(K)]
        pots += [("X%d" % i, (["W%d" % ((i - 1) % K)] if self.coupled else []) + ["W%d" % i], self.A[i]) for i in range(K)]
        pots += [("W%d" % i, [], self.pi[i][None, :]) for i in range(K)]
        return net_text_generic(nodes, pots)

    def flat(self) -> FlatModel:
        K, ns, ny = self.K, self.ns, self.ny
        idx = {n: i for i, n in enumerate(self.names)}
        card = np.array([ny] * K + [ns] * (2 * K))
        flags = np.array([0] * K + [IF_INCOMING | IF_OUTGOING] * K + [IF_OLD_OUTGOING] * K)
        parents, poff = [], [0]
        for v in range(3 * K):
            if v < K:
                parents += [idx["X%d" % v]]
            elif v < 2 * K:
                i = v - K
                parents += [idx["W%d" % i]] + ([idx["W%d" % ((i - 1) % K)]] if self.coupled else [])   # last written parent first
            poff.append(len(parents))
        # CPTs as the parser sees them (child fastest, blocks normalised left to right)
        cpt = {}
        for i in range(K):
            cpt[i] = _seq_normalise_blocks(self.E[i].reshape(-1), ny).reshape(ns, ny)
            cpt[K + i] = _seq_normalise_blocks(self.A[i].reshape(-1), ns).reshape(self.A[i].shape)
        cvars = [[idx[n] for n in c] for c in self._CLIQUES]
        tables, toff = [], [0]
        for c, vs in enumerate(cvars):
            shape = [int(card[v]) for v in reversed(vs)]           # numpy C order: last variable slowest
            t = np.ones(shape)
            for v in range(2 * K):
                if self._FAMILY[v] != c:
                    continue
                fam = [v] + parents[poff[v]:poff[v + 1]]          # child fastest in the CPT
                src = cpt[v]                                       # axes: reversed(fam)
                axes_in_t = [len(vs) - 1 - vs.index(u) for u in reversed(fam)]
                order = np.argsort(axes_in_t)
                shp = [1] * len(vs)
                for a in axes_in_t:
                    shp[a] = int(card[vs[len(vs) - 1 - a]])
                t = t * np.transpose(src, order).reshape(shp)
            tables.append(t.reshape(-1))
            toff.append(toff[-1] + t.size)
        prior, proff = [], [0]
        for v in range(3 * K):
            if v >= 2 * K:
                p = np.array(self.pi[v - 2 * K], dtype=np.float64)
                s = np.cumsum(p)[-1]
                prior.append(p / s if s != 0 else p)
            proff.append(proff[-1] + (ns if v >= 2 * K else 0))
        adj = self._ADJ
        return FlatModel(
            var_card=card, var_flags=flags, var_parent_off=np.array(poff), var_parents=np.array(parents),
            var_family=np.array(self._FAMILY), var_prior_off=np.array(proff), var_prior=np.concatenate(prior),
            clique_var_off=np.cumsum([0] + [len(c) for c in cvars]), clique_vars=np.concatenate(cvars),
            clique_tab_off=np.array(toff, dtype=np.int64), clique_tables=np.concatenate(tables),
            sepset_cliques=np.array([x for a, b, _ in self._SEPSETS for x in (a, b)]),
            sepset_var_off=np.cumsum([0] + [len(v) for _, _, v in self._SEPSETS]),
            sepset_vars=np.array([idx[n] for _, _, v in self._SEPSETS for n in v]),
            clique_adj_off=np.cumsum([0] + [len(a) for a in adj]), clique_adj=np.concatenate(adj),
            outgoing=np.arange(K, 2 * K), prev_outgoing=np.arange(2 * K, 3 * K),
            in_clique=self._IN_OUT[0], out_clique=self._IN_OUT[1],
            var_names=list(self.names),
        ).normalise_dtypes()

    obs_vars = [0, 1, 2, 3]

    def sample(self, n_series: int, T: int, seed: int = 2, missing: float = 0.0) -> np.ndarray:
        """[n_series, T, 4] int32 observations of Y0..Y3 (-1 = missing)."""
        K = self.K
        rng = np.random.default_rng(seed)
        nrm = lambda a: a / a.sum(-1, keepdims=True)
        pi, A, E = [nrm(a) for a in self.pi], [nrm(a) for a in self.A], [nrm(a) for a in self.E]
        draw = lambda p: np.minimum((np.cumsum(p, axis=-1) < rng.random(p.shape[:-1])[..., None]).sum(-1),
                                    p.shape[-1] - 1)
        w = [draw(np.broadcast_to(pi[i], (n_series, self.ns))) for i in range(K)]
        out = np.empty((n_series, T, K), dtype=np.int32)
        for t in range(T):
            x = [draw(A[i][w[(i - 1) % K], w[i]] if self.coupled else A[i][w[i]]) for i in range(K)]
            for i in range(K):
                out[:, t, i] = draw(E[i][x[i]])
            w = x
        if missing > 0:
            out[rng.random(out.shape) < missing] = -1
        return out
