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
