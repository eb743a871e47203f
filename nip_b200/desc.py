"""Flat model description shared by the device library and the test oracles.

`ModelDesc` mirrors `nipgpu_model_desc` (include/nipgpu.h) field for field;
`FlatModel` owns the numpy arrays behind one and can be saved/loaded as JSON
(the format of the fixtures under tests/golden/).  Layout conventions are the
reference's: tables flat with dimension 0 fastest (src/nippotential.c:58-68),
variables numbered in `model->variables[]` order.
"""
from __future__ import annotations

import ctypes as C
import json
from dataclasses import dataclass, field

import numpy as np

IF_INCOMING, IF_OUTGOING, IF_OLD_OUTGOING = 1, 2, 4

_I32P = C.POINTER(C.c_int32)
_I64P = C.POINTER(C.c_int64)
_F64P = C.POINTER(C.c_double)


class ModelDesc(C.Structure):
    _fields_ = [
        ("n_vars", C.c_int32),
        ("var_card", _I32P), ("var_flags", _I32P), ("var_parent_off", _I32P),
        ("var_parents", _I32P), ("var_family", _I32P), ("var_prior_off", _I32P),
        ("var_prior", _F64P),
        ("n_cliques", C.c_int32),
        ("clique_var_off", _I32P), ("clique_vars", _I32P), ("clique_tab_off", _I64P),
        ("clique_tables", _F64P),
        ("n_sepsets", C.c_int32),
        ("sepset_cliques", _I32P), ("sepset_var_off", _I32P), ("sepset_vars", _I32P),
        ("clique_adj_off", _I32P), ("clique_adj", _I32P),
        ("n_interface", C.c_int32),
        ("outgoing", _I32P), ("prev_outgoing", _I32P),
        ("in_clique", C.c_int32), ("out_clique", C.c_int32),
    ]


_I32_FIELDS = ["var_card", "var_flags", "var_parent_off", "var_parents", "var_family",
               "var_prior_off", "clique_var_off", "clique_vars", "sepset_cliques",
               "sepset_var_off", "sepset_vars", "clique_adj_off", "clique_adj",
               "outgoing", "prev_outgoing"]


def _arr(ptr, n, dtype):
    if n <= 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


@dataclass
class FlatModel:
    var_card: np.ndarray
    var_flags: np.ndarray
    var_parent_off: np.ndarray
    var_parents: np.ndarray
    var_family: np.ndarray
    var_prior_off: np.ndarray
    var_prior: np.ndarray
    clique_var_off: np.ndarray
    clique_vars: np.ndarray
    clique_tab_off: np.ndarray
    clique_tables: np.ndarray
    sepset_cliques: np.ndarray
    sepset_var_off: np.ndarray
    sepset_vars: np.ndarray
    clique_adj_off: np.ndarray
    clique_adj: np.ndarray
    outgoing: np.ndarray
    prev_outgoing: np.ndarray
    in_clique: int
    out_clique: int
    var_names: list = field(default_factory=list)

    # ---- sizes -----------------------------------------------------------
    @property
    def n_vars(self): return len(self.var_card)
    @property
    def n_cliques(self): return len(self.clique_var_off) - 1
    @property
    def n_sepsets(self): return len(self.sepset_var_off) - 1
    @property
    def n_interface(self): return len(self.outgoing)

    def parents(self, v):
        return [int(x) for x in self.var_parents[self.var_parent_off[v]:self.var_parent_off[v + 1]]]

    def clique(self, c):
        return [int(x) for x in self.clique_vars[self.clique_var_off[c]:self.clique_var_off[c + 1]]]

    def clique_table(self, c):
        return self.clique_tables[self.clique_tab_off[c]:self.clique_tab_off[c + 1]]

    def sepset(self, s):
        return [int(x) for x in self.sepset_vars[self.sepset_var_off[s]:self.sepset_var_off[s + 1]]]

    def prior(self, v):
        return self.var_prior[self.var_prior_off[v]:self.var_prior_off[v + 1]]

    def family_sizes(self):
        """size of each variable's family table (child, parents...) — the EM count layout"""
        out = []
        for v in range(self.n_vars):
            n = int(self.var_card[v])
            for p in self.parents(v):
                n *= int(self.var_card[p])
            out.append(n)
        return out

    def counts_offsets(self):
        return np.concatenate([[0], np.cumsum(self.family_sizes())]).astype(np.int64)

    def interface_size(self):
        n = 1
        for v in self.outgoing:
            n *= int(self.var_card[v])
        return n

    # ---- ctypes ------------------------------------------------------------
    def normalise_dtypes(self):
        for f in _I32_FIELDS:
            setattr(self, f, np.ascontiguousarray(getattr(self, f), dtype=np.int32))
        self.clique_tab_off = np.ascontiguousarray(self.clique_tab_off, dtype=np.int64)
        self.clique_tables = np.ascontiguousarray(self.clique_tables, dtype=np.float64)
        self.var_prior = np.ascontiguousarray(self.var_prior, dtype=np.float64)
        return self

    def to_ctypes(self) -> ModelDesc:
        """The returned struct borrows this object's arrays: keep `self` alive."""
        self.normalise_dtypes()
        d = ModelDesc()
        d.n_vars, d.n_cliques = self.n_vars, self.n_cliques
        d.n_sepsets, d.n_interface = self.n_sepsets, self.n_interface
        d.in_clique, d.out_clique = int(self.in_clique), int(self.out_clique)
        for f in _I32_FIELDS:
            setattr(d, f, getattr(self, f).ctypes.data_as(_I32P))
        d.clique_tab_off = self.clique_tab_off.ctypes.data_as(_I64P)
        d.clique_tables = self.clique_tables.ctypes.data_as(_F64P)
        d.var_prior = self.var_prior.ctypes.data_as(_F64P)
        d._owner = self
        return d

    @staticmethod
    def from_ctypes(d: ModelDesc, var_names=None) -> "FlatModel":
        nv, nc, ns, ni = d.n_vars, d.n_cliques, d.n_sepsets, d.n_interface
        poff = _arr(d.var_parent_off, nv + 1, np.int32)
        proff = _arr(d.var_prior_off, nv + 1, np.int32)
        cvoff = _arr(d.clique_var_off, nc + 1, np.int32)
        toff = _arr(d.clique_tab_off, nc + 1, np.int64)
        svoff = _arr(d.sepset_var_off, ns + 1, np.int32)
        aoff = _arr(d.clique_adj_off, nc + 1, np.int32)
        return FlatModel(
            var_card=_arr(d.var_card, nv, np.int32), var_flags=_arr(d.var_flags, nv, np.int32),
            var_parent_off=poff, var_parents=_arr(d.var_parents, int(poff[-1]), np.int32),
            var_family=_arr(d.var_family, nv, np.int32), var_prior_off=proff,
            var_prior=_arr(d.var_prior, int(proff[-1]), np.float64),
            clique_var_off=cvoff, clique_vars=_arr(d.clique_vars, int(cvoff[-1]), np.int32),
            clique_tab_off=toff, clique_tables=_arr(d.clique_tables, int(toff[-1]), np.float64),
            sepset_cliques=_arr(d.sepset_cliques, 2 * ns, np.int32), sepset_var_off=svoff,
            sepset_vars=_arr(d.sepset_vars, int(svoff[-1]), np.int32),
            clique_adj_off=aoff, clique_adj=_arr(d.clique_adj, int(aoff[-1]), np.int32),
            outgoing=_arr(d.outgoing, ni, np.int32), prev_outgoing=_arr(d.prev_outgoing, ni, np.int32),
            in_clique=int(d.in_clique), out_clique=int(d.out_clique),
            var_names=list(var_names or []),
        ).normalise_dtypes()

    # ---- JSON (fixtures) -----------------------------------------------------
    def to_json(self) -> dict:
        out = {f: getattr(self, f).tolist() for f in _I32_FIELDS}
        out["clique_tab_off"] = self.clique_tab_off.tolist()
        out["clique_tables"] = [float.hex(float(x)) for x in self.clique_tables]
        out["var_prior"] = [float.hex(float(x)) for x in self.var_prior]
        out["in_clique"], out["out_clique"] = int(self.in_clique), int(self.out_clique)
        out["var_names"] = list(self.var_names)
        return out

    @staticmethod
    def from_json(j: dict) -> "FlatModel":
        kw = {f: np.array(j[f], dtype=np.int32) for f in _I32_FIELDS}
        kw["clique_tab_off"] = np.array(j["clique_tab_off"], dtype=np.int64)
        kw["clique_tables"] = np.array([float.fromhex(x) for x in j["clique_tables"]], dtype=np.float64)
        kw["var_prior"] = np.array([float.fromhex(x) for x in j["var_prior"]], dtype=np.float64)
        return FlatModel(in_clique=j["in_clique"], out_clique=j["out_clique"],
                         var_names=j.get("var_names", []), **kw).normalise_dtypes()

    def structure_equal(self, other: "FlatModel") -> bool:
        for f in _I32_FIELDS + ["clique_tab_off"]:
            if not np.array_equal(getattr(self, f), getattr(other, f)):
                return False
        return self.in_clique == other.in_clique and self.out_clique == other.out_clique


def save_json(obj, path):
    with open(path, "w") as fh:
        json.dump(obj, fh, separators=(",", ":"))


def load_json(path):
    with open(path) as fh:
        return json.load(fh)
