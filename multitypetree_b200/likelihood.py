"""Host-side mirror of the BEAST 2 objects the reference wires around TreeLikelihood
(examples/structuredCoalescent.xml:532-540, 568-572; fxtemplates/MultiTypeTree.xml:184-197):
``Alignment`` (site-pattern compression), ``HKY``, ``SiteModel`` and ``TreeLikelihood`` with the
same Input names (``data``, ``tree``, ``siteModel``, ``branchRateModel``).  [EXT: BEAST 2 core is
not under the reference tree; semantics per SURVEY.md Appendix C.]

Only input preparation lives here; peeling runs in csrc/k2_peel.cu through the C-ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib

# BEAST Nucleotide data type: A C G T(U) are states 0..3; every other code is an ambiguity or a
# gap and is treated as missing data when useAmbiguities=false (the default).
_NUC = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}


def encode_nucleotides(seq: str) -> np.ndarray:
    return np.array([_NUC.get(ch, 4) for ch in seq.upper()], np.int32)


class Alignment:
    """Site patterns of a nucleotide alignment: distinct columns + integer weights
    (BEAST ``Alignment.calcPatterns``: columns are sorted and duplicates merged)."""

    def __init__(self, sequences: Dict[str, str] = None, taxa: Sequence[str] = None,
                 states: np.ndarray = None, compress: bool = True):
        if states is None:
            taxa = list(taxa) if taxa is not None else list(sequences.keys())
            states = np.stack([encode_nucleotides(sequences[t]) for t in taxa])
        self.taxa = list(taxa) if taxa is not None else [str(i) for i in range(states.shape[0])]
        states = np.minimum(np.asarray(states, np.int32), 4)
        self.siteCount = states.shape[1]
        if compress:
            pat, inv, cnt = np.unique(states, axis=1, return_inverse=True, return_counts=True)
            self.patterns = np.ascontiguousarray(pat, np.int32)       # [taxa][patterns]
            self.patternWeight = cnt.astype(np.int32)
            self.patternIndex = np.asarray(inv).ravel()
        else:
            self.patterns = np.ascontiguousarray(states, np.int32)
            self.patternWeight = np.ones(states.shape[1], np.int32)
            self.patternIndex = np.arange(states.shape[1])

    def getPatternCount(self) -> int:
        return self.patterns.shape[1]

    def getTaxonCount(self) -> int:
        return self.patterns.shape[0]

    def shard(self, rank: int, world: int) -> "Alignment":
        """Contiguous pattern shard for rank `rank` of `world`."""
        P = self.getPatternCount()
        lo, hi = (P * rank) // world, (P * (rank + 1)) // world
        a = Alignment.__new__(Alignment)
        a.taxa, a.siteCount = self.taxa, self.siteCount
        a.patterns = np.ascontiguousarray(self.patterns[:, lo:hi])
        a.patternWeight = self.patternWeight[lo:hi].copy()
        a.patternIndex = None
        return a

    @staticmethod
    def from_fasta(path: str) -> "Alignment":
        seqs, name = {}, None
        for line in open(path):
            line = line.strip()
            if line.startswith(">"):
                name = line[1:].strip()
                seqs[name] = ""
            elif name is not None:
                seqs[name] += line
        return Alignment(seqs)


class HKY:
    """BEAST ``HKY``: Inputs ``kappa`` and ``frequencies``."""

    def __init__(self, kappa: float = 2.0, frequencies: Sequence[float] = (0.25,) * 4):
        self.kappa = float(kappa)
        self.frequencies = np.asarray(frequencies, np.float64)


class SiteModel:
    """BEAST ``SiteModel``: Inputs ``substModel``, ``mutationRate``, ``gammaCategoryCount``,
    ``shape``, ``proportionInvariant``.  Category rates/proportions follow
    SiteModel.calculateCategoryRates with the invariant class handled through constant
    patterns (TreeLikelihood calls setPropInvariantIsCategory(false)) [EXT, recalled]."""

    def __init__(self, substModel: HKY, mutationRate: float = 1.0, gammaCategoryCount: int = 0,
                 shape: float = 1.0, proportionInvariant: float = 0.0):
        self.substModel = substModel
        self.mutationRate = float(mutationRate)
        self.gammaCategoryCount = int(gammaCategoryCount)
        self.shape = float(shape)
        self.proportionInvariant = float(proportionInvariant)

    def getCategoryCount(self) -> int:
        return max(self.gammaCategoryCount, 1)

    def category_rates_and_proportions(self):
        k = self.getCategoryCount()
        propVariable = 1.0 - self.proportionInvariant
        if self.gammaCategoryCount > 1:
            from scipy.stats import gamma
            a = self.shape
            q = (2.0 * np.arange(k) + 1.0) / (2.0 * k)
            rates = gamma.ppf(q, a, scale=1.0 / a)
            mean = propVariable * rates.sum() / k
            rates = rates / mean
        else:
            rates = np.array([1.0 / propVariable])
        props = np.full(k, propVariable / k)
        return rates * self.mutationRate, props


class StrictClockModel:
    def __init__(self, rate: float = 1.0):
        self.rate = float(rate)


class TreeLikelihood:
    """BEAST ``TreeLikelihood`` over the C-ABI: Inputs ``data``, ``tree``, ``siteModel``,
    ``branchRateModel`` (strict clock; default rate 1 as in config 1)."""

    def __init__(self, data: Alignment, tree=None, siteModel: SiteModel = None,
                 branchRateModel: Optional[StrictClockModel] = None, ctx=None):
        self.data, self.tree, self.siteModel = data, tree, siteModel
        self.branchRateModel = branchRateModel
        self.ctx = ctx or _lib.default_context()
        self._h = _lib.VP()
        self.logP = float("nan")
        self.initAndValidate()

    def initAndValidate(self):
        L = _lib.lib()
        nT, P = self.data.getTaxonCount(), self.data.getPatternCount()
        if self.tree is not None and self.tree.getLeafNodeCount() != nT:
            raise ValueError("The number of nodes in the tree does not match the number of "
                             "sequences")
        self.nCat = self.siteModel.getCategoryCount()
        self.ctx.check(L.mtt_peel_create(self.ctx.handle, nT, P, self.nCat, C.byref(self._h)))
        st = _lib.as_c(self.data.patterns, np.int32)
        w = _lib.as_c(self.data.patternWeight, np.int32)
        self.ctx.check(L.mtt_peel_set_tip_states(self._h, _lib.ptr(st)))
        self.ctx.check(L.mtt_peel_set_pattern_weights(self._h, _lib.ptr(w)))
        self.update_models()

    def update_models(self):
        """Push site-model / clock / substitution parameters (after a parameter move)."""
        L = _lib.lib()
        rates, props = self.siteModel.category_rates_and_proportions()
        fr = _lib.as_c(self.siteModel.substModel.frequencies, np.float64)
        rate = 1.0 if self.branchRateModel is None else self.branchRateModel.rate
        self.ctx.check(L.mtt_peel_set_site_model(
            self._h, _lib.ptr(_lib.as_c(rates, np.float64)), _lib.ptr(_lib.as_c(props, np.float64)),
            _lib.ptr(fr), self.siteModel.proportionInvariant, rate))
        self.ctx.check(L.mtt_peel_set_hky(self._h, self.siteModel.substModel.kappa))

    def calculateLogP(self) -> float:
        t = self.tree
        self.logP = float(self.calculateLogPBatch(np.asarray(t.parent)[None],
                                                  np.asarray(t.height)[None])[0])
        return self.logP

    def calculateLogPBatch(self, parent: np.ndarray, height: np.ndarray) -> np.ndarray:
        """parent/height: [B][2*nTips-1] -> logL[B] over this handle's patterns."""
        par = _lib.as_c(parent, np.int32)
        h = _lib.as_c(height, np.float64)
        B = par.shape[0]
        out = np.empty(B, np.float64)
        self.ctx.check(_lib.lib().mtt_peel_logl_batch(self._h, B, _lib.ptr(par), _lib.ptr(h),
                                                      _lib.ptr(out)))
        return out

    def calculateLogPParts(self, parent: np.ndarray, height: np.ndarray) -> np.ndarray:
        """parts[3][B] over this handle's patterns (a pattern shard): the sum of BEAST's unscaled
        evaluation, the sum of its rescaled evaluation, and the number of tiles in which the
        unscaled one underflowed (mtt_peel_logl_parts_batch).  Summed over the shards,
        ``parts[2] > 0 ? parts[1] : parts[0]`` is the alignment's logL."""
        par = _lib.as_c(parent, np.int32)
        h = _lib.as_c(height, np.float64)
        B = par.shape[0]
        out = np.empty((3, B), np.float64)
        self.ctx.check(_lib.lib().mtt_peel_logl_parts_batch(self._h, B, _lib.ptr(par),
                                                            _lib.ptr(h), _lib.ptr(out)))
        return out

    def calculateLogPBatchParams(self, parent: np.ndarray, height: np.ndarray, kappa,
                                 branchRate=None) -> np.ndarray:
        """One (kappa, clock rate) per tree: B parameter proposals or B chains in one call
        (mtt_peel_logl_batch_params).  branchRate None = the handle's rate for every tree; it
        multiplies the category rates exactly as ``mutationRate * clock rate`` does in
        ``update_models``."""
        par = _lib.as_c(parent, np.int32)
        h = _lib.as_c(height, np.float64)
        B = par.shape[0]
        kap = _lib.as_c(np.broadcast_to(np.asarray(kappa, np.float64), (B,)), np.float64)
        rate = None if branchRate is None else _lib.as_c(
            np.broadcast_to(np.asarray(branchRate, np.float64), (B,)), np.float64)
        out = np.empty(B, np.float64)
        self.ctx.check(_lib.lib().mtt_peel_logl_batch_params(
            self._h, B, _lib.ptr(par), _lib.ptr(h), _lib.ptr(kap), _lib.ptr(rate), _lib.ptr(out)))
        return out

    def calculateLogPBatchRates(self, parent: np.ndarray, height: np.ndarray,
                                branchRates: np.ndarray) -> np.ndarray:
        """Relaxed clock: branchRates[B][2*nTips-1] = branchRateModel.getRateForBranch(node) for
        the branch above every node (mtt_peel_logl_batch_rates); replaces the strict-clock
        rate of the handle."""
        par = _lib.as_c(parent, np.int32)
        h = _lib.as_c(height, np.float64)
        r = _lib.as_c(branchRates, np.float64)
        if r.shape != h.shape:
            raise ValueError("branchRates must have the shape of height")
        B = par.shape[0]
        out = np.empty(B, np.float64)
        self.ctx.check(_lib.lib().mtt_peel_logl_batch_rates(self._h, B, _lib.ptr(par), _lib.ptr(h),
                                                            _lib.ptr(r), _lib.ptr(out)))
        return out

    def calculateLogPBatchMats(self, parent: np.ndarray, mats: np.ndarray) -> np.ndarray:
        """Explicit transition matrices mats[B][2n-1][nCat][16] (any substitution model)."""
        par = _lib.as_c(parent, np.int32)
        mt = _lib.as_c(mats, np.float64)
        B = par.shape[0]
        out = np.empty(B, np.float64)
        self.ctx.check(_lib.lib().mtt_peel_logl_batch_mats(self._h, B, _lib.ptr(par),
                                                           _lib.ptr(mt), _lib.ptr(out)))
        return out

    def getPatternLogLikelihoods(self) -> np.ndarray:
        t = self.tree
        out = np.empty(self.data.getPatternCount(), np.float64)
        self.ctx.check(_lib.lib().mtt_peel_pattern_logl(
            self._h, _lib.ptr(_lib.as_c(t.parent, np.int32)),
            _lib.ptr(_lib.as_c(t.height, np.float64)), _lib.ptr(out)))
        return out

    def chains(self, nChains: int) -> "PeelChains":
        """Resident chains with incremental updates over this handle's alignment."""
        return PeelChains(self, nChains)

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            _lib.lib().mtt_peel_destroy(self._h)
            self._h = _lib.VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _flags(a) -> np.ndarray:
    """Byte flags (nonzero = set) without a copy when `a` already is a byte / bool array."""
    a = np.asarray(a)
    if a.dtype == np.bool_ or a.dtype == np.uint8:
        return np.ascontiguousarray(a).view(np.uint8)
    return _lib.as_c(a != 0, np.uint8)


class PeelChains:
    """What BEAST's ``TreeLikelihood`` does between MCMC steps (dirty-node updates of the
    partials ``BeerLikelihoodCore`` keeps per node, ``store`` / ``restore``), for ``nChains``
    chains at once: the partials of every chain stay in HBM, ``update`` evaluates one proposal per
    chain by recomputing only the nodes above what changed, ``accept`` keeps or drops the
    proposals (mtt_peel_chains_*)."""

    def __init__(self, likelihood: TreeLikelihood, nChains: int):
        self.likelihood, self.ctx, self.nChains = likelihood, likelihood.ctx, int(nChains)
        self._s = _lib.VP()
        self.ctx.check(_lib.lib().mtt_peel_chains_create(likelihood.handle, self.nChains,
                                                         C.byref(self._s)))

    def update(self, parent: np.ndarray, height: np.ndarray, branchRates=None, kappa=None,
               clockRate=None, dirty=None, parts: bool = False) -> np.ndarray:
        """parent/height [nChains][2*nTips-1] of the proposals -> logL[nChains].  `dirty`
        (optional, same shape, nonzero = the node may differ from the chain's accepted tree,
        ``Node.isDirty()``) restricts the comparison to the flagged nodes.  ``parts=True`` (the
        likelihood holds a pattern shard): parts[3][nChains] for the all-reduce over shards, see
        ``parallel.pattern_sharded_logl_parts``."""
        K = self.nChains
        par = _lib.as_c(parent, np.int32)
        h = _lib.as_c(height, np.float64)
        if par.shape != h.shape or par.shape[0] != K:
            raise ValueError("parent/height must be [nChains][2*nTips-1]")
        r = None if branchRates is None else _lib.as_c(branchRates, np.float64)
        if r is not None and r.shape != h.shape:
            raise ValueError("branchRates must have the shape of height")
        kap = None if kappa is None else _lib.as_c(
            np.broadcast_to(np.asarray(kappa, np.float64), (K,)), np.float64)
        clk = None if clockRate is None else _lib.as_c(
            np.broadcast_to(np.asarray(clockRate, np.float64), (K,)), np.float64)
        d = None if dirty is None else _flags(dirty)
        if d is not None and d.shape != h.shape:
            raise ValueError("dirty must have the shape of height")
        out = np.empty((3, K) if parts else K, np.float64)
        fn = _lib.lib().mtt_peel_chains_update_parts if parts else _lib.lib().mtt_peel_chains_update
        self.ctx.check(fn(self._s, _lib.ptr(par), _lib.ptr(h), _lib.ptr(r), _lib.ptr(kap),
                          _lib.ptr(clk), _lib.ptr(d), _lib.ptr(out)))
        return out

    def update_sparse(self, off, node, parent, height, branchRates=None, kappa=None,
                      clockRate=None) -> np.ndarray:
        """The proposals as lists of changed nodes: chain k names node[off[k]:off[k+1]] with
        their new parent / height (/ clock rate) per list entry (mtt_peel_chains_update_sparse);
        everything else stays as accepted.  -> logL[nChains]."""
        K = self.nChains
        off = _lib.as_c(off, np.int64)
        node = _lib.as_c(node, np.int32)
        par = _lib.as_c(parent, np.int32)
        h = _lib.as_c(height, np.float64)
        if off.shape != (K + 1,) or not (node.shape == par.shape == h.shape == (int(off[-1]),)):
            raise ValueError("off must be [nChains+1]; node/parent/height one entry per listed node")
        r = None if branchRates is None else _lib.as_c(branchRates, np.float64)
        if r is not None and r.shape != h.shape:
            raise ValueError("branchRates must have one entry per listed node")
        kap = None if kappa is None else _lib.as_c(
            np.broadcast_to(np.asarray(kappa, np.float64), (K,)), np.float64)
        clk = None if clockRate is None else _lib.as_c(
            np.broadcast_to(np.asarray(clockRate, np.float64), (K,)), np.float64)
        out = np.empty(K, np.float64)
        self.ctx.check(_lib.lib().mtt_peel_chains_update_sparse(
            self._s, _lib.ptr(off), _lib.ptr(node), _lib.ptr(par), _lib.ptr(h), _lib.ptr(r),
            _lib.ptr(kap), _lib.ptr(clk), _lib.ptr(out)))
        return out

    def accept(self, mask=None) -> None:
        """mask[nChains] (None = all): nonzero keeps the chain's proposal, zero drops it."""
        m = None if mask is None else _flags(mask)
        if m is not None and m.shape != (self.nChains,):
            raise ValueError("mask must be [nChains]")
        self.ctx.check(_lib.lib().mtt_peel_chains_accept(self._s, _lib.ptr(m)))

    @property
    def last_entries(self) -> int:
        """Internal nodes recomputed by the last update, over all chains."""
        return int(_lib.lib().mtt_peel_chains_last_entries(self._s))

    def set_verify(self, on: bool = True) -> None:
        """Debugging aid for bindings: fail an update whose dirty mask misses a changed node."""
        self.ctx.check(_lib.lib().mtt_peel_chains_set_verify(self._s, int(on)))

    def set_timing(self, on: bool = True) -> None:
        self.ctx.check(_lib.lib().mtt_peel_chains_set_timing(self._s, int(on)))

    @property
    def last_device_ms(self) -> float:
        """Device time of the last update's launches (after ``set_timing``)."""
        return float(_lib.lib().mtt_peel_chains_last_device_ms(self._s))

    def close(self):
        if self._s:
            _lib.lib().mtt_peel_chains_destroy(self._s)
            self._s = _lib.VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def simulate_alignment(parent: np.ndarray, height: np.ndarray, nSites: int, kappa: float,
                       freqs: Sequence[float], clockRate: float, seed: int = 0,
                       missing: float = 0.0) -> np.ndarray:
    """Synthetic nucleotide alignment evolved down one tree under HKY (host-side benchmark
    input).  Returns states [nTips][nSites] (4 = missing with probability
    `missing`)."""
    rng = np.random.default_rng(seed)
    parent = np.asarray(parent)
    N = len(parent)
    n = (N + 1) // 2
    freqs = np.asarray(freqs, float)
    pR, pY = freqs[0] + freqs[2], freqs[1] + freqs[3]
    beta = 1.0 / (2.0 * (pR * pY + kappa * (freqs[0] * freqs[2] + freqs[1] * freqs[3])))
    Q = np.array([[0 if i == j else freqs[j] * (kappa if (i % 2) == (j % 2) else 1.0)
                   for j in range(4)] for i in range(4)]) * beta
    np.fill_diagonal(Q, -Q.sum(1))
    w, V = np.linalg.eig(Q)
    Vi = np.linalg.inv(V)
    root = int(np.nonzero(parent < 0)[0][0])
    kids = [[] for _ in range(N)]
    for v in range(N):
        if parent[v] >= 0:
            kids[int(parent[v])].append(v)
    order, i = [root], 0                               # parents before children
    while i < len(order):
        order.extend(kids[order[i]])
        i += 1
    seqs = np.zeros((N, nSites), np.int8)
    seqs[root] = rng.choice(4, size=nSites, p=freqs)
    for v in order:
        v = int(v)
        if v == root:
            continue
        p = int(parent[v])
        t = (height[p] - height[v]) * clockRate
        Pm = np.real((V * np.exp(w * t)) @ Vi)
        Pm = np.clip(Pm, 0, None)
        Pm /= Pm.sum(1, keepdims=True)
        cdf = np.cumsum(Pm, axis=1)
        u = rng.random(nSites)
        seqs[v] = (u[:, None] > cdf[seqs[p]]).sum(1).clip(0, 3)
    out = seqs[:n].astype(np.int32)
    if missing > 0:
        out[rng.random(out.shape) < missing] = 4
    return out
