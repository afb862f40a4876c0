"""Seeded synthetic loci for the benchmark configs of BASELINE.json.

Pure numpy; no reference code, no GPU.  The same arrays feed the CUDA engine,
the CPU checker and (in tests/bench) the compiled reference, so every path sees
identical inputs.

Conventions (the reference's, gtree.c:2395-2399,2664-2675):
  node ids 0..T-1 are tips, T..2T-2 inner nodes in creation order (children are
  created before parents, the root is node 2T-2).  Initially clv_index =
  pmatrix_index = node id and scaler_index = id-T for inner nodes, -1 for tips.
"""
from dataclasses import dataclass

import numpy as np

SEED = 20261017

NT_ALPHABET = b"ACGT"
AA_ALPHABET = b"ARNDCQEGHILKMFPSTWYV"   # bit order of the reference's pll_map_aa (maps.c:126)

# discrete-Gamma(alpha=0.5, 4 cats, mean method) rates as printed by the reference's
# pll_compute_gamma_cats (gamma.c:221) -- see tests/golden/make_golden.py, which re-derives
# them; bpp_b200.gamma.discrete_gamma_rates() reproduces them on the host.
GAMMA4_ALPHA_0_5 = np.array([0.03338775338361239, 0.2519159176270096,
                             0.8202684819537311, 2.894427847035647])


def iupac_nt_map():
    """256-entry char -> 4-bit state mask, the standard IUPAC nucleotide code
    (A=1 C=2 G=4 T=8; same convention as the reference's pll_map_nt, maps.c:26)."""
    m = np.zeros(256, dtype=np.uint32)
    code = {"A": 1, "C": 2, "G": 4, "T": 8, "U": 8, "M": 3, "R": 5, "W": 9, "S": 6, "Y": 10,
            "K": 12, "V": 7, "H": 11, "D": 13, "B": 14, "N": 15, "O": 15, "X": 15, "-": 15, "?": 15}
    for ch, v in code.items():
        m[ord(ch)] = v
        m[ord(ch.lower())] = v
    return m


def aa_map():
    """256-entry char -> 20-bit state mask (reference: pll_map_aa, maps.c:126)."""
    m = np.zeros(256, dtype=np.uint32)
    for k, ch in enumerate(AA_ALPHABET.decode()):
        m[ord(ch)] = 1 << k
        m[ord(ch.lower())] = 1 << k
    amb = {"B": (1 << 2) | (1 << 3), "Z": (1 << 5) | (1 << 6)}
    for ch, v in amb.items():
        m[ord(ch)] = v
        m[ord(ch.lower())] = v
    for ch in "X*-?":
        m[ord(ch)] = 0xFFFFF
    m[ord("x")] = 0xFFFFF
    return m


@dataclass
class Workload:
    """N uniform loci: T tips, P site patterns, R rate categories, S states."""
    name: str
    n_loci: int
    tips: int
    sites: int
    states: int
    rate_cats: int
    model: str                 # "JC69" | "GTR" | "LG"
    scaling: bool
    left: np.ndarray           # [N, T-1] int32 node ids
    right: np.ndarray          # [N, T-1] int32
    times: np.ndarray          # [N, 2T-1] float64 node ages (tips 0)
    rate_mui: np.ndarray       # [N] float64
    tip_chars: np.ndarray      # [N, T, P] uint8 sequence characters
    weights: np.ndarray        # [N, P] uint32 pattern weights
    freqs: np.ndarray          # [N, S]
    subst: np.ndarray          # [N, S(S-1)/2] exchangeabilities (unused for JC69)
    rates: np.ndarray          # [R] category rates
    seed: int = SEED

    @property
    def inner(self):
        return self.tips - 1

    @property
    def edges(self):
        return 2 * self.tips - 2

    def b_pass(self):
        """Canonical algorithmic bytes of one locus full-tree pass."""
        T, P, R, S = self.tips, self.sites, self.rate_cats, self.states
        b = (T - 1) * 3 * (P * R * S * 8) + P * R * S * 8 + P * 4 + 2 * (2 * T - 2) * R * S * S * 8
        if self.scaling:
            b += (T - 1) * 3 * P * 4 + P * 4
        return b

    def b_min(self):
        """Compulsory traffic of the tree-fused kernel with packed tips."""
        T, P, R, S = self.tips, self.sites, self.rate_cats, self.states
        b = (T - 1) * P * R * S * 8 + T * P * (1 if S == 4 else 4) + P * 4 + (2 * T - 2) * R * S * S * 8
        if self.scaling:
            b += (T - 1) * P * 4
        return b

    def post_order(self, i):
        """Recursive left,right,node traversal of locus i (prop_mixing.c:28-50)."""
        T = self.tips
        out, stack = [], [(2 * T - 2, 0)]
        L, Rr = self.left[i], self.right[i]
        while stack:
            node, state = stack.pop()
            if node < T:
                continue
            if state == 0:
                stack.append((node, 1))
                stack.append((int(Rr[node - T]), 0))
                stack.append((int(L[node - T]), 0))
            else:
                out.append(node)
        return out

    def subset(self, n):
        """First n loci (used for the bounded CPU-baseline sample)."""
        n = min(n, self.n_loci)
        return Workload(self.name, n, self.tips, self.sites, self.states, self.rate_cats, self.model,
                        self.scaling, self.left[:n], self.right[:n], self.times[:n], self.rate_mui[:n],
                        self.tip_chars[:n], self.weights[:n], self.freqs[:n], self.subst[:n],
                        self.rates, self.seed)


def random_trees(rng, n, tips, dt_lo=0.0005, dt_hi=0.025):
    """n random coalescent-join trees; join k creates node tips+k at an age that
    increases with k (increments U(dt_lo, dt_hi))."""
    T = tips
    left = np.zeros((n, T - 1), dtype=np.int32)
    right = np.zeros((n, T - 1), dtype=np.int32)
    times = np.zeros((n, 2 * T - 1))
    active = np.tile(np.arange(T, dtype=np.int32), (n, 1))
    rows = np.arange(n)
    age = np.zeros(n)
    for k in range(T - 1):
        m = T - k
        a = rng.integers(0, m, size=n)
        b = rng.integers(0, m - 1, size=n)
        b = np.where(b >= a, b + 1, b)
        left[:, k] = active[rows, a]
        right[:, k] = active[rows, b]
        age = age + rng.uniform(dt_lo, dt_hi, size=n)
        times[:, T + k] = age
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        active[rows, lo] = T + k
        active[rows, hi] = active[rows, m - 1]
    return left, right, times


def make_workload(name, n_loci, tips, sites, states=4, rate_cats=1, model="JC69", scaling=False,
                  seed=SEED, ambiguity=0.02, dt_lo=0.0005, dt_hi=0.025, lg=None, rates=None):
    rng = np.random.Generator(np.random.PCG64(seed))
    left, right, times = random_trees(rng, n_loci, tips, dt_lo, dt_hi)
    alpha = np.frombuffer(NT_ALPHABET if states == 4 else AA_ALPHABET, dtype=np.uint8)
    idx = rng.integers(0, states, size=(n_loci, tips, sites), dtype=np.uint8)
    chars = alpha[idx]
    if ambiguity > 0:
        amb = rng.random(size=chars.shape, dtype=np.float32) < ambiguity
        chars = np.where(amb, np.uint8(ord("N") if states == 4 else ord("X")), chars)
    weights = rng.integers(1, 6, size=(n_loci, sites)).astype(np.uint32)
    nsub = states * (states - 1) // 2
    if model == "JC69":
        freqs = np.full((n_loci, states), 1.0 / states)
        subst = np.ones((n_loci, nsub))
    elif model == "GTR":
        f = rng.uniform(0.8, 1.2, size=(n_loci, states))
        freqs = f / f.sum(axis=1, keepdims=True)
        subst = rng.uniform(0.5, 1.5, size=(n_loci, nsub))
        subst[:, -1] = 1.0
    elif model in ("K80", "F81", "HKY", "T92", "TN93", "F84"):
        # closed-form DNA models (locus.c:1981-2324): qrates[0..2] carry kappa-like parameters the way the
        # reference reads them; K80 has equal base frequencies, T92 is parameterised by its GC content
        if model == "K80":
            freqs = np.full((n_loci, states), 0.25)
        elif model == "T92":
            gc = rng.uniform(0.35, 0.65, size=n_loci)
            freqs = np.stack([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2], axis=1)   # T C A G
        else:
            f = rng.uniform(0.6, 1.4, size=(n_loci, states))
            freqs = f / f.sum(axis=1, keepdims=True)
        subst = rng.uniform(0.5, 4.0, size=(n_loci, nsub))
    elif model == "LG":
        assert lg is not None, "pass lg=(rates190, freqs20)"
        subst = np.tile(np.asarray(lg[0], dtype=np.float64), (n_loci, 1))
        freqs = np.tile(np.asarray(lg[1], dtype=np.float64), (n_loci, 1))
    else:
        raise ValueError(model)
    if rates is not None:
        rates = np.array(rates, dtype=np.float64)
        assert rates.size == rate_cats
    elif rate_cats == 1:
        rates = np.ones(1)
    elif rate_cats == 4:
        rates = GAMMA4_ALPHA_0_5.copy()
    else:
        raise ValueError("pass rates= for rate_cats other than 1 or 4")
    return Workload(name, n_loci, tips, sites, states, rate_cats, model, scaling, left, right, times,
                    np.ones(n_loci), np.ascontiguousarray(chars), weights, freqs, subst, rates, seed)


# BASELINE.json configs (T for configs 4 and 5 is not given there; 8 and 16 are assumed)
CONFIGS = {
    "config2": dict(n_loci=10000, tips=8, sites=1000, states=4, rate_cats=1, model="JC69"),
    "config3": dict(n_loci=10000, tips=16, sites=1000, states=4, rate_cats=4, model="GTR"),
    "config4": dict(n_loci=2000, tips=8, sites=500, states=20, rate_cats=4, model="LG"),
    "config5": dict(n_loci=50000, tips=16, sites=2000, states=4, rate_cats=4, model="GTR"),
}


def make_config(name, n_loci=None, scaling=False, seed=SEED, lg=None):
    kw = dict(CONFIGS[name])
    if n_loci is not None:
        kw["n_loci"] = n_loci
    return make_workload(name, scaling=scaling, seed=seed, lg=lg, **kw)
