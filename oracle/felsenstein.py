"""CPU restatement of the reference's Felsenstein-pruning path -- TEST INFRASTRUCTURE ONLY.

This is the oracle of the parity tests: a numpy (IEEE double, no FMA) restatement of
what bpp v4.8.7 computes on the path  P-matrix build -> CLV update -> root lnL.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it;
the product path (bpp_b200/) never does and fails loudly without its CUDA library.

Parity of this oracle is PINNED: tests/test_oracle_golden.py checks every function
below against outputs of the reference itself (compiled unmodified by oracle/Makefile
into oracle/_ref/libbppref.so), stored as fixtures under tests/golden/ by
tests/golden/make_golden.py, and -- when oracle/_ref is present -- against the live
reference on fresh random inputs.  The reference ships no golden vectors of its own
for this path.

All file:line citations are relative to /root/reference/src.
numpy elementwise arithmetic is correctly rounded, so writing the sums in the
association order of the reference's AVX kernels reproduces their CLVs bit for bit
(given bit-identical P-matrices).
"""
import numpy as np

SCALE_FACTOR = 2.0 ** 256          # PLL_SCALE_FACTOR, bpp.h:376
SCALE_THRESHOLD = 2.0 ** -256      # PLL_SCALE_THRESHOLD, bpp.h:377
SCALE_NONE = -1                    # PLL_SCALE_BUFFER_NONE, bpp.h:380


# ---------------------------------------------------------------- tips (a2)
def tip_clv(masks, states, rate_cats):
    """set_tipclv, locus.c:525-559: clv[i,r,j] = (mask_i >> j) & 1, replicated per category."""
    masks = np.asarray(masks, dtype=np.uint32)
    if np.any(masks == 0):
        raise ValueError("Illegal state code in tip")     # locus.c:538 fatal()
    bits = ((masks[:, None] >> np.arange(states, dtype=np.uint32)[None, :]) & 1).astype(np.float64)
    return np.repeat(bits[:, None, :], rate_cats, axis=1).copy()       # [P, R, S]


def tip_clv_from_values(values, states, rate_cats):
    """pll_set_tip_clv, locus.c:596-619: copy S doubles per site into every category."""
    v = np.asarray(values, dtype=np.float64).reshape(-1, states)
    return np.repeat(v[:, None, :], rate_cats, axis=1).copy()


# ---------------------------------------------------------------- P-matrices
def branch_length(parent_time, node_time, rate_mui):
    """strict clock, core_pmatrix.c:711-715 / locus.c:2347-2351."""
    return (parent_time - node_time) * rate_mui


def pmatrix_jc69(t, rates):
    """locus_update_matrices_jc69, locus.c:2325-2415 (exp form, not expm1). -> [R,4,4]"""
    R = len(rates)
    out = np.zeros((R, 4, 4))
    for n in range(R):
        bl = t * rates[n]
        if bl < 1e-100:
            out[n] = np.eye(4)
        else:
            a = (1 + 3 * np.exp(-4 * bl / 3)) / 4
            b = (1 - a) / 3
            out[n] = np.full((4, 4), b)
            np.fill_diagonal(out[n], a)
    return out


def pmatrix_closed(model, t, rates, freqs, subst):
    """Closed-form DNA models other than JC69: locus_update_matrices_k80 (locus.c:2256-2323), _f81
    (:2190-2255), _tn93 for HKY / F84 / TN93 (:2068-2164), _t92 (:1981-2066).  Same expressions, same
    order; none of them has a zero-length special case. -> [R,4,4]"""
    R = len(rates)
    out = np.zeros((R, 16))
    f, q = np.asarray(freqs, dtype=np.float64), np.asarray(subst, dtype=np.float64)
    for n in range(R):
        bl = t * rates[n]
        pm = out[n]
        if model == "K80":
            kappa = q[1] / q[0]
            e1 = np.expm1(-4 * bl / (kappa + 2))
            if abs(kappa - 1) < 1e-20:
                pm[:] = -e1 / 4
                pm[[0, 5, 10, 15]] = 1. + 3 / 4. * e1
            else:
                e2 = np.expm1(-2 * bl * (kappa + 1) / (kappa + 2))
                pm[:] = -e1 / 4
                pm[[0, 5, 10, 15]] = 1 + (e1 + 2 * e2) / 4
                pm[[2, 7, 8, 13]] = (e1 - 2 * e2) / 4
        elif model == "F81":
            beta = 1.0
            for j in range(4):
                beta -= f[j] * f[j]
            beta = 1. / beta
            e, em1 = np.exp(-beta * bl), np.expm1(-beta * bl)
            for j in range(4):
                for k in range(4):
                    pm[4 * j + k] = e - f[k] * em1 if j == k else -f[k] * em1
        elif model == "T92":
            GC = f[3] + f[2]
            e1 = np.expm1(-bl)
            e2 = np.expm1(-(q[0] / q[1] + 1) * bl / 2)
            pm[0] = -(1 - GC) / 2 * e1
            pm[1] = GC / 2 * e1 - GC * e2
            pm[2] = -GC / 2 * e1
            pm[3] = 1 + 0.5 * (1 - GC) * e1 + GC * e2
            pm[4] = -(1 - GC) / 2 * e1
            pm[5] = 1 + GC / 2 * e1 + (1 - GC) * e2
            pm[6] = -GC / 2 * e1
            pm[7] = (1 - GC) / 2 * e1 - (1 - GC) * e2
            pm[8] = 1 + 0.5 * (1 - GC) * e1 + GC * e2
            pm[9] = -GC / 2 * e1
            pm[10] = GC / 2 * e1 - GC * e2
            pm[11] = -(1 - GC) / 2 * e1
            pm[12] = (1 - GC) / 2 * e1 - (1 - GC) * e2
            pm[13] = -GC / 2 * e1
            pm[14] = 1 + GC / 2 * e1 + (1 - GC) * e2
            pm[15] = -(1 - GC) / 2 * e1
        elif model in ("HKY", "F84", "TN93"):
            A, C, G, T = f[0], f[1], f[2], f[3]
            Y, Rr = T + C, A + G
            if model == "HKY":
                kappa = q[1] / q[0]
                mr = 1 / (2 * T * C * kappa + 2 * A * G * kappa + 2 * Y * Rr)
                bt = bl * mr
                a1t = a2t = kappa * bt
            elif model == "F84":
                kappa = q[0] / q[1]
                mr = 1 / (2 * T * C * kappa + 2 * A * G * kappa + 2 * Y * Rr)
                bt = bl * mr
                a1t = (1 + kappa / Y) * bt
                a2t = (1 + kappa / Rr) * bt
            else:
                mr = 1 / (2 * T * C * q[0] + 2 * A * G * q[1] + 2 * Y * Rr)
                bt = bl * mr
                a1t = (q[0] / q[2]) * bt
                a2t = (q[1] / q[2]) * bt
            e1 = np.expm1(-bt)
            e2 = np.expm1(-(Rr * a2t + Y * bt))
            e3 = np.expm1(-(Y * a1t + Rr * bt))
            pm[0] = 1 + Y * A / Rr * e1 + G / Rr * e2
            pm[1] = -C * e1
            pm[2] = Y * G / Rr * e1 - G / Rr * e2
            pm[3] = -T * e1
            pm[4] = -A * e1
            pm[5] = 1 + (Rr * C * e1 + T * e3) / Y
            pm[6] = -G * e1
            pm[7] = (Rr * e1 - e3) * T / Y
            pm[8] = Y * A / Rr * e1 - A / Rr * e2
            pm[9] = -C * e1
            pm[10] = 1 + Y * G / Rr * e1 + A / Rr * e2
            pm[11] = -T * e1
            pm[12] = -A * e1
            pm[13] = (Rr * e1 - e3) * C / Y
            pm[14] = -G * e1
            pm[15] = 1 + (Rr * T * e1 + C * e3) / Y
        else:
            raise ValueError(model)
    return out.reshape(R, 4, 4)


CLOSED_FORM_MODELS = ("K80", "F81", "HKY", "T92", "TN93", "F84")


def ratematrix_sym(subst, freqs):
    """create_ratematrix, core_pmatrix.c:186-237: symmetrised Q with mean rate 1."""
    S = len(freqs)
    p = np.array(subst, dtype=np.float64)
    if p[-1] > 0.0:
        p = p / p[-1]                                   # :199-201
    q = np.zeros((S, S))
    k = 0
    for i in range(S):
        for j in range(i + 1, S):
            f = p[k]
            k += 1
            q[i, j] = q[j, i] = f * np.sqrt(freqs[i] * freqs[j])
            q[i, i] -= f * freqs[j]
            q[j, j] -= f * freqs[i]
    mean = 0.0
    for i in range(S):
        mean += freqs[i] * (-q[i, i])                   # :227-229
    return q / mean


def update_eigen(subst, freqs):
    """pll_update_eigen, core_pmatrix.c:239-297.

    The reference diagonalises with Householder + QL (mytred2/mytqli, :28-182); any
    orthonormal eigenbasis of the same symmetric matrix gives the same P(t) up to
    rounding, so the oracle uses LAPACK (numpy.linalg.eigh).  Returns
    (eigenvecs[S,S], inv_eigenvecs[S,S], eigenvals[S]) in the reference's layout:
    eigenvecs[i][j] = a[i][j]*sqrt(pi_j) (:288-290), inv_eigenvecs = a^T / sqrt(pi_i)
    (:271-285), where row i of `a` is eigenvector i."""
    freqs = np.asarray(freqs, dtype=np.float64)
    q = ratematrix_sym(subst, freqs)
    lam, vec = np.linalg.eigh(q)
    a = vec.T.copy()                                    # rows = eigenvectors
    sq = np.sqrt(freqs)
    eigenvecs = a * sq[None, :]
    inv_eigenvecs = a.T / sq[:, None]
    return eigenvecs, inv_eigenvecs, lam


def pmatrix_eigen(eigenvecs, inv_eigenvecs, eigenvals, t, rates):
    """bpp_core_update_pmatrix, core_pmatrix.c:674-783: P = I + (V^-1 diag(expm1(lambda*bl))) V,
    m-sum sequential starting from delta_jk (:760-771); bl < 1e-100 -> identity (:738-743)."""
    S = len(eigenvals)
    R = len(rates)
    out = np.zeros((R, S, S))
    for n in range(R):
        bl = t * rates[n]
        if bl < 1e-100:
            out[n] = np.eye(S)
            continue
        expd = np.expm1(eigenvals * bl)
        temp = inv_eigenvecs * expd[None, :]
        acc = np.eye(S)
        for m in range(S):
            acc = acc + temp[:, m][:, None] * eigenvecs[m, :][None, :]
        out[n] = acc
    return out


# ---------------------------------------------------------------- CLV update (a9)
def _matvec_avx(mat, clv):
    """rows of mat [R,S,S] times clv [P,R,S] in the association order of the AVX kernels:
    S == 4: (p0+p1)+(p2+p3), separate mul/add (core_partials_avx.c:423-473);
    generic S: four lane sums over columns == 0..3 (mod 4), accumulated in column order with
    mul then add, combined as (s0+s1)+(s2+s3) (core_partials_avx.c:1330-1567; the AVX2
    variant core_partials_avx2.c:666-726 fuses the mul+add, a <=1 ulp difference)."""
    P, R, S = clv.shape
    out = np.empty((P, R, S))
    for i in range(S):
        row = mat[None, :, i, :]                        # [1,R,S]
        prod = row * clv                                # [P,R,S]
        if S == 4:
            out[:, :, i] = (prod[:, :, 0] + prod[:, :, 1]) + (prod[:, :, 2] + prod[:, :, 3])
        else:
            lanes = np.zeros((P, R, 4))
            for j in range(0, S, 4):
                lanes = lanes + prod[:, :, j:j + 4]
            out[:, :, i] = (lanes[:, :, 0] + lanes[:, :, 1]) + (lanes[:, :, 2] + lanes[:, :, 3])
    return out


def _matvec_scalar(mat, clv):
    """portable C order, core_partials.c:713-717: ((0+p0)+p1)+..."""
    P, R, S = clv.shape
    out = np.zeros((P, R, S))
    for j in range(S):
        out = out + mat[None, :, :, j] * clv[:, :, j][:, :, None]
    return out


def update_partial_ii(lclv, rclv, lmat, rmat, lscaler=None, rscaler=None, scaling=False,
                      order="avx"):
    """pll_core_update_partial_ii, core_partials.c:585-756.

    lclv/rclv [P,R,S]; lmat/rmat [R,S,S] (row = parent state).  Returns (parent_clv,
    parent_scaler or None).  With scaling: parent_scaler = l + r (missing = 0,
    fill_parent_scaler :24-46); a site whose S*R entries are ALL < 2^-256 (strict, on the
    unscaled product; :720,739,747) is multiplied by 2^256 and its scaler incremented."""
    mv = _matvec_avx if order == "avx" else _matvec_scalar
    parent = mv(lmat, lclv) * mv(rmat, rclv)
    if not scaling:
        return parent, None
    P = parent.shape[0]
    sc = np.zeros(P, dtype=np.uint32)
    if lscaler is not None:
        sc = sc + np.asarray(lscaler, dtype=np.uint32)
    if rscaler is not None:
        sc = sc + np.asarray(rscaler, dtype=np.uint32)
    below = np.all(parent.reshape(P, -1) < SCALE_THRESHOLD, axis=1)
    parent[below] = parent[below] * SCALE_FACTOR
    sc = sc + below.astype(np.uint32)
    return parent, sc


# ---------------------------------------------------------------- root (a11, a12, a10)
def site_likelihoods(clv, freqs, rate_weights):
    """inner loops of pll_core_root_loglikelihood (core_likelihood.c:179-195; 4-state AVX
    hadd order (p0+p1)+(p2+p3), core_likelihood_avx.c:121-130): term = sum_j rw_j * (clv_j . pi)."""
    P, R, S = clv.shape
    prod = clv * np.asarray(freqs)[None, None, :]
    if S == 4:
        dot = (prod[:, :, 0] + prod[:, :, 1]) + (prod[:, :, 2] + prod[:, :, 3])
    else:
        lanes = np.zeros((P, R, 4))
        for j in range(0, S, 4):
            lanes = lanes + prod[:, :, j:j + 4]
        dot = (lanes[:, :, 0] + lanes[:, :, 1]) + (lanes[:, :, 2] + lanes[:, :, 3])
    term = np.zeros(P)
    for j in range(R):
        term = term + dot[:, j] * rate_weights[j]
    return term


def root_loglikelihood(clv, freqs, rate_weights, pattern_weights, scaler=None, persite=False):
    """pll_core_root_loglikelihood, core_likelihood.c:24-212: log(term) + scaler*log(2^-256)
    (:199-201), times the pattern weight, summed sequentially over sites."""
    with np.errstate(divide="ignore"):
        site = np.log(site_likelihoods(clv, freqs, rate_weights))
    if scaler is not None:
        s = np.asarray(scaler, dtype=np.float64)
        site = np.where(s != 0, site + s * np.log(SCALE_THRESHOLD), site)
    site = site * np.asarray(pattern_weights, dtype=np.float64)
    logl = 0.0
    for v in site:                                      # sequential, :211
        logl += v
    return (logl, site) if persite else logl


def root_likelihood_vector(clv, freqs, rate_weights):
    """pll_core_root_likelihood_vector, core_likelihood.c:214-408: per-site likelihood, no log,
    scaler and weights ignored (:390)."""
    return site_likelihoods(clv, freqs, rate_weights)


def diploid_loglikelihood(lh_vector, resolution_count, mapping, pattern_weights):
    """diploid branch of locus_root_loglikelihood, locus.c:2586-2615."""
    logl, k = 0.0, 0
    for i, cnt in enumerate(resolution_count):
        mean = 0.0
        for _ in range(int(cnt)):
            mean += lh_vector[int(mapping[k])]
            k += 1
        mean /= float(cnt)
        logl += np.log(mean) * float(pattern_weights[i])
    return logl


# ---------------------------------------------------------------- whole-locus driver
class OracleLocus:
    """A numpy locus with the reference's buffer and index scheme (locus_create
    locus.c:622-870 as called from method.c:4137-4147; index flips locus.c:24-26)."""

    def __init__(self, tips, sites, states, rate_cats, scaling, model="JC69"):
        T = self.tips = tips
        self.P, self.S, self.R = sites, states, rate_cats
        self.scaling = bool(scaling)
        self.model = model
        self.clv = [None] * (T + 2 * (T - 1))
        self.pmat = [None] * (2 * (2 * T - 2))
        self.scale = [np.zeros(sites, dtype=np.uint32) for _ in range(2 * (T - 1))] if scaling else []
        self.rates = np.ones(rate_cats)
        self.rate_weights = np.full(rate_cats, 1.0 / rate_cats)          # locus.c:845-848
        self.freqs = np.full(states, 1.0 / states)
        self.subst = np.ones(states * (states - 1) // 2)
        self.weights = np.ones(sites, dtype=np.uint32)
        self.eigen = None
        # gnode_t index fields, gtree.c:2395-2399,2664-2675
        nn = 2 * T - 1
        self.clv_index = list(range(nn))
        self.pmatrix_index = list(range(nn))
        self.scaler_index = [SCALE_NONE] * T + ([k for k in range(T - 1)] if scaling else [SCALE_NONE] * (T - 1))
        self.left = self.right = None
        self.parent = [-1] * nn
        self.times = np.zeros(nn)
        self.rate_mui = 1.0

    def set_tree(self, left, right, times, rate_mui=1.0):
        T = self.tips
        self.left, self.right = [int(x) for x in left], [int(x) for x in right]
        self.parent = [-1] * (2 * T - 1)
        for k in range(T - 1):
            self.parent[self.left[k]] = T + k
            self.parent[self.right[k]] = T + k
        self.times = np.array(times, dtype=np.float64)
        self.rate_mui = rate_mui
        self.root = [n for n in range(2 * T - 1) if self.parent[n] < 0][-1]

    def set_tip_masks(self, tip, masks):
        self.clv[tip] = tip_clv(masks, self.S, self.R)

    def set_tip_values(self, tip, values):
        self.clv[tip] = tip_clv_from_values(values, self.S, self.R)

    def set_model(self, freqs=None, subst=None, rates=None):
        if freqs is not None:
            self.freqs = np.array(freqs, dtype=np.float64)
            self.eigen = None
        if subst is not None:
            self.subst = np.array(subst, dtype=np.float64)
            self.eigen = None
        if rates is not None:
            self.rates = np.array(rates, dtype=np.float64)

    # SWAP_* macros, locus.c:24-26
    def flip_clv(self, node):
        T = self.tips
        self.clv_index[node] = T + (self.clv_index[node] - 1) % (2 * T - 2)
        if self.scaling:
            self.scaler_index[node] = (T + self.scaler_index[node] - 1) % (2 * T - 2)

    def flip_pmatrix(self, node):
        e = 2 * self.tips - 2
        self.pmatrix_index[node] = (e + self.pmatrix_index[node]) % (2 * e)

    def update_matrices(self, nodes):
        for n in nodes:
            t = branch_length(self.times[self.parent[n]], self.times[n], self.rate_mui)
            if self.model == "JC69":
                pm = pmatrix_jc69(t, self.rates)
            elif self.model in CLOSED_FORM_MODELS:
                pm = pmatrix_closed(self.model, t, self.rates, self.freqs, self.subst)
            else:
                if self.eigen is None:
                    self.eigen = update_eigen(self.subst, self.freqs)
                pm = pmatrix_eigen(*self.eigen, t, self.rates)
            self.pmat[self.pmatrix_index[n]] = pm

    def update_partials(self, nodes, order="avx"):
        T = self.tips
        for n in nodes:
            l, r = self.left[n - T], self.right[n - T]
            ls = self.scale[self.scaler_index[l]] if self.scaler_index[l] != SCALE_NONE else None
            rs = self.scale[self.scaler_index[r]] if self.scaler_index[r] != SCALE_NONE else None
            use = self.scaler_index[n] != SCALE_NONE
            clv, sc = update_partial_ii(self.clv[self.clv_index[l]], self.clv[self.clv_index[r]],
                                        self.pmat[self.pmatrix_index[l]], self.pmat[self.pmatrix_index[r]],
                                        ls, rs, scaling=use, order=order)
            self.clv[self.clv_index[n]] = clv
            if use:
                self.scale[self.scaler_index[n]] = sc

    def root_loglikelihood(self, persite=False):
        sc = self.scale[self.scaler_index[self.root]] if self.scaler_index[self.root] != SCALE_NONE else None
        return root_loglikelihood(self.clv[self.clv_index[self.root]], self.freqs, self.rate_weights,
                                  self.weights, sc, persite)

    def root_likelihood_vector(self):
        return root_likelihood_vector(self.clv[self.clv_index[self.root]], self.freqs, self.rate_weights)

    def post_order(self):
        T = self.tips
        out, stack = [], [(self.root, 0)]
        while stack:
            node, st = stack.pop()
            if node < T:
                continue
            if st == 0:
                stack.append((node, 1))
                stack.append((self.right[node - T], 0))
                stack.append((self.left[node - T], 0))
            else:
                out.append(node)
        return out

    def full_pass(self):
        edges = [n for n in range(2 * self.tips - 1) if self.parent[n] >= 0]
        self.update_matrices(edges)
        self.update_partials(self.post_order())
        return self.root_loglikelihood()


def locus_from_workload(w, i, char_map):
    """Build an OracleLocus for locus i of a bpp_b200.synth.Workload."""
    o = OracleLocus(w.tips, w.sites, w.states, w.rate_cats, w.scaling, w.model)
    o.set_tree(w.left[i], w.right[i], w.times[i], float(w.rate_mui[i]))
    for t in range(w.tips):
        o.set_tip_masks(t, char_map[w.tip_chars[i, t]])
    o.weights = w.weights[i].copy()
    o.set_model(freqs=w.freqs[i], subst=w.subst[i], rates=w.rates)
    return o
