"""bpp v4.8.7 ITSELF on the engine: the unmodified reference (oracle/_ref/libbppref.so) with the product's
interposer bpp_b200/host/locus_cuda.c in front of its locus seam (locus.c:2417,2523,2530,2573,
prop_mixing.c:52), built by oracle/Makefile into oracle/_ref/bpp_b200 (+ bpp_b200_check with the reference's own
CHECK_LOGL validator, method.c:30,4699-4717).

The reference's acceptance test is a byte-identical mcmc.txt across --arch values (test/runtest.py:291-299);
here: frogs A00 (BASELINE.json config 1; seed 12345, burnin 200, nsample 500 x sampfreq 2 = 1 200 iterations)
stock --arch avx2 against BPP_B200=1.
"""
import hashlib
import os
import re
import shutil
import subprocess
import tempfile

import pytest

from helpers import ROOT

REF = os.path.join(ROOT, "oracle", "_ref")
BIN = os.path.join(REF, "bpp_b200")
BIN_CHECK = os.path.join(REF, "bpp_b200_check")
DATA = os.path.join(REF, "examples", "frogs")
CTL = os.path.join(ROOT, "tests", "data", "frogs_A00.ctl")
LOG_L0 = "-7320.932289"
LOG_PG0 = "1714.812053"
MCMC_MD5_AVX2 = "adf21e02ff4a6b2f735f665effbcaf58"      # SURVEY.md 8c, reproduced here by the stock path

pytestmark = pytest.mark.skipif(not (os.path.exists(BIN) and os.path.isdir(DATA)),
                                reason="oracle/_ref/bpp_b200 not built (needs /root/reference at build time)")


# the full-length chain (burnin 200 + 500 samples x 2 = 1 200 iterations, ~950 000 seam triplets) takes about a minute
# per run on the engine; the default GPU suite runs 240 iterations, BPP_LONG=1 the full length (profiles/ has the
# builder's full-length report)
LONG = os.environ.get("BPP_LONG", "0") != "0"
BURNIN, NSAMPLE = (200, 500) if LONG else (40, 100)


CTL_GTR = os.path.join(ROOT, "tests", "data", "frogs_gtr_g4.ctl")      # the same data read as haploid, GTR+G4, scaling on


def run_bpp(binary, env_extra, extra_args=(), timeout=1500, full_length=False, ctl_path=None):
    d = tempfile.mkdtemp(prefix="bpp_run_")
    for f in os.listdir(DATA):
        shutil.copy(os.path.join(DATA, f), d)
    ctl = open(ctl_path or CTL).read()
    if not (LONG or full_length):
        ctl = re.sub(r"burnin = \d+", "burnin = %d" % BURNIN, ctl)
        ctl = re.sub(r"nsample = \d+", "nsample = %d" % NSAMPLE, ctl)
    with open(os.path.join(d, "frogs_A00.ctl"), "w") as f:
        f.write(ctl)
    env = dict(os.environ, **env_extra)
    r = subprocess.run([binary, "--cfile", "frogs_A00.ctl"] + list(extra_args), cwd=d, env=env, capture_output=True,
                       text=True, timeout=timeout)
    mcmc = open(os.path.join(d, "out.mcmc.txt")).read() if os.path.exists(os.path.join(d, "out.mcmc.txt")) else ""
    shutil.rmtree(d, ignore_errors=True)
    return r, mcmc


def log_l0(stdout):
    m = re.search(r"log-PG0 = (\S+)\s+log-L0 = (\S+)", stdout)
    return (m.group(1), m.group(2)) if m else (None, None)


def report(text):
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "bpp_binary_report.txt"), "a") as f:
            f.write(text + "\n")
    except OSError:
        pass
    print(text)


def test_stock_path_of_the_interposed_binary_is_the_reference():
    """Without BPP_B200 the binary must be bpp v4.8.7 bit for bit: known log-L0 and mcmc.txt md5."""
    r, mcmc = run_bpp(BIN, {}, ["--arch", "avx2"], full_length=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert log_l0(r.stdout) == (LOG_PG0, LOG_L0)
    assert hashlib.md5(mcmc.encode()).hexdigest() == MCMC_MD5_AVX2


def test_cuda_path_fails_loudly_without_a_device():
    if os.path.exists("/dev/nvidia0"):
        pytest.skip("a GPU is present")
    r, _ = run_bpp(BIN, {"BPP_B200": "1"})
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def first_divergence(a, b):
    la, lb = a.splitlines(), b.splitlines()
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return i, x, y
    if len(la) != len(lb):
        return min(len(la), len(lb)), "<eof>", "<eof>"
    return None


@pytest.mark.gpu
@pytest.mark.parametrize("batch,fuse", [("1", "1"), ("0", "1"), ("0", "0")])
def test_frogs_a00_mcmc_on_the_engine(batch, fuse):
    """log-L0 as printed by the reference; the whole chain against the stock AVX2 run.  mcmc.txt is byte-identical
    as long as no accept/reject decision falls inside the last-bits difference of the two lnL evaluations; if it
    ever does, the first divergent sample is reported and the chain must still agree up to there and stay sane."""
    import time
    t0 = time.perf_counter()
    ref, ref_mcmc = run_bpp(BIN, {}, ["--arch", "avx2"])
    ref_secs = time.perf_counter() - t0
    assert ref.returncode == 0
    t0 = time.perf_counter()
    r, mcmc = run_bpp(BIN, {"BPP_B200": "1", "BPP_B200_BATCH": batch, "BPP_B200_FUSE": fuse, "BPP_B200_VERBOSE": "1"})
    secs = time.perf_counter() - t0
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert log_l0(r.stdout) == (LOG_PG0, LOG_L0)
    stats = [l for l in r.stderr.splitlines() if l.startswith("[bpp_b200]")]
    assert stats, "the CUDA path did not report: was the engine used at all?"
    m = re.search(r"update_partials (\d+), root_loglikelihood (\d+) calls; (\d+) batched passes over (\d+) loci; (\d+) kernels", stats[0])
    assert m and int(m.group(1)) > 1000 and int(m.group(5)) > 3000
    if batch == "1":
        assert int(m.group(3)) > 50 and int(m.group(3)) <= int(m.group(4)) <= 5 * int(m.group(3))
    else:
        assert int(m.group(3)) == 0
    stats[0] += " | %d iterations in %.1f s wall (whole process), stock avx2 %.1f s" % (BURNIN + 2 * NSAMPLE, secs, ref_secs)
    div = first_divergence(ref_mcmc, mcmc)
    n = len(ref_mcmc.splitlines())
    if div is None:
        report("frogs A00, BPP_B200=1 BATCH=%s FUSE=%s: mcmc.txt byte-identical to --arch avx2 over %d lines (md5 %s); %s"
               % (batch, fuse, n, hashlib.md5(mcmc.encode()).hexdigest(), stats[0]))
        if LONG:
            assert hashlib.md5(mcmc.encode()).hexdigest() == MCMC_MD5_AVX2
    else:
        i, x, y = div
        report("frogs A00, BPP_B200=1 BATCH=%s FUSE=%s: mcmc.txt diverges from --arch avx2 at line %d of %d\n  avx2: %s\n  b200: %s\n  %s"
               % (batch, fuse, i, n, x, y, stats[0]))
        # the samples before the divergence are identical; afterwards the chain is a different but valid one:
        # same number of samples, lnL in the same range
        assert i > 20, "diverged almost at once: that is a wrong likelihood, not a borderline accept/reject"
        assert len(mcmc.splitlines()) == n
        lnl_ref = [float(l.split()[-1]) for l in ref_mcmc.splitlines()[1:]]
        lnl = [float(l.split()[-1]) for l in mcmc.splitlines()[1:]]
        assert abs(sum(lnl) / len(lnl) - sum(lnl_ref) / len(lnl_ref)) < 25.0


@pytest.mark.gpu
@pytest.mark.parametrize("batch", ["1", "0"])
def test_frogs_a00_check_logl_clean(batch):
    """The reference's own validator (CHECK_LOGL): after every move of every iteration the incrementally maintained
    per-locus lnL must equal a full recomputation within 1e-9 -- with BOTH computed by the engine through the
    seam, index flips, rejections and partial updates included."""
    if not os.path.exists(BIN_CHECK):
        pytest.skip("bpp_b200_check not built")
    r, mcmc = run_bpp(BIN_CHECK, {"BPP_B200": "1", "BPP_B200_BATCH": batch, "BPP_B200_VERBOSE": "1"})
    assert "FATAL" not in r.stdout and "Invalid logl" not in r.stderr, (r.stdout[-1500:], r.stderr[-1500:])
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert log_l0(r.stdout) == (LOG_PG0, LOG_L0)
    assert len(mcmc.splitlines()) == NSAMPLE + 1
    report("frogs A00, CHECK_LOGL build, BPP_B200=1 BATCH=%s: %d iterations clean" % (batch, BURNIN + 2 * NSAMPLE))


@pytest.mark.gpu
def test_frogs_gtr_gamma_with_scaling_on_the_engine():
    """The same sequences read as haploid data under GTR+G4 with per-site scaling: exercises the eigen-form P-matrices
    (the reference's own decomposition is handed over), the model moves (alpha, qrates, freqs write the struct
    directly) and the scaler buffers through the seam.  The chain must start from the reference's log-L0 and agree
    with the stock AVX2 chain sample by sample; CUDA's expm1 and glibc's differ by <= 1 ulp, so a byte-identical
    file is expected but not guaranteed -- a divergence is reported with the first differing sample."""
    ref, ref_mcmc = run_bpp(BIN, {}, ["--arch", "avx2"], ctl_path=CTL_GTR)
    assert ref.returncode == 0, ref.stderr[-1500:]
    r, mcmc = run_bpp(BIN, {"BPP_B200": "1", "BPP_B200_VERBOSE": "1"}, ctl_path=CTL_GTR)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert log_l0(r.stdout) == log_l0(ref.stdout) and log_l0(ref.stdout)[1] is not None
    div = first_divergence(ref_mcmc, mcmc)
    n = len(ref_mcmc.splitlines())
    if div is None:
        report("frogs haploid GTR+G4 scaling, BPP_B200=1: mcmc.txt byte-identical to --arch avx2 over %d lines, log-L0 %s"
               % (n, log_l0(r.stdout)[1]))
    else:
        i, x, y = div
        report("frogs haploid GTR+G4 scaling, BPP_B200=1: mcmc.txt diverges from --arch avx2 at line %d of %d\n  avx2: %s\n  b200: %s"
               % (i, n, x, y))
        assert i > 10 and len(mcmc.splitlines()) == n


@pytest.mark.gpu
def test_batched_alpha_move_check_logl_and_chain():
    """BPP_B200_BATCH_ALPHA=1: the alpha move of all loci as one batch (propose all, evaluate once, decide per locus;
    prop_gamma.c:53-226).  Draw order differs from the reference, so the chain is another realisation: it must pass
    the reference's CHECK_LOGL validator at every move and sample the same posterior region as the stock chain."""
    if not os.path.exists(BIN_CHECK):
        pytest.skip("bpp_b200_check not built")
    ref, ref_mcmc = run_bpp(BIN, {}, ["--arch", "avx2"], ctl_path=CTL_GTR)
    r, mcmc = run_bpp(BIN_CHECK, {"BPP_B200": "1", "BPP_B200_BATCH_ALPHA": "1", "BPP_B200_VERBOSE": "1"}, ctl_path=CTL_GTR)
    assert "FATAL" not in r.stdout and "Invalid logl" not in r.stderr, (r.stdout[-1500:], r.stderr[-1500:])
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert len(mcmc.splitlines()) == len(ref_mcmc.splitlines())
    lnl_ref = [float(l.split()[-1]) for l in ref_mcmc.splitlines()[1:]]
    lnl = [float(l.split()[-1]) for l in mcmc.splitlines()[1:]]
    half = len(lnl) // 2
    m_ref, m = sum(lnl_ref[half:]) / (len(lnl_ref) - half), sum(lnl[half:]) / (len(lnl) - half)
    report("frogs haploid GTR+G4, batched alpha move, CHECK_LOGL clean over %d iterations; mean lnL of the second half "
           "%.2f (stock chain %.2f)" % (BURNIN + 2 * NSAMPLE, m, m_ref))
    assert abs(m - m_ref) < 40.0
