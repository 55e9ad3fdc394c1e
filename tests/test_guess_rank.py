"""CPU check of the rule behind the batched path's guessed phase thresholds (vrod_b200/csrc/guess_rank.hpp, host-only C++
compiled here on its own).  After a phase the k' best keys of the n0 rows seen are known; the next phase filters at the
key of rank r < k'.  If the rows to come resemble the rows seen, the number of NEW keys below an r-th order statistic
is negative binomial NB(r, 1/g) (g = rows after / rows before), and the guess holds when r + that count >= k'.
guess_rank(k', g) must return the smallest r whose failure probability is below 1e-9 -- plus its margin -- and the
rule itself must hold up in a simulation with real order statistics."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from scipy.stats import nbinom

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "guess_rank.hpp"
extern "C" double probe_nb_cdf(int m, int r, double p) { return vrod::nb_cdf(m, r, p); }
extern "C" int probe_guess_rank(int kprime, double g) { return vrod::guess_rank(kprime, g); }
'''


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("guess_rank")
    src, so = d / "probe.cpp", d / "probe.so"
    src.write_text(SRC)
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-I{ROOT}/vrod_b200/csrc", str(src), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    lib.probe_nb_cdf.restype = ctypes.c_double
    lib.probe_nb_cdf.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double]
    lib.probe_guess_rank.restype = ctypes.c_int
    lib.probe_guess_rank.argtypes = [ctypes.c_int, ctypes.c_double]
    return lib


def test_nb_cdf_matches_scipy(lib):
    for r in (1, 7, 60, 200, 250):
        for g in (1.5, 3.5, 8.0, 64.0):
            for m in (0, 5, 100, 400):
                got, want = lib.probe_nb_cdf(m, r, 1.0 / g), nbinom.cdf(m, r, 1.0 / g)
                assert got == pytest.approx(want, rel=1e-9, abs=1e-300), (r, g, m)
    assert lib.probe_nb_cdf(-1, 5, 0.5) == 0.0


@pytest.mark.parametrize("kprime", [32, 64, 192, 256])
@pytest.mark.parametrize("g", [2.0, 3.5, 8.0, 16.0, 64.0])
def test_guess_rank_is_the_smallest_safe_rank_plus_margin(lib, kprime, g):
    r = lib.probe_guess_rank(kprime, g)
    smallest = next(x for x in range(1, kprime + 1) if nbinom.cdf(kprime - x - 1, x, 1.0 / g) < 1e-9)
    want = smallest + smallest // 16 + 2
    if want * 10 > kprime * 9:
        assert r == 0, "a guess that is nearly the k'-th key anyway is not worth the second select"
    else:
        assert r == want
        assert nbinom.cdf(kprime - r - 1, r, 1.0 / g) < 1e-9
        assert r < kprime


def test_no_guess_without_growth(lib):
    assert lib.probe_guess_rank(256, 1.0) == 0 and lib.probe_guess_rank(256, 1.2) == 0


def test_the_rule_holds_for_real_order_statistics():
    """Simulation: n0 rows seen, (g - 1) n0 rows to come, all i.i.d.; the guess is the r-th smallest of the seen ones.  With
    the failure probability relaxed to what 20000 trials can resolve, the observed failure rate must stay within it."""
    rng = np.random.default_rng(1)
    kprime, g, n0, trials = 64, 8.0, 4096, 20000
    r = next(x for x in range(1, kprime + 1) if nbinom.cdf(kprime - x - 1, x, 1.0 / g) < 1e-3)
    seen = np.sort(rng.random((trials, n0)), axis=1)[:, r - 1]              # the guessed thresholds
    new_below = rng.binomial(int((g - 1) * n0), seen)                         # new rows under each of them
    failures = int((r + new_below < kprime).sum())
    assert failures <= trials * 1e-3 * 3 + 3, failures                        # NB is the n0 -> infinity limit: binomial counts vary less
    assert new_below.mean() == pytest.approx(r * (g - 1), rel=0.05)           # ~ r (g - 1) candidates per query and phase
