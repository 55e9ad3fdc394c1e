"""CPU model of the batched path's phase logic with GUESSED thresholds (knn_batched.cu: launch_batched_search,
batched_finish_kernel), in plain numpy: what is kept, what is dropped, and why a guess is "verified, not trusted".

A phase filters its rows at a threshold published by the previous finish: the k'-th best surrogate so far (safe) or
the key of a smaller rank r (a guess).  The finish of the phase merges the survivors into the best k' keys and checks
that the new k'-th key is not above the threshold the rows were filtered with.  The model shows the two properties the
CUDA path relies on:
  (1) whenever every check passes, the final list IS the exact k' smallest surrogates -- guesses never cost a neighbour;
  (2) a guess that was too tight (rows arrive nearest-first) is always caught by the check of its own phase."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def guess_rank(tmp_path_factory):
    d = tmp_path_factory.mktemp("guess_model")
    src, so = d / "probe.cpp", d / "probe.so"
    src.write_text('#include "guess_rank.hpp"\nextern "C" int probe_guess_rank(int kprime, double g) { return vrod::guess_rank(kprime, g); }\n')
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-I{ROOT}/vrod_b200/csrc", str(src), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    lib.probe_guess_rank.restype = ctypes.c_int
    lib.probe_guess_rank.argtypes = [ctypes.c_int, ctypes.c_double]
    return lib.probe_guess_rank


def run_phases(surrogates, kprime, first, growth, guess_rank):
    """One query.  Returns (kept keys, per-phase candidate counts, index of the first phase whose check failed or None)."""
    n = len(surrogates)
    kept = np.sort(surrogates[:first])[:kprime]                    # start phase: everything passes, the finish keeps k'
    seen, thr, guessed, counts, failed = first, np.inf, False, [], None
    while seen < n:
        nxt = min(n, int(seen * growth))
        if n <= 64 * seen:                                          # the rest in one phase (launch_batched_search)
            nxt = n
        r = guess_rank(kprime, nxt / seen) if seen >= 8 * kprime and len(kept) == kprime else 0
        thr, guessed = (kept[r - 1], True) if r > 0 else (kept[-1] if len(kept) == kprime else np.inf, False)
        new = surrogates[seen:nxt]
        cand = new[new < thr]                                       # the tile kernel's filter: D = thr - v > 0
        counts.append(len(cand))
        kept = np.sort(np.concatenate([kept, cand]))[:kprime]       # the finish kernel's merge + select
        kth = kept[-1] if len(kept) == kprime else np.inf
        if guessed and not kth <= thr and failed is None:           # the check of batched_finish_kernel (prev_guess)
            failed = len(counts) - 1
        seen = nxt
    return kept, counts, failed


@pytest.mark.parametrize("kprime", [32, 192])
def test_guessed_phases_keep_the_exact_best_keys_on_exchangeable_rows(guess_rank, kprime):
    rng = np.random.default_rng(kprime)
    n, first = 400_000, 4736
    for trial in range(20):
        v = rng.standard_normal(n).astype(np.float32) * 3 + 40      # surrogates of one query over the rows, random order
        kept, counts, failed = run_phases(v, kprime, first, 8.0, guess_rank)
        assert failed is None
        assert np.array_equal(kept, np.sort(v)[:kprime])
        safe = run_phases(v, kprime, first, 8.0, lambda kp, g: 0)[1]
        assert sum(counts) < (0.8 if kprime < 64 else 0.6) * sum(safe), (counts, safe)   # fewer candidates than filtering at the k'-th key


def test_a_guess_that_was_too_tight_is_caught_in_its_own_phase(guess_rank):
    rng = np.random.default_rng(3)
    v = np.sort(rng.standard_normal(400_000).astype(np.float32))    # rows arrive nearest-first: every guess is too tight
    kept, counts, failed = run_phases(v, 64, 4736, 8.0, guess_rank)
    assert failed == 0, "the first guessed phase must flag the query"
    # ... and the same rows filtered at the k'-th key (what the collection falls back to) lose nothing
    kept, counts, failed = run_phases(v, 64, 4736, 8.0, lambda kp, g: 0)
    assert failed is None and np.array_equal(kept, v[:64])


def test_whenever_every_check_passes_nothing_was_lost(guess_rank):
    """Property (1) on hostile inputs: drifting row distributions, reckless guesses (rank 1 = the best key so far).  The
    check may fail -- but if it does not, the kept keys are exactly the k' smallest."""
    rng = np.random.default_rng(9)
    passed = 0
    for trial in range(200):
        n = int(rng.integers(20_000, 60_000))
        drift = rng.uniform(-2.0, 2.0)
        v = (rng.standard_normal(n) + drift * np.linspace(0, 1, n)).astype(np.float32)
        rank = int(rng.integers(1, 40))
        kept, counts, failed = run_phases(v, 48, 1024, float(rng.uniform(2, 16)), lambda kp, g, rank=rank: min(rank, kp - 1))
        if failed is None:
            passed += 1
            assert np.array_equal(kept, np.sort(v)[:48])
    assert 0 < passed < 200, "the hostile inputs should produce both outcomes"
