// guess_rank.hpp -- the statistical rule behind the batched path's guessed phase thresholds (host only, no CUDA: the CPU
// tests compile it on its own, tests/test_guess_rank.py).
#pragma once
#include <cmath>

namespace vrod {

// Guessed thresholds.  After a phase the kprime best keys of the n0 rows seen so far are known exactly.  If the next phase
// takes the rows seen to g * n0 and the new rows resemble the old ones (exchangeable row order), the key of rank r among
// the old ones is a threshold below which the new rows contribute NB(r, 1/g) keys (negative binomial: Poisson counts
// whose rate has the Gamma(r) uncertainty of an r-th order statistic).  The guess holds when r + that count >= kprime.
// guess_rank returns the smallest r whose failure probability is below 1e-9 per query and phase, plus a margin for rows
// that are only roughly exchangeable; 0 when guessing gains nothing.  (kprime = 256, g = 8: r = 72 -- the phase collects
// ~500 candidates per query where the kprime-th best key as threshold lets ~1800 through.)
inline double nb_cdf(int m, int r, double pr) {   // P(NB(r, pr) <= m), in the log domain
    if (m < 0) return 0.0;
    const double lq = log1p(-pr);
    double lt = (double)r * log(pr), top = lt, acc = 1.0;   // the sum so far = exp(top) * acc
    for (int i = 0; i < m; ++i) {
        lt += log((double)(i + r) / (double)(i + 1)) + lq;
        if (lt > top) {
            acc = acc * exp(top - lt) + 1.0;
            top = lt;
        } else {
            acc += exp(lt - top);
        }
    }
    return exp(top) * acc;
}
inline int guess_rank(int kprime, double g) {
    if (!(g > 1.25)) return 0;
    struct Memo { int kprime; double g; int r; };
    static thread_local Memo memo[8] = {};
    static thread_local int memo_next = 0;
    for (const Memo &m : memo)
        if (m.kprime == kprime && m.g == g) return m.r;
    int lo = 1, hi = kprime;   // smallest r with P(r + NB(r, 1/g) < kprime) < tol; the probability falls with r
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (nb_cdf(kprime - mid - 1, mid, 1.0 / g) < 1e-9) hi = mid;
        else lo = mid + 1;
    }
    int r = lo + lo / 16 + 2;
    if (r * 10 > kprime * 9) r = 0;   // nearly the kprime-th key anyway
    memo[memo_next] = Memo{kprime, g, r};
    memo_next = (memo_next + 1) % 8;
    return r;
}

}  // namespace vrod
