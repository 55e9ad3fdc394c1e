/*
 * knn_oracle.c -- CPU oracle for vRod's SEARCH hot path (exact top-k nearest neighbour scan).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. The shipped library
 * (vrod_b200/libvrod_knn.so) never links, loads or calls anything in this directory.
 *
 * PARITY UNPINNED: the mounted reference (/root/reference) has no search arithmetic to restate.
 * The call site this oracle stands in for is the empty body of SearchCommand::execute
 * (reference src/command/types.rs:114-119) over a Database collection that is a TODO comment
 * (reference src/database/mod.rs:6-10).  The reference holds no test, golden vector or fixture
 * (SURVEY.md section 4), the path's arithmetic lives in no third-party dependency
 * (fastembed 3.4.0, Cargo.toml:11, only produces embeddings), and the Rust toolchain is absent,
 * so nothing of the reference can be run here.  What the reference does pin is the element type
 * (f32 rows, `Vec<Vec<f32>>`, reference src/utils/embeddings.rs:29-31) and the record text
 * format (`f32,f32,...;payload`, reference src/utils/embeddings.rs:52-62).  The semantics below
 * are therefore this repo's own written contract (DESIGN.md "Search semantics"), authored from
 * SURVEY.md section 8(c); the KATs in tests/golden/ are authored here as well.
 *
 * Canonical arithmetic (what ids and distances are graded against):
 *   - all sums are accumulated in f64 over the f32 inputs with fused multiply-add, in a FIXED
 *     128-way interleaved order: partial[j mod 128] takes element j (increasing j), and the 128
 *     partials are then combined by an adjacent-pair tree (p[i] = p[2i] + p[2i+1], 7 levels).
 *     The order is part of the contract, so every implementation that follows it is bit-identical
 *     (f64 add/fma/sqrt/div are correctly rounded everywhere).
 *   - Euclidean: dist = (f32) sqrt( SUM (x_j - q_j)^2 )
 *   - Cosine:    dist = (f32) (1 - dot / (sqrt(nx) * sqrt(nq))), nx = SUM x_j^2, nq = SUM q_j^2;
 *                if nx == 0 or nq == 0 the similarity is defined as 0, i.e. dist = 1.
 *   - ranking key = (f32 dist ascending, id ascending); slots beyond min(k, N) are padded with
 *     id = UINT64_MAX, dist = +inf.
 * Naive-f32 mode (informational only): sequential f32 accumulation without contraction, the way
 * an unoptimised `iter().zip().map().sum::<f32>()` Rust loop would run.
 *
 * Build: see oracle/Makefile (gcc -O3 -ffp-contract=off -mfma ..., OpenMP for the threaded scan).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))
#define NPART 128

/* ------------------------------------------------------------------------------------------
 * Synthetic data: Philox4x32-10 counter-based generator (Salmon et al., SC'11), restated from
 * the published algorithm.  Element (row i, col j) of a collection with logical dim d:
 *   e = i*d + j;  block = philox(counter = (lo32(e>>2), hi32(e>>2), 0, 0), key = (lo32(seed), hi32(seed)))
 *   word = block[e & 3];  value = ((int32)(word >> 8) - 2^23) * 2^-23   in [-1, 1)
 * SURVEY.md section 8(d).
 * ---------------------------------------------------------------------------------------- */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                 uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

ORACLE_API void vrod_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

static inline float word_to_unit(uint32_t w) {
    return (float)((int32_t)(w >> 8) - (1 << 23)) * 0x1p-23f;
}

/* rows[0 .. n*d) <- synthetic rows row0 .. row0+n of the collection seeded `seed`. */
ORACLE_API void vrod_oracle_fill(float *rows, uint64_t row0, uint64_t n, uint32_t d, uint64_t seed) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint64_t e0 = row0 * d, e1 = (row0 + n) * d;
    const uint64_t b0 = e0 >> 2, b1 = (e1 + 3) >> 2;
#pragma omp parallel for schedule(static)
    for (uint64_t b = b0; b < b1; ++b) {
        uint32_t w[4];
        philox4x32_10((uint32_t)b, (uint32_t)(b >> 32), 0, 0, k0, k1, w);
        for (int c = 0; c < 4; ++c) {
            uint64_t e = (b << 2) + c;
            if (e >= e0 && e < e1) rows[e - e0] = word_to_unit(w[c]);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Canonical f64 sums, 128-way interleaved + adjacent-pair tree.
 * ---------------------------------------------------------------------------------------- */
static inline double tree128(double *p) {
    for (int n = NPART; n > 1; n >>= 1)
        for (int i = 0; i < n / 2; ++i) p[i] = p[2 * i] + p[2 * i + 1];
    return p[0];
}

static double canon_sqdist(const float *x, const float *q, uint32_t d) {
    double p[NPART];
    for (int i = 0; i < NPART; ++i) p[i] = 0.0;
    uint32_t j = 0;
    for (; j + NPART <= d; j += NPART)
        for (int i = 0; i < NPART; ++i) {
            double diff = (double)x[j + i] - (double)q[j + i];
            p[i] = __builtin_fma(diff, diff, p[i]);
        }
    for (int i = 0; j + i < d; ++i) {
        double diff = (double)x[j + i] - (double)q[j + i];
        p[i] = __builtin_fma(diff, diff, p[i]);
    }
    return tree128(p);
}

static double canon_dot(const float *x, const float *q, uint32_t d) {
    double p[NPART];
    for (int i = 0; i < NPART; ++i) p[i] = 0.0;
    uint32_t j = 0;
    for (; j + NPART <= d; j += NPART)
        for (int i = 0; i < NPART; ++i) p[i] = __builtin_fma((double)x[j + i], (double)q[j + i], p[i]);
    for (int i = 0; j + i < d; ++i) p[i] = __builtin_fma((double)x[j + i], (double)q[j + i], p[i]);
    return tree128(p);
}

static inline float canon_l2(const float *x, const float *q, uint32_t d) {
    return (float)sqrt(canon_sqdist(x, q, d)) + 0.0f;
}

static inline float canon_cos(const float *x, const float *q, uint32_t d, double nq) {
    double nx = canon_dot(x, x, d);
    if (nx == 0.0 || nq == 0.0) return 1.0f;
    double dot = canon_dot(x, q, d);
    double den = sqrt(nx) * sqrt(nq);
    double sim = dot / den;
    return (float)(1.0 - sim) + 0.0f;
}

ORACLE_API float vrod_oracle_distance(const float *x, const float *q, uint32_t d, int metric) {
    if (metric == 0) return canon_l2(x, q, d);
    return canon_cos(x, q, d, canon_dot(q, q, d));
}

/* Naive f32: sequential, no fma (compiled with -ffp-contract=off). */
static inline float naive_l2(const float *x, const float *q, uint32_t d) {
    float acc = 0.0f;
    for (uint32_t j = 0; j < d; ++j) { float t = x[j] - q[j]; acc = acc + t * t; }
    return sqrtf(acc);
}
static inline float naive_cos(const float *x, const float *q, uint32_t d) {
    float dot = 0.0f, nx = 0.0f, nq = 0.0f;
    for (uint32_t j = 0; j < d; ++j) { dot = dot + x[j] * q[j]; nx = nx + x[j] * x[j]; nq = nq + q[j] * q[j]; }
    if (nx == 0.0f || nq == 0.0f) return 1.0f;
    return 1.0f - dot / (sqrtf(nx) * sqrtf(nq));
}

/* ------------------------------------------------------------------------------------------
 * Bounded top-k list ordered by (dist asc, id asc).
 * ---------------------------------------------------------------------------------------- */
typedef struct { float dist; uint64_t id; } hit_t;

static inline int hit_less(float da, uint64_t ia, float db, uint64_t ib) {
    return da < db || (da == db && ia < ib);
}

typedef struct { hit_t *h; uint32_t k, n; } topk_t;

static inline void topk_push(topk_t *t, float dist, uint64_t id) {
    if (t->k == 0) return;
    if (t->n == t->k && !hit_less(dist, id, t->h[t->n - 1].dist, t->h[t->n - 1].id)) return;
    uint32_t pos = t->n < t->k ? t->n : t->k - 1;
    while (pos > 0 && hit_less(dist, id, t->h[pos - 1].dist, t->h[pos - 1].id)) {
        t->h[pos] = t->h[pos - 1];
        --pos;
    }
    t->h[pos].dist = dist; t->h[pos].id = id;
    if (t->n < t->k) t->n++;
}

/*
 * Exact top-k of `b` queries over `n` rows (row-major n x d, ids = id_base + row unless `ids`).
 * mode 0 = canonical f64, mode 1 = naive f32.  nthreads <= 0 -> all OpenMP threads.
 * Returns 0, or 1 on invalid arguments.
 */
ORACLE_API int vrod_oracle_search(const float *rows, const uint64_t *ids, uint64_t n, uint32_t d, int metric,
                                  const float *queries, uint32_t b, uint32_t k, uint64_t id_base, int mode,
                                  int nthreads, uint64_t *out_ids, float *out_dist) {
    if (d == 0 || (metric != 0 && metric != 1) || (mode != 0 && mode != 1)) return 1;
    if (n > 0 && rows == NULL) return 1;
    int nt = 1;
#ifdef _OPENMP
    nt = nthreads > 0 ? nthreads : omp_get_max_threads();
#endif
    if ((uint64_t)nt > n) nt = n ? (int)n : 1;
    hit_t *scratch = (hit_t *)malloc(sizeof(hit_t) * (size_t)(k ? k : 1) * (size_t)nt);
    if (!scratch) return 1;
    for (uint32_t qi = 0; qi < b; ++qi) {
        const float *q = queries + (size_t)qi * d;
        const double nq = (metric == 1 && mode == 0) ? canon_dot(q, q, d) : 0.0;
#pragma omp parallel num_threads(nt)
        {
            int t = 0, T = 1;
#ifdef _OPENMP
            t = omp_get_thread_num(); T = omp_get_num_threads();
#endif
            topk_t tk = { scratch + (size_t)t * k, k, 0 };
            uint64_t lo = n * (uint64_t)t / T, hi = n * (uint64_t)(t + 1) / T;
            for (uint64_t r = lo; r < hi; ++r) {
                const float *x = rows + (size_t)r * d;
                float dist;
                if (mode == 0) dist = metric == 0 ? canon_l2(x, q, d) : canon_cos(x, q, d, nq);
                else           dist = metric == 0 ? naive_l2(x, q, d) : naive_cos(x, q, d);
                topk_push(&tk, dist, ids ? ids[r] : id_base + r);
            }
            /* pad the thread's list so the merge below can read k slots */
            for (uint32_t s = tk.n; s < k; ++s) { tk.h[s].dist = INFINITY; tk.h[s].id = UINT64_MAX; }
        }
        hit_t *fin = (hit_t *)malloc(sizeof(hit_t) * (k ? k : 1));
        topk_t m = { fin, k, 0 };
        for (int t = 0; t < nt; ++t)
            for (uint32_t s = 0; s < k; ++s) {
                hit_t h = scratch[(size_t)t * k + s];
                if (h.id == UINT64_MAX && isinf(h.dist)) break;
                topk_push(&m, h.dist, h.id);
            }
        for (uint32_t s = 0; s < k; ++s) {
            out_ids[(size_t)qi * k + s]  = s < m.n ? fin[s].id : UINT64_MAX;
            out_dist[(size_t)qi * k + s] = s < m.n ? fin[s].dist : INFINITY;
        }
        free(fin);
    }
    free(scratch);
    return 0;
}

/*
 * Merge `g` per-shard result lists (each b x k, padded as above) into one b x k list with the
 * same (dist, id) order -- the CPU statement of the cross-GPU merge.
 */
ORACLE_API void vrod_oracle_merge(const uint64_t *ids, const float *dist, uint32_t g, uint32_t b, uint32_t k,
                                  uint64_t *out_ids, float *out_dist) {
    hit_t *fin = (hit_t *)malloc(sizeof(hit_t) * (k ? k : 1));
    for (uint32_t qi = 0; qi < b; ++qi) {
        topk_t m = { fin, k, 0 };
        for (uint32_t s = 0; s < g; ++s)
            for (uint32_t j = 0; j < k; ++j) {
                size_t o = ((size_t)s * b + qi) * k + j;
                if (ids[o] == UINT64_MAX && isinf(dist[o])) continue;
                topk_push(&m, dist[o], ids[o]);
            }
        for (uint32_t j = 0; j < k; ++j) {
            out_ids[(size_t)qi * k + j]  = j < m.n ? fin[j].id : UINT64_MAX;
            out_dist[(size_t)qi * k + j] = j < m.n ? fin[j].dist : INFINITY;
        }
    }
    free(fin);
}

/* OpenMP team size for the calls that take no nthreads argument (vrod_oracle_fill): torchrun exports
 * OMP_NUM_THREADS=1 to its workers, bench.py's rank 0 sets the real core count back with this. */
ORACLE_API void vrod_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORACLE_API int vrod_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
