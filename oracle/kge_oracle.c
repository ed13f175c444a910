/*
 * kge_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the reference's link-prediction hot path
 * (OpenKE native runtime + the scoring maths of the OpenKE / paper Python models).
 * It exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs can CHECK the CUDA product path.  Nothing in the product path may call into this file.
 *
 * Pinning: this restatement is checked (tests/test_oracle_cpu.py, tests/golden/make_golden.py)
 * against (a) the unmodified reference library oracle/_ref/Base.so built from
 * /root/reference/OpenKE/openke/base/Base.cpp (index tables, tph/hpt, filtered ranks, metric tuple,
 * and -- through the reference's own LCG stream -- the Bernoulli sampler output, bit for bit), and
 * (b) golden vectors produced by importing the reference's OpenKE PyTorch models in the build
 * container (tests/golden/ npz files).  The Philox stream itself is pinned by the Random123 known-answer
 * vectors (orc_philox_selftest).
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int64_t h, r, t; } orc_triple;

typedef struct orc_index {
    int64_t E, R;
    /* train side: OpenKE/openke/base/Reader.h:53-160 */
    int64_t n_train;            /* after de-duplication */
    orc_triple *train_head;     /* sorted (h,r,t)  == trainList == trainHead */
    orc_triple *train_tail;     /* sorted (t,r,h) */
    int64_t *lef_head, *rig_head, *lef_tail, *rig_tail;   /* inclusive ranges per entity */
    float *left_mean, *right_mean;                        /* tph, hpt */
    /* test side: Reader.h:167-257 */
    int64_t n_all;   orc_triple *all;    /* train(raw) + valid + test, sorted (h,r,t) */
    int64_t n_test;  orc_triple *test;   /* sorted (r,h,t) */
    int64_t n_valid; orc_triple *valid;  /* sorted (r,h,t) */
} orc_index;

/* ---- comparators: OpenKE/openke/base/Triple.h:9-23 ---- */
static int cmp_hrt(const void *pa, const void *pb) {
    const orc_triple *a = pa, *b = pb;
    if (a->h != b->h) return a->h < b->h ? -1 : 1;
    if (a->r != b->r) return a->r < b->r ? -1 : 1;
    if (a->t != b->t) return a->t < b->t ? -1 : 1;
    return 0;
}
static int cmp_trh(const void *pa, const void *pb) {
    const orc_triple *a = pa, *b = pb;
    if (a->t != b->t) return a->t < b->t ? -1 : 1;
    if (a->r != b->r) return a->r < b->r ? -1 : 1;
    if (a->h != b->h) return a->h < b->h ? -1 : 1;
    return 0;
}
static int cmp_rht(const void *pa, const void *pb) {
    const orc_triple *a = pa, *b = pb;
    if (a->r != b->r) return a->r < b->r ? -1 : 1;
    if (a->h != b->h) return a->h < b->h ? -1 : 1;
    if (a->t != b->t) return a->t < b->t ? -1 : 1;
    return 0;
}

static orc_triple *pack(const int64_t *h, const int64_t *t, const int64_t *r, int64_t n) {
    orc_triple *a = (orc_triple *)calloc((size_t)(n > 0 ? n : 1), sizeof(orc_triple));
    for (int64_t i = 0; i < n; i++) { a[i].h = h[i]; a[i].t = t[i]; a[i].r = r[i]; }
    return a;
}

/*
 * Build every table the hot path needs.
 * Train side restates importTrainFiles (Reader.h:53-160): sort by (h,r,t), drop duplicates (:92-105),
 * second copy sorted by (t,r,h) (:107-109), per-entity inclusive ranges (:112-140, rig defaults to -1),
 * tph/hpt (:142-159).  Test side restates importTestFiles (Reader.h:167-257): the membership list is
 * test + RAW train + valid sorted by (h,r,t) (:201-226); test and valid sorted by (r,h,t) (:227-228).
 */
orc_index *orc_index_create(int64_t E, int64_t R,
                            const int64_t *trh, const int64_t *trt, const int64_t *trr, int64_t ntr,
                            const int64_t *vah, const int64_t *vat, const int64_t *var_, int64_t nva,
                            const int64_t *teh, const int64_t *tet, const int64_t *ter, int64_t nte) {
    orc_index *ix = (orc_index *)calloc(1, sizeof(orc_index));
    ix->E = E; ix->R = R;

    orc_triple *raw = pack(trh, trt, trr, ntr);
    qsort(raw, (size_t)ntr, sizeof(orc_triple), cmp_hrt);
    int64_t n = 0;
    for (int64_t i = 0; i < ntr; i++)
        if (i == 0 || cmp_hrt(&raw[i], &raw[i - 1]) != 0) raw[n++] = raw[i];
    ix->n_train = n;
    ix->train_head = raw;
    ix->train_tail = (orc_triple *)calloc((size_t)(n > 0 ? n : 1), sizeof(orc_triple));
    memcpy(ix->train_tail, raw, (size_t)n * sizeof(orc_triple));
    qsort(ix->train_tail, (size_t)n, sizeof(orc_triple), cmp_trh);

    int64_t *freq_rel = (int64_t *)calloc((size_t)R, sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) freq_rel[raw[i].r]++;

    ix->lef_head = (int64_t *)calloc((size_t)E, sizeof(int64_t));
    ix->rig_head = (int64_t *)calloc((size_t)E, sizeof(int64_t));
    ix->lef_tail = (int64_t *)calloc((size_t)E, sizeof(int64_t));
    ix->rig_tail = (int64_t *)calloc((size_t)E, sizeof(int64_t));
    for (int64_t e = 0; e < E; e++) { ix->rig_head[e] = -1; ix->rig_tail[e] = -1; }
    for (int64_t i = 0; i < n; i++) {
        int64_t h = ix->train_head[i].h, t = ix->train_tail[i].t;
        if (i == 0 || ix->train_head[i - 1].h != h) ix->lef_head[h] = i;
        if (i == n - 1 || ix->train_head[i + 1].h != h) ix->rig_head[h] = i;
        if (i == 0 || ix->train_tail[i - 1].t != t) ix->lef_tail[t] = i;
        if (i == n - 1 || ix->train_tail[i + 1].t != t) ix->rig_tail[t] = i;
    }

    /* tph = freq[r] / #distinct heads of r ; hpt = freq[r] / #distinct tails of r, all in float32
       (Reader.h:142-159: REAL counters incremented by 1.0, then INT / REAL). */
    ix->left_mean = (float *)calloc((size_t)R, sizeof(float));
    ix->right_mean = (float *)calloc((size_t)R, sizeof(float));
    for (int64_t i = 0; i < n; i++) {
        const orc_triple *a = &ix->train_head[i];
        if (i == 0 || a->h != a[-1].h || a->r != a[-1].r) ix->left_mean[a->r] += 1.0f;
        const orc_triple *b = &ix->train_tail[i];
        if (i == 0 || b->t != b[-1].t || b->r != b[-1].r) ix->right_mean[b->r] += 1.0f;
    }
    for (int64_t r = 0; r < R; r++) {
        ix->left_mean[r] = (float)freq_rel[r] / ix->left_mean[r];
        ix->right_mean[r] = (float)freq_rel[r] / ix->right_mean[r];
    }
    free(freq_rel);

    ix->n_test = nte;  ix->test = pack(teh, tet, ter, nte);
    ix->n_valid = nva; ix->valid = pack(vah, vat, var_, nva);
    ix->n_all = nte + ntr + nva;
    ix->all = (orc_triple *)calloc((size_t)(ix->n_all > 0 ? ix->n_all : 1), sizeof(orc_triple));
    memcpy(ix->all, ix->test, (size_t)nte * sizeof(orc_triple));
    for (int64_t i = 0; i < ntr; i++) {
        ix->all[nte + i].h = trh[i]; ix->all[nte + i].t = trt[i]; ix->all[nte + i].r = trr[i];
    }
    memcpy(ix->all + nte + ntr, ix->valid, (size_t)nva * sizeof(orc_triple));
    qsort(ix->all, (size_t)ix->n_all, sizeof(orc_triple), cmp_hrt);
    qsort(ix->test, (size_t)nte, sizeof(orc_triple), cmp_rht);
    qsort(ix->valid, (size_t)nva, sizeof(orc_triple), cmp_rht);
    return ix;
}

void orc_index_destroy(orc_index *ix) {
    if (!ix) return;
    free(ix->train_head); free(ix->train_tail);
    free(ix->lef_head); free(ix->rig_head); free(ix->lef_tail); free(ix->rig_tail);
    free(ix->left_mean); free(ix->right_mean);
    free(ix->all); free(ix->test); free(ix->valid);
    free(ix);
}

int64_t orc_train_total(const orc_index *ix) { return ix->n_train; }
int64_t orc_test_total(const orc_index *ix) { return ix->n_test; }
int64_t orc_valid_total(const orc_index *ix) { return ix->n_valid; }
int64_t orc_triple_total(const orc_index *ix) { return ix->n_all; }

static void unpack(const orc_triple *a, int64_t n, int64_t *h, int64_t *t, int64_t *r) {
    for (int64_t i = 0; i < n; i++) { h[i] = a[i].h; t[i] = a[i].t; r[i] = a[i].r; }
}
void orc_get_test(const orc_index *ix, int64_t *h, int64_t *t, int64_t *r) { unpack(ix->test, ix->n_test, h, t, r); }
void orc_get_train(const orc_index *ix, int64_t *h, int64_t *t, int64_t *r) { unpack(ix->train_head, ix->n_train, h, t, r); }
void orc_get_means(const orc_index *ix, float *left_mean, float *right_mean) {
    memcpy(left_mean, ix->left_mean, (size_t)ix->R * sizeof(float));
    memcpy(right_mean, ix->right_mean, (size_t)ix->R * sizeof(float));
}

/* Membership of (h,r,t) in train+valid+test: OpenKE/openke/base/Corrupt.h:166-177 (_find). */
int orc_find(const orc_index *ix, int64_t h, int64_t t, int64_t r) {
    orc_triple key = { h, r, t };
    int64_t lo = 0, hi = ix->n_all;            /* first element >= key */
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (cmp_hrt(&ix->all[mid], &key) < 0) lo = mid + 1; else hi = mid;
    }
    return lo < ix->n_all && cmp_hrt(&ix->all[lo], &key) == 0;
}

/*
 * Raw and filtered counts for one query from a materialised score vector:
 * OpenKE/openke/base/Test.h:65-127 (testHead, side=0: candidates replace the head) and
 * :130-192 (testTail, side=1).  Strict '<' on float32, the true entity excluded by index.
 */
void orc_rank_from_scores(const orc_index *ix, const float *con, int side,
                          int64_t h, int64_t t, int64_t r, int64_t *raw, int64_t *filt) {
    int64_t truth = side == 0 ? h : t;
    float minimal = con[truth];
    int64_t s = 0, fs = 0;
    for (int64_t j = 0; j < ix->E; j++) {
        if (j == truth) continue;
        if (con[j] < minimal) {
            s++;
            int known = side == 0 ? orc_find(ix, j, t, r) : orc_find(ix, h, j, r);
            if (!known) fs++;
        }
    }
    *raw = s; *filt = fs;
}

/*
 * Type-constrained counts (Test.h:88-98 / 153-163): the same comparison restricted to the relation's head (side 0) or
 * tail (side 1) candidate set from type_constrain.txt (Reader.h:267-317: sorted per relation), walked with the
 * reference's merge pointer, so duplicates in the list count once and the true entity is excluded by index.
 */
void orc_rank_from_scores_constrained(const orc_index *ix, const float *con, int side, int64_t h, int64_t t, int64_t r,
                                      const int64_t *type_sorted, int64_t n_type, int64_t *raw, int64_t *filt) {
    int64_t truth = side == 0 ? h : t;
    float minimal = con[truth];
    int64_t s = 0, fs = 0, lef = 0;
    for (int64_t j = 0; j < ix->E; j++) {
        if (j == truth) continue;
        while (lef < n_type && type_sorted[lef] < j) lef++;
        if (lef < n_type && j == type_sorted[lef]) {
            if (con[j] < minimal) {
                s++;
                int known = side == 0 ? orc_find(ix, j, t, r) : orc_find(ix, h, j, r);
                if (!known) fs++;
            }
        }
    }
    *raw = s; *filt = fs;
}

/*
 * Metric accumulation exactly as the reference does it -- float32 globals (Test.h:13-20),
 * thresholds <10/<3/<1 (:102-107), rank = count+1 and 1.0/rank added in double then rounded to
 * float (:109-112), division by testTotal and (head+tail)/2 of the FILTERED values (:232-277).
 * acc layout (per side, 10 floats): filt{hit10,hit3,hit1,rank,rrank}, raw{hit10,hit3,hit1,rank,rrank}.
 */
void orc_metrics_add(float *acc, int64_t raw, int64_t filt) {
    acc[0] += filt < 10 ? 1.f : 0.f;
    acc[1] += filt < 3 ? 1.f : 0.f;
    acc[2] += filt < 1 ? 1.f : 0.f;
    acc[3] += (float)(filt + 1);
    acc[4] = (float)((double)acc[4] + 1.0 / (double)(filt + 1));
    acc[5] += raw < 10 ? 1.f : 0.f;
    acc[6] += raw < 3 ? 1.f : 0.f;
    acc[7] += raw < 1 ? 1.f : 0.f;
    acc[8] += (float)(raw + 1);
    acc[9] = (float)((double)acc[9] + 1.0 / (double)(raw + 1));
}
/* out = (mrr, mr, hit10, hit3, hit1), the order Tester.run_link_prediction returns (Tester.py:83-91). */
void orc_metrics_final(const float *head_acc, const float *tail_acc, int64_t test_total, float *out) {
    float l[5], r[5];
    for (int i = 0; i < 5; i++) { l[i] = head_acc[i] / (float)test_total; r[i] = tail_acc[i] / (float)test_total; }
    out[0] = (l[4] + r[4]) / 2; out[1] = (l[3] + r[3]) / 2;
    out[2] = (l[0] + r[0]) / 2; out[3] = (l[1] + r[1]) / 2; out[4] = (l[2] + r[2]) / 2;
}

/* ------------------------------------------------------------------------------------------
 * Filtered corruption: OpenKE/openke/base/Corrupt.h:7-44 (corrupt_head: keep (h,r), draw a tail
 * that completes no train triple) and :46-83 (corrupt_tail: keep (t,r), draw a head).
 * `rnd` is the raw 64-bit random word; the reference reduces it with '% x' (Random.h:24-29).
 * The draw is the tmp-th entity NOT in the sorted true run (order-statistic skip) => exactly
 * uniform over the complement, never a train triple, no rejection loop.
 * ------------------------------------------------------------------------------------------ */
static int64_t skip_draw(const orc_triple *a, int use_t, int64_t ll, int64_t rr, int64_t E, uint64_t rnd) {
    int64_t k = rr - ll + 1;
    int64_t tmp = (int64_t)(rnd % (uint64_t)(E - k));
#define V(i) (use_t ? a[i].t : a[i].h)
    if (tmp < V(ll)) return tmp;
    if (tmp > V(rr) - k) return tmp + k;
    /* last position p in [ll,rr] with V(p) - (p-ll) - 1 < tmp, i.e. p-ll+1 true entities lie below the draw */
    int64_t lo = ll, hi = rr + 1;
    while (lo + 1 < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (V(mid) - (mid - ll) - 1 < tmp) lo = mid; else hi = mid;
    }
#undef V
    return tmp + (lo - ll) + 1;
}

static void run_of(const orc_triple *a, int by_h, int64_t lo0, int64_t hi0, int64_t r, int64_t *ll, int64_t *rr) {
    (void)by_h;
    int64_t lo = lo0, hi = hi0 + 1;               /* first index in [lo0,hi0] with .r >= r */
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a[m].r >= r) hi = m; else lo = m + 1; }
    *ll = lo;
    lo = lo0; hi = hi0 + 1;                        /* first index with .r > r */
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a[m].r > r) hi = m; else lo = m + 1; }
    *rr = lo - 1;
}

int64_t orc_corrupt_head(const orc_index *ix, int64_t h, int64_t r, uint64_t rnd) {
    int64_t ll, rr;
    run_of(ix->train_head, 1, ix->lef_head[h], ix->rig_head[h], r, &ll, &rr);
    return skip_draw(ix->train_head, 1, ll, rr, ix->E, rnd);
}
int64_t orc_corrupt_tail(const orc_index *ix, int64_t t, int64_t r, uint64_t rnd) {
    int64_t ll, rr;
    run_of(ix->train_tail, 0, ix->lef_tail[t], ix->rig_tail[t], r, &ll, &rr);
    return skip_draw(ix->train_tail, 0, ll, rr, ix->E, rnd);
}

/* Bernoulli threshold in float32: Base.cpp:101,112-113 (prob = 1000*hpt/(hpt+tph); 500 without bern). */
static float bern_prob(const orc_index *ix, int64_t r, int bern) {
    if (!bern) return 500.0f;
    return 1000 * ix->right_mean[r] / (ix->right_mean[r] + ix->left_mean[r]);
}

/*
 * The reference sampler with the reference's own RNG, one worker thread's slice:
 * Base.cpp:78-159 (getBatch, val_loss=false, negRelRate=0) + Random.h:18-29 (64-bit LCG).
 * `state` is next_random[id]; rows [lef,rig) of the batch are produced.  Used only to pin this
 * restatement against oracle/_ref/Base.so bit for bit (the product uses Philox below).
 */
static uint64_t lcg(uint64_t *s) { *s = *s * 25214903917ULL + 11ULL; return *s; }
void orc_sample_lcg(const orc_index *ix, uint64_t *state, int64_t lef, int64_t rig,
                    int64_t B, int64_t neg, int mode, int bern,
                    int64_t *bh, int64_t *bt, int64_t *br, float *by) {
    for (int64_t b = lef; b < rig; b++) {
        int64_t i = (int64_t)(lcg(state) % (uint64_t)ix->n_train);
        orc_triple p = ix->train_head[i];
        bh[b] = p.h; bt[b] = p.t; br[b] = p.r; by[b] = 1;
        int64_t last = B;
        for (int64_t k = 0; k < neg; k++, last += B) {
            int keep_head;
            if (mode == 0) keep_head = (float)(lcg(state) % 1000ULL) < bern_prob(ix, p.r, bern);
            else keep_head = mode != -1;
            if (keep_head) { bh[b + last] = p.h; bt[b + last] = orc_corrupt_head(ix, p.h, p.r, lcg(state)); }
            else           { bh[b + last] = orc_corrupt_tail(ix, p.t, p.r, lcg(state)); bt[b + last] = p.t; }
            br[b + last] = p.r; by[b + last] = -1;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
 * generator Random123 and cuRAND ship).  The product's sampler replaces Random.h's LCG with it
 * (north_star: "counter-based Philox kernel"); this is its CPU replay.
 * ------------------------------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Random123 kat_vectors for philox4x32-10; returns 0 when all three match. */
int orc_philox_selftest(void) {
    static const uint32_t C[3][4] = { {0,0,0,0}, {~0u,~0u,~0u,~0u}, {0x243f6a88u,0x85a308d3u,0x13198a2eu,0x03707344u} };
    static const uint32_t K[3][2] = { {0,0}, {~0u,~0u}, {0xa4093822u,0x299f31d0u} };
    static const uint32_t X[3][4] = { {0x6627e8d5u,0xe169c58du,0xbc57ac4cu,0x9b00dbd8u},
                                      {0x408f276du,0x41c83b0eu,0xa20bc7c6u,0x6d5451fdu},
                                      {0xd16cfe09u,0x94fdccebu,0x5001e420u,0x24126ea1u} };
    int bad = 0;
    for (int i = 0; i < 3; i++) {
        uint32_t o[4]; orc_philox4x32_10(C[i], K[i], o);
        for (int j = 0; j < 4; j++) bad += o[j] != X[i][j];
    }
    return bad;
}

/*
 * CPU replay of the product's Philox sampler (same batch layout and Bernoulli rule as
 * Base.cpp:101-124; same corruption as Corrupt.h:7-83).  Stream definition (DESIGN.md "sampler"):
 *   key = (seed_lo, seed_hi);  ctr = (row b, slot, step_lo, (step_hi & 0xffff) | stream << 16)
 *   slot 0      : x -> positive index  i = (x1:x0) % trainTotal
 *   slot k+1    : x -> negative k:  keep_head = float(x0 % 1000) < prob_r ;  draw word = (x2:x1)
 */
void orc_sample_philox(const orc_index *ix, uint64_t seed, uint64_t step, uint32_t stream,
                       int64_t B, int64_t neg, int mode, int bern,
                       int64_t *bh, int64_t *bt, int64_t *br, float *by) {
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t c2 = (uint32_t)step, c3 = ((uint32_t)(step >> 32) & 0xffffu) | (stream << 16);
    for (int64_t b = 0; b < B; b++) {
        uint32_t ctr[4] = { (uint32_t)b, 0, c2, c3 }, x[4];
        orc_philox4x32_10(ctr, key, x);
        int64_t i = (int64_t)((((uint64_t)x[1] << 32) | x[0]) % (uint64_t)ix->n_train);
        orc_triple p = ix->train_head[i];
        bh[b] = p.h; bt[b] = p.t; br[b] = p.r; by[b] = 1;
        float prob = bern_prob(ix, p.r, bern);
        for (int64_t k = 0; k < neg; k++) {
            int64_t o = b + (k + 1) * B;
            ctr[1] = (uint32_t)(k + 1);
            orc_philox4x32_10(ctr, key, x);
            uint64_t word = ((uint64_t)x[2] << 32) | x[1];
            int keep_head = mode == 0 ? ((float)(x[0] % 1000u) < prob) : (mode != -1);
            if (keep_head) { bh[o] = p.h; bt[o] = orc_corrupt_head(ix, p.h, p.r, word); }
            else           { bh[o] = orc_corrupt_tail(ix, p.t, p.r, word); bt[o] = p.t; }
            br[o] = p.r; by[o] = -1;
        }
    }
}

/*
 * CPU replay of the product's SUBGRAPH sampler (the paper's neg_sample_fn, module/NegativeSampling.py:114-140 with
 * __normal_batch :321-349 and __corrupt_head/__corrupt_tail :351-375), same Philox stream as the kernel:
 *   ctr = (edge b, slot | attempt << 16, step_lo, (step_hi & 0xffff) | stream << 16)
 *   attempt 0, slot j = 1..neg : decision j : to_head = (x0 >> 8) * 2^-24 < prob   (prob 0.5, or hpt/(hpt+tph) with bern)
 *   nh = number of to_head decisions; output slot k <= nh corrupts the head, k > nh the tail (heads first, :128-133)
 *   attempt a >= 1, slot k : position (x1:x0) % n_nodes of node_list; rejected while the candidate's GLOBAL id completes a
 *   train triple with the kept entity (np.in1d against h_of_tr / t_of_hr, :360,373); 64 attempts, then a scan.
 * Layout: slot k of edge b at k * n_edges + b (the reference transposes its [n, 1+neg] arrays, :134-136); int32 (:138-139).
 */
static int train_has(const orc_index *ix, int64_t h, int64_t r, int64_t t) {
    orc_triple key = { h, r, t };
    return bsearch(&key, ix->train_head, (size_t)ix->n_train, sizeof(orc_triple), cmp_hrt) != NULL;
}
void orc_sample_subgraph_philox(const orc_index *ix, uint64_t seed, uint64_t step, uint32_t stream,
                                const int64_t *eh, const int64_t *et, const int64_t *er, int64_t n_edges,
                                const int64_t *nodes, int64_t n_nodes, const int64_t *l2g, int64_t n_local,
                                int64_t neg, int bern, int filter, int32_t *oh, int32_t *ot, int32_t *orel) {
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t c2 = (uint32_t)step, c3 = ((uint32_t)(step >> 32) & 0xffffu) | (stream << 16);
    for (int64_t b = 0; b < n_edges; b++) {
        int64_t h = eh[b], t = et[b], r = er[b];
        oh[b] = (int32_t)h; ot[b] = (int32_t)t; orel[b] = (int32_t)r;
        float prob = 0.5f;
        if (bern && r >= 0 && r < ix->R) prob = bern_prob(ix, r, 1) * 1e-3f;
        uint32_t ctr[4] = { (uint32_t)b, 0, c2, c3 }, x[4];
        int64_t nh = 0;
        for (int64_t j = 1; j <= neg; j++) {
            ctr[1] = (uint32_t)j;
            orc_philox4x32_10(ctr, key, x);
            nh += ((float)(x[0] >> 8) * 5.9604644775390625e-8f < prob) ? 1 : 0;
        }
        for (int64_t k = 1; k <= neg; k++) {
            int64_t o = k * n_edges + b, nhd = h, ntl = t;
            if (n_nodes > 0) {
                int corrupt_head = k <= nh;
                int64_t fixed_local = corrupt_head ? t : h;
                int64_t fg = (fixed_local >= 0 && fixed_local < n_local && l2g) ? l2g[fixed_local] : -1;
                int64_t pick = -1, pos = 0;
#define KNOWN(cand) ({ int64_t cg_ = ((cand) >= 0 && (cand) < n_local && l2g) ? l2g[(cand)] : -1; \
                       (fg >= 0 && fg < ix->E && cg_ >= 0) ? (corrupt_head ? train_has(ix, cg_, r, fg) : train_has(ix, fg, r, cg_)) : 0; })
                for (int a = 0; a < 64 && pick < 0; a++) {
                    ctr[1] = (uint32_t)k | ((uint32_t)(a + 1) << 16);
                    orc_philox4x32_10(ctr, key, x);
                    pos = (int64_t)((((uint64_t)x[1] << 32) | x[0]) % (uint64_t)n_nodes);
                    int64_t cand = nodes[pos];
                    if (!filter || !KNOWN(cand)) pick = cand;
                }
                for (int64_t st = 1; st <= n_nodes && pick < 0; st++) {
                    int64_t cand = nodes[(pos + st) % n_nodes];
                    if (!KNOWN(cand)) pick = cand;
                }
#undef KNOWN
                if (pick >= 0) { if (corrupt_head) nhd = pick; else ntl = pick; }
            }
            oh[o] = (int32_t)nhd; ot[o] = (int32_t)ntl; orel[o] = (int32_t)r;
        }
    }
}

/*
 * corrupt(h, r), OpenKE/openke/base/Corrupt.h:179-195, as a function of the random words it consumes: a tail drawn from the
 * relation's tail-type list [ll, rr) (importTypeFiles, Reader.h:267-317; `rand(ll, rr)` = word % (rr - ll) + ll,
 * Random.h:32-34), redrawn while _find says (h, r, tail) is a triple of any split; after 1000 failed draws
 * corrupt_head(0, h, r) (Corrupt.h:7-44) with one more word.  words[0..999] = the attempts, words[1000] = the fallback's
 * draw word; *used = attempts consumed.  Returns -1 when the fallback is needed and `fallback_word` is NULL.
 * Pinned to the compiled reference's own corrupt() through the libc rand() stream (tests/golden/make_golden_corrupt.py).
 */
typedef uint64_t (*orc_word_fn)(void *state, int attempt);
static int64_t corrupt_typed_core(const orc_index *ix, const int64_t *tail_ptr, const int64_t *tail_idx, int64_t h, int64_t r,
                                  orc_word_fn next, void *state, int64_t *used) {
    int64_t ll = tail_ptr[r], cnt = tail_ptr[r + 1] - ll;
    int loop = 0;
    for (; loop < 1000 && cnt > 0; loop++) {
        int64_t t = tail_idx[ll + (int64_t)(next(state, loop) % (uint64_t)cnt)];
        if (!orc_find(ix, h, t, r)) { if (used) *used = loop + 1; return t; }
    }
    if (used) *used = loop;
    uint64_t word = next(state, 1000);
    int64_t ll2, rr2;
    run_of(ix->train_head, 1, ix->lef_head[h], ix->rig_head[h], r, &ll2, &rr2);   /* empty run: rr2 < ll2 */
    return rr2 >= ll2 ? orc_corrupt_head(ix, h, r, word) : (int64_t)(word % (uint64_t)ix->E);
}

/* the reference's streams: a caller-provided array of libc rand() values consumed in order for the list draws, and thread 0's
   LCG (randd(0), Random.h:18-21; *lcg_state = next_random[0]) for the corrupt_head fallback */
typedef struct { const int64_t *w; int64_t pos, n; uint64_t *lcg_state; } word_array;
static uint64_t next_from_array(void *state, int attempt) {
    word_array *a = state;
    if (attempt == 1000) return lcg(a->lcg_state);
    return a->pos < a->n ? (uint64_t)a->w[a->pos++] : 0;
}
int64_t orc_corrupt_typed_words(const orc_index *ix, const int64_t *tail_ptr, const int64_t *tail_idx, const int64_t *qh,
                                const int64_t *qr, int64_t n, const int64_t *words, int64_t n_words, uint64_t *lcg_state,
                                int64_t *out) {
    word_array a = { words, 0, n_words, lcg_state };
    for (int64_t i = 0; i < n; i++) out[i] = corrupt_typed_core(ix, tail_ptr, tail_idx, qh[i], qr[i], next_from_array, &a, NULL);
    return a.pos;                                   /* words consumed */
}

/*
 * CPU replay of mre_corrupt_typed: the same function on the product's Philox stream
 *   ctr = (pair i lo, attempt | (i hi) << 16, step_lo, (step_hi & 0xffff) | stream << 16), key = seed, word = (x1:x0)
 *   attempt a = 0..999: the list draws;  attempt 1000: the draw word of the corrupt_head fallback.
 * tail_ptr [R+1] / tail_idx: the per-relation SORTED tail-type lists.
 */
typedef struct { uint32_t key[2], c0, hi16, c2, c3; } philox_pair;
static uint64_t next_from_philox(void *state, int attempt) {
    philox_pair *p = state;
    uint32_t ctr[4] = { p->c0, (uint32_t)attempt | p->hi16, p->c2, p->c3 }, x[4];
    orc_philox4x32_10(ctr, p->key, x);
    return ((uint64_t)x[1] << 32) | x[0];
}
void orc_corrupt_typed_philox(const orc_index *ix, uint64_t seed, uint64_t step, uint32_t stream, const int64_t *tail_ptr,
                              const int64_t *tail_idx, const int64_t *qh, const int64_t *qr, int64_t n, int64_t *out) {
    philox_pair p = { { (uint32_t)seed, (uint32_t)(seed >> 32) }, 0, 0, (uint32_t)step,
                      ((uint32_t)(step >> 32) & 0xffffu) | (stream << 16) };
    for (int64_t i = 0; i < n; i++) {
        p.c0 = (uint32_t)i; p.hi16 = (uint32_t)((uint64_t)i >> 32) << 16;
        out[i] = corrupt_typed_core(ix, tail_ptr, tail_idx, qh[i], qr[i], next_from_philox, &p, NULL);
    }
}

/* number of emitted negatives that are train triples (must be 0): property check helper */
int64_t orc_count_train_leaks(const orc_index *ix, const int64_t *bh, const int64_t *bt, const int64_t *br,
                              int64_t from, int64_t to) {
    int64_t leaks = 0;
    for (int64_t i = from; i < to; i++) {
        orc_triple key = { bh[i], br[i], bt[i] };
        if (bsearch(&key, ix->train_head, (size_t)ix->n_train, sizeof(orc_triple), cmp_hrt)) leaks++;
    }
    return leaks;
}

/* ------------------------------------------------------------------------------------------
 * Scoring maths in float32 with a FIXED, documented evaluation order (d = 0,1,2,... sequential,
 * no fused multiply-add except where written).  The reference expresses the same maths through
 * ATen reductions whose order it does not pin, so scores agree with torch to ~1 ulp
 * of the sum and ranks agree outside the 1e-5 tie band; the golden fixtures carry torch's own
 * scores for that comparison.
 * ------------------------------------------------------------------------------------------ */

/* F.normalize(x, 2, -1): x / max(||x||_2, 1e-12)   (OpenKE/openke/module/model/TransE.py:47-50) */
void orc_l2_normalize_rows(const float *x, int64_t n, int64_t D, float *out) {
    for (int64_t i = 0; i < n; i++) {
        float ss = 0.f;
        for (int64_t d = 0; d < D; d++) ss = fmaf(x[i * D + d], x[i * D + d], ss);
        float nrm = sqrtf(ss);
        if (nrm < 1e-12f) nrm = 1e-12f;
        for (int64_t d = 0; d < D; d++) out[i * D + d] = x[i * D + d] / nrm;
    }
}

/*
 * TransE 1-vs-all scores for one query: TransE.py:46-60 (_calc) with mode head_batch
 * (side 0: score_j = || e_j + (r - t) ||_p) or tail_batch (side 1: || (h + r) - e_j ||_p);
 * module/NegativeSampling.py:294-305 (paper evaluate) is the side-1, un-normalised, p=1 case.
 * `ent`, `rel` are the (already normalised, if norm_flag) tables.
 */
void orc_transe_scores(const float *ent, const float *rel, int64_t E, int64_t D, int p_norm, int side,
                       int64_t h, int64_t t, int64_t r, float *out) {
    const float *vr = rel + r * D, *vh = ent + h * D, *vt = ent + t * D;
    for (int64_t j = 0; j < E; j++) {
        const float *e = ent + j * D;
        float acc = 0.f;
        for (int64_t d = 0; d < D; d++) {
            float u = side == 0 ? e[d] + (vr[d] - vt[d]) : (vh[d] + vr[d]) - e[d];
            if (p_norm == 1) acc = acc + fabsf(u); else acc = fmaf(u, u, acc);
        }
        out[j] = p_norm == 1 ? acc : sqrtf(acc);
    }
}

/* DistMult.py:34-44,70-72: predict = -sum_d h*r*t  (head_batch: h*(r*t); tail_batch: (h*r)*t) */
void orc_distmult_scores(const float *ent, const float *rel, int64_t E, int64_t D, int side,
                         int64_t h, int64_t t, int64_t r, float *out) {
    const float *vr = rel + r * D, *vh = ent + h * D, *vt = ent + t * D;
    for (int64_t j = 0; j < E; j++) {
        const float *e = ent + j * D;
        float acc = 0.f;
        for (int64_t d = 0; d < D; d++)
            acc += side == 0 ? e[d] * (vr[d] * vt[d]) : (vh[d] * vr[d]) * e[d];
        out[j] = -acc;
    }
}

/* ComplEx.py:20-27,60-61: predict = -sum(h_re t_re r_re + h_im t_im r_re + h_re t_im r_im - h_im t_re r_im) */
void orc_complex_scores(const float *ent_re, const float *ent_im, const float *rel_re, const float *rel_im,
                        int64_t E, int64_t D, int side, int64_t h, int64_t t, int64_t r, float *out) {
    for (int64_t j = 0; j < E; j++) {
        int64_t hh = side == 0 ? j : h, tt = side == 0 ? t : j;
        const float *hr = ent_re + hh * D, *hi = ent_im + hh * D, *tr = ent_re + tt * D, *ti = ent_im + tt * D;
        const float *rr = rel_re + r * D, *ri = rel_im + r * D;
        float acc = 0.f;
        for (int64_t d = 0; d < D; d++)
            acc += hr[d] * tr[d] * rr[d] + hi[d] * ti[d] * rr[d] + hr[d] * ti[d] * ri[d] - hi[d] * tr[d] * ri[d];
        out[j] = -acc;
    }
}

/* The same ComplEx score in CONTRACTION form: tail query => -[a;b].[t_re;t_im] with a = h_re r_re - h_im r_im,
   b = h_im r_re + h_re r_im; head query => -[a';b'].[h_re;h_im] with a' = t_re r_re + t_im r_im, b' = t_im r_re - t_re r_im.
   Algebraically ComplEx.py:20-27; the roundings are those of a K = 2D dot product of a pre-multiplied query vector with the
   [re|im] entity rows, accumulated sequentially in float32 -- the association the tcgen05 path's exact re-score uses, so its
   COUNTS can be asserted bit for bit (the 4-term form above differs from it, and from torch's own reduction, by ~1e-7 relative). */
void orc_complex_scores_contracted(const float *ent_re, const float *ent_im, const float *rel_re, const float *rel_im,
                                   int64_t E, int64_t D, int side, int64_t h, int64_t t, int64_t r, float *out) {
    const int64_t f = side == 0 ? t : h;   /* the entity that stays fixed */
    const float *fr = ent_re + f * D, *fi = ent_im + f * D, *rr = rel_re + r * D, *ri = rel_im + r * D;
    float *v = (float *)malloc((size_t)(2 * D) * sizeof(float));
    for (int64_t d = 0; d < D; d++) {
        if (side) { v[d] = fr[d] * rr[d] - fi[d] * ri[d]; v[D + d] = fi[d] * rr[d] + fr[d] * ri[d]; }
        else      { v[d] = fr[d] * rr[d] + fi[d] * ri[d]; v[D + d] = fi[d] * rr[d] - fr[d] * ri[d]; }
    }
    for (int64_t j = 0; j < E; j++) {
        const float *er = ent_re + j * D, *ei = ent_im + j * D;
        float acc = 0.f;
        for (int64_t d = 0; d < D; d++) acc = acc + v[d] * er[d];
        for (int64_t d = 0; d < D; d++) acc = acc + v[D + d] * ei[d];
        out[j] = -acc;
    }
    free(v);
}

/* Paper candidate rank, true candidate at index 0: main.py:245-250
   rank = #(n < p) + floor(#(n == p) / 2) + 1 over candidates 1..C-1. */
int64_t orc_rank_ties_half(const float *scores, int64_t C) {
    int64_t lt = 0, eq = 0;
    for (int64_t i = 1; i < C; i++) { lt += scores[i] < scores[0]; eq += scores[i] == scores[0]; }
    return lt + eq / 2 + 1;
}

/* MarginLoss.py:24-28 on the strategy's layout (strategy/NegativeSampling.py:13-21):
   loss = mean_{b,k} max(p_b - n_{b,k}, -m) + m with n_{b,k} = score[B + k*B + b]; double accumulation. */
double orc_margin_loss(const float *score, int64_t B, int64_t neg, float margin) {
    double s = 0;
    for (int64_t k = 0; k < neg; k++)
        for (int64_t b = 0; b < B; b++) {
            float v = score[b] - score[B + k * B + b];
            s += v > -margin ? v : -margin;
        }
    return s / (double)(B * neg) + margin;
}
