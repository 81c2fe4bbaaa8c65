/*
 * hnsw.c -- HNSW insert + search with DistHamming (test infrastructure, see gso.h).
 *
 * Follows hnsw_rs 0.3 `hnsw.rs` [U, high for the algorithm, medium for tie-breaks;
 * SURVEY A.10] as driven by the reference:
 *   Hnsw::new(M, capacity, 16, ef_c, DistHamming{})       src/dna/dnasketch.rs:139
 *   modify_level_scale(scale)                              src/dna/dnasketch.rs:141
 *   set_extend_candidates(true), set_keeping_pruned(false) src/dna/dnasketch.rs:159-160
 *   parallel_insert(&[(&Vec<Sig>, id)])                    src/dna/dnasketch.rs:435
 *   parallel_search(&[Vec<Sig>], knbn, ef=5000)            src/dna/dnarequest.rs:353,
 *                                                          src/bin/gsearch.rs:893
 *
 * Restated pieces: LayerGenerator::generate (level = floor(-ln(U) * scale), scale =
 * scale_modification / ln(M), re-drawn uniformly if >= max_layer), insert_slice,
 * search_layer, select_neighbours (Malkov Alg. 4 with extendCandidates only on layer 0),
 * reverse_update_neighborhood_simple (push, sort, drop the farthest when over
 * M / 2M), search (one greedy hop per upper layer, then search_layer(max(ef,knbn), 0),
 * into_sorted_vec, first knbn).
 *
 * Heaps reproduce Rust's std::collections::BinaryHeap exactly (push = sift_up; pop =
 * swap-with-last + sift_down_to_bottom + sift_up; into_sorted_vec = sift_down_range), with
 * the ordering of hnsw_rs PointWithOrder (distance only), because Hamming distances tie
 * often and the result of a search depends on how ties leave the heap.
 *
 * Deterministic where the reference is not: the level RNG is xoshiro256++ seeded by
 * `level_seed` (reference: entropy), insertion is sequential in the given order (reference:
 * rayon parallel_insert), sets that the reference iterates in HashMap order are iterated in
 * discovery order, and neighbour lists are sorted by (distance, internal index).
 */
#include "gso.h"
#ifdef _OPENMP
#include <omp.h>
#endif

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float d;
    uint32_t p;
} hitem; /* PointWithOrder: ordered by d only */

typedef struct {
    hitem *a;
    uint32_t n, cap;
} heap;

static void heap_reserve(heap *h, uint32_t cap) {
    if (cap > h->cap) {
        h->a = (hitem *)realloc(h->a, (size_t)cap * sizeof(hitem));
        h->cap = cap;
    }
}
static void heap_sift_up(heap *h, uint32_t start, uint32_t pos) {
    hitem e = h->a[pos];
    while (pos > start) {
        uint32_t parent = (pos - 1) / 2;
        if (e.d <= h->a[parent].d) break;
        h->a[pos] = h->a[parent];
        pos = parent;
    }
    h->a[pos] = e;
}
static void heap_push(heap *h, hitem x) {
    if (h->n == h->cap) heap_reserve(h, h->cap ? 2 * h->cap : 64);
    h->a[h->n] = x;
    h->n++;
    heap_sift_up(h, 0, h->n - 1);
}
static void heap_sift_down_to_bottom(heap *h, uint32_t pos) {
    const uint32_t end = h->n, start = pos;
    hitem e = h->a[pos];
    uint32_t child = 2 * pos + 1;
    while (end >= 2 && child <= end - 2) {
        child += (h->a[child].d <= h->a[child + 1].d) ? 1u : 0u;
        h->a[pos] = h->a[child];
        pos = child;
        child = 2 * pos + 1;
    }
    if (end >= 1 && child == end - 1) {
        h->a[pos] = h->a[child];
        pos = child;
    }
    h->a[pos] = e;
    heap_sift_up(h, start, pos);
}
static hitem heap_pop(heap *h) {
    hitem item = h->a[h->n - 1];
    h->n--;
    if (h->n > 0) {
        hitem t = h->a[0];
        h->a[0] = item;
        item = t;
        heap_sift_down_to_bottom(h, 0);
    }
    return item;
}
static void heap_sift_down_range(heap *h, uint32_t pos, uint32_t end) {
    hitem e = h->a[pos];
    uint32_t child = 2 * pos + 1;
    while (end >= 2 && child <= end - 2) {
        child += (h->a[child].d <= h->a[child + 1].d) ? 1u : 0u;
        if (e.d >= h->a[child].d) {
            h->a[pos] = e;
            return;
        }
        h->a[pos] = h->a[child];
        pos = child;
        child = 2 * pos + 1;
    }
    if (end >= 1 && child == end - 1 && e.d < h->a[child].d) {
        h->a[pos] = h->a[child];
        pos = child;
    }
    h->a[pos] = e;
}
static void heap_into_sorted(heap *h) {
    uint32_t end = h->n;
    while (end > 1) {
        end--;
        hitem t = h->a[0];
        h->a[0] = h->a[end];
        h->a[end] = t;
        heap_sift_down_range(h, 0, end);
    }
}

typedef struct {
    uint32_t *idx;
    float *dist;
    uint32_t n;
} nlist;

struct gso_hnsw {
    uint32_t M, max_layer, ef_c, sig_type, S, extend_candidates, keep_pruned;
    uint64_t capacity;
    double scale;
    uint32_t esz;
    uint64_t row;
    gso_xoshiro level_rng;
    uint64_t n;        /* points */
    uint8_t *data;     /* n x row */
    uint64_t *ids;     /* origin ids */
    uint8_t *level;    /* level of each point */
    uint32_t *rank;    /* rank in its layer */
    nlist **nbrs;      /* nbrs[p][l], l = 0..level[p] */
    uint32_t layer_count[256];
    int64_t entry;     /* -1 if none */
    uint64_t cap_pts;
    uint32_t *stamp;   /* visited stamps */
    uint32_t cur_stamp;
    uint64_t nb_eval;
};

static uint32_t esize(uint32_t t) { return t == GSO_SIG_U64 ? 8u : (t == GSO_SIG_U16 ? 2u : 4u); }

gso_hnsw *gso_hnsw_new(uint32_t max_nb_conn, uint64_t capacity, uint32_t max_layer,
                       uint32_t ef_c, double scale_modification, uint32_t sig_type, uint32_t S,
                       uint32_t extend_candidates, uint32_t keep_pruned, uint64_t level_seed) {
    gso_hnsw *h = (gso_hnsw *)calloc(1, sizeof *h);
    if (!h) return NULL;
    h->M = max_nb_conn;
    h->capacity = capacity;
    h->max_layer = max_layer > 255 ? 255 : max_layer;
    h->ef_c = ef_c;
    h->sig_type = sig_type;
    h->S = S;
    h->extend_candidates = extend_candidates;
    h->keep_pruned = keep_pruned;
    h->scale = scale_modification / log((double)max_nb_conn);
    h->esz = esize(sig_type);
    h->row = (uint64_t)S * h->esz;
    gso_xoshiro_seed_from_u64(&h->level_rng, level_seed);
    h->entry = -1;
    return h;
}

void gso_hnsw_free(gso_hnsw *h) {
    if (!h) return;
    for (uint64_t p = 0; p < h->n; p++) {
        for (uint32_t l = 0; l <= h->level[p]; l++) {
            free(h->nbrs[p][l].idx);
            free(h->nbrs[p][l].dist);
        }
        free(h->nbrs[p]);
    }
    free(h->nbrs);
    free(h->data);
    free(h->ids);
    free(h->level);
    free(h->rank);
    free(h->stamp);
    free(h);
}

uint64_t gso_hnsw_nb_point(const gso_hnsw *h) { return h->n; }
uint64_t gso_hnsw_nb_eval(const gso_hnsw *h) { return h->nb_eval; }

static inline const void *pt(const gso_hnsw *h, uint32_t p) { return h->data + (uint64_t)p * h->row; }
static inline float dist_qp(gso_hnsw *h, const void *q, uint32_t p) {
    h->nb_eval++;
    return gso_hamming(q, pt(h, p), h->S, h->sig_type);
}

static uint32_t gen_level(gso_hnsw *h) {
    double xsi = gso_uniform_f64(&h->level_rng);
    double lv = -log(xsi) * h->scale;
    if (!(lv < (double)h->max_layer)) /* also catches +inf */
        return (uint32_t)gso_uniform_usize(&h->level_rng, h->max_layer);
    return (uint32_t)floor(lv);
}

static uint32_t new_stamp(gso_hnsw *h) {
    h->cur_stamp++;
    if (h->cur_stamp == 0) {
        memset(h->stamp, 0, h->cap_pts * sizeof(uint32_t));
        h->cur_stamp = 1;
    }
    return h->cur_stamp;
}

/* search_layer: returns max-heap `ret` (positive distances) of at most ef points */
static void search_layer(gso_hnsw *h, const void *q, uint32_t ep, uint32_t ef, uint32_t layer,
                         heap *ret, heap *cand) {
    ret->n = 0;
    cand->n = 0;
    const uint32_t st = new_stamp(h);
    float d0 = dist_qp(h, q, ep);
    h->stamp[ep] = st;
    heap_push(cand, (hitem){-d0, ep});
    heap_push(ret, (hitem){d0, ep});
    while (cand->n > 0) {
        hitem c = heap_pop(cand);
        hitem f = ret->a[0];
        if (-c.d > f.d) return;
        const nlist *nl = &h->nbrs[c.p][layer];
        for (uint32_t i = 0; i < nl->n; i++) {
            uint32_t e = nl->idx[i];
            if (h->stamp[e] == st) continue;
            h->stamp[e] = st;
            float fd = ret->a[0].d;
            float ed = dist_qp(h, q, e);
            if (ed < fd || ret->n < ef) {
                heap_push(cand, (hitem){-ed, e});
                heap_push(ret, (hitem){ed, e});
                if (ret->n > ef) (void)heap_pop(ret);
            }
        }
    }
}

typedef struct {
    uint32_t p;
    float d;
} sel;

static int cmp_sel(const void *a, const void *b) {
    const sel *x = (const sel *)a, *y = (const sel *)b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return (x->p > y->p) - (x->p < y->p);
}

/* select_neighbours: `cand` is a max-heap on negated distances (pops nearest first) */
static uint32_t select_neighbours(gso_hnsw *h, const void *q, heap *cand, uint32_t nb_asked,
                                  int extend_asked, uint32_t layer, sel *out) {
    uint32_t nout = 0;
    int extend = 0;
    if (cand->n <= nb_asked) {
        if (!extend_asked) {
            while (cand->n > 0) {
                hitem p = heap_pop(cand);
                out[nout].p = p.p;
                out[nout].d = -p.d;
                nout++;
            }
            return nout;
        }
        extend = 1;
    }
    if (extend) {
        const uint32_t st = new_stamp(h);
        const uint32_t n0 = cand->n;
        for (uint32_t i = 0; i < n0; i++) h->stamp[cand->a[i].p] = st;
        uint32_t nnew = 0, capnew = 256;
        uint32_t *newc = (uint32_t *)malloc(capnew * sizeof(uint32_t));
        for (uint32_t i = 0; i < n0; i++) {
            const nlist *nl = &h->nbrs[cand->a[i].p][layer];
            for (uint32_t j = 0; j < nl->n; j++) {
                uint32_t e = nl->idx[j];
                if (h->stamp[e] == st) continue;
                h->stamp[e] = st;
                if (nnew == capnew) {
                    capnew *= 2;
                    newc = (uint32_t *)realloc(newc, capnew * sizeof(uint32_t));
                }
                newc[nnew++] = e;
            }
        }
        for (uint32_t i = 0; i < nnew; i++) {
            float d = dist_qp(h, q, newc[i]);
            heap_push(cand, (hitem){-d, newc[i]});
        }
        free(newc);
    }
    while (cand->n > 0 && nout < nb_asked) {
        hitem e = heap_pop(cand);
        const float ed = -e.d;
        int insert = 1;
        for (uint32_t i = 0; i < nout; i++) {
            h->nb_eval++;
            float dd = gso_hamming(pt(h, e.p), pt(h, out[i].p), h->S, h->sig_type);
            if (dd <= ed) {
                insert = 0;
                break;
            }
        }
        if (insert) {
            out[nout].p = e.p;
            out[nout].d = ed;
            nout++;
        }
        /* keep_pruned = false in the reference call (dnasketch.rs:160): discarded points
         * are dropped */
    }
    return nout;
}

static int grow(gso_hnsw *h, uint64_t need) {
    if (need <= h->cap_pts) return 0;
    uint64_t nc = h->cap_pts ? h->cap_pts : 1024;
    while (nc < need) nc *= 2;
    uint8_t *d = (uint8_t *)realloc(h->data, nc * h->row);
    if (!d) return 4;
    h->data = d;
    h->ids = (uint64_t *)realloc(h->ids, nc * sizeof(uint64_t));
    h->level = (uint8_t *)realloc(h->level, nc);
    h->rank = (uint32_t *)realloc(h->rank, nc * sizeof(uint32_t));
    h->nbrs = (nlist **)realloc(h->nbrs, nc * sizeof(nlist *));
    uint32_t *s = (uint32_t *)realloc(h->stamp, nc * sizeof(uint32_t));
    if (!h->ids || !h->level || !h->rank || !h->nbrs || !s) return 4;
    memset(s + h->cap_pts, 0, (nc - h->cap_pts) * sizeof(uint32_t));
    h->stamp = s;
    h->cap_pts = nc;
    return 0;
}

static void list_add_sorted_shrink(gso_hnsw *h, uint32_t qp, uint32_t l, uint32_t newp, float d) {
    nlist *nl = &h->nbrs[qp][l];
    for (uint32_t i = 0; i < nl->n; i++)
        if (nl->idx[i] == newp) return;
    /* push + sort by (dist, index): insertion keeps the list sorted */
    uint32_t pos = nl->n;
    while (pos > 0 && (nl->dist[pos - 1] > d || (nl->dist[pos - 1] == d && nl->idx[pos - 1] > newp))) {
        nl->dist[pos] = nl->dist[pos - 1];
        nl->idx[pos] = nl->idx[pos - 1];
        pos--;
    }
    nl->dist[pos] = d;
    nl->idx[pos] = newp;
    nl->n++;
    const uint32_t thr = l > 0 ? h->M : 2 * h->M;
    if (nl->n > thr) nl->n--; /* pop the farthest */
}

static int insert_one(gso_hnsw *h, const void *sig, uint64_t id, heap *ret, heap *cand,
                      sel *selbuf) {
    if (h->n >= h->capacity) return 8;
    if (grow(h, h->n + 1)) return 4;
    const uint32_t np = (uint32_t)h->n;
    const uint32_t level = gen_level(h);
    memcpy(h->data + (uint64_t)np * h->row, sig, h->row);
    h->ids[np] = id;
    h->level[np] = (uint8_t)level;
    h->rank[np] = h->layer_count[level]++;
    h->nbrs[np] = (nlist *)calloc(level + 1, sizeof(nlist));
    for (uint32_t l = 0; l <= level; l++) {
        uint32_t cap = (l > 0 ? h->M : 2 * h->M) + 1;
        h->nbrs[np][l].idx = (uint32_t *)malloc(cap * sizeof(uint32_t));
        h->nbrs[np][l].dist = (float *)malloc(cap * sizeof(float));
        h->nbrs[np][l].n = 0;
    }
    h->n++;
    if (h->entry < 0) {
        h->entry = np;
        return 0;
    }
    const void *q = pt(h, np);
    uint32_t ep = (uint32_t)h->entry;
    const uint32_t max_level_observed = h->level[ep];
    float dist_to_entry = dist_qp(h, q, ep);
    for (int l = (int)max_level_observed; l >= (int)level + 1; l--) {
        search_layer(h, q, ep, 1, (uint32_t)l, ret, cand);
        if (ret->n > 0) {
            hitem e = heap_pop(ret);
            float tmp = dist_qp(h, q, e.p);
            if (tmp < dist_to_entry) {
                ep = e.p;
                dist_to_entry = tmp;
            }
        }
    }
    const int top = (int)(level < max_level_observed ? level : max_level_observed);
    for (int l = top; l >= 0; l--) {
        search_layer(h, q, ep, h->ef_c, (uint32_t)l, ret, cand);
        /* from_positive_binaryheap_to_negative_binary_heap: push in underlying-vec order */
        cand->n = 0;
        for (uint32_t i = 0; i < ret->n; i++) heap_push(cand, (hitem){-ret->a[i].d, ret->a[i].p});
        if (cand->n > 0) {
            const uint32_t nb_conn = l == 0 ? 2 * h->M : h->M;
            const int extend_c = l == 0 ? (int)h->extend_candidates : 0;
            uint32_t ns = select_neighbours(h, q, cand, nb_conn, extend_c, (uint32_t)l, selbuf);
            qsort(selbuf, ns, sizeof(sel), cmp_sel);
            nlist *nl = &h->nbrs[np][l];
            nl->n = ns;
            for (uint32_t i = 0; i < ns; i++) {
                nl->idx[i] = selbuf[i].p;
                nl->dist[i] = selbuf[i].d;
            }
            if (ns > 0) ep = selbuf[0].p;
        }
    }
    /* reverse_update_neighborhood_simple */
    for (int l = (int)level; l >= 0; l--) {
        const nlist *nl = &h->nbrs[np][l];
        for (uint32_t i = 0; i < nl->n; i++) {
            uint32_t qp = nl->idx[i];
            if (qp == np) continue;
            if ((uint32_t)l > h->level[qp]) continue;
            list_add_sorted_shrink(h, qp, (uint32_t)l, np, nl->dist[i]);
        }
    }
    /* check_entry_point */
    if (level > h->level[h->entry]) h->entry = np;
    return 0;
}

int gso_hnsw_insert(gso_hnsw *h, const void *sigs, const uint64_t *ids, uint64_t n) {
    heap ret = {0}, cand = {0};
    uint32_t selcap = 2 * h->M + 2;
    sel *selbuf = (sel *)malloc(selcap * sizeof(sel));
    int rc = 0;
    for (uint64_t i = 0; i < n && !rc; i++)
        rc = insert_one(h, (const uint8_t *)sigs + i * h->row, ids[i], &ret, &cand, selbuf);
    free(selbuf);
    free(ret.a);
    free(cand.a);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Wave insertion: the deterministic restatement of `parallel_insert` that the GPU builder
 * follows (gsearch_b200/csrc/hnsw_insert.cuh).  The reference inserts with rayon threads that
 * see each other's partial updates in scheduling order (src/dna/dnasketch.rs:435); here the
 * points are taken in waves, in the given order:
 *   wave size  W = clamp(nb_point / 4, 1, wave_max) (so a wave of 1 == sequential insertion);
 *   phase A    every point of the wave searches the graph as it was BEFORE the wave (greedy
 *              descent, search_layer(ef_c) per layer) and then also sees the EARLIER points
 *              of its own wave, by exact distance, as if search_layer had met them last;
 *              select_neighbours as usual (points of the wave have no lists yet); the entry of
 *              the next lower layer is the nearest selected point that is not of this wave;
 *   phase B    the lists of all points of the wave are written, then, in order, reverse updates
 *              are applied (list_add_sorted_shrink keeps the M / 2M smallest by (distance,
 *              index), so the result does not depend on the order of arrivals) and the entry
 *              point is updated.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t n[17];
    sel *l[17];
} wave_sel;

static int add_point(gso_hnsw *h, const void *sig, uint64_t id) {
    if (h->n >= h->capacity) return 8;
    if (grow(h, h->n + 1)) return 4;
    const uint32_t np = (uint32_t)h->n;
    const uint32_t level = gen_level(h);
    memcpy(h->data + (uint64_t)np * h->row, sig, h->row);
    h->ids[np] = id;
    h->level[np] = (uint8_t)level;
    h->rank[np] = h->layer_count[level]++;
    h->nbrs[np] = (nlist *)calloc(level + 1, sizeof(nlist));
    for (uint32_t l = 0; l <= level; l++) {
        uint32_t cap = (l > 0 ? h->M : 2 * h->M) + 1;
        h->nbrs[np][l].idx = (uint32_t *)malloc(cap * sizeof(uint32_t));
        h->nbrs[np][l].dist = (float *)malloc(cap * sizeof(float));
        h->nbrs[np][l].n = 0;
    }
    h->n++;
    return 0;
}

static void wave_phase_a(gso_hnsw *h, uint32_t np, uint32_t first, uint32_t entry, heap *ret,
                         heap *cand, sel *selbuf, wave_sel *ws) {
    const uint32_t level = h->level[np];
    const void *q = pt(h, np);
    uint32_t ep = entry;
    const uint32_t max_level_observed = h->level[ep];
    float dist_to_entry = dist_qp(h, q, ep);
    for (int l = (int)max_level_observed; l >= (int)level + 1; l--) {
        search_layer(h, q, ep, 1, (uint32_t)l, ret, cand);
        if (ret->n > 0) {
            hitem e = heap_pop(ret);
            float tmp = dist_qp(h, q, e.p);
            if (tmp < dist_to_entry) {
                ep = e.p;
                dist_to_entry = tmp;
            }
        }
    }
    const int top = (int)(level < max_level_observed ? level : max_level_observed);
    for (int l = top; l >= 0; l--) {
        search_layer(h, q, ep, h->ef_c, (uint32_t)l, ret, cand);
        for (uint32_t m = first; m < np; m++) { /* earlier points of this wave */
            if (h->level[m] < (uint32_t)l) continue;
            const float fd = ret->a[0].d;
            const float ed = dist_qp(h, q, m);
            if (ed < fd || ret->n < h->ef_c) {
                heap_push(ret, (hitem){ed, m});
                if (ret->n > h->ef_c) (void)heap_pop(ret);
            }
        }
        cand->n = 0;
        for (uint32_t i = 0; i < ret->n; i++) heap_push(cand, (hitem){-ret->a[i].d, ret->a[i].p});
        const uint32_t nb_conn = l == 0 ? 2 * h->M : h->M;
        const int extend_c = l == 0 ? (int)h->extend_candidates : 0;
        uint32_t ns = select_neighbours(h, q, cand, nb_conn, extend_c, (uint32_t)l, selbuf);
        qsort(selbuf, ns, sizeof(sel), cmp_sel);
        ws->n[l] = ns;
        ws->l[l] = (sel *)malloc((ns ? ns : 1) * sizeof(sel));
        memcpy(ws->l[l], selbuf, ns * sizeof(sel));
        for (uint32_t i = 0; i < ns; i++)
            if (selbuf[i].p < first) {
                ep = selbuf[i].p;
                break;
            }
    }
}

/* phase B, step 1: the point's own selection becomes its lists (already sorted by (d, index)) */
static void wave_phase_b_own(gso_hnsw *h, uint32_t np, wave_sel *ws) {
    const uint32_t level = h->level[np];
    for (uint32_t l = 0; l <= level && l < 17; l++) {
        nlist *nl = &h->nbrs[np][l];
        nl->n = ws->n[l];
        for (uint32_t i = 0; i < ws->n[l]; i++) {
            nl->idx[i] = ws->l[l][i].p;
            nl->dist[i] = ws->l[l][i].d;
        }
    }
}

/* phase B, step 2: reverse updates and entry point, after ALL own lists of the wave are set */
static void wave_phase_b_reverse(gso_hnsw *h, uint32_t np, wave_sel *ws) {
    const uint32_t level = h->level[np];
    for (int l = (int)level; l >= 0; l--) {
        for (uint32_t i = 0; i < ws->n[l]; i++) {
            const uint32_t qp = ws->l[l][i].p;
            if (qp == np) continue;
            if ((uint32_t)l > h->level[qp]) continue;
            list_add_sorted_shrink(h, qp, (uint32_t)l, np, ws->l[l][i].d);
        }
    }
    if (level > h->level[h->entry]) h->entry = np;
}

uint32_t gso_hnsw_wave_size(uint64_t nb_point, uint32_t wave_max) {
    uint64_t w = nb_point / 4;
    if (w < 1) w = 1;
    if (w > wave_max) w = wave_max;
    return (uint32_t)w;
}

int gso_hnsw_insert_waves(gso_hnsw *h, const void *sigs, const uint64_t *ids, uint64_t n,
                          uint32_t wave_max) {
    heap ret = {0}, cand = {0};
    sel *selbuf = (sel *)malloc((2 * h->M + 2) * sizeof(sel));
    int rc = 0;
    uint64_t i = 0;
    if (wave_max < 1) wave_max = 1;
    while (i < n && !rc) {
        if (h->entry < 0) { /* first point of the index: becomes the entry point */
            rc = add_point(h, (const uint8_t *)sigs + i * h->row, ids[i]);
            if (!rc) h->entry = (int64_t)(h->n - 1);
            i++;
            continue;
        }
        uint64_t W = gso_hnsw_wave_size(h->n, wave_max);
        if (W > n - i) W = n - i;
        const uint32_t first = (uint32_t)h->n;
        for (uint64_t t = 0; t < W && !rc; t++)
            rc = add_point(h, (const uint8_t *)sigs + (i + t) * h->row, ids[i + t]);
        if (rc) break;
        wave_sel *ws = (wave_sel *)calloc(W, sizeof(wave_sel));
        const uint32_t entry = (uint32_t)h->entry;
        for (uint64_t t = 0; t < W; t++)
            wave_phase_a(h, first + (uint32_t)t, first, entry, &ret, &cand, selbuf, &ws[t]);
        for (uint64_t t = 0; t < W; t++) wave_phase_b_own(h, first + (uint32_t)t, &ws[t]);
        for (uint64_t t = 0; t < W; t++) {
            wave_phase_b_reverse(h, first + (uint32_t)t, &ws[t]);
            for (int l = 0; l < 17; l++) free(ws[t].l[l]);
        }
        free(ws);
        i += W;
    }
    free(selbuf);
    free(ret.a);
    free(cand.a);
    return rc;
}

/* The same wave insertion with phase A spread over host threads (one point per task): phase A
 * only READS the pre-wave graph and the signatures / levels of the wave, so every worker runs it
 * on a private shallow copy of the handle (own visited stamps, heaps, eval counter) and the graph
 * that comes out is bit-identical to gso_hnsw_insert_waves.  This is how bench.py's CPU reference
 * arm builds its index within minutes (the reference's parallel_insert is multi-threaded too,
 * src/dna/dnasketch.rs:435). */
int gso_hnsw_insert_waves_mt(gso_hnsw *h, const void *sigs, const uint64_t *ids, uint64_t n,
                             uint32_t wave_max, int nthreads) {
    if (nthreads <= 1) return gso_hnsw_insert_waves(h, sigs, ids, n, wave_max);
    if (h->n + n > h->capacity) return 8;
    if (grow(h, h->n + n)) return 4; /* no reallocation while workers hold copies of the handle */
    if (wave_max < 1) wave_max = 1;
    const uint64_t cap = h->cap_pts;
    uint32_t **tstamp = (uint32_t **)calloc((size_t)nthreads, sizeof(uint32_t *));
    uint32_t *tcur = (uint32_t *)calloc((size_t)nthreads, sizeof(uint32_t));
    heap *tret = (heap *)calloc((size_t)nthreads, sizeof(heap));
    heap *tcand = (heap *)calloc((size_t)nthreads, sizeof(heap));
    sel **tsel = (sel **)calloc((size_t)nthreads, sizeof(sel *));
    uint64_t *teval = (uint64_t *)calloc((size_t)nthreads, sizeof(uint64_t));
    for (int t = 0; t < nthreads; t++) {
        tstamp[t] = (uint32_t *)calloc(cap, sizeof(uint32_t));
        tsel[t] = (sel *)malloc((2 * h->M + 2) * sizeof(sel));
    }
    int rc = 0;
    uint64_t i = 0;
    while (i < n && !rc) {
        if (h->entry < 0) {
            rc = add_point(h, (const uint8_t *)sigs + i * h->row, ids[i]);
            if (!rc) h->entry = (int64_t)(h->n - 1);
            i++;
            continue;
        }
        uint64_t W = gso_hnsw_wave_size(h->n, wave_max);
        if (W > n - i) W = n - i;
        const uint32_t first = (uint32_t)h->n;
        for (uint64_t t = 0; t < W && !rc; t++)
            rc = add_point(h, (const uint8_t *)sigs + (i + t) * h->row, ids[i + t]);
        if (rc) break;
        wave_sel *ws = (wave_sel *)calloc(W, sizeof(wave_sel));
        const uint32_t entry = (uint32_t)h->entry;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
        for (int64_t t = 0; t < (int64_t)W; t++) {
#ifdef _OPENMP
            const int me = omp_get_thread_num();
#else
            const int me = 0;
#endif
            gso_hnsw local = *h;
            local.stamp = tstamp[me];
            local.cur_stamp = tcur[me];
            local.nb_eval = 0;
            wave_phase_a(&local, first + (uint32_t)t, first, entry, &tret[me], &tcand[me], tsel[me], &ws[t]);
            tcur[me] = local.cur_stamp;
            teval[me] += local.nb_eval;
        }
        for (uint64_t t = 0; t < W; t++) wave_phase_b_own(h, first + (uint32_t)t, &ws[t]);
        for (uint64_t t = 0; t < W; t++) {
            wave_phase_b_reverse(h, first + (uint32_t)t, &ws[t]);
            for (int l = 0; l < 17; l++) free(ws[t].l[l]);
        }
        free(ws);
        i += W;
    }
    for (int t = 0; t < nthreads; t++) {
        h->nb_eval += teval[t];
        free(tstamp[t]);
        free(tsel[t]);
        free(tret[t].a);
        free(tcand[t].a);
    }
    free(tstamp);
    free(tcur);
    free(tret);
    free(tcand);
    free(tsel);
    free(teval);
    return rc;
}

static uint32_t search_one(gso_hnsw *h, const void *q, uint32_t knbn, uint32_t ef_arg,
                           gso_neighbour *out, heap *ret, heap *cand) {
    if (h->entry < 0) return 0;
    uint32_t pivot = (uint32_t)h->entry;
    float dist_to_entry = dist_qp(h, q, pivot);
    for (int layer = (int)h->level[pivot]; layer >= 1; layer--) {
        /* the entry point's level bounds the loop; the pivot may move to a point that
         * still has this layer because links are within a layer */
        const nlist *nl = &h->nbrs[pivot][layer];
        int changed = 0;
        uint32_t newp = pivot;
        for (uint32_t i = 0; i < nl->n; i++) {
            float tmp = dist_qp(h, q, nl->idx[i]);
            if (tmp < dist_to_entry) {
                changed = 1;
                dist_to_entry = tmp;
                newp = nl->idx[i];
            }
        }
        if (changed) pivot = newp;
    }
    const uint32_t ef = ef_arg > knbn ? ef_arg : knbn;
    search_layer(h, q, pivot, ef, 0, ret, cand);
    heap_into_sorted(ret);
    uint32_t last = knbn < ef ? knbn : ef;
    if (ret->n < last) last = ret->n;
    for (uint32_t i = 0; i < last; i++) {
        uint32_t p = ret->a[i].p;
        out[i].d_id = h->ids[p];
        out[i].distance = ret->a[i].d;
        out[i].layer = h->level[p];
        out[i].pad_[0] = out[i].pad_[1] = out[i].pad_[2] = 0;
        out[i].rank = (int32_t)h->rank[p];
    }
    return last;
}

uint32_t gso_hnsw_search(gso_hnsw *h, const void *q, uint32_t knbn, uint32_t ef,
                         gso_neighbour *out, uint64_t *nb_eval_out) {
    heap ret = {0}, cand = {0};
    uint64_t e0 = h->nb_eval;
    uint32_t n = search_one(h, q, knbn, ef, out, &ret, &cand);
    if (nb_eval_out) *nb_eval_out = h->nb_eval - e0;
    free(ret.a);
    free(cand.a);
    return n;
}

/* parallel_search: one query per worker.  Each worker uses a private shallow copy of the
 * handle (own visited stamps and eval counter); the graph itself is read-only here. */
void gso_hnsw_search_batch(gso_hnsw *h, const void *queries, uint32_t nq, uint32_t knbn,
                           uint32_t ef, gso_neighbour *out, uint32_t *counts,
                           uint64_t *nb_eval_out, int nthreads) {
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        gso_hnsw local = *h;
        local.stamp = (uint32_t *)calloc(h->cap_pts ? h->cap_pts : 1, sizeof(uint32_t));
        local.cur_stamp = 0;
        heap ret = {0}, cand = {0};
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < (int64_t)nq; i++) {
            uint64_t e0 = local.nb_eval;
            counts[i] = search_one(&local, (const uint8_t *)queries + (uint64_t)i * h->row, knbn,
                                   ef, out + (uint64_t)i * knbn, &ret, &cand);
            if (nb_eval_out) nb_eval_out[i] = local.nb_eval - e0;
        }
        free(local.stamp);
        free(ret.a);
        free(cand.a);
    }
}

/* graph image import (inverse of gso_hnsw_export): lets the CPU search run on a graph built
 * elsewhere (bench.py times the CPU search on the graph the GPU built) */
int gso_hnsw_import(gso_hnsw *h, const void *sigs, const uint64_t *ids, uint64_t n,
                    const uint8_t *levels, const uint32_t *ranks, const uint64_t *nbr_offsets,
                    const uint32_t *nbr_index, const float *nbr_dist, uint64_t entry_point) {
    if (h->n != 0) return 1;
    if (grow(h, n ? n : 1)) return 4;
    memcpy(h->data, sigs, n * h->row);
    uint64_t li = 0;
    for (uint64_t p = 0; p < n; p++) {
        h->ids[p] = ids[p];
        h->level[p] = levels[p];
        h->rank[p] = ranks[p];
        h->layer_count[levels[p]]++;
        h->nbrs[p] = (nlist *)calloc((size_t)levels[p] + 1, sizeof(nlist));
        for (uint32_t l = 0; l <= levels[p]; l++, li++) {
            const uint32_t cap = (l > 0 ? h->M : 2 * h->M) + 1;
            const uint64_t b = nbr_offsets[li], e = nbr_offsets[li + 1];
            if (e - b >= cap) return 1;
            nlist *nl = &h->nbrs[p][l];
            nl->idx = (uint32_t *)malloc(cap * sizeof(uint32_t));
            nl->dist = (float *)malloc(cap * sizeof(float));
            nl->n = (uint32_t)(e - b);
            for (uint64_t i = b; i < e; i++) {
                nl->idx[i - b] = nbr_index[i];
                nl->dist[i - b] = nbr_dist ? nbr_dist[i] : 0.f;
            }
        }
    }
    h->n = n;
    h->entry = n ? (int64_t)entry_point : -1;
    return 0;
}

uint64_t gso_hnsw_total_lists(const gso_hnsw *h) {
    uint64_t t = 0;
    for (uint64_t p = 0; p < h->n; p++) t += (uint64_t)h->level[p] + 1;
    return t;
}
uint64_t gso_hnsw_total_nbrs(const gso_hnsw *h) {
    uint64_t t = 0;
    for (uint64_t p = 0; p < h->n; p++)
        for (uint32_t l = 0; l <= h->level[p]; l++) t += h->nbrs[p][l].n;
    return t;
}

/* graph image: list index of (p, l) = list_base[p] + l where list_base is the prefix sum of
 * (level+1); nbr_offsets has total_lists+1 entries. */
void gso_hnsw_export(const gso_hnsw *h, uint8_t *levels, uint32_t *ranks, uint64_t *ids,
                     uint64_t *nbr_offsets, uint32_t *nbr_index, float *nbr_dist,
                     uint64_t *entry_point) {
    uint64_t li = 0, off = 0;
    for (uint64_t p = 0; p < h->n; p++) {
        levels[p] = h->level[p];
        ranks[p] = h->rank[p];
        ids[p] = h->ids[p];
        for (uint32_t l = 0; l <= h->level[p]; l++) {
            nbr_offsets[li++] = off;
            const nlist *nl = &h->nbrs[p][l];
            for (uint32_t i = 0; i < nl->n; i++) {
                nbr_index[off] = nl->idx[i];
                if (nbr_dist) nbr_dist[off] = nl->dist[i];
                off++;
            }
        }
    }
    nbr_offsets[li] = off;
    *entry_point = h->entry < 0 ? UINT64_MAX : (uint64_t)h->entry;
}
