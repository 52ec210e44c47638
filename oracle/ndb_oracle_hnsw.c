/*
 * ndb_oracle_hnsw.c -- CPU oracle for the HNSW path (TEST INFRASTRUCTURE ONLY).
 *
 * Restates NeuronDB/src/index/hnsw_am.c over an in-memory graph:
 *   hnswGetRandomLevel  :1143-1161
 *   hnswSearch          :1545-2080   (search_mode 0, "literal": greedy descent, then
 *                                     level-0 BFS that stops at ef candidates -- SURVEY Q12)
 *   hnswInsertNode      :2091-2670   (insert mode 0, "literal": the same level-0 search
 *                                     result is linked at every level -- SURVEY Q13)
 * Node ids are 0-based insertion indices (the reference's BlockNumber minus the meta
 * block); ORC_INVALID stands for InvalidBlockNumber.
 *
 * Two deliberate additions, both flagged where they appear:
 *   search_mode 1 -- best-first search_layer (Malkov & Yashunin, Alg. 2), the algorithm
 *                    the reference's dead src/scan/hnsw_scan.c intends; this is the mode
 *                    the recall@10 >= 0.95 target is measured in.
 *   insert mode 1 -- per-level search at that level and no self hits; what the GPU build
 *                    implements (the literal mode cannot complete in the reference
 *                    itself: SURVEY Q22).
 *   insert mode 2 -- EXTENSION, experiment only: mode 1 + a full neighbour swaps its farthest
 *                    link for a closer new node (the intent of the reference's unreachable
 *                    prune block, :2515-2612).  Measured worse than mode 3; no GPU twin.
 *   insert mode 3 -- EXTENSION: mode 1 with the diversity heuristic of Malkov & Yashunin
 *                    (Alg. 4) for the forward links and re-selection of a full neighbour's
 *                    links; the GPU build's NDB_HNSW_SELECT_HEURISTIC, link for link.
 * Out-of-bounds behaviour of the reference (back-links written at a level the
 * neighbour node does not have, :2493-2513 via HnswGetNeighborsSafe) is undefined
 * there and skipped here.
 */
#include "ndb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

struct OrcHnsw {
    int dim, m, efc, efs;
    float ml;
    int64_t n, cap;
    float *vec;        /* cap * dim */
    int *level;        /* cap */
    int16_t *cnt;      /* cap * 16   neighborCount[level] */
    uint32_t *nbr0;    /* cap * 2m   level-0 slots */
    uint32_t **upper;  /* cap        level>=1 slots: level * 2m entries, or NULL */
    uint32_t entry;
    int entry_level, max_level;
};

static int64_t g_evals;
#ifdef _OPENMP
#pragma omp threadprivate(g_evals)
#endif
static int64_t g_evals_total;

int64_t orc_hnsw_distance_evals(void) { return g_evals_total; }

/* hnswInitMetaPage, :1090-1110 */
OrcHnsw *orc_hnsw_create(int dim, int m, int efc, int efs, float ml, int64_t capacity)
{
    OrcHnsw *g = (OrcHnsw *) calloc(1, sizeof(OrcHnsw));
    g->dim = dim; g->m = m; g->efc = efc; g->efs = efs; g->ml = ml;
    g->cap = capacity > 0 ? capacity : 1;
    g->vec = (float *) malloc(sizeof(float) * (size_t) g->cap * dim);
    g->level = (int *) malloc(sizeof(int) * (size_t) g->cap);
    g->cnt = (int16_t *) calloc((size_t) g->cap * ORC_HNSW_MAX_LEVEL, sizeof(int16_t));
    g->nbr0 = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) g->cap * 2 * m);
    g->upper = (uint32_t **) calloc((size_t) g->cap, sizeof(uint32_t *));
    g->entry = ORC_INVALID;
    g->entry_level = -1;
    g->max_level = -1;
    return g;
}

void orc_hnsw_free(OrcHnsw *g)
{
    if (!g) return;
    for (int64_t i = 0; i < g->n; i++) free(g->upper[i]);
    free(g->upper); free(g->nbr0); free(g->cnt); free(g->level); free(g->vec); free(g);
}

/* hnswGetRandomLevel, :1143-1161 (libc random(); callers seed with srandom) */
int orc_hnsw_random_level(float ml)
{
    double r = (double) random() / (double) RAND_MAX;
    while (r == 0.0)
        r = (double) random() / (double) RAND_MAX;
    int level = (int) (-log(r) * ml);
    if (level > ORC_HNSW_MAX_LEVEL - 1) level = ORC_HNSW_MAX_LEVEL - 1;
    if (level < 0) level = 0;
    return level;
}

static inline uint32_t *slots(const OrcHnsw *g, uint32_t node, int lev)
{
    if (lev == 0) return g->nbr0 + (size_t) node * 2 * g->m;
    if (lev > g->level[node]) return NULL;
    return g->upper[node] + (size_t) (lev - 1) * 2 * g->m;
}

/* hnswValidateNeighborCount, :1167-1187 */
static inline int clamp_cnt(const OrcHnsw *g, uint32_t node, int lev)
{
    int c = g->cnt[(size_t) node * ORC_HNSW_MAX_LEVEL + lev];
    if (c < 0) return 0;
    if (c > 2 * g->m) return 2 * g->m;
    return c;
}

static inline float dist_to(const OrcHnsw *g, const float *q, uint32_t node, int strategy)
{
    g_evals++;
    return orc_hnsw_distance(q, g->vec + (size_t) node * g->dim, g->dim, strategy);
}

/* greedy descent on the upper levels, hnsw_am.c:1638-1750.  The reference re-evaluates
 * the current node and every neighbour on each pass; a neighbour replaces current on
 * strict <. */
static uint32_t greedy_descent(const OrcHnsw *g, const float *q, int strategy, uint32_t current,
                               int from_level, int to_level_exclusive)
{
    for (int level = from_level; level > to_level_exclusive; level--) {
        int foundBetter;
        do {
            foundBetter = 0;
            if (current == ORC_INVALID || current >= (uint32_t) g->n) break;
            float currentDist = dist_to(g, q, current, strategy);
            if (g->level[current] >= level) {
                const uint32_t *nb = slots(g, current, level);
                int nc = clamp_cnt(g, current, level);
                for (int i = 0; i < nc; i++) {
                    if (nb[i] == ORC_INVALID || nb[i] >= (uint32_t) g->n) continue;
                    float nd = dist_to(g, q, nb[i], strategy);
                    if (nd < currentDist) {
                        current = nb[i];
                        currentDist = nd;
                        foundBetter = 1;
                    }
                }
            }
        } while (foundBetter);
    }
    return current;
}

/* top-k by selection sort over an index array with swaps, strict <, :1984-2013 */
static int select_topk(const uint32_t *cand, const float *cd, int cc, int k,
                       uint32_t *out_nodes, float *out_dist)
{
    int *idx = (int *) malloc(sizeof(int) * (size_t) (cc > 0 ? cc : 1));
    for (int i = 0; i < cc; i++) idx[i] = i;
    for (int i = 0; i < k && i < cc; i++) {
        int minIdx = i;
        float minDist = cd[idx[i]];
        for (int j = i + 1; j < cc; j++)
            if (cd[idx[j]] < minDist) { minDist = cd[idx[j]]; minIdx = j; }
        if (minIdx != i) { int t = idx[i]; idx[i] = idx[minIdx]; idx[minIdx] = t; }
    }
    int tk = k < cc ? k : cc;
    for (int i = 0; i < tk; i++) { out_nodes[i] = cand[idx[i]]; out_dist[i] = cd[idx[i]]; }
    free(idx);
    return tk;
}

/* level-0 phase of hnswSearch, :1752-1975 */
static int search_level0_literal(const OrcHnsw *g, const float *q, int strategy, uint32_t current,
                                 int ef, int k, uint8_t *visited, uint32_t *out_nodes, float *out_dist)
{
    uint32_t *cand = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) ef);
    float *cd = (float *) malloc(sizeof(float) * (size_t) ef);
    int cc;
    cand[0] = current;
    cd[0] = dist_to(g, q, current, strategy);
    cc = 1;
    visited[current] = 1;
    for (int i = 0; i < cc && cc < ef; i++) {
        uint32_t c = cand[i];
        if (c == ORC_INVALID || c >= (uint32_t) g->n) continue;
        const uint32_t *nb = slots(g, c, 0);
        int nc = clamp_cnt(g, c, 0);
        for (int j = 0; j < nc; j++) {
            if (nb[j] == ORC_INVALID || nb[j] >= (uint32_t) g->n) continue;
            if (visited[nb[j]]) continue;
            float nd = dist_to(g, q, nb[j], strategy);
            visited[nb[j]] = 1;
            if (cc < ef) {
                cand[cc] = nb[j];
                cd[cc] = nd;
                cc++;
            } else {
                int worstIdx = 0;
                float worstDist = cd[0];
                for (int l = 1; l < cc && l < ef; l++)
                    if (cd[l] > worstDist) { worstDist = cd[l]; worstIdx = l; }
                if (nd < worstDist) { cand[worstIdx] = nb[j]; cd[worstIdx] = nd; }
            }
        }
    }
    int tk = select_topk(cand, cd, cc, k, out_nodes, out_dist);
    free(cand); free(cd);
    return tk;
}

typedef struct { float d; uint32_t id; uint8_t expanded; } BfEnt;

static inline int ent_less(float d1, uint32_t i1, float d2, uint32_t i2)
{
    return d1 < d2 || (d1 == d2 && i1 < i2);
}

/* ADDITION (not hnswSearch): best-first search_layer at `lev` with beam ef.
 * W is kept as one array sorted by (dist,id); the nearest unexpanded entry is expanded
 * until none is left.  This equals Alg. 2's two-heap loop: a candidate evicted from W is
 * lexicographically beyond W's furthest entry, which is exactly the break condition. */
static int search_layer_bestfirst(const OrcHnsw *g, const float *q, int strategy, uint32_t ep,
                                  int lev, int ef, uint8_t *visited, uint32_t exclude,
                                  uint32_t *out_nodes, float *out_dist, int kmax)
{
    BfEnt *W = (BfEnt *) malloc(sizeof(BfEnt) * (size_t) (ef + 1));
    int wn = 0;
    visited[ep] = 1;
    W[0].d = dist_to(g, q, ep, strategy); W[0].id = ep; W[0].expanded = 0; wn = 1;
    for (;;) {
        int ci = -1;
        for (int i = 0; i < wn; i++) if (!W[i].expanded) { ci = i; break; }
        if (ci < 0) break;
        W[ci].expanded = 1;
        uint32_t c = W[ci].id;
        const uint32_t *nb = slots(g, c, lev);
        if (!nb) continue;
        int nc = clamp_cnt(g, c, lev);
        for (int j = 0; j < nc; j++) {
            uint32_t e = nb[j];
            if (e == ORC_INVALID || e >= (uint32_t) g->n) continue;
            if (visited[e]) continue;
            visited[e] = 1;
            if (e == exclude) continue;
            float d = dist_to(g, q, e, strategy);
            if (wn == ef && !ent_less(d, e, W[wn - 1].d, W[wn - 1].id)) continue;
            int pos = wn < ef ? wn : ef - 1;
            while (pos > 0 && ent_less(d, e, W[pos - 1].d, W[pos - 1].id)) { W[pos] = W[pos - 1]; pos--; }
            W[pos].d = d; W[pos].id = e; W[pos].expanded = 0;
            if (wn < ef) wn++;
        }
    }
    int out = 0;
    for (int i = 0; i < wn && out < kmax; i++) {
        if (W[i].id == exclude) continue;
        out_nodes[out] = W[i].id; out_dist[out] = W[i].d; out++;
    }
    free(W);
    return out;
}

int orc_hnsw_search_one(const OrcHnsw *g, const float *q, int strategy, int ef, int k,
                        int search_mode, uint32_t *out_nodes, float *out_dist)
{
    if (g->entry == ORC_INVALID || g->n == 0) return 0;
    uint8_t *visited = (uint8_t *) calloc((size_t) g->n, 1);   /* bool[numBlocks], :1619-1631 */
    int currentLevel = g->entry_level;
    if (currentLevel < 0 || currentLevel >= ORC_HNSW_MAX_LEVEL) currentLevel = 0;
    uint32_t cur = greedy_descent(g, q, strategy, g->entry, currentLevel, 0);
    int r;
    if (search_mode == 0)
        r = search_level0_literal(g, q, strategy, cur, ef, k, visited, out_nodes, out_dist);
    else
        r = search_layer_bestfirst(g, q, strategy, cur, 0, ef, visited, ORC_INVALID, out_nodes, out_dist, k);
    free(visited);
    return r;
}

void orc_hnsw_search(const OrcHnsw *g, const float *Q, int nq, int strategy, int ef, int k,
                     int search_mode, uint32_t *out_nodes, float *out_dist, int *out_count,
                     int nthreads)
{
    int64_t total = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads > 0 ? nthreads : 1) reduction(+:total)
#endif
    for (int qi = 0; qi < nq; qi++) {
        g_evals = 0;
        uint32_t *on = out_nodes + (size_t) qi * k;
        float *od = out_dist + (size_t) qi * k;
        int c = orc_hnsw_search_one(g, Q + (size_t) qi * g->dim, strategy, ef, k, search_mode, on, od);
        for (int j = c; j < k; j++) { on[j] = ORC_INVALID; od[j] = INFINITY; }
        if (out_count) out_count[qi] = c;
        total += g_evals;
    }
    g_evals_total = total;
}

/* back-link append, hnsw_am.c:2493-2513: first InvalidBlockNumber hole below the count,
 * else the count itself; written only if < 2m.  (The prune block after it, :2515-2612,
 * is unreachable: count can never exceed 2m.) */
/* EXTENSION (insert mode 3): neighbour selection by the diversity heuristic of Malkov & Yashunin
 * (Alg. 4, as in hnswlib's getNeighborsByHeuristic2).  cand/cd: candidates sorted ascending by
 * (distance to the base vector, id); a candidate is kept if it is closer to the base than to every
 * neighbour kept so far; at most `want` are kept.  Returns the number kept (in place, order kept). */
static int select_heuristic(OrcHnsw *g, uint32_t *cand, float *cd, int cc, int want)
{
    int kept = 0;
    for (int i = 0; i < cc && kept < want; i++) {
        const float *cv = g->vec + (size_t) cand[i] * g->dim;
        int good = 1;
        for (int j = 0; j < kept; j++) {
            float d = orc_hnsw_distance(cv, g->vec + (size_t) cand[j] * g->dim, g->dim, 1);
            g_evals++;
            if (d < cd[i]) { good = 0; break; }
        }
        if (good) { cand[kept] = cand[i]; cd[kept] = cd[i]; kept++; }
    }
    return kept;
}

/* ascending by (distance, id): insertion sort, the lists are short */
static void sort_by_dist_id(uint32_t *cand, float *cd, int cc)
{
    for (int i = 1; i < cc; i++) {
        uint32_t c = cand[i]; float d = cd[i];
        int j = i - 1;
        while (j >= 0 && (cd[j] > d || (cd[j] == d && cand[j] > c))) { cand[j + 1] = cand[j]; cd[j + 1] = cd[j]; j--; }
        cand[j + 1] = c; cd[j + 1] = d;
    }
}

static void backlink(OrcHnsw *g, uint32_t nb, int lev, uint32_t newnode, int mode, float d_new)
{
    uint32_t *s = slots(g, nb, lev);
    if (!s) return;     /* reference writes out of the node's bounds here: skipped */
    int nc = clamp_cnt(g, nb, lev);
    int insertPos = nc;
    for (int j = 0; j < nc; j++)
        if (s[j] == ORC_INVALID) { insertPos = j; break; }
    if (insertPos < g->m * 2) {
        s[insertPos] = newnode;
        if (insertPos >= nc)
            g->cnt[(size_t) nb * ORC_HNSW_MAX_LEVEL + lev] = (int16_t) (insertPos + 1);
    } else if (mode == 3) {
        /* full: re-select nb's neighbours among the current ones and the new node (hnswlib's
         * mutuallyConnectNewElement) */
        const int m2 = g->m * 2;
        uint32_t tc[ORC_HNSW_MAX_M2 + 1];
        float td[ORC_HNSW_MAX_M2 + 1];
        const float *nv = g->vec + (size_t) nb * g->dim;
        int cc = 0;
        for (int j = 0; j < nc; j++) {
            tc[cc] = s[j];
            td[cc] = orc_hnsw_distance(nv, g->vec + (size_t) s[j] * g->dim, g->dim, 1);
            g_evals++;
            cc++;
        }
        tc[cc] = newnode; td[cc] = d_new; cc++;
        sort_by_dist_id(tc, td, cc);
        int kept = select_heuristic(g, tc, td, cc, m2);
        for (int j = 0; j < m2; j++) s[j] = j < kept ? tc[j] : ORC_INVALID;
        g->cnt[(size_t) nb * ORC_HNSW_MAX_LEVEL + lev] = (int16_t) kept;
    } else if (mode == 2) {
        /* EXTENSION (insert mode 2): what the reference's unreachable prune block (:2515-2612,
         * "prune to at most m*2 nearest neighbors") is after, restated as a replacement: the new
         * node takes the slot of the farthest current neighbour (hnswComputeDistance, strategy 1;
         * ties -> the later slot) if it is strictly closer to nb than that neighbour. */
        const float *nv = g->vec + (size_t) nb * g->dim;
        int far = -1;
        float fd = 0.0f;
        for (int j = 0; j < nc; j++) {
            float d = orc_hnsw_distance(nv, g->vec + (size_t) s[j] * g->dim, g->dim, 1);
            g_evals++;
            if (far < 0 || d >= fd) { far = j; fd = d; }
        }
        if (far >= 0 && d_new < fd) s[far] = newnode;
    }
}

void orc_hnsw_insert(OrcHnsw *g, const float *vec, int level, int mode)
{
    if (g->n >= g->cap) return;
    if (level >= ORC_HNSW_MAX_LEVEL) level = ORC_HNSW_MAX_LEVEL - 1;
    if (level < 0) level = 0;
    const int m = g->m;
    uint32_t blkno = (uint32_t) g->n;

    /* Step 2/4: node with empty neighbour slots (memset 0xFF, :2149), one node per page */
    memcpy(g->vec + (size_t) blkno * g->dim, vec, sizeof(float) * (size_t) g->dim);
    g->level[blkno] = level;
    memset(g->cnt + (size_t) blkno * ORC_HNSW_MAX_LEVEL, 0, sizeof(int16_t) * ORC_HNSW_MAX_LEVEL);
    memset(g->nbr0 + (size_t) blkno * 2 * m, 0xFF, sizeof(uint32_t) * 2 * (size_t) m);
    g->upper[blkno] = NULL;
    if (level > 0) {
        g->upper[blkno] = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) level * 2 * m);
        memset(g->upper[blkno], 0xFF, sizeof(uint32_t) * (size_t) level * 2 * m);
    }
    g->n++;
    /* Step 3 (:2156-2286) computes bestEntry and never uses it: no effect. */

    /* Step 5, :2334-2645 */
    if (g->entry != ORC_INVALID && g->entry_level >= 0) {
        int maxLevel = level < g->entry_level ? level : g->entry_level;
        int efc = g->efc;
        uint32_t *cand = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) efc);
        float *cd = (float *) malloc(sizeof(float) * (size_t) efc);
        uint8_t *visited = (uint8_t *) malloc((size_t) g->n);
        uint32_t ep = g->entry;
        if (mode >= 1)  /* ADDITION: descend once to maxLevel+1, then carry ep level to level */
            ep = greedy_descent(g, vec, 1, g->entry, g->entry_level, maxLevel);

        for (int lev = maxLevel; lev >= 0; lev--) {
            int cc;
            memset(visited, 0, (size_t) g->n);
            if (mode == 0) {
                /* hnswSearch(index, meta, vector, dim, 1, efC, efC): full descent to level 0
                 * on every iteration */
                int cl = g->entry_level;
                if (cl < 0 || cl >= ORC_HNSW_MAX_LEVEL) cl = 0;
                uint32_t cur = greedy_descent(g, vec, 1, g->entry, cl, 0);
                cc = search_level0_literal(g, vec, 1, cur, efc, efc, visited, cand, cd);
            } else {
                cc = search_layer_bestfirst(g, vec, 1, ep, lev, efc, visited, blkno, cand, cd, efc);
                if (cc > 0) ep = cand[0];
            }
            int selectedCount = m < cc ? m : cc;
            if (mode == 3) {
                sort_by_dist_id(cand, cd, cc);
                selectedCount = select_heuristic(g, cand, cd, cc, m);
            }
            /* "closest m" selection sort with swaps, :2386-2424 */
            for (int idx = 0; idx < selectedCount && mode != 3; idx++) {
                int bestIdx = idx;
                float bestDist = cd[idx];
                for (int j = idx + 1; j < cc; j++)
                    if (cd[j] < bestDist) { bestDist = cd[j]; bestIdx = j; }
                if (bestIdx != idx) {
                    uint32_t tb = cand[idx]; float td = cd[idx];
                    cand[idx] = cand[bestIdx]; cd[idx] = cd[bestIdx];
                    cand[bestIdx] = tb; cd[bestIdx] = td;
                }
            }
            /* forward links + capped back-links, :2452-2513 */
            uint32_t *mine = slots(g, blkno, lev);
            for (int idx = 0; idx < selectedCount; idx++) {
                if (idx < m) {
                    mine[idx] = cand[idx];
                    g->cnt[(size_t) blkno * ORC_HNSW_MAX_LEVEL + lev] = (int16_t) (idx + 1);
                }
                backlink(g, cand[idx], lev, blkno, mode, cd[idx]);
            }
        }
        free(cand); free(cd); free(visited);
    }

    /* Step 6, :2649-2666 */
    if (g->entry == ORC_INVALID || level > g->entry_level) {
        g->entry = blkno;
        g->entry_level = level;
    }
    if (level > g->max_level) g->max_level = level;
}

void orc_hnsw_build(OrcHnsw *g, const float *X, int64_t n, const int *levels, int mode)
{
    for (int64_t i = 0; i < n; i++)
        orc_hnsw_insert(g, X + (size_t) i * g->dim, levels ? levels[i] : orc_hnsw_random_level(g->ml), mode);
}

int64_t orc_hnsw_size(const OrcHnsw *g) { return g->n; }

void orc_hnsw_meta(const OrcHnsw *g, uint32_t *entry_point, int *entry_level, int *max_level)
{
    *entry_point = g->entry; *entry_level = g->entry_level; *max_level = g->max_level;
}

int64_t orc_hnsw_upper_slots(const OrcHnsw *g)
{
    int64_t s = 0;
    for (int64_t i = 0; i < g->n; i++) s += (int64_t) g->level[i] * 2 * g->m;
    return s;
}

void orc_hnsw_export(const OrcHnsw *g, int *levels, uint32_t *nbr0, int16_t *cnt,
                     int64_t *upper_off, uint32_t *upper)
{
    int64_t off = 0;
    memcpy(levels, g->level, sizeof(int) * (size_t) g->n);
    memcpy(nbr0, g->nbr0, sizeof(uint32_t) * (size_t) g->n * 2 * g->m);
    memcpy(cnt, g->cnt, sizeof(int16_t) * (size_t) g->n * ORC_HNSW_MAX_LEVEL);
    for (int64_t i = 0; i < g->n; i++) {
        upper_off[i] = off;
        int64_t c = (int64_t) g->level[i] * 2 * g->m;
        if (c > 0) memcpy(upper + off, g->upper[i], sizeof(uint32_t) * (size_t) c);
        off += c;
    }
    upper_off[g->n] = off;
}
