/*
 * ndb_oracle_pages.c -- index relation images in the reference's on-disk layout
 * (TEST INFRASTRUCTURE ONLY).
 *
 * Builds, in memory, the 8 KB blocks that ivfbuild + a sequence of ivfinsert calls
 * (NeuronDB/src/index/ivf_am.c:556-577, 640-711, 982-1157) and hnswInsertNode
 * (NeuronDB/src/index/hnsw_am.c:1090-1110, 2121-2153, 2288-2332) leave behind, so that the
 * library's relation loaders (ndb_b200_ivf_load_relation / ndb_b200_hnsw_load_relation) can
 * be tested against byte-exact pages.  The page primitives below restate the documented
 * PostgreSQL page format (PageHeaderData 24 B, ItemIdData {lp_off:15, lp_flags:2, lp_len:15},
 * tuples packed downwards from pd_upper, special space at the page tail).
 */
#include "ndb_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <stddef.h>

#define BLCKSZ 8192
#define PAGE_HEADER 24
#define MAXALIGN8(x) (((size_t) (x) + 7) & ~(size_t) 7)
#define LP_NORMAL 1
#define LP_DEAD 3

typedef struct {
    uint64_t pd_lsn;
    uint16_t pd_checksum, pd_flags, pd_lower, pd_upper, pd_special, pd_pagesize_version;
    uint32_t pd_prune_xid;
} PgPageHeader;

static void page_init(uint8_t *page, size_t special)
{
    PgPageHeader *h = (PgPageHeader *) page;
    memset(page, 0, BLCKSZ);
    special = MAXALIGN8(special);
    h->pd_lower = PAGE_HEADER;
    h->pd_upper = (uint16_t) (BLCKSZ - special);
    h->pd_special = (uint16_t) (BLCKSZ - special);
    h->pd_pagesize_version = BLCKSZ | 4;
}

/* PageGetFreeSpace: room for one more item including its line pointer */
static size_t page_free_space(const uint8_t *page)
{
    const PgPageHeader *h = (const PgPageHeader *) page;
    int space = (int) h->pd_upper - (int) h->pd_lower;
    if (space < 4) return 0;
    return (size_t) space - 4;
}

/* PageAddItem(page, item, size, InvalidOffsetNumber, false, false): returns the 1-based offset, 0 on failure */
static int page_add_item(uint8_t *page, const void *item, size_t size)
{
    PgPageHeader *h = (PgPageHeader *) page;
    size_t aligned = MAXALIGN8(size);
    int lower = h->pd_lower + 4;
    int upper = (int) h->pd_upper - (int) aligned;
    if (lower > upper) return 0;
    int off = (h->pd_lower - PAGE_HEADER) / 4 + 1;
    uint32_t lp = ((uint32_t) upper & 0x7fff) | ((uint32_t) LP_NORMAL << 15) | ((uint32_t) size << 17);
    memcpy(page + h->pd_lower, &lp, 4);
    memcpy(page + upper, item, size);
    h->pd_lower = (uint16_t) lower;
    h->pd_upper = (uint16_t) upper;
    return off;
}

static uint8_t *page_item(uint8_t *page, int offnum)
{
    uint32_t lp;
    memcpy(&lp, page + PAGE_HEADER + 4 * (offnum - 1), 4);
    return page + (lp & 0x7fff);
}

/* heap TID packed as (block << 16) | offset  ->  ItemPointerData {bi_hi, bi_lo, ip_posid} */
static void tid_pack(int64_t tid, uint8_t *out6)
{
    uint32_t blk = (uint32_t) (tid >> 16);
    uint16_t v[3] = { (uint16_t) (blk >> 16), (uint16_t) (blk & 0xffff), (uint16_t) (tid & 0xffff) };
    memcpy(out6, v, 6);
}

/* item layouts the encoders below write; orc_page_layout() reports them (and the struct sizes / offsets)
 * in the order of the reference-side ref_layout() of oracle/extract_ref_leafs.py, so that the test can hold
 * them against the reference's own struct definitions */
#define IVF_ENTRY_TID_OFF 0      /* IvfListEntryData.heapPtr (ivf_am.c:251-256) */
#define IVF_ENTRY_DIM_OFF 6      /* IvfListEntryData.dim */
#define IVF_ENTRY_HDR 8          /* sizeof(IvfListEntryData); the vector follows at MAXALIGN of it */
#define HNSW_NODE_TID_OFF 0      /* HnswNodeData.heapPtr (hnsw_am.c:124-136) */
#define HNSW_NODE_LEVEL_OFF 8
#define HNSW_NODE_DIM_OFF 12
#define HNSW_NODE_CNT_OFF 14     /* int16 neighborCount[16] */
#define HNSW_NODE_HDR 48         /* sizeof(HnswNodeData) = MAXALIGN of it; vector, then neighbours[level+1][2m] */

/* ---- IVF ------------------------------------------------------------------------------- */
typedef struct { uint32_t magic, version; int32_t nlists, nprobe, dim; uint32_t centroidsBlock; int64_t insertedVectors; } IvfMeta;           /* ivf_am.c:75-84 */
typedef struct { int32_t listId, dim; int64_t memberCount; uint32_t firstBlock; uint32_t pad; } IvfCentroidHdr;                                /* :94-103, 24 B  */
typedef struct { uint32_t nextBlock; int32_t entryCount; } IvfListSpecial;                                                                       /* :241-246       */

/* Returns the number of blocks written, or -1 (centroids do not fit one page and
 * multi_page_centroids == 0: the reference's "ivf: failed to add centroid to page"),
 * or -2 (cap_blocks too small). */
int64_t orc_ivf_encode_relation(const float *X, const int64_t *tids, int64_t n, int dim,
                                const float *C, int nlists, int nprobe, const int *assign,
                                uint8_t *blocks, int64_t cap_blocks, int multi_page_centroids)
{
    if (cap_blocks < 2) return -2;
    int64_t nblocks = 0;
    /* meta page, block 0 (ivf_am.c:556-577) */
    uint8_t *metaPage = blocks;
    page_init(metaPage, sizeof(IvfMeta));
    IvfMeta *meta = (IvfMeta *) (metaPage + PAGE_HEADER);
    meta->magic = 0x49564646u;
    meta->version = 1;
    meta->nlists = nlists;
    meta->nprobe = nprobe;
    meta->dim = dim;
    meta->centroidsBlock = 1;
    meta->insertedVectors = 0;
    nblocks = 1;

    /* centroid page(s) (ivf_am.c:640-711).  The reference uses exactly one page; when
     * multi_page_centroids is set the overflow continues on consecutive blocks. */
    const size_t csize = MAXALIGN8(sizeof(IvfCentroidHdr) + (size_t) dim * 4);
    uint8_t *item = (uint8_t *) calloc(1, csize + (size_t) dim * 4 + 64);
    uint32_t *cblk = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) nlists);   /* where centroid i lives */
    int *coff = (int *) malloc(sizeof(int) * (size_t) nlists);
    uint8_t *cpage = blocks + (size_t) nblocks * BLCKSZ;
    page_init(cpage, sizeof(IvfCentroidHdr));
    nblocks++;
    for (int i = 0; i < nlists; i++) {
        IvfCentroidHdr hdr = { i, dim, 0, ORC_INVALID, 0 };
        memset(item, 0, csize);
        memcpy(item, &hdr, sizeof(hdr));
        memcpy(item + MAXALIGN8(sizeof(IvfCentroidHdr)), C + (size_t) i * dim, (size_t) dim * 4);
        int off = page_add_item(cpage, item, csize);
        if (!off) {
            if (!multi_page_centroids) { free(item); free(cblk); free(coff); return -1; }
            if (nblocks >= cap_blocks) { free(item); free(cblk); free(coff); return -2; }
            cpage = blocks + (size_t) nblocks * BLCKSZ;
            page_init(cpage, sizeof(IvfCentroidHdr));
            nblocks++;
            off = page_add_item(cpage, item, csize);
        }
        cblk[i] = (uint32_t) (nblocks - 1);
        coff[i] = off;
    }

    /* ivfinsert for every row in order (ivf_am.c:982-1157) */
    const size_t entrySize = MAXALIGN8(8) + MAXALIGN8((size_t) dim * 4);
    uint32_t *last = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) nlists);
    for (int i = 0; i < nlists; i++) last[i] = ORC_INVALID;
    uint8_t *entry = (uint8_t *) calloc(1, entrySize + 16);
    for (int64_t r = 0; r < n; r++) {
        const int l = assign[r];
        IvfCentroidHdr *cen = (IvfCentroidHdr *) page_item(blocks + (size_t) cblk[l] * BLCKSZ, coff[l]);
        uint8_t *lpage;
        if (last[l] == ORC_INVALID) {
            if (nblocks >= cap_blocks) { nblocks = -2; break; }
            lpage = blocks + (size_t) nblocks * BLCKSZ;
            page_init(lpage, sizeof(IvfListSpecial));
            ((IvfListSpecial *) (lpage + BLCKSZ - 8))->nextBlock = ORC_INVALID;
            cen->firstBlock = (uint32_t) nblocks;
            last[l] = (uint32_t) nblocks;
            nblocks++;
        } else {
            lpage = blocks + (size_t) last[l] * BLCKSZ;
            if (page_free_space(lpage) < entrySize) {
                if (nblocks >= cap_blocks) { nblocks = -2; break; }
                uint8_t *np = blocks + (size_t) nblocks * BLCKSZ;
                page_init(np, sizeof(IvfListSpecial));
                ((IvfListSpecial *) (np + BLCKSZ - 8))->nextBlock = ORC_INVALID;
                ((IvfListSpecial *) (lpage + BLCKSZ - 8))->nextBlock = (uint32_t) nblocks;
                last[l] = (uint32_t) nblocks;
                lpage = np;
                nblocks++;
            }
        }
        memset(entry, 0, entrySize);
        tid_pack(tids ? tids[r] : r, entry + IVF_ENTRY_TID_OFF);
        int16_t d16 = (int16_t) dim;
        memcpy(entry + IVF_ENTRY_DIM_OFF, &d16, 2);
        memcpy(entry + MAXALIGN8(IVF_ENTRY_HDR), X + (size_t) r * dim, (size_t) dim * 4);
        if (!page_add_item(lpage, entry, entrySize)) { nblocks = -2; break; }
        ((IvfListSpecial *) (lpage + BLCKSZ - 8))->entryCount++;
        cen->memberCount++;
        meta->insertedVectors++;
    }
    free(entry); free(last); free(item); free(cblk); free(coff);
    return nblocks;
}

/* ItemIdSetDead on (block, 1-based offset): what bulkdelete leaves and scans skip (ivf_am.c:1816) */
void orc_page_mark_dead(uint8_t *blocks, int64_t block, int offnum)
{
    uint8_t *page = blocks + (size_t) block * BLCKSZ;
    uint32_t lp;
    memcpy(&lp, page + PAGE_HEADER + 4 * (offnum - 1), 4);
    lp = (lp & ~(3u << 15)) | ((uint32_t) LP_DEAD << 15);
    memcpy(page + PAGE_HEADER + 4 * (offnum - 1), &lp, 4);
}

/* ---- HNSW ------------------------------------------------------------------------------ */
typedef struct { uint32_t magic, version, entryPoint; int32_t entryLevel, maxLevel; int16_t m, efConstruction, efSearch; int16_t pad; float ml; int64_t insertedVectors; } HnswMeta;   /* hnsw_am.c:108-120, 40 B */

/* one node per page, node i -> block i + 1; neighbour slots hold block numbers (hnsw_am.c:124-181) */
int64_t orc_hnsw_encode_relation(const OrcHnsw *g, const float *X, const int64_t *tids, int dim, int m, int efc,
                                 int efs, uint8_t *blocks, int64_t cap_blocks)
{
    const int64_t n = orc_hnsw_size(g);
    if (cap_blocks < n + 1) return -2;
    const int m2 = 2 * m;
    uint32_t ep; int el, ml;
    orc_hnsw_meta(g, &ep, &el, &ml);
    int *levels = (int *) malloc(sizeof(int) * (size_t) n);
    int64_t ups = orc_hnsw_upper_slots(g);
    int64_t *uoff = (int64_t *) malloc(sizeof(int64_t) * (size_t) (n + 1));
    uint32_t *upper = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) (ups > 0 ? ups : 1));
    int16_t *cnt = (int16_t *) malloc(sizeof(int16_t) * (size_t) n * ORC_HNSW_MAX_LEVEL);
    uint32_t *nbr0 = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) n * (size_t) m2);
    orc_hnsw_export(g, levels, nbr0, cnt, uoff, upper);

    page_init(blocks, sizeof(HnswMeta));
    HnswMeta *meta = (HnswMeta *) (blocks + PAGE_HEADER);
    meta->magic = 0x48534E57u;
    meta->version = 1;
    meta->entryPoint = ep == ORC_INVALID ? ORC_INVALID : ep + 1;
    meta->entryLevel = el;
    meta->maxLevel = ml;
    meta->m = (int16_t) (m2 / 2);
    meta->efConstruction = (int16_t) efc;
    meta->efSearch = (int16_t) efs;
    meta->ml = 0.36f;
    meta->insertedVectors = n;

    const size_t hdr = HNSW_NODE_HDR;
    uint8_t *item = (uint8_t *) malloc(BLCKSZ);
    int64_t rc = n + 1;
    for (int64_t i = 0; i < n; i++) {
        const int level = levels[i];
        const size_t size = MAXALIGN8(hdr + (size_t) dim * 4 + (size_t) (level + 1) * m2 * 4);
        if (size > BLCKSZ - PAGE_HEADER - 4) { rc = -1; break; }
        memset(item, 0, size);
        tid_pack(tids ? tids[i] : i, item + HNSW_NODE_TID_OFF);
        int32_t lv = level;
        int16_t d16 = (int16_t) dim;
        memcpy(item + HNSW_NODE_LEVEL_OFF, &lv, 4);
        memcpy(item + HNSW_NODE_DIM_OFF, &d16, 2);
        memcpy(item + HNSW_NODE_CNT_OFF, cnt + (size_t) i * ORC_HNSW_MAX_LEVEL, 2 * ORC_HNSW_MAX_LEVEL);
        memcpy(item + hdr, X + (size_t) i * dim, (size_t) dim * 4);
        uint32_t *nb = (uint32_t *) (item + hdr + (size_t) dim * 4);
        for (int l = 0; l <= level; l++)
            for (int j = 0; j < m2; j++) {
                uint32_t v = l == 0 ? nbr0[(size_t) i * m2 + j] : upper[uoff[i] + (size_t) (l - 1) * m2 + j];
                nb[(size_t) l * m2 + j] = v == ORC_INVALID ? ORC_INVALID : v + 1;
            }
        uint8_t *page = blocks + (size_t) (i + 1) * BLCKSZ;
        page_init(page, 0);
        if (!page_add_item(page, item, size)) { rc = -1; break; }
    }
    free(item); free(nbr0); free(cnt); free(upper); free(uoff); free(levels);
    return rc;
}


/* sizes and offsets the encoders above use, in the order of ref_layout() (oracle/extract_ref_leafs.py) */
int orc_page_layout(int64_t *o)
{
    int n = 0;
    o[n++] = sizeof(IvfMeta);
    o[n++] = offsetof(IvfMeta, magic); o[n++] = offsetof(IvfMeta, version);
    o[n++] = offsetof(IvfMeta, nlists); o[n++] = offsetof(IvfMeta, nprobe); o[n++] = offsetof(IvfMeta, dim);
    o[n++] = offsetof(IvfMeta, centroidsBlock); o[n++] = offsetof(IvfMeta, insertedVectors);
    o[n++] = 0x49564646;
    o[n++] = sizeof(IvfCentroidHdr); o[n++] = MAXALIGN8(sizeof(IvfCentroidHdr));
    o[n++] = offsetof(IvfCentroidHdr, listId); o[n++] = offsetof(IvfCentroidHdr, dim);
    o[n++] = offsetof(IvfCentroidHdr, memberCount); o[n++] = offsetof(IvfCentroidHdr, firstBlock);
    o[n++] = sizeof(IvfListSpecial); o[n++] = offsetof(IvfListSpecial, nextBlock); o[n++] = offsetof(IvfListSpecial, entryCount);
    o[n++] = IVF_ENTRY_HDR; o[n++] = MAXALIGN8(IVF_ENTRY_HDR);
    o[n++] = IVF_ENTRY_TID_OFF; o[n++] = IVF_ENTRY_DIM_OFF;
    o[n++] = sizeof(HnswMeta);
    o[n++] = offsetof(HnswMeta, magic); o[n++] = offsetof(HnswMeta, version); o[n++] = offsetof(HnswMeta, entryPoint);
    o[n++] = offsetof(HnswMeta, entryLevel); o[n++] = offsetof(HnswMeta, maxLevel); o[n++] = offsetof(HnswMeta, m);
    o[n++] = offsetof(HnswMeta, efConstruction); o[n++] = offsetof(HnswMeta, efSearch); o[n++] = offsetof(HnswMeta, ml);
    o[n++] = offsetof(HnswMeta, insertedVectors);
    o[n++] = 0x48534E57;
    o[n++] = HNSW_NODE_HDR; o[n++] = MAXALIGN8(HNSW_NODE_HDR);
    o[n++] = HNSW_NODE_TID_OFF; o[n++] = HNSW_NODE_LEVEL_OFF; o[n++] = HNSW_NODE_DIM_OFF; o[n++] = HNSW_NODE_CNT_OFF;
    o[n++] = (int64_t) MAXALIGN8(HNSW_NODE_HDR + (size_t) 768 * 4 + (size_t) (0 + 1) * 32 * 4);
    o[n++] = (int64_t) MAXALIGN8(HNSW_NODE_HDR + (size_t) 128 * 4 + (size_t) (2 + 1) * 16 * 4);
    o[n++] = (int64_t) (HNSW_NODE_HDR + (size_t) 768 * 4 + (size_t) 1 * 32 * 4);       /* level-1 neighbours, 768-d, m = 16 */
    o[n++] = (int64_t) (HNSW_NODE_HDR + (size_t) 768 * 4);
    return n;
}
