/*
 * index_am_b200.c -- the reference-side glue for the index access methods (SURVEY.md 8b surface b1, 8f-1):
 * what ivf_am.c / hnsw_am.c call instead of their own scan loops.  Two pieces:
 *
 *   1. Staging shim.  The device copy of an index is derived and rebuildable; after a restart (or an
 *      invalidation) the AM streams the relation's 8 KB blocks -- the loop ivfCollectCandidates and hnswSearch
 *      run per query today (ReadBuffer + LockBuffer(BUFFER_LOCK_SHARE) + BufferGetPage, ivf_am.c:1795-1840,
 *      hnsw_am.c:1843-1920) -- ONCE into pinned host memory and hands them to ndb_b200_{ivf,hnsw}_load_relation,
 *      which decodes the page layouts on the device.  The buffer manager stays on the PostgreSQL side: the AM
 *      passes a read callback, so this file needs no bufmgr.h.
 *
 *   2. Scan state.  ambeginscan / amrescan / amgettuple / amendscan keep their signatures and their state
 *      machine (IvfScanOpaqueData ivf_am.c:266-281, callbacks :1412-2048; HnswScanOpaqueData hnsw_am.c:205-216,
 *      :904-1056): rescan stores the ORDER BY vector, the first gettuple runs the search, every gettuple hands
 *      back the next heap TID in ascending distance with its distance for xs_orderbyvals.
 *
 * Compiled against oracle/pgshim (ItemPointerData, elog levels) by oracle/Makefile, target `glue`; exercised on
 * the GPU by tests/test_gpu_boundary.py against the oracle's scan of the same relation image.
 */
#include "postgres.h"
#include "ndb_b200.h"

#define NDB_BLCKSZ 8192

/* ---- 1. staging shim ---------------------------------------------------------------------------------- */

/* returns the page of block `blkno`, pinned and share-locked, or NULL; release undoes both */
typedef const void *(*ndb_b200_read_block_fn) (void *arg, uint32 blkno);
typedef void (*ndb_b200_release_block_fn) (void *arg, uint32 blkno);

static int
stage_blocks(ndb_b200_read_block_fn rd, ndb_b200_release_block_fn rel, void *arg, uint32 nblocks, void **staging)
{
	uint32		b;
	int			rc;

	*staging = NULL;
	if (rd == NULL || nblocks == 0)
		return NDB_B200_EINVAL;
	/* pinned: the H2D copy inside load_relation is then a true DMA of the whole image */
	rc = ndb_b200_host_alloc_pinned(staging, (size_t) nblocks * NDB_BLCKSZ);
	if (rc != NDB_B200_OK)
		return rc;
	for (b = 0; b < nblocks; b++)
	{
		const void *page = rd(arg, b);

		if (page == NULL)
		{
			ndb_b200_host_free_pinned(*staging);
			*staging = NULL;
			return NDB_B200_ESTATE;
		}
		memcpy((char *) *staging + (size_t) b * NDB_BLCKSZ, page, NDB_BLCKSZ);
		if (rel != NULL)
			rel(arg, b);
	}
	return NDB_B200_OK;
}

int
ndb_b200_am_stage_ivf(ndb_b200_ivf *ix, ndb_b200_read_block_fn rd, ndb_b200_release_block_fn rel, void *arg,
					  uint32 nblocks)
{
	void	   *staging;
	int			rc = stage_blocks(rd, rel, arg, nblocks, &staging);

	if (rc != NDB_B200_OK)
		return rc;
	rc = ndb_b200_ivf_load_relation(ix, staging, nblocks);
	ndb_b200_host_free_pinned(staging);
	return rc;
}

int
ndb_b200_am_stage_hnsw(ndb_b200_hnsw *h, ndb_b200_read_block_fn rd, ndb_b200_release_block_fn rel, void *arg,
					   uint32 nblocks)
{
	void	   *staging;
	int			rc = stage_blocks(rd, rel, arg, nblocks, &staging);

	if (rc != NDB_B200_OK)
		return rc;
	rc = ndb_b200_hnsw_load_relation(h, staging, nblocks);
	ndb_b200_host_free_pinned(staging);
	return rc;
}

/* a reader over an in-memory relation image (tests; also what a CREATE INDEX that still holds its pages uses) */
const void *
ndb_b200_am_image_reader(void *arg, uint32 blkno)
{
	return (const char *) arg + (size_t) blkno * NDB_BLCKSZ;
}

/* ---- 2. scan state ------------------------------------------------------------------------------------ */

#define NDB_AM_MAX_K 128

typedef struct NdbB200Scan
{
	ndb_b200_ivf *ivf;			/* exactly one of the two */
	ndb_b200_hnsw *hnsw;
	int			dim;
	int			k;				/* so->k: 10 by default (ivf_am.c:1422,1543; hnsw_am.c:929) */
	int			nprobe;			/* IVF: meta->nprobe / neurondb.ivf_probes */
	int			ef_search;		/* HNSW: meta->efSearch / neurondb.hnsw_ef_search */
	int			mode;			/* NDB_IVF_LITERAL = ivfCollectCandidates as written; NDB_IVF_FULL scans the probed lists */
	int			arith;
	float	   *query;			/* so->queryVector, owned */
	bool		have_query;
	bool		executed;		/* so->firstCall inverted */
	int			n_results;
	int			cursor;			/* so->currentResult */
	int64_t		tids[NDB_AM_MAX_K];
	float		dist[NDB_AM_MAX_K];
} NdbB200Scan;

static NdbB200Scan *
scan_new(int dim)
{
	NdbB200Scan *so = (NdbB200Scan *) calloc(1, sizeof(NdbB200Scan));	/* palloc0 in the extension */

	if (so == NULL)
		return NULL;
	so->dim = dim;
	so->k = 10;
	so->query = (float *) calloc((size_t) (dim > 0 ? dim : 1), sizeof(float));
	if (so->query == NULL)
	{
		free(so);
		return NULL;
	}
	return so;
}

/* ivfbeginscan (:1412-1437) */
NdbB200Scan *
ndb_b200_am_ivf_beginscan(ndb_b200_ivf *ix, int nprobe, int mode, int arith)
{
	NdbB200Scan *so;

	if (ix == NULL || nprobe < 1)
		return NULL;
	so = scan_new(ndb_b200_ivf_dim(ix));
	if (so == NULL)
		return NULL;
	so->ivf = ix;
	so->nprobe = nprobe;
	so->mode = mode;
	so->arith = arith;
	return so;
}

/* hnswbeginscan (:878-902) */
NdbB200Scan *
ndb_b200_am_hnsw_beginscan(ndb_b200_hnsw *h, int dim, int ef_search, int mode)
{
	NdbB200Scan *so;

	if (h == NULL || ef_search < 1)
		return NULL;
	so = scan_new(dim);
	if (so == NULL)
		return NULL;
	so->hnsw = h;
	so->ef_search = ef_search;
	so->mode = mode;
	return so;
}

/* ivfrescan (:1439-1545) / hnswrescan (:904-976): store the ORDER BY vector, reset the cursor.
 * A vector of another dimension is the reference's "dimensions must match" error. */
int
ndb_b200_am_rescan(NdbB200Scan *so, const float *query, int dim, int k)
{
	if (so == NULL || query == NULL)
		return NDB_B200_EINVAL;
	if (dim != so->dim)
		return NDB_B200_EDIM;
	if (k < 1 || k > NDB_AM_MAX_K)
		return NDB_B200_EINVAL;
	memcpy(so->query, query, (size_t) dim * sizeof(float));
	so->k = k;
	so->have_query = true;
	so->executed = false;
	so->n_results = 0;
	so->cursor = 0;
	return NDB_B200_OK;
}

/* ivfgettuple (:1911-2027) / hnswgettuple (:978-1056): 1 = a tuple was produced, 0 = no more, < 0 = error
 * (the AM turns it into ereport(ERROR)).  The first call after rescan runs the search. */
int
ndb_b200_am_gettuple(NdbB200Scan *so, ItemPointerData *tid, float *orderby_distance)
{
	int64_t		packed;

	if (so == NULL || tid == NULL)
		return NDB_B200_EINVAL;
	if (!so->have_query)
		return 0;				/* no ORDER BY key: nothing to return (:1927-1934) */
	if (!so->executed)
	{
		int			rc;
		int			i;

		if (so->ivf != NULL)
			rc = ndb_b200_ivf_search(so->ivf, so->query, 1, so->nprobe, so->k, so->mode, so->arith, so->dist, so->tids);
		else					/* the AMs pass strategy 1 whatever the opclass */
			rc = ndb_b200_hnsw_search(so->hnsw, so->query, 1, 1, so->ef_search, so->k, so->mode, so->dist, so->tids);
		if (rc != NDB_B200_OK)
			return rc;
		so->n_results = 0;
		for (i = 0; i < so->k && so->tids[i] >= 0; i++)
			so->n_results++;
		so->executed = true;
		so->cursor = 0;
	}
	if (so->cursor >= so->n_results)
		return 0;
	packed = so->tids[so->cursor];
	/* ids carry the heap TID as (block << 16) | offset (the relation loaders pack it so) */
	tid->bi_hi = (uint16) ((packed >> 32) & 0xffff);
	tid->bi_lo = (uint16) ((packed >> 16) & 0xffff);
	tid->ip_posid = (uint16) (packed & 0xffff);
	if (orderby_distance != NULL)
		*orderby_distance = so->dist[so->cursor];
	so->cursor++;
	return 1;
}

/* ivfendscan (:2029-2048) / hnswendscan */
void
ndb_b200_am_endscan(NdbB200Scan *so)
{
	if (so == NULL)
		return;
	free(so->query);
	free(so);
}
