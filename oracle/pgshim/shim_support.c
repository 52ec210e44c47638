/*
 * pgshim/shim_support.c -- runtime half of the PostgreSQL stand-in (TEST INFRASTRUCTURE).
 *
 * Linked with the reference's unmodified NeuronDB/src/vector/vector_distance.c and
 * vector_distance_simd.c (compiled where they lie under /root/reference; never copied)
 * into oracle/_ref/libndb_ref_distance*.so.  Exposes plain-C entry points that build
 * `Vector` varlenas and call the reference's fmgr functions -- the PROCEDUREs behind
 * <->, <=>, <#> (NeuronDB/src/index/opclass.c:53-161 forward to the same
 * l2_distance_simd / cosine_distance_simd / inner_product_simd dispatchers).
 */
#include "postgres.h"
#include "fmgr.h"
#include "neurondb.h"

#include <stdarg.h>

MemoryContext CurrentMemoryContext = NULL;

static __thread jmp_buf ndb_jmp;
static __thread int ndb_jmp_armed = 0;
static __thread char ndb_msg[512];
static __thread char ndb_last_error[512];

const char *ndb_shim_fmt(const char *fmt, ...)
{
	va_list		ap;

	va_start(ap, fmt);
	vsnprintf(ndb_msg, sizeof(ndb_msg), fmt, ap);
	va_end(ap);
	return ndb_msg;
}

void ndb_shim_raise(int elevel, const char *msg)
{
	if (elevel < ERROR)
		return;
	snprintf(ndb_last_error, sizeof(ndb_last_error), "%s", msg ? msg : "");
	if (ndb_jmp_armed)
		longjmp(ndb_jmp, 1);
	fprintf(stderr, "pgshim: ERROR outside a guarded call: %s\n", ndb_last_error);
	abort();
}

void *palloc(Size n) { return malloc(n ? n : 1); }
void *palloc0(Size n) { return calloc(1, n ? n : 1); }
void *repalloc(void *p, Size n) { return realloc(p, n); }
void pfree(void *p) { free(p); }
void *MemoryContextAlloc(MemoryContext c, Size n) { (void) c; return palloc(n); }
void *MemoryContextAllocZero(MemoryContext c, Size n) { (void) c; return palloc0(n); }
MemoryContext MemoryContextSwitchTo(MemoryContext c) { MemoryContext o = CurrentMemoryContext; CurrentMemoryContext = c; return o; }
char *pstrdup(const char *s) { return strdup(s); }

/* helpers some reference headers declare extern */
bool ndb_memory_context_validate(MemoryContext context) { (void) context; return true; }
bool ndb_ensure_memory_context(MemoryContext context) { (void) context; return true; }
void ndb_safe_context_cleanup(MemoryContext context, MemoryContext oldcontext) { (void) context; (void) oldcontext; }
void ndb_track_allocation(void *ptr, const char *alloc_func) { (void) ptr; (void) alloc_func; }
void ndb_untrack_allocation(void *ptr) { (void) ptr; }

extern Datum vector_l2_distance(PG_FUNCTION_ARGS);
extern Datum vector_cosine_distance(PG_FUNCTION_ARGS);
extern Datum vector_inner_product(PG_FUNCTION_ARGS);

static Vector *make_vector(const float *v, int dim)
{
	Vector	   *r = (Vector *) malloc(VECTOR_SIZE(dim > 0 ? dim : 0) + 8);

	SET_VARSIZE(r, VECTOR_SIZE(dim > 0 ? dim : 0));
	r->dim = (int16) dim;
	r->unused = 0;
	if (dim > 0)
		memcpy(r->data, v, sizeof(float) * (size_t) dim);
	return r;
}

const char *ndb_ref_last_error(void) { return ndb_last_error; }

static int call2(int metric, Vector *a, Vector *b, float *out)
{
	FunctionCallInfoBaseData fc;
	Datum		d;

	memset(&fc, 0, sizeof(fc));
	fc.nargs = 2;
	fc.args[0].value = PointerGetDatum(a);
	fc.args[1].value = PointerGetDatum(b);
	ndb_jmp_armed = 1;
	if (setjmp(ndb_jmp))
	{
		ndb_jmp_armed = 0;
		return -1;
	}
	d = metric == 1 ? vector_l2_distance(&fc)
		: metric == 2 ? vector_cosine_distance(&fc)
		: vector_inner_product(&fc);
	ndb_jmp_armed = 0;
	*out = DatumGetFloat4(d);
	return 0;
}

/* one pair through the fmgr entry point; 0 = ok, -1 = the reference raised ERROR */
int ndb_ref_distance(int metric, const float *a, int dim_a, const float *b, int dim_b, float *out)
{
	Vector	   *va = make_vector(a, dim_a);
	Vector	   *vb = make_vector(b, dim_b);
	int			rc = call2(metric, va, vb, out);

	free(va);
	free(vb);
	return rc;
}

/* n pairs (row i of A with row i of B) */
int ndb_ref_distance_pairs(int metric, const float *A, const float *B, float *out, long n, int dim)
{
	Vector	   *va = make_vector(A, dim);
	Vector	   *vb = make_vector(B, dim);
	int			rc = 0;

	for (long i = 0; i < n && rc == 0; i++)
	{
		memcpy(va->data, A + (size_t) i * dim, sizeof(float) * (size_t) dim);
		memcpy(vb->data, B + (size_t) i * dim, sizeof(float) * (size_t) dim);
		rc = call2(metric, va, vb, out + i);
	}
	free(va);
	free(vb);
	return rc;
}

/* SeqScan leg: every row of X against one query, one fmgr call per row;
 * rows are pre-built varlenas so the timing is the operator, not the wrapper */
typedef struct RefTable { long n; int dim; Vector **rows; } RefTable;

RefTable *ndb_ref_table_create(const float *X, long n, int dim)
{
	RefTable   *t = (RefTable *) malloc(sizeof(RefTable));

	t->n = n;
	t->dim = dim;
	t->rows = (Vector **) malloc(sizeof(Vector *) * (size_t) n);
	for (long i = 0; i < n; i++)
		t->rows[i] = make_vector(X + (size_t) i * dim, dim);
	return t;
}

void ndb_ref_table_free(RefTable *t)
{
	for (long i = 0; i < t->n; i++)
		free(t->rows[i]);
	free(t->rows);
	free(t);
}

int ndb_ref_seqscan(const RefTable *t, int metric, const float *q, float *out)
{
	Vector	   *vq = make_vector(q, t->dim);
	int			rc = 0;

	for (long i = 0; i < t->n && rc == 0; i++)
		rc = call2(metric, t->rows[i], vq, out + i);
	free(vq);
	return rc;
}
