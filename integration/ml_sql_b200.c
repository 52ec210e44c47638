/*
 * ml_sql_b200.c -- the reference-side glue for the SQL functions that sit on the same kernels as the index path
 * (SURVEY.md 8f-3, 8f-4): what cluster_kmeans, cluster_minibatch_kmeans, train_pq_codebook and pq_encode_vector run
 * instead of their own loops once they hold the fetched rows.
 *
 * Each function takes the reference's own locals at the point where its arithmetic starts -- `float **data` as
 * neurondb_fetch_vectors_from_table returns it (ml_kmeans.c:178, ml_minibatch_kmeans.c:263, ml_product_quantization.c:239),
 * the codebook as the bytea payload carries it -- and returns an NDB_B200_* code; the SQL function turns a failure into
 * ereport(ERROR, (errmsg("%s", ndb_b200_last_error()))) (the library's messages are the reference's where it has one).
 *
 * rand(): these functions seed from the backend's rand() stream.  The glue draws from rand() itself, exactly as often and
 * in the order the reference does, so a session that calls other rand() users afterwards sees the same stream as before:
 *   cluster_kmeans            k draws (kmeanspp_init :59, :85), all up front -- the count is fixed
 *   train_pq_codebook         ksub draws per subspace (train_subspace_kmeans :98-103), subspace-major
 *   cluster_minibatch_kmeans  a data-dependent number: the library calls back for each one
 *
 * Compiled against oracle/pgshim by oracle/Makefile, target `glue` (-Werror=implicit-function-declaration); the flatten /
 * draw helpers are tested on the CPU (tests/test_boundary_cpu.py), the whole functions on the GPU against the direct ABI
 * calls with the same draws (tests/test_gpu_boundary.py).
 */
#include "postgres.h"
#include "ndb_b200.h"

#include <stdlib.h>

/* data[i] (one palloc'd row each) -> one contiguous float[nvec * dim]; NULL when out of memory */
float *
ndb_b200_sql_flatten_rows(float **data, int nvec, int dim)
{
	float	   *X;
	int			i;

	if (data == NULL || nvec <= 0 || dim <= 0)
		return NULL;
	X = (float *) malloc(sizeof(float) * (size_t) nvec * (size_t) dim);
	if (X == NULL)
		return NULL;
	for (i = 0; i < nvec; i++)
		memcpy(X + (size_t) i * dim, data[i], sizeof(float) * (size_t) dim);
	return X;
}

/* the next `count` values of the backend's rand() */
int *
ndb_b200_sql_draw(int count)
{
	int		   *draws;
	int			i;

	if (count <= 0)
		return NULL;
	draws = (int *) malloc(sizeof(int) * (size_t) count);
	if (draws == NULL)
		return NULL;
	for (i = 0; i < count; i++)
		draws[i] = rand();
	return draws;
}

static int
sql_rand_cb(void *unused)
{
	(void) unused;
	return rand();
}

/* cluster_kmeans (ml_kmeans.c:146-303) from :188 on; labels[nvec] come back 1-based as :286 builds them */
int
ndb_b200_sql_cluster_kmeans(float **data, int nvec, int dim, int num_clusters, int max_iters, int *labels)
{
	float	   *X;
	int		   *draws;
	int			rc;

	/* (the SQL function has made its argument checks, :170-186, before it gets here -- and before any draw; the library
	 * repeats them and reports the same messages) */
	if (num_clusters <= 1 || nvec < num_clusters)
		return NDB_B200_EINVAL;
	X = ndb_b200_sql_flatten_rows(data, nvec, dim);
	draws = ndb_b200_sql_draw(num_clusters);
	if (X == NULL || draws == NULL)
	{
		free(X);
		free(draws);
		return NDB_B200_ENOMEM;
	}
	rc = ndb_b200_cluster_kmeans(X, nvec, dim, num_clusters, max_iters, draws, RAND_MAX, labels, NULL, NULL, NULL);
	free(X);
	free(draws);
	return rc;
}

/* cluster_minibatch_kmeans (ml_minibatch_kmeans.c:206-449) from :306 on */
int
ndb_b200_sql_cluster_minibatch_kmeans(float **data, int nvec, int dim, int num_clusters, int batch_size, int max_iters,
									  int *labels)
{
	float	   *X = ndb_b200_sql_flatten_rows(data, nvec, dim);
	int			rc;

	if (X == NULL)
		return NDB_B200_ENOMEM;
	rc = ndb_b200_cluster_minibatch_kmeans(X, nvec, dim, num_clusters, batch_size, max_iters, sql_rand_cb, NULL, RAND_MAX,
										   labels, NULL);
	free(X);
	return rc;
}

/* train_pq_codebook (ml_product_quantization.c:195-415) from :303 on: centroids[m][ksub][dsub] is the float block the
 * result bytea carries after its three ints (:362-378) */
int
ndb_b200_sql_train_pq_codebook(float **data, int nvec, int dim, int m, int ksub, float *centroids)
{
	float	   *X;
	int		   *draws;
	int			rc;

	if (m < 1 || m > 128 || ksub < 2 || ksub > 65536 || dim <= 0 || dim % m != 0)	/* :218-227, :270-276: checked before any draw */
		return NDB_B200_EINVAL;
	X = ndb_b200_sql_flatten_rows(data, nvec, dim);
	draws = ndb_b200_sql_draw(m * ksub);
	if (X == NULL || draws == NULL)
	{
		free(X);
		free(draws);
		return NDB_B200_ENOMEM;
	}
	rc = ndb_b200_pq_train(X, nvec, dim, m, ksub, 100, draws, centroids);		/* 100 iterations: :340-346 */
	free(X);
	free(draws);
	return rc;
}

/* pq_encode_vector (:421-536): cb_payload = VARDATA(codebook_bytea): int m, int ksub, int dsub, float centroids[m][ksub][dsub] */
int
ndb_b200_sql_pq_encode_vector(const float4 *vec_data, int dim, const char *cb_payload, int16 *codes)
{
	int			m,
				ksub,
				dsub;

	memcpy(&m, cb_payload, sizeof(int));
	memcpy(&ksub, cb_payload + sizeof(int), sizeof(int));
	memcpy(&dsub, cb_payload + 2 * sizeof(int), sizeof(int));
	if (dim != m * dsub)						/* "Vector dimension (%d) does not match codebook definition" (:456-463) */
		return NDB_B200_EDIM;
	return ndb_b200_pq_encode(vec_data, 1, dim, (const float *) (cb_payload + 3 * sizeof(int)), m, ksub, codes);
}

#ifdef NDB_B200_GLUE_STANDALONE
/* ---- test harness entry points: rows arrive flat, the float ** the SQL functions hold is built here ------------------ */
static float **
rows_of(const float *X, int nvec, int dim)
{
	float	  **data = (float **) malloc(sizeof(float *) * (size_t) (nvec > 0 ? nvec : 1));
	int			i;

	for (i = 0; i < nvec; i++)
		data[i] = (float *) (X + (size_t) i * dim);
	return data;
}
int ndb_b200_glue_flatten_check(const float *X, int nvec, int dim)
{
	float	  **data = rows_of(X, nvec, dim);
	float	   *flat = ndb_b200_sql_flatten_rows(data, nvec, dim);
	int			same = flat != NULL && memcmp(flat, X, sizeof(float) * (size_t) nvec * dim) == 0;

	free(flat);
	free(data);
	return same;
}
void ndb_b200_glue_draw(int count, int *out)
{
	int		   *d = ndb_b200_sql_draw(count);

	memcpy(out, d, sizeof(int) * (size_t) count);
	free(d);
}
int ndb_b200_glue_cluster_kmeans(const float *X, int nvec, int dim, int k, int max_iters, int *labels)
{
	float	  **data = rows_of(X, nvec, dim);
	int			rc = ndb_b200_sql_cluster_kmeans(data, nvec, dim, k, max_iters, labels);

	free(data);
	return rc;
}
int ndb_b200_glue_cluster_minibatch_kmeans(const float *X, int nvec, int dim, int k, int batch, int max_iters, int *labels)
{
	float	  **data = rows_of(X, nvec, dim);
	int			rc = ndb_b200_sql_cluster_minibatch_kmeans(data, nvec, dim, k, batch, max_iters, labels);

	free(data);
	return rc;
}
int ndb_b200_glue_train_pq_codebook(const float *X, int nvec, int dim, int m, int ksub, float *centroids)
{
	float	  **data = rows_of(X, nvec, dim);
	int			rc = ndb_b200_sql_train_pq_codebook(data, nvec, dim, m, ksub, centroids);

	free(data);
	return rc;
}
int ndb_b200_glue_pq_encode_vector(const float *vec, int dim, const char *payload, int16 *codes)
{
	return ndb_b200_sql_pq_encode_vector(vec, dim, payload, codes);
}
#endif
