/*
 * gpu_backend_b200.c -- the reference-side glue for surface b2: an `ndb_gpu_backend`
 * instance (NeuronDB/include/neurondb_gpu_backend.h:28-354) whose members are libndb_b200.so's entry
 * points.  A maintainer drops this file into NeuronDB/src/gpu/b200/ and calls
 * neurondb_gpu_register_b200_backend() next to the CUDA backend's registration in _PG_init
 * (NeuronDB/src/worker/worker_init.c:176-186); nothing else on the reference side changes.
 *
 * It is compiled here against the reference's REAL header (oracle/Makefile, target `glue`, with the
 * PostgreSQL stand-ins of oracle/pgshim) so that every member's signature is checked by the compiler,
 * and tests/test_gpu_boundary.py calls the kernels THROUGH the struct, the way
 * src/gpu/common/gpu_distance.c:50,78 and gpu_clustering.c:52,68 do.
 */
#include "postgres.h"
#include "neurondb_gpu_backend.h"
#include "ndb_b200.h"

int			neurondb_gpu_device = 0;	/* GUC neurondb.gpu_device (src/util/neurondb_guc.c) in the extension */

static int
b200_init(void)
{
	return ndb_b200_init(neurondb_gpu_device);
}

static void
b200_shutdown(void)
{
	ndb_b200_shutdown();
}

static int
b200_is_available(void)
{
	return ndb_b200_is_available();
}

static int
b200_device_count(void)
{
	return ndb_b200_device_count();
}

static int
b200_set_device(int device_id)
{
	return ndb_b200_init(device_id);
}

static int
b200_device_info(int device_id, NDBGpuDeviceInfo *info)
{
	size_t		total = 0,
				freeb = 0;
	int			maj = 0,
				min = 0,
				sms = 0;
	int			rc;

	if (info == NULL)
		return -1;
	memset(info, 0, sizeof(*info));
	rc = ndb_b200_device_info(device_id, info->name, sizeof(info->name), &total, &freeb, &maj, &min, &sms);
	info->device_id = device_id;
	info->total_memory_bytes = total;
	info->free_memory_bytes = freeb;
	info->compute_major = maj;
	info->compute_minor = min;
	info->is_available = rc == 0;
	return rc;
}

/* ndb_stream_t is an opaque pointer; the library takes the same pointer as void * */
static int
b200_stream_create(ndb_stream_t *stream)
{
	return ndb_b200_stream_create((void **) stream);
}

static int
b200_stream_destroy(ndb_stream_t stream)
{
	return ndb_b200_stream_destroy((void *) stream);
}

static int
b200_stream_synchronize(ndb_stream_t stream)
{
	return ndb_b200_stream_synchronize((void *) stream);
}

static int
b200_launch_l2_distance(const float *A, const float *B, float *out, int n, int d, ndb_stream_t stream)
{
	return ndb_b200_launch_l2_distance(A, B, out, n, d, (void *) stream);
}

static int
b200_launch_cosine(const float *A, const float *B, float *out, int n, int d, ndb_stream_t stream)
{
	return ndb_b200_launch_cosine(A, B, out, n, d, (void *) stream);
}

static int
b200_launch_kmeans_assign(const float *X, const float *C, int *idx, int n, int d, int k, ndb_stream_t stream)
{
	return ndb_b200_launch_kmeans_assign(X, C, idx, n, d, k, (void *) stream);
}

static int
b200_launch_kmeans_update(const float *X, const int *idx, float *C, int n, int d, int k, ndb_stream_t stream)
{
	return ndb_b200_launch_kmeans_update(X, idx, C, n, d, k, (void *) stream);
}

static int
b200_launch_pq_encode(const float *X, const float *codebooks, uint8_t *codes, int n, int d, int m, int ks, ndb_stream_t stream)
{
	return ndb_b200_launch_pq_encode(X, codebooks, codes, n, d, m, ks, (void *) stream);
}

/* Members not named stay NULL: scalar quantisation and the ML trainers are outside this path, and the
 * registry's callers treat a NULL launcher as "not supported by this backend". */
static const ndb_gpu_backend ndb_b200_backend = {
	.name = "b200",
	.provider = "NVIDIA",
	.kind = NDB_GPU_BACKEND_CUDA,
	.features = 0,
	.priority = 100,			/* above ndb_cuda_backend (.priority = 90, gpu_backend_cuda.c:734-740) */
	.init = b200_init,
	.shutdown = b200_shutdown,
	.is_available = b200_is_available,
	.device_count = b200_device_count,
	.device_info = b200_device_info,
	.set_device = b200_set_device,
	.mem_alloc = ndb_b200_mem_alloc,
	.mem_free = ndb_b200_mem_free,
	.memcpy_h2d = ndb_b200_memcpy_h2d,
	.memcpy_d2h = ndb_b200_memcpy_d2h,
	.launch_l2_distance = b200_launch_l2_distance,
	.launch_cosine = b200_launch_cosine,
	.launch_kmeans_assign = b200_launch_kmeans_assign,
	.launch_kmeans_update = b200_launch_kmeans_update,
	.launch_pq_encode = b200_launch_pq_encode,
	.stream_create = b200_stream_create,
	.stream_destroy = b200_stream_destroy,
	.stream_synchronize = b200_stream_synchronize,
};

const ndb_gpu_backend *
neurondb_gpu_b200_backend(void)
{
	return &ndb_b200_backend;
}

#ifndef NDB_B200_GLUE_STANDALONE
extern int	ndb_gpu_register_backend(const ndb_gpu_backend *backend);	/* gpu_backend_registry.c:91-131 */

void
neurondb_gpu_register_b200_backend(void)
{
	(void) ndb_gpu_register_backend(&ndb_b200_backend);
}
#endif

#ifdef NDB_B200_GLUE_STANDALONE
/*
 * Test hooks (tests/test_gpu_boundary.py): a stand-in for the registry's selection loop
 * (gpu_backend_registry.c:91-131, highest priority wins) and calls that go THROUGH the struct's members, the way
 * neurondb_gpu_l2_distance (gpu_distance.c:50) and neurondb_gpu_kmeans (gpu_clustering.c:52,68) reach a backend.
 */
static const ndb_gpu_backend *glue_registry[NDB_GPU_MAX_BACKENDS];
static int	glue_count = 0;

int
ndb_gpu_register_backend(const ndb_gpu_backend *backend)
{
	if (backend == NULL || backend->name == NULL || glue_count >= NDB_GPU_MAX_BACKENDS)
		return -1;
	glue_registry[glue_count++] = backend;
	return 0;
}

static const ndb_gpu_backend *
glue_active(void)
{
	const ndb_gpu_backend *best = NULL;
	int			i;

	if (glue_count == 0)
		ndb_gpu_register_backend(&ndb_b200_backend);
	for (i = 0; i < glue_count; i++)
		if (best == NULL || glue_registry[i]->priority > best->priority)
			best = glue_registry[i];
	return best;
}

const char *ndb_b200_glue_name(void) { return glue_active()->name; }
int ndb_b200_glue_priority(void) { return glue_active()->priority; }
int ndb_b200_glue_init(void) { return glue_active()->init(); }
int ndb_b200_glue_device_info(int id, char *name, size_t len, size_t *total, int *major)
{
	NDBGpuDeviceInfo info;
	int			rc = glue_active()->device_info(id, &info);

	snprintf(name, len, "%s", info.name);
	*total = info.total_memory_bytes;
	*major = info.compute_major;
	return rc;
}
int ndb_b200_glue_l2(const float *A, const float *B, float *out, int n, int d)
{
	return glue_active()->launch_l2_distance(A, B, out, n, d, NULL);
}
int ndb_b200_glue_cosine(const float *A, const float *B, float *out, int n, int d)
{
	return glue_active()->launch_cosine(A, B, out, n, d, NULL);
}
int ndb_b200_glue_kmeans_assign(const float *X, const float *C, int *idx, int n, int d, int k)
{
	return glue_active()->launch_kmeans_assign(X, C, idx, n, d, k, NULL);
}
int ndb_b200_glue_kmeans_update(const float *X, const int *idx, float *C, int n, int d, int k)
{
	return glue_active()->launch_kmeans_update(X, idx, C, n, d, k, NULL);
}
int ndb_b200_glue_pq_encode(const float *X, const float *codebooks, uint8_t *codes, int n, int d, int m, int ks)
{
	return glue_active()->launch_pq_encode(X, codebooks, codes, n, d, m, ks, NULL);
}
int ndb_b200_glue_stream_roundtrip(void)
{
	ndb_stream_t s = NULL;
	int			rc = glue_active()->stream_create(&s);

	if (rc == 0)
		rc = glue_active()->stream_synchronize(s);
	if (rc == 0)
		rc = glue_active()->stream_destroy(s);
	return rc;
}
int ndb_b200_glue_unsupported_members_are_null(void)
{
	const ndb_gpu_backend *b = glue_active();

	return b->launch_quant_fp16 == NULL && b->launch_quant_binary == NULL && b->rf_train == NULL;
}
#endif
