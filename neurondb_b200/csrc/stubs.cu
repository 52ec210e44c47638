// stubs.cu -- TEMPORARY: entry points still under construction this round.
#include "common.cuh"
using namespace ndb;
extern "C" {
#define NOTYET(name) set_error(name ": not implemented yet"); return NDB_B200_ESTATE
int ndb_b200_ivf_load_relation(ndb_b200_ivf *, const void *, uint32_t) { NOTYET("ivf_load_relation"); }
int ndb_b200_hnsw_create(int, int, int, int, int, ndb_b200_hnsw **) { NOTYET("hnsw_create"); }
void ndb_b200_hnsw_free(ndb_b200_hnsw *) {}
int ndb_b200_hnsw_build(ndb_b200_hnsw *, const float *, const int64_t *, int64_t, const int *, unsigned, int) { NOTYET("hnsw_build"); }
int ndb_b200_hnsw_load_graph(ndb_b200_hnsw *, const float *, const int64_t *, int64_t, const int *, const uint32_t *, const int16_t *, const int64_t *, const uint32_t *, uint32_t, int) { NOTYET("hnsw_load_graph"); }
int ndb_b200_hnsw_export_graph(const ndb_b200_hnsw *, int *, uint32_t *, int16_t *, int64_t *, uint32_t *, int64_t, uint32_t *, int *) { NOTYET("hnsw_export_graph"); }
int64_t ndb_b200_hnsw_size(const ndb_b200_hnsw *) { return 0; }
int ndb_b200_hnsw_load_relation(ndb_b200_hnsw *, const void *, uint32_t) { NOTYET("hnsw_load_relation"); }
int ndb_b200_hnsw_search(ndb_b200_hnsw *, const float *, int, int, int, int, int, float *, int64_t *) { NOTYET("hnsw_search"); }
int ndb_b200_hnsw_search_dev(ndb_b200_hnsw *, const float *, int, int, int, int, int, float *, int64_t *, void *) { NOTYET("hnsw_search_dev"); }
int64_t ndb_b200_hnsw_last_evals(const ndb_b200_hnsw *) { return 0; }
}
