// stubs.cu -- TEMPORARY: entry points still under construction this round.
#include "common.cuh"
using namespace ndb;
extern "C" {
#define NOTYET(name) set_error(name ": not implemented yet"); return NDB_B200_ESTATE
int ndb_b200_ivf_load_relation(ndb_b200_ivf *, const void *, uint32_t) { NOTYET("ivf_load_relation"); }
int ndb_b200_hnsw_load_relation(ndb_b200_hnsw *, const void *, uint32_t) { NOTYET("hnsw_load_relation"); }
}
