// comm.cu -- multi-GPU exchange behind the C ABI: one process per GPU, one NCCL communicator per
// process.
//
// What it replaces.  The reference's only scale-out code is NeuronDB/src/util/distributed.c: the
// coordinator opens a libpq connection per shard, runs the same kNN query on each (:56-300), collects the
// per-shard (id, distance) rows and merges them on the host by (distance ASC, id ASC) (:323-487, order
// :425-438).  Here the shards are the GPUs of one box: every rank answers the query batch against the rows
// it holds and writes its top-k as packed 12-byte (dist, id) records (4 bytes of distance block + 8 bytes of id block per
// result) into its slot of an exchange window; one kernel stores that slot into every peer's window over NVLink
// (CUDA IPC peer memory; no collective call in the step) and raises a flag, and every rank's merge kernel waits for the
// world's flags and merges the lists on the device in the same (dist, id) order.  Where IPC windows cannot be had (or
// with NDB_B200_EXCHANGE=nccl) the records travel in ONE ncclAllGather instead.  k-means training exchanges per-cluster sums and counts
// with ncclAllReduce once per Lloyd iteration; an HNSW graph built on one rank reaches the replicas with
// ncclBroadcast.
//
// NCCL is bound at run time (dlopen "libnccl.so.2", the SONAME both the system package and the PyTorch
// wheel install): a PostgreSQL backend that never calls ndb_b200_comm_init does not need the library, and a
// process that already loaded PyTorch's copy shares that copy instead of loading a second one.
#include "comm.cuh"
#include "layout.cuh"
#include "scan.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <vector>

namespace ndb {

void ivf_set_coarse_by_rank(ndb_b200_ivf *ix, bool on);          // ivf.cu

struct NcclApi {
    void *dl = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

struct Comm {
    NcclApi api;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    bool ready = false;
    int device = -1;
    DevBuf gather;                 // [world][record block] of the sharded searches (NCCL exchange)
    // Peer-memory exchange (NVLink loads / stores, no NCCL call in the step): one window per rank, opened by every
    // other rank through CUDA IPC.  Window = [flags: world x u32, padded to 256 B][parity 0: world slots][parity 1: ...]
    bool p2p_ok = false, p2p_tried = false;
    void *p2p_win = nullptr;                       // this rank's window (cudaMalloc)
    void *p2p_peer[16] = {nullptr};                // every rank's window as seen from here (own entry = p2p_win)
    void **p2p_peer_dev = nullptr;                 // the same table on the device
    unsigned *p2p_done = nullptr;                  // CTA counter of the push kernel
    size_t p2p_slot = 0;                           // bytes per (rank, parity) slot
    uint32_t p2p_step = 0;
    DevBuf qbuf, outd, outi;       // staging of the host-pointer sharded entry points
    DevBuf red, cst;               // k-means: [k*d f32 sums | k i32 counts], [1 f32 cost]
    void release_scratch() { gather.release(); qbuf.release(); outd.release(); outi.release(); red.release(); cst.release(); }
    void release_p2p()
    {
        for (int r = 0; r < 16; r++) {
            if (p2p_peer[r] && p2p_peer[r] != p2p_win) cudaIpcCloseMemHandle(p2p_peer[r]);
            p2p_peer[r] = nullptr;
        }
        if (p2p_win) cudaFree(p2p_win);
        if (p2p_peer_dev) cudaFree(p2p_peer_dev);
        if (p2p_done) cudaFree(p2p_done);
        p2p_win = nullptr; p2p_peer_dev = nullptr; p2p_done = nullptr;
        p2p_ok = false; p2p_slot = 0; p2p_step = 0;
    }
};

static Comm g_comm;

static int nccl_bind(NcclApi &a)
{
    if (a.dl) return NDB_B200_OK;
    const char *names[] = {getenv("NDB_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        a.dl = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.dl) break;
    }
    NDB_REQUIRE(a.dl, NDB_B200_ENOTINIT, "comm: cannot load libnccl.so.2 (%s); set NDB_B200_NCCL_LIB", dlerror());
#define NDB_SYM(field, name)                                                                        \
    do {                                                                                            \
        *reinterpret_cast<void **>(&a.field) = dlsym(a.dl, name);                                   \
        if (!a.field) {                                                                             \
            set_error("comm: %s not found in libnccl", name);                                       \
            dlclose(a.dl);                                                                          \
            a = NcclApi();                                                                          \
            return NDB_B200_ENOTINIT;                                                               \
        }                                                                                           \
    } while (0)
    NDB_SYM(GetUniqueId, "ncclGetUniqueId");
    NDB_SYM(CommInitRank, "ncclCommInitRank");
    NDB_SYM(CommDestroy, "ncclCommDestroy");
    NDB_SYM(AllGather, "ncclAllGather");
    NDB_SYM(AllReduce, "ncclAllReduce");
    NDB_SYM(Broadcast, "ncclBroadcast");
    NDB_SYM(GetErrorString, "ncclGetErrorString");
    NDB_SYM(GetVersion, "ncclGetVersion");
#undef NDB_SYM
    return NDB_B200_OK;
}

#define NDB_NCCL(call)                                                                              \
    do {                                                                                            \
        ncclResult_t r__ = (call);                                                                  \
        if (r__ != ncclSuccess) {                                                                   \
            set_error("%s failed: %s", #call, g_comm.api.GetErrorString(r__));                      \
            return NDB_B200_ECUDA;                                                                  \
        }                                                                                           \
    } while (0)

// called by ndb_b200_shutdown: device memory must not outlive the context it was allocated in
void comm_at_shutdown()
{
    if (g_comm.ready && g_comm.comm) g_comm.api.CommDestroy(g_comm.comm);
    g_comm.comm = nullptr;
    g_comm.release_scratch();
    g_comm.release_p2p();
    g_comm.p2p_tried = false;
    g_comm.ready = false;
    g_comm.rank = 0;
    g_comm.world = 1;
}

bool comm_ready() { return g_comm.ready && g_comm.device == ctx().device; }
int comm_rank() { return comm_ready() ? g_comm.rank : 0; }
int comm_nranks() { return comm_ready() ? g_comm.world : 1; }

int comm_allgather(const void *send_dev, void *recv_dev, size_t bytes, cudaStream_t s)
{
    if (!comm_ready() || g_comm.world == 1) {
        if (send_dev != recv_dev) NDB_CUDA(cudaMemcpyAsync(recv_dev, send_dev, bytes, cudaMemcpyDeviceToDevice, s));
        return NDB_B200_OK;
    }
    NDB_NCCL(g_comm.api.AllGather(send_dev, recv_dev, bytes, ncclInt8, g_comm.comm, s));
    return NDB_B200_OK;
}

int comm_allreduce_sum(void *buf_dev, size_t count, CommType t, cudaStream_t s)
{
    if (!comm_ready() || g_comm.world == 1) return NDB_B200_OK;
    const ncclDataType_t dt = t == COMM_F32 ? ncclFloat32 : t == COMM_I32 ? ncclInt32 : t == COMM_F64 ? ncclFloat64 : ncclInt64;
    NDB_NCCL(g_comm.api.AllReduce(buf_dev, buf_dev, count, dt, ncclSum, g_comm.comm, s));
    return NDB_B200_OK;
}

int comm_broadcast(void *buf_dev, size_t bytes, int root, cudaStream_t s)
{
    if (!comm_ready() || g_comm.world == 1) return NDB_B200_OK;
    NDB_NCCL(g_comm.api.Broadcast(buf_dev, buf_dev, bytes, ncclInt8, root, g_comm.comm, s));
    return NDB_B200_OK;
}

// ---- merge of the gathered record blocks -------------------------------------------------------
// rec = [world] blocks of `stride` bytes; block r = rank r's [nq][k] f32 distances, then (at ids_off)
// its [nq][k] int64 ids, both sorted by (dist, id) per query, missing = (+inf, -1).
template <int KR>
__global__ void merge_records_kernel(const unsigned char *__restrict__ rec, size_t stride, size_t ids_off, int nshards,
                                     int nq, int k, float *__restrict__ out_dist, int64_t *__restrict__ out_ids)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    WarpTopK<KR, int64_t> top;
    top.init();
    for (int s = 0; s < nshards; s++) {
        const float *d = reinterpret_cast<const float *>(rec + (size_t) s * stride) + (size_t) q * k;
        const int64_t *id = reinterpret_cast<const int64_t *>(rec + (size_t) s * stride + ids_off) + (size_t) q * k;
        for (int i = lane; i < round_up(k, 32); i += 32) {
            float cd = INFINITY;
            int64_t ci = -1;
            if (i < k) { cd = d[i]; ci = id[i]; }
            top.offer(cd, ci, i < k && ci >= 0, lane, k);
        }
    }
#pragma unroll
    for (int r = 0; r < KR; r++) {
        const int e = r * 32 + lane;
        if (e < k) {
            const bool have = top.key[r] != KeyMax<int64_t>::v;
            out_dist[(size_t) q * k + e] = have ? top.d[r] : INFINITY;
            out_ids[(size_t) q * k + e] = have ? top.key[r] : -1;
        }
    }
}

// ---- peer-memory exchange ---------------------------------------------------------------------------------------
// push: this rank's record block (already in its own window, slot `rank` of the step's parity) is stored into the same
// slot of every peer's window over NVLink (16-byte stores), then -- once every CTA's stores are fenced -- the step number
// is written to flag `rank` of every peer.  merge: waits until all `world` flags of the LOCAL window show the step, then
// merges the world's lists exactly like merge_records_kernel.  Two parities: a rank can run at most one step ahead of a
// peer that is still merging (its next push needs that peer's push of the current step).
constexpr size_t P2P_FLAG_BYTES = 256;

__global__ void __launch_bounds__(256) p2p_push_kernel(void *const *__restrict__ peer, size_t slot_off, size_t bytes, int rank, int world,
                                                        uint32_t step, unsigned *__restrict__ done)
{
    const uint4 *src = reinterpret_cast<const uint4 *>(static_cast<const unsigned char *>(peer[rank]) + slot_off);
    const size_t n16 = bytes / 16;
    for (int r = 0; r < world; r++) {
        if (r == rank) continue;
        uint4 *dst = reinterpret_cast<uint4 *>(static_cast<unsigned char *>(peer[r]) + slot_off);
        for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t) gridDim.x * blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *done = 0;
        if ((int) threadIdx.x < world) {
            volatile uint32_t *flag = reinterpret_cast<volatile uint32_t *>(peer[threadIdx.x]) + rank;
            *flag = step;
        }
        __threadfence_system();
    }
}

template <int KR>
__global__ void p2p_merge_kernel(const unsigned char *__restrict__ win, size_t par_off, size_t stride, size_t ids_off, int nshards,
                                 uint32_t step, int nq, int k, float *__restrict__ out_dist, int64_t *__restrict__ out_ids)
{
    // every rank's records of this step have arrived once its flag shows the step (flags only grow)
    if (threadIdx.x < (unsigned) nshards) {
        const volatile uint32_t *flag = reinterpret_cast<const volatile uint32_t *>(win) + threadIdx.x;
        const long long t0 = clock64();
        while ((int32_t) (*flag - step) < 0) {
            if (clock64() - t0 > 20000000000ll) __trap();        // ~10 s: a peer died; fail loudly rather than hang
        }
    }
    __threadfence_system();
    __syncthreads();
    const unsigned char *rec = win + par_off;
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    WarpTopK<KR, int64_t> top;
    top.init();
    for (int s = 0; s < nshards; s++) {
        const float *d = reinterpret_cast<const float *>(rec + (size_t) s * stride) + (size_t) q * k;
        const int64_t *id = reinterpret_cast<const int64_t *>(rec + (size_t) s * stride + ids_off) + (size_t) q * k;
        for (int i = lane; i < round_up(k, 32); i += 32) {
            float cd = INFINITY;
            int64_t ci = -1;
            if (i < k) { cd = __ldcv(d + i); ci = __ldcv(id + i); }      // (written by a peer: not through the read-only cache)
            top.offer(cd, ci, i < k && ci >= 0, lane, k);
        }
    }
#pragma unroll
    for (int r = 0; r < KR; r++) {
        const int e = r * 32 + lane;
        if (e < k) {
            const bool have = top.key[r] != KeyMax<int64_t>::v;
            out_dist[(size_t) q * k + e] = have ? top.d[r] : INFINITY;
            out_ids[(size_t) q * k + e] = have ? top.key[r] : -1;
        }
    }
}

// (re)create the windows for slots of at least `slot` bytes: collective over the communicator
static int p2p_setup(size_t slot, cudaStream_t s)
{
    Comm &c = g_comm;
    if (c.p2p_ok && c.p2p_slot >= slot) return NDB_B200_OK;
    static const bool disabled = getenv("NDB_B200_EXCHANGE") && !strcmp(getenv("NDB_B200_EXCHANGE"), "nccl");
    if (disabled || c.world > 16 || (c.p2p_tried && !c.p2p_ok)) return NDB_B200_ESTATE;
    c.p2p_tried = true;
    NDB_CUDA(cudaStreamSynchronize(s));
    // nobody may still be reading an old window: one all-reduce as a barrier
    DevBuf tmp;
    NDB_CHECK(tmp.reserve(256 + (size_t) c.world * sizeof(cudaIpcMemHandle_t)));
    NDB_CUDA(cudaMemsetAsync(tmp.p, 0, 256, s));
    NDB_CHECK(comm_allreduce_sum(tmp.p, 1, COMM_I32, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    c.release_p2p();
    size_t want = ((slot + slot / 2 + 65535) / 65536) * 65536;
    const size_t win_bytes = P2P_FLAG_BYTES + 2 * (size_t) c.world * want;
    NDB_CUDA(cudaMalloc(&c.p2p_win, win_bytes));
    NDB_CUDA(cudaMemset(c.p2p_win, 0, P2P_FLAG_BYTES));
    NDB_CUDA(cudaMalloc((void **) &c.p2p_peer_dev, 16 * sizeof(void *)));
    NDB_CUDA(cudaMalloc((void **) &c.p2p_done, 4));
    NDB_CUDA(cudaMemset(c.p2p_done, 0, 4));
    cudaIpcMemHandle_t mine;
    cudaError_t e = cudaIpcGetMemHandle(&mine, c.p2p_win);
    // exchange the handles (and whether everybody got one) through the communicator
    std::vector<unsigned char> hbuf((size_t) c.world * (sizeof(cudaIpcMemHandle_t) + 8), 0);
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
    unsigned char *d_all = tmp.as<unsigned char>();
    std::vector<unsigned char> my(rec, 0);
    memcpy(my.data(), &mine, sizeof(mine));
    my[sizeof(mine)] = e == cudaSuccess ? 1 : 0;
    NDB_CHECK(tmp.reserve((size_t) c.world * rec));
    d_all = tmp.as<unsigned char>();
    NDB_CUDA(cudaMemcpyAsync(d_all + (size_t) c.rank * rec, my.data(), rec, cudaMemcpyHostToDevice, s));
    NDB_CHECK(comm_allgather(d_all + (size_t) c.rank * rec, d_all, rec, s));
    NDB_CUDA(cudaMemcpyAsync(hbuf.data(), d_all, hbuf.size(), cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    bool all_ok = true;
    for (int r = 0; r < c.world; r++) all_ok = all_ok && hbuf[(size_t) r * rec + sizeof(mine)] == 1;
    if (all_ok) {
        for (int r = 0; r < c.world && all_ok; r++) {
            if (r == c.rank) { c.p2p_peer[r] = c.p2p_win; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, hbuf.data() + (size_t) r * rec, sizeof(h));
            if (cudaIpcOpenMemHandle(&c.p2p_peer[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { all_ok = false; c.p2p_peer[r] = nullptr; }
        }
    }
    (void) cudaGetLastError();
    // everybody must agree (a rank that failed to open a handle drags all to the NCCL exchange)
    int ok_i = all_ok ? 1 : 0;
    NDB_CUDA(cudaMemcpyAsync(tmp.p, &ok_i, 4, cudaMemcpyHostToDevice, s));
    NDB_CHECK(comm_allreduce_sum(tmp.p, 1, COMM_I32, s));
    NDB_CUDA(cudaMemcpyAsync(&ok_i, tmp.p, 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    if (ok_i != c.world) { c.release_p2p(); return NDB_B200_ESTATE; }
    NDB_CUDA(cudaMemcpy(c.p2p_peer_dev, c.p2p_peer, 16 * sizeof(void *), cudaMemcpyHostToDevice));
    c.p2p_slot = want;
    c.p2p_step = 0;
    c.p2p_ok = true;
    return NDB_B200_OK;
}

struct RecordLayout {
    size_t ids_off, stride;
    RecordLayout(int nq, int k)
    {
        ids_off = ((size_t) nq * k * 4 + 15) / 16 * 16;
        stride = ids_off + (size_t) nq * k * 8;
        stride = (stride + 15) / 16 * 16;
    }
};

// exchange + merge of a local [nq][k] result that `fill` writes into this rank's block
template <class Fill>
static int sharded_topk(int nq, int k, float *dist_dev, int64_t *ids_dev, cudaStream_t s, Fill fill)
{
    NDB_REQUIRE(k >= 1 && k <= 128, NDB_B200_EINVAL, "sharded search: k=%d out of range 1..128", k);
    const int world = comm_nranks(), rank = comm_rank();
    if (world == 1) return fill(dist_dev, ids_dev);
    const RecordLayout rl(nq, k);
    const unsigned grid = (unsigned) ((nq + 3) / 4);
    if (p2p_setup(rl.stride, s) == NDB_B200_OK) {
        // exchange through peer memory: local result straight into this rank's window, one push kernel storing it into the
        // peers' windows over NVLink, the merge kernel waits for the world's flags -- no collective call in the step
        Comm &c = g_comm;
        const uint32_t step = ++c.p2p_step;
        const size_t par_off = P2P_FLAG_BYTES + (size_t) (step & 1u) * world * c.p2p_slot;
        const size_t slot_off = par_off + (size_t) rank * c.p2p_slot;
        unsigned char *mine = static_cast<unsigned char *>(c.p2p_win) + slot_off;
        NDB_CHECK(fill(reinterpret_cast<float *>(mine), reinterpret_cast<int64_t *>(mine + rl.ids_off)));
        const unsigned pgrid = (unsigned) std::min<size_t>(64, (rl.stride / 16 + 255) / 256);
        p2p_push_kernel<<<pgrid ? pgrid : 1, 256, 0, s>>>(c.p2p_peer_dev, slot_off, rl.stride, rank, world, step, c.p2p_done);
        if (k <= 32)
            p2p_merge_kernel<1><<<grid, 128, 0, s>>>(static_cast<const unsigned char *>(c.p2p_win), par_off, c.p2p_slot, rl.ids_off, world, step,
                                                     nq, k, dist_dev, ids_dev);
        else
            p2p_merge_kernel<4><<<grid, 128, 0, s>>>(static_cast<const unsigned char *>(c.p2p_win), par_off, c.p2p_slot, rl.ids_off, world, step,
                                                     nq, k, dist_dev, ids_dev);
        count_launch(2);
        NDB_CUDA(cudaGetLastError());
        return NDB_B200_OK;
    }
    NDB_CHECK(g_comm.gather.reserve(rl.stride * world));
    unsigned char *rec = g_comm.gather.as<unsigned char>();
    unsigned char *mine = rec + rl.stride * rank;
    NDB_CHECK(fill(reinterpret_cast<float *>(mine), reinterpret_cast<int64_t *>(mine + rl.ids_off)));
    NDB_CHECK(comm_allgather(mine, rec, rl.stride, s));            // in place: block `rank` is already where it belongs
    if (k <= 32) merge_records_kernel<1><<<grid, 128, 0, s>>>(rec, rl.stride, rl.ids_off, world, nq, k, dist_dev, ids_dev);
    else merge_records_kernel<4><<<grid, 128, 0, s>>>(rec, rl.stride, rl.ids_off, world, nq, k, dist_dev, ids_dev);
    count_launch();
    NDB_CUDA(cudaGetLastError());
    return NDB_B200_OK;
}

__global__ void kmeans_divide_kernel(const float *__restrict__ sums, const int *__restrict__ counts, int k, int d,
                                     float *__restrict__ C)
{
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t) k * d) return;
    const int c = counts[t / d];
    // empty cluster -> zeros (ivf_am.c:2207-2211); otherwise sum / (float) count (:2203)
    C[t] = c > 0 ? __fdiv_rn(sums[t], (float) c) : 0.0f;
}

}  // namespace ndb

using namespace ndb;

extern "C" {

int ndb_b200_comm_unique_id(void *id, size_t len)
{
    NDB_REQUIRE(id && len >= sizeof(ncclUniqueId), NDB_B200_EINVAL, "comm_unique_id: need a %zu-byte buffer", sizeof(ncclUniqueId));
    NDB_CHECK(nccl_bind(g_comm.api));
    ncclUniqueId u;
    NDB_NCCL(g_comm.api.GetUniqueId(&u));
    memset(id, 0, len);
    memcpy(id, &u, sizeof(u));
    return NDB_B200_OK;
}

int ndb_b200_comm_init(int rank, int world, const void *id, size_t len)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(world >= 1 && rank >= 0 && rank < world, NDB_B200_EINVAL, "comm_init: bad rank %d / world %d", rank, world);
    NDB_REQUIRE(!g_comm.ready, NDB_B200_ESTATE, "comm_init: a communicator already exists (comm_shutdown first)");
    if (world == 1) {
        g_comm.rank = 0; g_comm.world = 1; g_comm.ready = true; g_comm.device = ctx().device; g_comm.comm = nullptr;
        return NDB_B200_OK;
    }
    NDB_REQUIRE(id && len >= sizeof(ncclUniqueId), NDB_B200_EINVAL, "comm_init: need the %zu-byte id of rank 0", sizeof(ncclUniqueId));
    NDB_CHECK(nccl_bind(g_comm.api));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    NDB_NCCL(g_comm.api.CommInitRank(&g_comm.comm, world, u, rank));
    g_comm.rank = rank;
    g_comm.world = world;
    g_comm.device = ctx().device;
    g_comm.ready = true;
    return NDB_B200_OK;
}

int ndb_b200_comm_shutdown(void)
{
    if (!g_comm.ready) return NDB_B200_OK;
    if (ctx().initialized) { cudaSetDevice(ctx().device); cudaDeviceSynchronize(); }
    if (g_comm.comm) g_comm.api.CommDestroy(g_comm.comm);
    g_comm.comm = nullptr;
    g_comm.release_scratch();
    g_comm.release_p2p();
    g_comm.p2p_tried = false;
    g_comm.ready = false;
    g_comm.rank = 0;
    g_comm.world = 1;
    return NDB_B200_OK;
}

int ndb_b200_comm_rank(void) { return comm_rank(); }
int ndb_b200_comm_nranks(void) { return comm_nranks(); }

/* 1 = the sharded searches exchange their records through peer memory (CUDA IPC windows, NVLink stores);
 * 0 = through ncclAllGather (no windows yet, IPC unavailable, or NDB_B200_EXCHANGE=nccl) */
int ndb_b200_comm_exchange_is_p2p(void) { return g_comm.ready && g_comm.p2p_ok ? 1 : 0; }

int ndb_b200_comm_nccl_version(void)
{
    if (nccl_bind(g_comm.api) != NDB_B200_OK) return 0;
    int v = 0;
    return g_comm.api.GetVersion(&v) == ncclSuccess ? v : 0;
}

int ndb_b200_comm_allgather_dev(const void *send_dev, void *recv_dev, size_t bytes_per_rank, void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(send_dev && recv_dev && bytes_per_rank, NDB_B200_EINVAL, "comm_allgather: NULL or empty buffer");
    return comm_allgather(send_dev, recv_dev, bytes_per_rank, stream ? (cudaStream_t) stream : ctx().stream);
}

int ndb_b200_comm_allreduce_sum_dev(void *buf_dev, size_t count, int type, void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(buf_dev && count && type >= 0 && type <= 3, NDB_B200_EINVAL, "comm_allreduce: bad argument");
    return comm_allreduce_sum(buf_dev, count, (CommType) type, stream ? (cudaStream_t) stream : ctx().stream);
}

int ndb_b200_comm_broadcast_dev(void *buf_dev, size_t bytes, int root, void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(buf_dev && bytes && root >= 0 && root < comm_nranks(), NDB_B200_EINVAL, "comm_broadcast: bad argument");
    return comm_broadcast(buf_dev, bytes, root, stream ? (cudaStream_t) stream : ctx().stream);
}

// ---- sharded searches: local top-k -> one all-gather of packed records -> device merge ---------
int ndb_b200_ivf_search_sharded_dev(ndb_b200_ivf *ix, const float *Q_dev, int nq, int nprobe, int k, int mode, int arith,
                                    float *dist_dev, int64_t *ids_dev, void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix && Q_dev && dist_dev && ids_dev && nq > 0, NDB_B200_EINVAL, "ivf_search_sharded: bad argument");
    cudaStream_t s = stream ? (cudaStream_t) stream : ctx().stream;
    return sharded_topk(nq, k, dist_dev, ids_dev, s, [&](float *d, int64_t *i) {
        ivf_set_coarse_by_rank(ix, true);          // the coarse quantiser is split by queries, its probe lists all-gathered
        const int rc = ndb_b200_ivf_search_dev(ix, Q_dev, nq, nprobe, k, mode, arith, d, i, s);
        ivf_set_coarse_by_rank(ix, false);
        return rc;
    });
}

int ndb_b200_knn_exact_sharded_dev(ndb_b200_dataset *ds, int metric, int arith, const float *Q_dev, int nq, int k,
                                   float *dist_dev, int64_t *ids_dev, void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ds && Q_dev && dist_dev && ids_dev && nq > 0, NDB_B200_EINVAL, "knn_exact_sharded: bad argument");
    cudaStream_t s = stream ? (cudaStream_t) stream : ctx().stream;
    return sharded_topk(nq, k, dist_dev, ids_dev, s, [&](float *d, int64_t *i) {
        return ndb_b200_knn_exact_dev(ds, metric, arith, Q_dev, nq, k, d, i, s);
    });
}

// host-pointer forms (the call an index AM makes): queries in, merged results out, on every rank
static int sharded_host(const float *Q, int nq, int dim, int k, float *dist, int64_t *ids,
                        int (*run)(void *, const float *, float *, int64_t *, cudaStream_t), void *arg)
{
    DevBuf &qbuf = g_comm.qbuf, &outd = g_comm.outd, &outi = g_comm.outi;
    cudaStream_t s = ctx().stream;
    const size_t qb = (size_t) nq * dim * 4, m = (size_t) nq * k;
    NDB_CHECK(qbuf.reserve(qb));
    NDB_CHECK(outd.reserve(m * 4));
    NDB_CHECK(outi.reserve(m * 8));
    NDB_CUDA(cudaMemcpyAsync(qbuf.p, Q, qb, cudaMemcpyHostToDevice, s));
    NDB_CHECK(validate_begin(qbuf.as<float>(), (int64_t) nq * dim, s));
    NDB_CHECK(run(arg, qbuf.as<float>(), outd.as<float>(), outi.as<int64_t>(), s));
    NDB_CUDA(cudaMemcpyAsync(dist, outd.p, m * 4, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaMemcpyAsync(ids, outi.p, m * 8, cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    NDB_REQUIRE(validate_end() < 0, NDB_B200_EVECTOR, "sharded search: NaN/Inf in query");
    return NDB_B200_OK;
}

int ndb_b200_ivf_search_sharded(ndb_b200_ivf *ix, const float *Q, int nq, int nprobe, int k, int mode, int arith,
                                float *dist, int64_t *ids)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ix && Q && dist && ids && nq > 0 && k >= 1 && k <= 128, NDB_B200_EINVAL, "ivf_search_sharded: bad argument");
    struct A { ndb_b200_ivf *ix; int nq, nprobe, k, mode, arith; } a = {ix, nq, nprobe, k, mode, arith};
    return sharded_host(Q, nq, ndb_b200_ivf_dim(ix), k, dist, ids, [](void *p, const float *q, float *d, int64_t *i, cudaStream_t s) {
        A *a = static_cast<A *>(p);
        return ndb_b200_ivf_search_sharded_dev(a->ix, q, a->nq, a->nprobe, a->k, a->mode, a->arith, d, i, s);
    }, &a);
}

// ---- row-sharded k-means: the whole Lloyd loop (kmeans_run, ivf_am.c:2117-2159) ------------------
// Every rank holds n_local rows and the replicated k*d centroids C_dev (in: the initial centroids, i.e.
// kmeans_init's first k rows of the GLOBAL order, which the caller broadcasts; out: the trained ones).
// Per iteration: local assignment + per-cluster f32 sums and counts, ONE all-reduce of the
// [k*d sums | k counts] block, centroid = sum / count (empty -> zeros), local cost, all-reduce of the
// cost, |prev - cost| < tol stops.  Sums are added across ranks in the reduction's order, so with more
// than one rank the centroids agree with the single-process result to fp32 rounding, not bit for bit.
int ndb_b200_kmeans_train_sharded_dev(const float *X_dev, int64_t n_local, int d, int k, int max_iter, float tol,
                                      float *C_dev, int *assign_dev, int *counts_dev, int *iters, float *cost_out,
                                      void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(X_dev && C_dev && assign_dev && n_local > 0 && d > 0 && k > 0 && max_iter >= 0, NDB_B200_EINVAL,
                "kmeans_train_sharded: bad argument");
    cudaStream_t s = stream ? (cudaStream_t) stream : ctx().stream;
    DevBuf &red = g_comm.red, &cst = g_comm.cst;
    NDB_CHECK(red.reserve(((size_t) k * d + k) * 4));
    NDB_CHECK(cst.reserve(16));
    float *sums = red.as<float>();
    int *cnts = reinterpret_cast<int *>(sums + (size_t) k * d);
    float prev = 3.402823466e+38f, cost = 0.0f;
    int iter;
    for (iter = 0; iter < max_iter; iter++) {
        NDB_CHECK(ndb_b200_kmeans_shard_step_dev(X_dev, n_local, d, k, C_dev, assign_dev, sums, cnts, s));
        // counts travel as int32 and sums as f32: two reductions over one contiguous block
        NDB_CHECK(comm_allreduce_sum(sums, (size_t) k * d, COMM_F32, s));
        NDB_CHECK(comm_allreduce_sum(cnts, (size_t) k, COMM_I32, s));
        kmeans_divide_kernel<<<(unsigned) (((size_t) k * d + 255) / 256), 256, 0, s>>>(sums, cnts, k, d, C_dev);
        count_launch();
        NDB_CUDA(cudaGetLastError());
        NDB_CHECK(ndb_b200_kmeans_shard_cost_dev(X_dev, n_local, d, C_dev, assign_dev, cst.as<float>(), s));
        NDB_CHECK(comm_allreduce_sum(cst.p, 1, COMM_F32, s));
        NDB_CUDA(cudaMemcpyAsync(&cost, cst.p, 4, cudaMemcpyDeviceToHost, s));
        NDB_CUDA(cudaStreamSynchronize(s));
        if (fabsf(prev - cost) < tol) { iter++; break; }
        prev = cost;
    }
    if (counts_dev) NDB_CUDA(cudaMemcpyAsync(counts_dev, cnts, (size_t) k * 4, cudaMemcpyDeviceToDevice, s));
    if (iters) *iters = iter;
    if (cost_out) *cost_out = cost;
    return NDB_B200_OK;
}

}  // extern "C"
