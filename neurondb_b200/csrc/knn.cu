// knn.cu -- resident dataset + exact kNN (the SeqScan + top-N sort of SURVEY 3.1 as one
// fused scan, evaluated with the arithmetic the `<->`/`<=>`/`<#>` operators really execute).
#include "layout.cuh"
#include "scan.cuh"
#include "tc.cuh"
#include <vector>
#include <algorithm>

namespace ndb {

struct NormCache {
    DevBuf buf;
    int64_t valid_for = -1;    // number of vectors the cache covers
};

}  // namespace ndb

using namespace ndb;

struct ndb_b200_dataset {
    int dim = 0, dimp = 0;
    int64_t n = 0;
    VecStore store;
    DevBuf ids;                 // int64 per slot
    NormCache norm_f32_ivf, norm_f32_fast, norm_f64;
    ScanScratch scratch;
    DevBuf tmp_rows, tmp_ids, qbuf, outd, outi;
    TcStore tc;                 // bf16 blocked copy for NDB_ARITH_TENSOR, built on first use
    TcScratch tcs;
};

namespace ndb {

__global__ void iota_i64_kernel(int64_t *out, int64_t base, int64_t n)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = base + i;
}

static int dataset_reserve(ndb_b200_dataset *ds, int64_t total, cudaStream_t s)
{
    const int64_t blocks = (total + 31) / 32;
    const size_t old_bytes = ds->store.bytes_for((ds->n + 31) / 32);
    const size_t need = ds->store.bytes_for(blocks);
    if (need > ds->store.data.cap) {
        const size_t before = ds->store.data.cap;
        NDB_CHECK(ds->store.data.grow(need, old_bytes, s));
        // fresh capacity must be zero: pad lanes of the last block are read (and masked) by the scan
        NDB_CUDA(cudaMemsetAsync((char *) ds->store.data.p + old_bytes, 0, ds->store.data.cap - old_bytes, s));
        (void) before;
    }
    NDB_CHECK(ds->ids.grow((size_t) total * sizeof(int64_t), (size_t) ds->n * sizeof(int64_t), s));
    return NDB_B200_OK;
}

static int dataset_append_dev(ndb_b200_dataset *ds, const float *rows_dev, const int64_t *ids_dev, int64_t n,
                              cudaStream_t s)
{
    NDB_REQUIRE(ds->n + n < (int64_t) 0xfffffff0ll, NDB_B200_EINVAL, "dataset too large for 32-bit slots");
    NDB_CHECK(dataset_reserve(ds, ds->n + n, s));
    NDB_CHECK(il32_scatter(rows_dev, n, ds->dim, ds->dimp, nullptr, (uint32_t) ds->n, ds->store.ptr(), s));
    if (ids_dev) {
        NDB_CUDA(cudaMemcpyAsync(ds->ids.as<int64_t>() + ds->n, ids_dev, (size_t) n * sizeof(int64_t),
                                 cudaMemcpyDeviceToDevice, s));
    } else {
        iota_i64_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(ds->ids.as<int64_t>() + ds->n, ds->n, n);
        count_launch();
        NDB_CUDA(cudaGetLastError());
    }
    ds->n += n;
    ds->store.nblk = (ds->n + 31) / 32;
    return NDB_B200_OK;
}

static int dataset_norms(ndb_b200_dataset *ds, int arith, const void **out, cudaStream_t s)
{
    NormCache &c = arith == NDB_ARITH_OP_F64 ? ds->norm_f64 : (arith == NDB_ARITH_FAST ? ds->norm_f32_fast : ds->norm_f32_ivf);
    if (c.valid_for != ds->n) {
        NDB_CHECK(c.buf.reserve((size_t) ds->n * norm_elem_size(arith)));
        NDB_CHECK(slot_norms(arith, ds->store.ptr(), ds->n, ds->dim, ds->dimp, c.buf.p, s));
        c.valid_for = ds->n;
    }
    *out = c.buf.p;
    return NDB_B200_OK;
}

// shared with ivf.cu / kmeans.cu: dense scan of `nq` row-major queries against an IL32 store
int dense_scan(const float *store, const void *vnorm, int64_t nvec, int dim, int dimp, int metric, int arith,
               const float *Q_dev, int nq, int k, ScanScratch &scr, int *out_nparts, cudaStream_t s)
{
    const int qt = scan_pick_qt(arith, dim, k);
    NDB_REQUIRE(qt > 0, NDB_B200_EINVAL, "unsupported scan shape: dim=%d k=%d (k must be 1..128)", dim, k);
    const int64_t B = (nvec + 31) / 32;
    const int ntiles = (nq + qt - 1) / qt;
    // cut the store into runs so that there are ~4 work items per resident CTA slot, but never
    // below 32 blocks (1024 vectors) per run: every run restarts the k-th-best filter from scratch,
    // and short runs turn most candidates into insertions (profiles/r01_scan_variants.txt)
    const int64_t target_items = (int64_t) ctx().sm_count * 12;
    int64_t nseg = (target_items + ntiles - 1) / ntiles;
    if (nseg < 1) nseg = 1;
    int64_t seg_blocks = (B + nseg - 1) / nseg;
    if (seg_blocks < 32) seg_blocks = 32;
    if (seg_blocks > B) seg_blocks = B > 0 ? B : 1;
    nseg = B > 0 ? (B + seg_blocks - 1) / seg_blocks : 1;
    const size_t norm_bytes = (metric == NDB_COSINE) ? norm_elem_size(arith) : 0;
    NDB_CHECK(scr.ensure((size_t) nq * nseg, k, nq, norm_bytes));
    if (norm_bytes) NDB_CHECK(row_norms(arith, Q_dev, nq, dim, scr.qnorm.p, s));
    NDB_CUDA(cudaMemsetAsync(scr.pslot.p, 0xff, (size_t) nq * nseg * k * sizeof(uint32_t), s));

    ScanParams p;
    memset(&p, 0, sizeof(p));
    p.vecs = reinterpret_cast<const float4 *>(store);
    p.vnorm = vnorm;
    p.Q = Q_dev;
    p.qnorm = norm_bytes ? scr.qnorm.p : nullptr;
    p.dim = dim; p.dimp = dimp; p.k = k;
    p.dense_ntiles = (uint32_t) ntiles;
    p.dense_items = (uint32_t) (nseg * ntiles);
    p.dense_seg_blocks = (uint32_t) seg_blocks;
    p.dense_nq = (uint32_t) nq;
    p.dense_nparts = (uint32_t) nseg;
    p.dense_nvec = (uint64_t) nvec;
    p.counter = scr.counter.as<uint32_t>();
    p.pdist = scr.pdist.as<float>();
    p.pslot = scr.pslot.as<uint32_t>();
    Context &c = ctx();
    if (c.timing) NDB_CUDA(cudaEventRecord(c.ev0, s));
    NDB_CHECK(launch_scan(metric, arith, qt, p, p.dense_items, s));
    if (c.timing) {
        NDB_CUDA(cudaEventRecord(c.ev1, s));
        c.last_bytes = (double) nvec * dim * 4.0 * ntiles;     // each tile streams the store once
        c.last_evals = (int64_t) nvec * nq;
        c.last_ms = -1.0;                                      // resolved lazily
    }
    *out_nparts = (int) nseg;
    return NDB_B200_OK;
}

// euclidean_distance as the double the reference sorts on: one thread per (query, candidate slot)
__global__ void knn_ml_rescore_kernel(const float *__restrict__ store, int dim, int dimp, const float *__restrict__ Q,
                                      const uint32_t *__restrict__ slots, int kq, int64_t total, double *__restrict__ out)
{
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const uint32_t slot = slots[t];
    if (slot == INVALID_SLOT) { out[t] = INFINITY; return; }
    const float *q = Q + (size_t) (t / kq) * dim;
    const float *v = store + (size_t) (slot >> 5) * IL * dimp + (size_t) (slot & 31) * 4;
    double sum = 0.0;
    for (int d = 0; d < dim; d++) {
        const double diff = (double) __fsub_rn(q[d], v[(size_t) (d >> 2) * (IL * 4) + (d & 3)]);      // double diff = a[i] - b[i] (:84)
        sum = __fma_rn(diff, diff, sum);                                                             // diff*diff is exact in fp64
    }
    out[t] = __dsqrt_rn(sum);
}

}  // namespace ndb

extern "C" {

int ndb_b200_dataset_create(int dim, ndb_b200_dataset **out)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(out && dim > 0 && dim <= 16000, NDB_B200_EINVAL, "dataset_create: dim must be 1..16000 (VECTOR_MAX_DIM)");
    ndb_b200_dataset *ds = new ndb_b200_dataset();
    ds->dim = dim;
    ds->dimp = round_up(dim, 4);
    ds->store.dim = dim;
    ds->store.dimp = ds->dimp;
    *out = ds;
    return NDB_B200_OK;
}

void ndb_b200_dataset_free(ndb_b200_dataset *ds)
{
    if (!ds) return;
    if (ctx().initialized) { cudaSetDevice(ctx().device); cudaStreamSynchronize(ctx().stream); }
    delete ds;
}

int64_t ndb_b200_dataset_size(const ndb_b200_dataset *ds) { return ds ? ds->n : 0; }

int ndb_b200_dataset_append(ndb_b200_dataset *ds, const float *rows, const int64_t *ids, int64_t n)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ds && rows && n > 0, NDB_B200_EINVAL, "dataset_append: NULL or empty input");
    const int64_t bad = find_nonfinite(rows, n * ds->dim);
    NDB_REQUIRE(bad < 0, NDB_B200_EVECTOR, "vector contains NaN or Infinity at row %lld index %lld",
                (long long) (bad / ds->dim), (long long) (bad % ds->dim));
    cudaStream_t s = ctx().stream;
    const int64_t chunk = 1 << 20;
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t m = n - off < chunk ? n - off : chunk;
        NDB_CHECK(ds->tmp_rows.reserve((size_t) m * ds->dim * sizeof(float)));
        NDB_CUDA(cudaMemcpyAsync(ds->tmp_rows.p, rows + (size_t) off * ds->dim, (size_t) m * ds->dim * sizeof(float),
                                 cudaMemcpyHostToDevice, s));
        const int64_t *idp = nullptr;
        if (ids) {
            NDB_CHECK(ds->tmp_ids.reserve((size_t) m * sizeof(int64_t)));
            NDB_CUDA(cudaMemcpyAsync(ds->tmp_ids.p, ids + off, (size_t) m * sizeof(int64_t), cudaMemcpyHostToDevice, s));
            idp = ds->tmp_ids.as<int64_t>();
        }
        NDB_CHECK(dataset_append_dev(ds, ds->tmp_rows.as<float>(), idp, m, s));
        NDB_CUDA(cudaStreamSynchronize(s));
    }
    return NDB_B200_OK;
}

int ndb_b200_dataset_append_dev(ndb_b200_dataset *ds, const float *rows_dev, const int64_t *ids_dev, int64_t n,
                                void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ds && rows_dev && n > 0, NDB_B200_EINVAL, "dataset_append_dev: NULL or empty input");
    return dataset_append_dev(ds, rows_dev, ids_dev, n, stream ? (cudaStream_t) stream : ctx().stream);
}

int ndb_b200_knn_exact_dev(ndb_b200_dataset *ds, int metric, int arith, const float *Q_dev, int nq, int k,
                           float *dist_dev, int64_t *ids_dev, void *stream)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ds && Q_dev && dist_dev && ids_dev && nq > 0, NDB_B200_EINVAL, "knn_exact: NULL or empty input");
    NDB_REQUIRE(k >= 1 && k <= 128, NDB_B200_EINVAL, "knn_exact: k=%d out of range 1..128", k);
    NDB_REQUIRE(metric >= NDB_L2 && metric <= NDB_IP, NDB_B200_EINVAL, "knn_exact: unknown metric %d", metric);
    NDB_REQUIRE(arith == NDB_ARITH_OP_F64 || arith == NDB_ARITH_IVF_F32 || arith == NDB_ARITH_FAST || arith == NDB_ARITH_TENSOR ||
                    (arith == NDB_ARITH_HNSW && metric == NDB_L2),
                NDB_B200_EINVAL, "knn_exact: arith %d not available for the scan (metric %d)", arith, metric);
    cudaStream_t s = stream ? (cudaStream_t) stream : ctx().stream;
    if (arith == NDB_ARITH_TENSOR) {
        // bf16 tcgen05 GEMM-form path (tolerance 1e-3): ||x||^2 - 2 x.q + ||q||^2 with fused top-k
        if (ds->tc.valid_for != ds->n) NDB_CHECK(tc_build_store(ds->tc, ds->store.ptr(), ds->n, ds->dim, ds->dimp, s));
        return tc_knn(ds->tc, ds->tcs, ds->dim, metric, Q_dev, nq, k, ds->ids.as<int64_t>(), dist_dev, ids_dev, nullptr, nullptr, false, s);
    }
    const void *vnorm = nullptr;
    if (metric == NDB_COSINE) NDB_CHECK(dataset_norms(ds, arith, &vnorm, s));
    int nparts = 1;
    NDB_CHECK(dense_scan(ds->store.ptr(), vnorm, ds->n, ds->dim, ds->dimp, metric, arith, Q_dev, nq, k, ds->scratch,
                         &nparts, s));
    return launch_merge_parts(ds->scratch.pdist.as<float>(), ds->scratch.pslot.as<uint32_t>(), ds->ids.as<int64_t>(), nq,
                              nparts, k, dist_dev, ids_dev, nullptr, s);
}

int ndb_b200_knn_exact(ndb_b200_dataset *ds, int metric, int arith, const float *Q, int nq, int k, float *dist,
                       int64_t *ids)
{
    NDB_CHECK(require_init());
    NDB_REQUIRE(ds && Q && dist && ids && nq > 0, NDB_B200_EINVAL, "knn_exact: NULL or empty input");
    cudaStream_t s = ctx().stream;
    const size_t qb = (size_t) nq * ds->dim * sizeof(float), m = (size_t) nq * k;
    NDB_CHECK(ds->qbuf.reserve(qb));
    NDB_CHECK(ds->outd.reserve(m * sizeof(float)));
    NDB_CHECK(ds->outi.reserve(m * sizeof(int64_t)));
    NDB_CUDA(cudaMemcpyAsync(ds->qbuf.p, Q, qb, cudaMemcpyHostToDevice, s));
    NDB_CHECK(validate_begin(ds->qbuf.as<float>(), (int64_t) nq * ds->dim, s));
    NDB_CHECK(ndb_b200_knn_exact_dev(ds, metric, arith, ds->qbuf.as<float>(), nq, k, ds->outd.as<float>(),
                                     ds->outi.as<int64_t>(), s));
    NDB_CUDA(cudaMemcpyAsync(dist, ds->outd.p, m * sizeof(float), cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaMemcpyAsync(ids, ds->outi.p, m * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    NDB_CUDA(cudaStreamSynchronize(s));
    const int64_t bad = validate_end();
    NDB_REQUIRE(bad < 0, NDB_B200_EVECTOR, "vector contains NaN or Infinity at index %lld", (long long) (bad % ds->dim));
    return NDB_B200_OK;
}

// ---- knn_classify / knn_regress (src/ml/ml_knn.c:112-357, 363-569) ---------------------------------------------
// The SQL functions read (feature, label) rows of a table, compute euclidean_distance (:76-90: f32 difference, f64
// sum, sqrt) of every row to the query, qsort by distance and vote among / average the k nearest.  Here the rows are a
// resident dataset, the distances and the top-k are the scan kernel's (NDB_ARITH_HNSW is that same arithmetic), and
// only the vote over k labels is left to the host.  labels[i] belongs to the i-th appended row.
// The scan kernel's top-k is ordered by the distance rounded to float; the reference sorts the doubles.  So kq = k +
// margin candidates are taken, re-evaluated as doubles and ordered by (double distance, row); the answer is certified
// when the float distance of the last candidate is strictly above that of the k-th (rounding is monotonic, so every row
// outside the candidates is then farther than the k-th), otherwise the query is repeated with the widest list (128).
static int knn_neighbours(ndb_b200_dataset *ds, const float *Q, int nq, int k, const char *who, std::vector<uint32_t> &slots)
{
    using namespace ndb;
    NDB_CHECK(require_init());
    NDB_REQUIRE(ds && Q && nq > 0, NDB_B200_EINVAL, "%s: NULL or empty input", who);
    NDB_REQUIRE(k >= 1, NDB_B200_EINVAL, "neurondb: %s: k must be at least 1, got %d", who, k);              // :149-154
    NDB_REQUIRE(k <= 128, NDB_B200_EINVAL, "%s: k=%d out of range 1..128", who, k);
    NDB_REQUIRE(ds->n >= k, NDB_B200_ERANGE, "neurondb: %s: need at least %d samples, got %lld", who, k, (long long) ds->n);  // :190-205
    cudaStream_t s = ctx().stream;
    const size_t qb = (size_t) nq * ds->dim * sizeof(float);
    NDB_CHECK(ds->qbuf.reserve(qb));
    NDB_CUDA(cudaMemcpyAsync(ds->qbuf.p, Q, qb, cudaMemcpyHostToDevice, s));
    NDB_CHECK(validate_begin(ds->qbuf.as<float>(), (int64_t) nq * ds->dim, s));
    slots.assign((size_t) nq * k, 0);
    DevBuf dslots, dexact;
    std::vector<uint32_t> cand;
    std::vector<float> cand_f;
    std::vector<double> cand_d;
    std::vector<int> order;
    std::vector<int> todo(nq), again;
    for (int q = 0; q < nq; q++) todo[q] = q;
    bool validated = false;
    for (int pass = 0; pass < 2 && !todo.empty(); pass++) {
        const int64_t cap = ds->n < 128 ? ds->n : 128;
        const int kq = (int) (pass == 0 ? (k + 8 < cap ? k + 8 : cap) : cap);
        const int m = (int) todo.size();
        const float *dq = ds->qbuf.as<float>();
        DevBuf qsub;
        if (pass > 0) {                                   // gather the queries that were not certified
            NDB_CHECK(qsub.reserve((size_t) m * ds->dim * sizeof(float)));
            for (int j = 0; j < m; j++)
                NDB_CUDA(cudaMemcpyAsync(qsub.as<float>() + (size_t) j * ds->dim, ds->qbuf.as<float>() + (size_t) todo[j] * ds->dim,
                                         (size_t) ds->dim * sizeof(float), cudaMemcpyDeviceToDevice, s));
            dq = qsub.as<float>();
        }
        const size_t mk = (size_t) m * kq;
        NDB_CHECK(ds->outd.reserve(mk * sizeof(float)));
        NDB_CHECK(dslots.reserve(mk * sizeof(uint32_t)));
        NDB_CHECK(dexact.reserve(mk * sizeof(double)));
        int nparts = 1;
        NDB_CHECK(dense_scan(ds->store.ptr(), nullptr, ds->n, ds->dim, ds->dimp, NDB_L2, NDB_ARITH_HNSW, dq, m, kq, ds->scratch, &nparts, s));
        // (no id table: the merged keys are the slots themselves -- labels[] is indexed by the row's position)
        NDB_CHECK(launch_merge_parts(ds->scratch.pdist.as<float>(), ds->scratch.pslot.as<uint32_t>(), nullptr, m, nparts, kq,
                                     ds->outd.as<float>(), nullptr, dslots.as<uint32_t>(), s));
        knn_ml_rescore_kernel<<<(unsigned) ((mk + 127) / 128), 128, 0, s>>>(ds->store.ptr(), ds->dim, ds->dimp, dq, dslots.as<uint32_t>(), kq,
                                                                            (int64_t) mk, dexact.as<double>());
        count_launch();
        cand.resize(mk); cand_f.resize(mk); cand_d.resize(mk);
        NDB_CUDA(cudaMemcpyAsync(cand.data(), dslots.p, mk * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        NDB_CUDA(cudaMemcpyAsync(cand_f.data(), ds->outd.p, mk * sizeof(float), cudaMemcpyDeviceToHost, s));
        NDB_CUDA(cudaMemcpyAsync(cand_d.data(), dexact.p, mk * sizeof(double), cudaMemcpyDeviceToHost, s));
        NDB_CUDA(cudaStreamSynchronize(s));
        if (!validated) {
            const int64_t bad = validate_end();
            NDB_REQUIRE(bad < 0, NDB_B200_EVECTOR, "vector contains NaN or Infinity at index %lld", (long long) (bad % ds->dim));   // NDB_CHECK_VECTOR_VALID
            validated = true;
        }
        again.clear();
        order.resize(kq);
        for (int j = 0; j < m; j++) {
            const size_t b = (size_t) j * kq;
            for (int i = 0; i < kq; i++) order[i] = i;
            std::sort(order.begin(), order.end(), [&](int x, int y) {
                return cand_d[b + x] < cand_d[b + y] || (cand_d[b + x] == cand_d[b + y] && cand[b + x] < cand[b + y]);
            });
            const bool all_rows = kq >= ds->n;
            const bool certified = all_rows || (kq > k && cand_f[b + kq - 1] > (float) cand_d[b + order[k - 1]]);
            if (!certified && pass == 0 && kq < cap) { again.push_back(todo[j]); continue; }
            for (int i = 0; i < k; i++) slots[(size_t) todo[j] * k + i] = cand[b + order[i]];
        }
        todo.swap(again);
    }
    return NDB_B200_OK;
}

int ndb_b200_knn_classify(ndb_b200_dataset *ds, const double *labels, const float *Q, int nq, int k, int *out_class)
{
    using namespace ndb;
    NDB_REQUIRE(labels && out_class, NDB_B200_EINVAL, "knn_classify: NULL labels or output");
    std::vector<uint32_t> slots;
    NDB_CHECK(knn_neighbours(ds, Q, nq, k, "knn_classify", slots));
    for (int q = 0; q < nq; q++) {
        double votes[2] = {0.0, 0.0};                               // :326-333: binary vote, other labels are ignored
        for (int i = 0; i < k; i++) {
            const int c = (int) labels[slots[(size_t) q * k + i]];
            if (c >= 0 && c < 2) votes[c] += 1.0;
        }
        out_class[q] = votes[1] > votes[0] ? 1 : 0;
    }
    return NDB_B200_OK;
}

int ndb_b200_knn_regress(ndb_b200_dataset *ds, const double *targets, const float *Q, int nq, int k, double *out)
{
    using namespace ndb;
    NDB_REQUIRE(targets && out, NDB_B200_EINVAL, "knn_regress: NULL targets or output");
    std::vector<uint32_t> slots;
    NDB_CHECK(knn_neighbours(ds, Q, nq, k, "knn_regress", slots));
    for (int q = 0; q < nq; q++) {
        double prediction = 0.0;                                    // :555-557: sum in neighbour order, divided by k
        for (int i = 0; i < k; i++) prediction += targets[slots[(size_t) q * k + i]];
        out[q] = prediction / k;
    }
    return NDB_B200_OK;
}

}  // extern "C"
