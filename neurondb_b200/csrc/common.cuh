// common.cuh -- shared host/device plumbing for libndb_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/ndb_b200.h"

namespace ndb {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t INVALID_SLOT = 0xffffffffu;
constexpr int IL = 32;               // vectors per interleaved block (one warp)

// ---- error plumbing ------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern thread_local char g_last_error[512];

#define NDB_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            ndb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return NDB_B200_ECUDA;                                                          \
        }                                                                                   \
    } while (0)

#define NDB_CHECK(expr)                                                                     \
    do {                                                                                    \
        int rc__ = (expr);                                                                  \
        if (rc__ != NDB_B200_OK) return rc__;                                               \
    } while (0)

#define NDB_REQUIRE(cond, code, ...)                                                        \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            ndb::set_error(__VA_ARGS__);                                                    \
            return (code);                                                                  \
        }                                                                                   \
    } while (0)

// ---- global context ------------------------------------------------------------------
struct Context {
    bool initialized = false;
    uint64_t generation = 0;              // bumped by every ndb_b200_init: per-device caches compare against it
    int device = -1;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;        // library stream for host-pointer entry points
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;   // copy streams of the pipelined (begin/end) entry points
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timing = false;
    double last_ms = 0.0, last_bytes = 0.0;
    int64_t last_evals = 0;
    int64_t launches = 0;
    const void *stats_src = nullptr;      // device counter of vectors scanned by the last list scan
    int stats_dim = 0;
    void *pinned = nullptr;               // staging for small D2H/H2D control traffic
    size_t pinned_bytes = 0;
    unsigned long long *d_badidx = nullptr;   // device cell of validate_begin / validate_end
};
Context &ctx();
int require_init();

inline void count_launch(int n = 1) { ctx().launches += n; }

// ---- growable device buffer ------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return NDB_B200_OK;
        size_t want = bytes + bytes / 4 + 256;
        void *np = nullptr;
        cudaError_t e = cudaMalloc(&np, want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? NDB_B200_ENOMEM : NDB_B200_ECUDA;
        }
        if (p) cudaFree(p);
        p = np;
        cap = want;
        return NDB_B200_OK;
    }
    // grow keeping the first `keep` bytes
    int grow(size_t bytes, size_t keep, cudaStream_t s)
    {
        if (bytes <= cap) return NDB_B200_OK;
        size_t want = bytes + bytes / 2 + 256;
        void *np = nullptr;
        cudaError_t e = cudaMalloc(&np, want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? NDB_B200_ENOMEM : NDB_B200_ECUDA;
        }
        if (p && keep) {
            e = cudaMemcpyAsync(np, p, keep, cudaMemcpyDeviceToDevice, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) {
                cudaFree(np);
                set_error("device copy failed: %s", cudaGetErrorString(e));
                return NDB_B200_ECUDA;
            }
        }
        if (p) cudaFree(p);
        p = np;
        cap = want;
        return NDB_B200_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- interleaved-32 ("IL32") vector store ------------------------------------------------
// Vectors live in blocks of 32.  Inside a block the layout is [dimp/4][32 lanes][4 floats]:
// lane l of a warp reads the float4 chunk c of "its" vector at float4 index
//   blk * 8*dimp + c*32 + l
// so one warp-wide LDG.128 covers 512 contiguous bytes, and each thread walks the
// dimensions of one vector in order -- which is what lets the kernels reproduce the
// reference's sequential f32 / Kahan-f64 accumulation bit for bit while staying coalesced.
// dimp = dim rounded up to 4; pad lanes/elements are zero.
struct VecStore {
    int dim = 0, dimp = 0;
    int64_t nblk = 0;                 // blocks in use
    DevBuf data;                      // nblk * 32 * dimp floats
    size_t block_floats() const { return (size_t) IL * dimp; }
    size_t bytes_for(int64_t blocks) const { return (size_t) blocks * block_floats() * sizeof(float); }
    float *ptr() const { return data.as<float>(); }
};

// host validation: NaN/Inf anywhere -> index of the first bad element, else -1
int64_t find_nonfinite(const float *v, int64_t n);
// The same check for input already staged on the device (the query batches of the search entry
// points): validate_begin launches the scan, validate_end is called after the entry point's final
// stream synchronisation and returns the first offending index or -1.  Keeps the 5 MB-per-batch
// host scan out of the end-to-end path.
int validate_begin(const float *v_dev, int64_t n, cudaStream_t s);
// the same scan into a caller-owned device cell (pre-set to all ones): first offending index
int validate_into(const float *v_dev, int64_t n, unsigned long long *cell_dev, cudaStream_t s);
int64_t validate_end();

}  // namespace ndb
