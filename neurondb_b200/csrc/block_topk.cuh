// block_topk.cuh -- the k best of a CTA whose warps each hold a WarpTopK list (quantised scans: pq.cu, quant.cu).
#pragma once
#include "arith.cuh"

namespace ndb {

// Merges the lists of the block's warps through `smem` (>= nwarps * KR * 32 * 8 bytes, all warps past their last use of
// it) into warp 0's and writes its first k entries to pdist / pslot at offset `ob` (INVALID_SLOT where fewer than k).
template <int KR>
__device__ __forceinline__ void block_topk_write(WarpTopK<KR, uint32_t> &top, void *smem, int warp, int lane, int nwarps, int k,
                                                 float *__restrict__ pdist, uint32_t *__restrict__ pslot, size_t ob)
{
    float *sd = reinterpret_cast<float *>(smem);
    uint32_t *ss = reinterpret_cast<uint32_t *>(sd + nwarps * KR * 32);
#pragma unroll
    for (int r = 0; r < KR; r++) {
        sd[(warp * KR + r) * 32 + lane] = top.d[r];
        ss[(warp * KR + r) * 32 + lane] = top.key[r];
    }
    __syncthreads();
    if (warp != 0) return;
    for (int w = 1; w < nwarps; w++)
        for (int r = 0; r < KR; r++) {
            const float cd = sd[(w * KR + r) * 32 + lane];
            const uint32_t ck = ss[(w * KR + r) * 32 + lane];
            top.offer(cd, ck, ck != INVALID_SLOT, lane, k);
        }
#pragma unroll
    for (int r = 0; r < KR; r++) {
        const int e = r * 32 + lane;
        if (e < k) { pdist[ob + e] = top.d[r]; pslot[ob + e] = top.key[r]; }
    }
}

}  // namespace ndb
