// pages.cuh -- the reference's on-disk page layouts, as read by the relation loaders.
//
// PostgreSQL page format (PG 16-18, x86-64): BLCKSZ 8192, PageHeaderData 24 B,
// ItemIdData 4 B {lp_off:15, lp_flags:2, lp_len:15}, LP_NORMAL = 1, LP_DEAD = 3, MAXALIGN 8,
// special space at the page tail (pd_special), meta structs at PageGetContents = page + 24.
//   IVF  : NeuronDB/src/index/ivf_am.c:75-84 (meta), :94-106 (centroid item), :241-261 (list page)
//   HNSW : NeuronDB/src/index/hnsw_am.c:108-120 (meta), :124-181 (node item)
#pragma once
#include <cstdint>
#include <cstring>

namespace ndb {
namespace pg {

constexpr uint32_t BLCKSZ = 8192;
constexpr uint32_t PAGE_HEADER = 24;
constexpr uint32_t LP_NORMAL = 1, LP_DEAD = 3;
constexpr uint32_t INVALID_BLOCK = 0xffffffffu;
constexpr uint32_t IVF_MAGIC = 0x49564646u, HNSW_MAGIC = 0x48534E57u;

struct PageHeader {
    uint64_t pd_lsn;
    uint16_t pd_checksum, pd_flags, pd_lower, pd_upper, pd_special, pd_pagesize_version;
    uint32_t pd_prune_xid;
};
struct IvfMeta { uint32_t magic, version; int32_t nlists, nprobe, dim; uint32_t centroidsBlock; int64_t insertedVectors; };
struct IvfCentroidHdr { int32_t listId, dim; int64_t memberCount; uint32_t firstBlock; uint32_t pad; };     // 24 B, vector follows
struct IvfListSpecial { uint32_t nextBlock; int32_t entryCount; };
struct HnswMeta { uint32_t magic, version, entryPoint; int32_t entryLevel, maxLevel; int16_t m, efConstruction, efSearch, pad; float ml; int64_t insertedVectors; };
constexpr uint32_t HNSW_NODE_HDR = 48;      // MAXALIGN(sizeof(HnswNodeData)): heapPtr 6, pad 2, level 4, dim 2, neighborCount[16] 32
constexpr uint32_t IVF_ENTRY_HDR = 8;       // MAXALIGN(sizeof(IvfListEntryData)): heapPtr 6, dim 2

inline const uint8_t *page_at(const void *blocks, uint32_t blk) { return (const uint8_t *) blocks + (size_t) blk * BLCKSZ; }
inline int max_offset(const uint8_t *page)
{
    PageHeader h;
    memcpy(&h, page, sizeof(h));
    return h.pd_lower <= PAGE_HEADER ? 0 : (int) ((h.pd_lower - PAGE_HEADER) / 4);
}
// line pointer `off` (1-based): byte offset of the item, flags, length
inline void item_id(const uint8_t *page, int off, uint32_t *lp_off, uint32_t *flags, uint32_t *len)
{
    uint32_t lp;
    memcpy(&lp, page + PAGE_HEADER + 4 * (off - 1), 4);
    *lp_off = lp & 0x7fff;
    *flags = (lp >> 15) & 3;
    *len = lp >> 17;
}
// A line pointer taken from a page image is untrusted input (a torn or corrupt page must not send the loaders
// outside the 8 KB block): item `off` is usable iff it is LP_NORMAL, lies between pd_upper and pd_special inside
// the block, and holds at least `min_len` bytes.  Returns 1 = usable, 0 = skip (unused / LP_DEAD), -1 = corrupt.
inline int checked_item(const uint8_t *page, int off, uint32_t min_len, uint32_t *lp_off, uint32_t *len)
{
    PageHeader h;
    memcpy(&h, page, sizeof(h));
    uint32_t lo, fl, ln;
    item_id(page, off, &lo, &fl, &ln);
    if (fl != LP_NORMAL) return 0;
    if (h.pd_special > BLCKSZ || h.pd_upper > h.pd_special || h.pd_lower > h.pd_upper) return -1;
    if (lo < h.pd_upper || lo < PAGE_HEADER || (uint64_t) lo + ln > h.pd_special || ln < min_len) return -1;
    *lp_off = lo;
    *len = ln;
    return 1;
}
inline uint16_t special_offset(const uint8_t *page)
{
    PageHeader h;
    memcpy(&h, page, sizeof(h));
    return h.pd_special;
}
// ItemPointerData {bi_hi, bi_lo, ip_posid} -> (block << 16) | offset
inline int64_t tid_unpack(const uint8_t *p)
{
    uint16_t v[3];
    memcpy(v, p, 6);
    return ((int64_t) (((uint32_t) v[0] << 16) | v[1]) << 16) | v[2];
}

}  // namespace pg
}  // namespace ndb
