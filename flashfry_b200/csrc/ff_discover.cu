// Off-target discovery on the GPU: the seed-and-verify scan kernel, hit ordering and the database-order overflow cut.
//
// What it replaces (FlashFry, src/main/scala/...):
//   OrderedBinTraversalFactory precompute   reference/traversal/OrderedBinTraversalFactory.scala:146-173
//   SeekTraverser / LinearTraverser scan    reference/traverser/SeekTraverser.scala:78-102
//   BlockManager.compareIndexedBlock        reference/binary/blocks/BlockManager.scala:143-201
//   BlockManager.compareLinearBlock         reference/binary/blocks/BlockManager.scala:212-254
//   BitEncoding.mismatches                  bitcoding/BitEncoding.scala:127-132
//   ResultsAggregator.updateOT / addOT      crispr/ResultsAggregator.scala:61-69, crispr/CRISPRSiteOT.scala:39-46
//
// Not a translation.  The reference prunes with one prefix (7-mer bin, then 11-mer sub-bin) and *tests every guide
// against every prefix*; it still compares ~3.3e6 targets per guide at k = 4 on a human-sized index.  Here:
//   * prefixes are ENUMERATED, not tested: substituting a base is an XOR of its 2-bit code with 1, 2 or 3, so the keys
//     within h mismatches of a guide's key are { key ^ m : m in M_h } for a fixed table M sorted by distance;
//   * the protospacer is split in two parts (first a bases | last b bases, a + b = P).  A pair within k mismatches has
//     d1 + d2 <= k, so either d1 <= hA (found through index A, keyed by the first part) or d1 > hA and then
//     d2 <= k - hA - 1 (found through index B, keyed by the second part and accepted only if d1 > hA).  The two cases
//     are disjoint and complete, so the hit SET equals the reference's brute-force set, with ~7e4 compares per guide.
// The kernel is guide-major: a warp takes (guide, batch of seeds), its lanes look the seeds' buckets up in the
// resident index (one batched, coalesced-latency lookup per 32 seeds), then the warp streams each bucket with
// 128-bit loads and compares the complementary part with XOR + fold + POPC.  Hits are staged per warp in shared
// memory and flushed with one global atomic per flush, as (guide, database index) keys; a radix sort and a per-guide
// warp scan then apply the reference's database-order overflow cut.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ff_common.cuh"
#include "ff_kernels.cuh"

namespace ff {

// ------------------------------------------------------------------------------------------------------------
// the scan kernel
struct SeedSide {
  const uint32_t *off;     // [4^w + 1]
  const uint32_t *other;   // complementary protospacer part per entry (padded, 16-byte aligned)
  const uint32_t *canon;   // database index per entry, or nullptr when entries are in database order
  const uint32_t *masks;   // mask | distance << 24, sorted by distance
  int n_seeds;             // masks [0, n_seeds) are within this pass's seed budget
  int seeds_per_item;      // seeds a warp takes at once
  int items;               // ceil(n_seeds / seeds_per_item)
};

struct ScanParams {
  const uint64_t *guides;
  int64_t n_guides;
  SeedSide A, B;
  int items_per_guide;     // A.items + B.items
  int proto_shift;         // protospacer position inside the target long
  int b_bits;              // 2 * (bases of part two)
  uint64_t proto_mask;
  int k;                   // max mismatches
  int hA;                  // pass A finds pairs with d1 <= hA, pass B the pairs with d1 > hA
  int tbits;               // bits of a database index: a hit key is guide << tbits | index (fewest radix-sort passes)
  uint64_t *hits;
  unsigned long long *hit_count;
  unsigned long long hit_cap;
  unsigned long long *n_compares;
};

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kHW = 32;  // staged hits per warp

struct WarpHits {  // everything the (rare) emit path needs, passed by value so the out-of-line path takes no pointer to the params
  uint64_t *buf;                 // this warp's staging slots (shared)
  unsigned int *count;           // this warp's counter (shared)
  uint64_t *hits;                // global candidate-hit buffer
  unsigned long long *hit_count;
  unsigned long long hit_cap;
  int hA;
};

__device__ __forceinline__ void emit_hit(const WarpHits &wh, uint64_t key) {
  const unsigned int slot = atomicAdd(wh.count, 1u);
  if (slot < kHW) {
    wh.buf[slot] = key;
  } else {  // staging full (a guide sitting in a repeat family): straight to global
    const unsigned long long g = atomicAdd(wh.hit_count, 1ull);
    if (g < wh.hit_cap) wh.hits[g] = key;
  }
}

__device__ __forceinline__ void flush_warp_hits(const WarpHits &wh, int lane) {
  __syncwarp();
  const unsigned int nh = min(*wh.count, (unsigned int)kHW);
  if (nh == 0) return;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(wh.hit_count, (unsigned long long)nh);
  base = __shfl_sync(0xffffffffu, base, 0);
  if ((unsigned int)lane < nh && base + lane < wh.hit_cap) wh.hits[base + lane] = wh.buf[lane];
  __syncwarp();
  if (lane == 0) *wh.count = 0;
  __syncwarp();
}

__device__ __forceinline__ uint4 ldg128(const uint32_t *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ int base_dist32(uint32_t x) {  // # non-zero 2-bit digits
  return __popc((x | (x >> 1)) & 0x55555555u);
}

// Compare one 128-bit chunk (4 entries starting at the 4-aligned index `base`) of a bucket [lo, hi): four
// XOR / fold / POPC and ONE branch in the common case.  The rare path (some entry within budget) is a compact rolled
// loop so that the hot loop stays small in the instruction cache: it applies the bucket-range check (the 16-byte
// aligned chunk may straddle the neighbouring buckets) and, in pass B, the d1 > hA rule, then emits.
// PAT (bulge patterns): the probed part may hold the pattern's wildcard base, which `pmask` drops from the count, and
// the d1 > hA threshold comes with the work item (it differs between pattern classes).
template <bool PASS_B, bool PAT>
__device__ __forceinline__ void verify_chunk(const ScanParams &p, const WarpHits &wh, const uint32_t *canon, uint4 v, uint32_t base,
                                             uint32_t lo, uint32_t hi, uint32_t probe, int budget, uint64_t guide_key,
                                             uint32_t pmask, int pat_lo_d) {
  if (PAT) { v.x = (v.x ^ probe) & pmask; v.y = (v.y ^ probe) & pmask; v.z = (v.z ^ probe) & pmask; v.w = (v.w ^ probe) & pmask; probe = 0; }
  const int d0 = base_dist32(v.x ^ probe), d1 = base_dist32(v.y ^ probe), d2 = base_dist32(v.z ^ probe), d3 = base_dist32(v.w ^ probe);
  if (min(min(d0, d1), min(d2, d3)) <= budget) {
    // rare path: which of the four entries are inside the bucket and within budget (and, in pass B, have d1 > hA)
    const int lo_d = PASS_B ? (PAT ? pat_lo_d : wh.hA) : -1;
    unsigned int ok = 0;
    ok |= (d0 <= budget && d0 > lo_d && base + 0 >= lo && base + 0 < hi) ? 1u : 0u;
    ok |= (d1 <= budget && d1 > lo_d && base + 1 >= lo && base + 1 < hi) ? 2u : 0u;
    ok |= (d2 <= budget && d2 > lo_d && base + 2 >= lo && base + 2 < hi) ? 4u : 0u;
    ok |= (d3 <= budget && d3 > lo_d && base + 3 >= lo && base + 3 < hi) ? 8u : 0u;
    while (ok) {
      const uint32_t idx = base + (uint32_t)(__ffs((int)ok) - 1);
      ok &= ok - 1u;
      emit_hit(wh, guide_key | (canon ? canon[idx] : idx));
    }
  }
}

#ifndef FF_GROUP
#define FF_GROUP 4
#endif
#ifndef FF_TAIL
#define FF_TAIL 4
#endif

// One pass (A or B) of one work item: `n` (<= 32) seeds starting at `seed0`.  Lanes look the buckets up (one batched
// index access per 32 seeds); the warp then streams the buckets FF_GROUP at a time: the first 128-entry chunk of every
// bucket of the group is requested back to back before any of them is verified, so each lane keeps FF_GROUP 128-bit
// loads in flight (tools/gather_bw.cu: random 288-byte runs reach 2.4 TB/s with one load in flight per warp, 4.3 TB/s
// with four).
// What one (pattern, pass) work item of the bulge / windowed scan adds to the plain scan.
struct PatItem {
  const uint32_t *masks;    // mask table of this item: full key width, or width - 1 when the key holds the wildcard base
  int wild_bit;             // bit offset of the wildcard base inside the key, -1 = none
  uint32_t pmask;           // applied to (entry ^ probe): drops the wildcard base when it lies in the probed part
  int lo_d;                 // pass B accepts only d1 > lo_d
  int c0, c1;               // database-order window: cells [c0, c1) of kCells
  int cell_shift;           // index-A key >> cell_shift = its cell
  const uint32_t *cell_off; // pass B: per-bucket offsets at the cell boundaries (nullptr = whole database)
};

template <bool PASS_B, bool PAT>
__device__ __forceinline__ void scan_seeds(const ScanParams &p, const SeedSide &sd, const WarpHits &wh, int lane, uint32_t key,
                                           uint32_t probe, int seed0, int n, uint64_t guide_key, unsigned long long &compares,
                                           const PatItem &pi) {
  uint32_t lo = 0, hi = 0;  // lanes >= n keep an empty bucket
  int budget = -1;
  if (!PAT) {
    if (lane < n) {
      const uint32_t m = sd.masks[seed0 + lane];
      const uint32_t kk = key ^ (m & 0xFFFFFFu);
      lo = sd.off[kk];
      hi = sd.off[kk + 1];
      budget = p.k - (int)(m >> 24);
    }
  } else if (lane < n) {
    uint32_t m, x;
    if (pi.wild_bit >= 0) {  // seed = (mask over the other key bases, value of the wildcard base)
      const int s = seed0 + lane;
      m = pi.masks[s >> 2];
      const uint32_t mm = m & 0xFFFFFFu, wb = (uint32_t)pi.wild_bit;
      x = ((mm >> wb) << (wb + 2)) | (mm & ((1u << wb) - 1u)) | ((uint32_t)(s & 3) << wb);
    } else {
      m = pi.masks[seed0 + lane];
      x = m & 0xFFFFFFu;
    }
    const uint32_t kk = key ^ x;
    budget = p.k - (int)(m >> 24);
    if (!pi.cell_off) {
      lo = sd.off[kk];
      hi = sd.off[kk + 1];
    } else if (PASS_B) {  // the part of the bucket whose database indices lie in the window
      lo = pi.cell_off[kk * (kCells + 1) + pi.c0];
      hi = pi.cell_off[kk * (kCells + 1) + pi.c1];
    } else {              // index A is in database order: a key is inside the window or not
      const int cell = (int)(kk >> pi.cell_shift);
      if (cell >= pi.c0 && cell < pi.c1) { lo = sd.off[kk]; hi = sd.off[kk + 1]; }
    }
  }
  compares += hi - lo;
  if (PAT) {
    // windows leave most buckets of a batch empty: move the non-empty ones to the front lanes so that the group loop
    // below runs over those only
    const unsigned int nz = __ballot_sync(0xffffffffu, hi > lo);
    n = __popc(nz);
    if (n == 0) return;
    const unsigned int src = __fns(nz, 0, lane + 1) & 31u;
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    budget = __shfl_sync(0xffffffffu, budget, src);
    if (lane >= n) { lo = 0; hi = 0; budget = -1; }
  }
  static_assert(32 % FF_GROUP == 0, "a group must not wrap around the warp");
  const uint32_t lane4 = 4u * lane;
  // does any bucket of this batch need more than its first 128-entry chunk?  (never, for part-one buckets of a
  // human-sized index; always, for part-two buckets)
  const bool any_long = __any_sync(0xffffffffu, hi - (lo & ~3u) > 128u);
  for (int l0 = 0; l0 < n; l0 += FF_GROUP) {
    uint32_t blo[FF_GROUP], bhi[FF_GROUP];
    uint4 v[FF_GROUP];
#pragma unroll
    for (int j = 0; j < FF_GROUP; ++j) {  // lanes beyond n hold empty buckets, so l0 + j (< 32) needs no bound check
      blo[j] = __shfl_sync(0xffffffffu, lo, l0 + j);
      bhi[j] = __shfl_sync(0xffffffffu, hi, l0 + j);
      v[j] = make_uint4(0, 0, 0, 0);  // (leaving idle lanes' registers undefined makes ptxas spill: measured 7.2 -> 10.2 ms)
      if ((blo[j] & ~3u) + lane4 < bhi[j]) v[j] = ldg128(sd.other + (blo[j] & ~3u) + lane4);
    }
#pragma unroll
    for (int j = 0; j < FF_GROUP; ++j) {
      const int bud = __shfl_sync(0xffffffffu, budget, l0 + j);
      const uint32_t base = (blo[j] & ~3u) + lane4;
      if (base < bhi[j]) verify_chunk<PASS_B, PAT>(p, wh, sd.canon, v[j], base, blo[j], bhi[j], probe, bud, guide_key, pi.pmask, pi.lo_d);
    }
    if (!any_long) continue;
    // buckets longer than 128 entries (pass B, repeat-rich part-one keys): stream the rest, FF_TAIL chunks in flight
#pragma unroll 1
    for (int j = 0; j < FF_GROUP; ++j) {
      const uint32_t jlo = __shfl_sync(0xffffffffu, lo, l0 + j), jhi = __shfl_sync(0xffffffffu, hi, l0 + j);
      if (jhi - (jlo & ~3u) <= 128u) continue;  // warp-uniform
      const int bud = __shfl_sync(0xffffffffu, budget, l0 + j);
      for (uint32_t c2 = (jlo & ~3u) + lane4 + 128u; c2 < jhi; c2 += 128u * FF_TAIL) {
        uint4 w[FF_TAIL];
#pragma unroll
        for (int c = 0; c < FF_TAIL; ++c) {
          w[c] = make_uint4(0, 0, 0, 0);
          if (c2 + 128u * c < jhi) w[c] = ldg128(sd.other + c2 + 128u * c);
        }
#pragma unroll
        for (int c = 0; c < FF_TAIL; ++c)
          if (c2 + 128u * c < jhi) verify_chunk<PASS_B, PAT>(p, wh, sd.canon, w[c], c2 + 128u * c, jlo, jhi, probe, bud, guide_key, pi.pmask, pi.lo_d);
      }
    }
  }
  __syncwarp();
  if (*wh.count >= kHW / 2) flush_warp_hits(wh, lane);
}

#ifndef FF_SCAN_MIN_BLOCKS
#define FF_SCAN_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(kScanThreads, FF_SCAN_MIN_BLOCKS) k_seed_scan(ScanParams p) {
  __shared__ uint64_t s_hits[kScanWarps * kHW];
  __shared__ unsigned int s_hitn[kScanWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpHits wh{s_hits + warp * kHW, &s_hitn[warp], p.hits, p.hit_count, p.hit_cap, p.hA};
  if (lane == 0) s_hitn[warp] = 0;
  __syncwarp();
  unsigned long long compares = 0;
  const long long n_items = p.n_guides * (long long)p.items_per_guide;
  const long long n_warps = (long long)gridDim.x * kScanWarps;
  for (long long item = (long long)blockIdx.x * kScanWarps + warp; item < n_items; item += n_warps) {
    // item -> (guide, batch); items of one guide are consecutive so neighbouring warps share the guide's cache lines
    const long long g = item / p.items_per_guide;
    const int bi = (int)(item - g * p.items_per_guide);
    const uint64_t guide = p.guides[g];
    const uint64_t proto = (guide >> p.proto_shift) & p.proto_mask;
    const uint32_t key_a = (uint32_t)(proto >> p.b_bits);
    const uint32_t key_b = (uint32_t)(proto & ((1ull << p.b_bits) - 1ull));
    const uint64_t guide_key = (uint64_t)g << p.tbits;
    if (bi < p.A.items) {
      const int seed0 = bi * p.A.seeds_per_item;
      scan_seeds<false, false>(p, p.A, wh, lane, key_a, key_b, seed0, min(p.A.seeds_per_item, p.A.n_seeds - seed0), guide_key, compares, PatItem{});
    } else {
      const int seed0 = (bi - p.A.items) * p.B.seeds_per_item;
      scan_seeds<true, false>(p, p.B, wh, lane, key_b, key_a, seed0, min(p.B.seeds_per_item, p.B.n_seeds - seed0), guide_key, compares, PatItem{});
    }
  }
  flush_warp_hits(wh, lane);
  for (int o = 16; o > 0; o >>= 1) compares += __shfl_down_sync(0xffffffffu, compares, o);
  if (lane == 0 && compares) atomicAdd(p.n_compares, compares);
}

// ------------------------------------------------------------------------------------------------------------
// ordering + overflow cut
// seg_start[g] = first sorted key whose guide index >= g   (g in [0, n_guides])
__global__ void k_segments(const uint64_t *__restrict__ keys, int64_t n_hits, int64_t n_guides, int tbits, int64_t *__restrict__ seg_start) {
  int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g > n_guides) return;
  const uint64_t want = (uint64_t)g << tbits;
  int64_t lo = 0, hi = n_hits;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < want) lo = mid + 1; else hi = mid;
  }
  seg_start[g] = lo;
}

// One warp per guide: walk its hits in database order and keep the shortest prefix whose summed occurrence count
// reaches max_ot (ResultsAggregator.scala:61-69 / CRISPRSiteOT.scala:39-46: append while currentTotal < overflow).
__global__ void k_overflow_cut(const uint64_t *__restrict__ keys, const int64_t *__restrict__ seg_start,
                               const uint64_t *__restrict__ targets, int64_t n_guides, int max_ot, int tbits,
                               int64_t *__restrict__ n_keep, int32_t *__restrict__ total_count,
                               uint8_t *__restrict__ overflowed) {
  const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_guides) return;
  const int64_t s0 = seg_start[g], s1 = seg_start[g + 1];
  long long running = 0;
  int64_t kept = 0;
  for (int64_t base = s0; base < s1 && running < max_ot; base += 32) {
    const int64_t i = base + lane;
    long long c = 0;
    if (i < s1) c = (long long)(targets[keys[i] & ((1ull << tbits) - 1ull)] >> 48);
    long long incl = c;
    for (int o = 1; o < 32; o <<= 1) {
      long long v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const bool keep = (i < s1) && (running + incl - c < max_ot);
    const unsigned int km = __ballot_sync(0xffffffffu, keep);
    const int nk = __popc(km);  // kept hits form a prefix of the chunk
    kept += nk;
    long long chunk_total = __shfl_sync(0xffffffffu, incl, nk > 0 ? nk - 1 : 0);
    if (nk > 0) running += chunk_total;
    if (nk < 32) break;
  }
  if (lane == 0) {
    n_keep[g] = kept;
    total_count[g] = (int32_t)running;
    overflowed[g] = running >= max_ot ? 1 : 0;
  }
}

// One warp per guide: copy the kept prefix out (target long, mismatch count, target index).
__global__ void k_gather(const uint64_t *__restrict__ keys, const int64_t *__restrict__ seg_start,
                         const int64_t *__restrict__ row_ptr, const uint64_t *__restrict__ targets,
                         const uint64_t *__restrict__ guides, uint64_t cmp_mask, int64_t n_guides, int tbits,
                         uint64_t *__restrict__ out_targets, uint8_t *__restrict__ out_mm, uint32_t *__restrict__ out_tidx) {
  const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_guides) return;
  const int64_t s0 = seg_start[g];
  const int64_t r0 = row_ptr[g], r1 = row_ptr[g + 1];
  const uint64_t guide = guides[g];
  for (int64_t i = lane; i < r1 - r0; i += 32) {
    const uint32_t t = (uint32_t)(keys[s0 + i] & ((1ull << tbits) - 1ull));
    const uint64_t tl = targets[t];
    out_targets[r0 + i] = tl;
    out_mm[r0 + i] = (uint8_t)mismatches64(guide, tl, cmp_mask);
    out_tidx[r0 + i] = t;
  }
}

__global__ void k_pos_counts(const uint64_t *__restrict__ out_targets, int64_t n_hits, int64_t *__restrict__ cnt) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_hits) cnt[i] = (int64_t)(out_targets[i] >> 48);
}

__global__ void k_gather_positions(const uint32_t *__restrict__ out_tidx, const int64_t *__restrict__ pos_ptr,
                                   const uint64_t *__restrict__ pos_off, const uint64_t *__restrict__ positions,
                                   int64_t n_hits, uint64_t *__restrict__ out_positions) {
  // one warp per hit
  const int64_t h = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (h >= n_hits) return;
  const int64_t o0 = pos_ptr[h], n = pos_ptr[h + 1] - o0;
  const uint64_t src = pos_off[out_tidx[h]];
  for (int64_t i = lane; i < n; i += 32) out_positions[o0 + i] = positions[src + i];
}

// ------------------------------------------------------------------------------------------------------------
// Ordering the candidate hits without a full radix sort, and without the host in the loop.  A typical guide has ~10^2
// hits, so the keys are grouped by guide with a counting sort (per-guide counts, prefix sum, scatter of the 32-bit
// database indices) and every guide's short segment is sorted, cut by the overflow rule and gathered inside ONE warp
// (k_sort_cut: bitonic network in shared memory, then the walk of k_overflow_cut, then the row itself written at the
// segment's position).  Segments of 257..16 384 candidates (a guide inside a repeat family) are sorted beforehand by one
// CTA each (k_sort_long); only a longer segment, or more than kLongCap long ones, sends the call to the radix sort.
// Every kernel reads the number of candidates from device memory: the host synchronises once, at the end.
constexpr int kSegSortMax = 256;
constexpr int kLongMax = 16384;   // longest segment k_sort_long takes (64 KB of shared memory)
constexpr int kLongCap = 1024;    // long segments per call

// status words the pipeline leaves for the host (mirrored into pinned memory with one copy)
struct PlainStatus {
  unsigned long long n_cand, n_compares;
  long long n_hits;
  unsigned int flag, n_long;
  unsigned long long n_compares_b;  // entries streamed by the part-two kernel of the bin scan (n_compares: part one)
  unsigned int seq;                 // host mirror only: sequence number of the publication (written last)
};

__global__ void k_guide_hist(const uint64_t *__restrict__ keys, const unsigned long long *__restrict__ n_ptr, unsigned long long cap, int tbits,
                             unsigned int *__restrict__ cnt) {
  const unsigned long long n = min(*n_ptr, cap);
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
    atomicAdd(cnt + (keys[i] >> tbits), 1u);
}

__global__ void k_guide_scatter(const uint64_t *__restrict__ keys, const unsigned long long *__restrict__ n_ptr, unsigned long long cap, int tbits,
                                const int64_t *__restrict__ seg_start, unsigned int *__restrict__ cursor, uint32_t *__restrict__ out) {
  if (*n_ptr > cap) return;  // the candidate buffer overflowed: the call is repeated with a larger one
  const unsigned long long n = *n_ptr;
  const uint64_t low = (1ull << tbits) - 1ull;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint64_t key = keys[i];
    const uint64_t g = key >> tbits;
    out[seg_start[g] + (int64_t)atomicAdd(cursor + g, 1u)] = (uint32_t)(key & low);
  }
}

// The same placement when the scan has recorded every candidate's rank among its guide's: no atomics.
__global__ void k_guide_place(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ ranks, const unsigned long long *__restrict__ n_ptr,
                              unsigned long long cap, int tbits, const int64_t *__restrict__ seg_start, uint32_t *__restrict__ out) {
  if (*n_ptr > cap) return;
  const unsigned long long n = *n_ptr;
  const uint64_t low = (1ull << tbits) - 1ull;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint64_t key = keys[i];
    out[seg_start[key >> tbits] + (int64_t)ranks[i]] = (uint32_t)(key & low);
  }
}

__global__ void k_widen_counts(const unsigned int *__restrict__ cnt, int64_t n, int64_t *__restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int64_t)cnt[i];
  if (i == n) out[i] = 0;
}

__global__ void k_mark_long(const int64_t *__restrict__ seg_start, int64_t n_guides, uint32_t *__restrict__ long_list, PlainStatus *__restrict__ stt) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= n_guides) return;
  const int64_t n = seg_start[g + 1] - seg_start[g];
  if (n <= kSegSortMax) return;
  if (n > kLongMax) { atomicOr(&stt->flag, 1u); return; }
  const unsigned int pos = atomicAdd(&stt->n_long, 1u);
  if (pos < (unsigned int)kLongCap) long_list[pos] = (uint32_t)g; else atomicOr(&stt->flag, 1u);
}

// one CTA per long segment: bitonic sort of up to kLongMax 32-bit indices in shared memory
__global__ void __launch_bounds__(512) k_sort_long(uint32_t *__restrict__ idx, const int64_t *__restrict__ seg_start,
                                                   const uint32_t *__restrict__ long_list, const PlainStatus *__restrict__ stt) {
  extern __shared__ uint32_t s_long[];
  if (blockIdx.x >= min(stt->n_long, (unsigned int)kLongCap) || stt->flag) return;
  const uint32_t g = long_list[blockIdx.x];
  const int64_t s0 = seg_start[g];
  const int n = (int)(seg_start[g + 1] - s0);
  int np = 512;
  while (np < n) np <<= 1;
  for (int i = threadIdx.x; i < np; i += blockDim.x) s_long[i] = i < n ? idx[s0 + i] : 0xFFFFFFFFu;
  __syncthreads();
  for (int k = 2; k <= np; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np; i += blockDim.x) {
        const int x = i ^ j;
        if (x > i) {
          const uint32_t a = s_long[i], b = s_long[x];
          if ((a > b) == ((i & k) == 0)) { s_long[i] = b; s_long[x] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < n; i += blockDim.x) idx[s0 + i] = s_long[i];
}

// What the walk over a guide's candidates in database order carries from chunk to chunk.
struct CutState {
  long long running;
  int64_t kept;
  bool done;
};

// One chunk of 32 candidates (lane i holds database index t, valid when i < n): keep the prefix while the running sum of
// occurrence counts is below max_ot (ResultsAggregator.scala:61-69 / CRISPRSiteOT.scala:39-46: append while
// currentTotal < overflow) and write the kept rows -- target long, mismatch count, database index -- at the segment's
// own position (k_compact_rows closes the gaps between segments).
__device__ __forceinline__ void cut_chunk(CutState &cs, int64_t i, int64_t n, uint32_t t, uint64_t tl, int lane,
                                          uint64_t guide, uint64_t cmp_mask, int max_ot, int64_t s0, uint64_t *__restrict__ st_targets,
                                          uint8_t *__restrict__ st_mm, uint32_t *__restrict__ idx) {
  const int c = (int)(tl >> 48);  // tl = targets[t], 0 past the end of the segment
  int incl = c;  // a chunk sums to < 2^21
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const bool keep = (i < n) && (cs.running + (long long)(incl - c) < (long long)max_ot);
  const unsigned int km = __ballot_sync(0xffffffffu, keep);
  const int nk = __popc(km);  // kept hits form a prefix of the chunk
  if (keep) {
    st_targets[s0 + i] = tl;
    st_mm[s0 + i] = (uint8_t)mismatches64(guide, tl, cmp_mask);
    idx[s0 + i] = t;
  }
  cs.kept += nk;
  const int chunk_total = __shfl_sync(0xffffffffu, incl, nk > 0 ? nk - 1 : 0);
  if (nk > 0) cs.running += chunk_total;
  if (nk < 32 || cs.running >= max_ot) cs.done = true;
}

// A segment of at most 32 E candidates: bitonic sort in registers (element e * 32 + lane lives in v[e] of the lane:
// partners closer than 32 are reached with one shuffle, the others are in the same lane), then the walk.
template <int E>
__device__ __forceinline__ void sort_cut_regs(CutState &cs, int64_t n, int lane, int64_t s0, uint32_t *__restrict__ idx,
                                              const uint64_t *__restrict__ targets, uint64_t guide, uint64_t cmp_mask, int max_ot,
                                              uint64_t *__restrict__ st_targets, uint8_t *__restrict__ st_mm) {
  uint32_t v[E];
#pragma unroll
  for (int e = 0; e < E; ++e) v[e] = (e * 32 + lane < n) ? idx[s0 + e * 32 + lane] : 0xFFFFFFFFu;
#pragma unroll
  for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int je = j >> 5;
#pragma unroll
        for (int e = 0; e < E; ++e)
          if ((e & je) == 0) {
            const bool up = ((e * 32) & k) == 0;  // k >= 64 here: the direction depends on e only
            const uint32_t a = v[e], b = v[e | je];
            const uint32_t lo = min(a, b), hi = max(a, b);
            v[e] = up ? lo : hi; v[e | je] = up ? hi : lo;
          }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int i = e * 32 + lane;
          const bool up = (i & k) == 0, lower = (lane & j) == 0;
          const uint32_t o = __shfl_xor_sync(0xffffffffu, v[e], j);
          v[e] = (lower == up) ? min(v[e], o) : max(v[e], o);
        }
      }
    }
  }
  // all the gathers of the segment are issued together (one round trip to HBM per guide instead of one per chunk)
  uint64_t tl[E];
#pragma unroll
  for (int e = 0; e < E; ++e) tl[e] = (e * 32 + lane < n) ? targets[v[e]] : 0ull;
#pragma unroll
  for (int e = 0; e < E; ++e)
    if (!cs.done && e * 32 < n) cut_chunk(cs, e * 32 + lane, n, v[e], tl[e], lane, guide, cmp_mask, max_ot, s0, st_targets, st_mm, idx);
}

// One warp per guide: sort its candidates by database index (up to kSegSortMax: in registers; longer segments arrive
// sorted from k_sort_long and are streamed), cut, and stage the kept row.
__global__ void __launch_bounds__(256) k_sort_cut(uint32_t *__restrict__ idx, const int64_t *__restrict__ seg_start, int64_t n_guides,
                                                  const uint64_t *__restrict__ targets, const uint64_t *__restrict__ guides, uint64_t cmp_mask,
                                                  int max_ot, uint64_t *__restrict__ st_targets, uint8_t *__restrict__ st_mm,
                                                  int64_t *__restrict__ n_keep, int32_t *__restrict__ total_count, uint8_t *__restrict__ overflowed,
                                                  const PlainStatus *__restrict__ stt, unsigned long long cap) {
  const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_guides) return;
  if (stt->n_cand > cap || stt->flag) {  // overflowed buffer / a segment for the radix sort: this attempt is void
    if (lane == 0) n_keep[g] = 0;
    return;
  }
  const int64_t s0 = seg_start[g];
  const int64_t n = seg_start[g + 1] - s0;
  const uint64_t guide = guides[g];
  CutState cs{0, 0, max_ot <= 0 || n == 0};
  if (n <= 32) sort_cut_regs<1>(cs, n, lane, s0, idx, targets, guide, cmp_mask, max_ot, st_targets, st_mm);
  else if (n <= 64) sort_cut_regs<2>(cs, n, lane, s0, idx, targets, guide, cmp_mask, max_ot, st_targets, st_mm);
  else if (n <= 128) sort_cut_regs<4>(cs, n, lane, s0, idx, targets, guide, cmp_mask, max_ot, st_targets, st_mm);
  else if (n <= kSegSortMax) sort_cut_regs<8>(cs, n, lane, s0, idx, targets, guide, cmp_mask, max_ot, st_targets, st_mm);
  else
    for (int64_t base = 0; base < n && !cs.done; base += 32) {
      const int64_t i = base + lane;
      const uint32_t t = i < n ? idx[s0 + i] : 0u;
      cut_chunk(cs, i, n, t, i < n ? targets[t] : 0ull, lane, guide, cmp_mask, max_ot, s0, st_targets, st_mm, idx);
    }
  if (lane == 0) {
    n_keep[g] = cs.kept;
    total_count[g] = (int32_t)cs.running;
    overflowed[g] = cs.running >= max_ot ? 1 : 0;
  }
}

// the status words, written into mapped host memory (no copy engine involved)
__global__ void k_publish_status(const PlainStatus *__restrict__ stt, const int64_t *__restrict__ row_ptr_end, volatile PlainStatus *host,
                                 unsigned int seq) {
  if (threadIdx.x == 0) {
    host->n_cand = stt->n_cand; host->n_compares = stt->n_compares;
    host->n_hits = row_ptr_end ? *row_ptr_end : stt->n_hits;
    host->flag = stt->flag; host->n_long = stt->n_long; host->n_compares_b = stt->n_compares_b;
    __threadfence_system();
    host->seq = seq;  // the host spins on this word: no driver round trip between the last kernel and the D2H that follows
    __threadfence_system();
  }
}

// A few 64-bit words from device memory to the host without the copy engine (general path: per-window counters).
__global__ void k_publish_words(const unsigned long long *a, const unsigned long long *b, const unsigned long long *c,
                                volatile unsigned long long *host, volatile unsigned int *host_seq, unsigned int seq) {
  if (threadIdx.x == 0) {
    host[0] = a ? *a : 0ull; host[1] = b ? *b : 0ull; host[2] = c ? *c : 0ull;
    __threadfence_system();
    *host_seq = seq;
    __threadfence_system();
  }
}

// Wait for publication `seq` of the status words.  Spinning on the mapped word costs a few microseconds less than
// cudaStreamSynchronize and, unlike it, is not delayed by copies in flight on the context's other stream; after a
// generous number of polls the stream is synchronised anyway (a kernel fault must not hang the host).
static int wait_status(ff_ctx *ctx, cudaStream_t st, unsigned int seq) {
  volatile PlainStatus *h = static_cast<volatile PlainStatus *>(ctx->h_status);
  for (long spin = 0; spin < 20000000L; ++spin) {
    if (h->seq == seq) return FF_OK;
    if ((spin & 1023) == 1023 && cudaStreamQuery(st) != cudaErrorNotReady) break;  // finished (or failed) meanwhile
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  FF_CUDA(cudaStreamSynchronize(st));
  if (h->seq != seq) { set_error("status words were not published"); return FF_ECUDA; }
  return FF_OK;
}

// one warp per guide: move the kept row from its segment to its place in the CSR
__global__ void k_compact_rows(const int64_t *__restrict__ seg_start, const int64_t *__restrict__ row_ptr, int64_t n_guides,
                               const uint64_t *__restrict__ st_targets, const uint8_t *__restrict__ st_mm, const uint32_t *__restrict__ idx,
                               uint64_t *__restrict__ out_targets, uint8_t *__restrict__ out_mm, uint32_t *__restrict__ out_tidx,
                               PlainStatus *__restrict__ stt) {
  const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_guides) return;
  const int64_t s0 = seg_start[g], r0 = row_ptr[g], n = row_ptr[g + 1] - r0;
  for (int64_t i0 = 0; i0 < n; i0 += 128) {  // four chunks of 32 at a time: their loads are in flight together
    uint64_t t[4];
    uint32_t x[4];
    uint8_t m[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * 32 + lane;
      if (i < n) { t[u] = st_targets[s0 + i]; m[u] = st_mm[s0 + i]; x[u] = idx[s0 + i]; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * 32 + lane;
      if (i < n) { out_targets[r0 + i] = t[u]; out_mm[r0 + i] = m[u]; out_tidx[r0 + i] = x[u]; }
    }
  }
  if (g == n_guides - 1 && lane == 0) stt->n_hits = row_ptr[n_guides];
}

// Read up to three device words on the host: a kernel writes them into the mapped status block (words 16..18, i.e. past
// PlainStatus), the host spins on the shared sequence word.
static int fetch_words(ff_ctx *ctx, cudaStream_t st, const void *a, unsigned long long *ha, const void *b = nullptr,
                       unsigned long long *hb = nullptr, const void *c = nullptr, unsigned long long *hc = nullptr) {
  volatile unsigned long long *hw = static_cast<volatile unsigned long long *>(ctx->h_status) + 16;
  unsigned long long *dw = static_cast<unsigned long long *>(ctx->h_status_dev) + 16;
  PlainStatus *hs_dev = static_cast<PlainStatus *>(ctx->h_status_dev);
  k_publish_words<<<1, 32, 0, st>>>(static_cast<const unsigned long long *>(a), static_cast<const unsigned long long *>(b),
                                    static_cast<const unsigned long long *>(c), dw, &hs_dev->seq, ++ctx->status_seq);
  FF_TRY(wait_status(ctx, st, ctx->status_seq));
  if (ha) *ha = hw[0];
  if (hb) *hb = hw[1];
  if (hc) *hc = hw[2];
  return FF_OK;
}

static inline unsigned int blocks_for(int64_t n, int threads) { return (unsigned int)((n + threads - 1) / threads); }

// Positions of the emitted hits (only with want_positions and a database image that holds positions).
static int gather_positions(ff_ctx *ctx, ff_ctx::OutSlot &os, bool want_positions, int64_t n_hits, DeviceResult *res, int64_t *n_pos_out,
                            int *launches) {
  Database &db = ctx->db;
  cudaStream_t st = ctx->stream;
  const int64_t Hp = n_hits > 0 ? n_hits : 1;
  size_t tmp_bytes = 0;
  int64_t n_pos = 0;
  res->d_pos_ptr = nullptr; res->d_positions = nullptr;
  if (want_positions && db.d_positions) {
    FF_TRY(ctx->pos_cnt.reserve((Hp + 1) * 8));
    FF_TRY(ctx->pos_ptr.reserve((Hp + 1) * 8));
    FF_CUDA(cudaMemsetAsync(ctx->pos_cnt.p, 0, (Hp + 1) * 8, st));
    if (n_hits > 0) {
      k_pos_counts<<<blocks_for(n_hits, 256), 256, 0, st>>>(os.out_targets.as<uint64_t>(), n_hits, ctx->pos_cnt.as<int64_t>());
      (*launches)++;
    }
    FF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ctx->pos_cnt.as<int64_t>(), ctx->pos_ptr.as<int64_t>(), n_hits + 1, st));
    FF_TRY(ctx->cub_tmp.reserve(tmp_bytes));
    FF_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, ctx->pos_cnt.as<int64_t>(), ctx->pos_ptr.as<int64_t>(), n_hits + 1, st));
    *launches += 2;
    FF_CUDA(cudaMemcpyAsync(&n_pos, ctx->pos_ptr.as<int64_t>() + n_hits, 8, cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    FF_TRY(ctx->out_positions.reserve((n_pos > 0 ? n_pos : 1) * 8));
    if (n_hits > 0) {
      k_gather_positions<<<blocks_for(n_hits * 32, 256), 256, 0, st>>>(os.out_tidx.as<uint32_t>(), ctx->pos_ptr.as<int64_t>(), db.d_pos_off,
                                                                      db.d_positions, n_hits, ctx->out_positions.as<uint64_t>());
      (*launches)++;
    }
    res->d_pos_ptr = ctx->pos_ptr.as<int64_t>();
    res->d_positions = ctx->out_positions.as<uint64_t>();
  }
  *n_pos_out = n_pos;
  return FF_OK;
}

// Choose hA (pass A covers d1 <= hA, pass B covers d1 > hA, i.e. d2 <= k - hA - 1) by the expected number of entries
// streamed per guide: seeds x (average bucket + a fixed per-seed cost).
static void plan_passes(const Database &db, int k, int *hA_out, int *nA, int *nB) {
  const double n = (double)db.n_targets;
  const int a = db.A.key_bases, b = db.B.key_bases;
  const double bucket_a = n / (double)(1ull << (2 * a)), bucket_b = n / (double)(1ull << (2 * b));
  const double per_seed = 24.0;  // index lookup + partially filled 128-entry chunk, in entry-equivalents
  double best = -1.0;
  for (int h = 0; h <= std::min(k, a); ++h) {
    const int hb = k - h - 1;
    const double sa = (double)db.A.cum[std::min(h, a)];
    const double sb = hb < 0 ? 0.0 : (double)db.B.cum[std::min(hb, b)];
    const double cost = sa * (bucket_a + per_seed) + sb * (bucket_b + per_seed);
    if (best < 0 || cost < best) { best = cost; *hA_out = h; *nA = (int)sa; *nB = (int)sb; }
  }
}

#include "ff_binscan.inl"

struct U32ToI64 {
  __host__ __device__ int64_t operator()(unsigned int v) const { return (int64_t)v; }
};

// The scan's view of a call: both index halves, the pass plan for this k, the key layout.
static void fill_scan_params(ff_ctx *ctx, const uint64_t *d_guides, int64_t G, int max_mm, ScanParams *spp, int *hA_out, int *nA_out, int *nB_out) {
  Database &db = ctx->db;
  ScanParams &sp = *spp;
  const int k_eff = std::min(max_mm, db.proto_bases);  // more mismatches than compared bases changes nothing
  int hA = 0, nA = 1, nB = 0;
  plan_passes(db, k_eff, &hA, &nA, &nB);
  sp.guides = d_guides; sp.n_guides = G;
  sp.A.off = db.A.d_off; sp.A.other = db.A.d_other; sp.A.canon = db.A.d_canon; sp.A.masks = db.A.d_masks;
  sp.A.n_seeds = nA; sp.A.seeds_per_item = 32; sp.A.items = (nA + 31) / 32;
  sp.B.off = db.B.d_off; sp.B.other = db.B.d_other; sp.B.canon = db.B.d_canon; sp.B.masks = db.B.d_masks;
  sp.B.n_seeds = nB;
  {  // part-two buckets are 4^(a-b) times longer: hand them out in smaller batches
    const double bucket_b = (double)db.n_targets / (double)(1ull << (2 * db.B.key_bases));
    int spi = bucket_b > 2048 ? 1 : bucket_b > 512 ? 4 : bucket_b > 128 ? 8 : 32;
    if (ctx->opt.b_spi > 0) spi = ctx->opt.b_spi;
    sp.B.seeds_per_item = spi; sp.B.items = (nB + spi - 1) / spi;
  }
  sp.items_per_guide = sp.A.items + sp.B.items;
  sp.proto_shift = db.proto_shift; sp.b_bits = 2 * db.B.key_bases; sp.proto_mask = (1ull << (2 * db.proto_bases)) - 1ull;
  sp.k = k_eff; sp.hA = hA;
  int tbits = 1;
  while ((1ull << tbits) < db.n_targets + 1) tbits++;
  sp.tbits = tbits;
  sp.hits = nullptr; sp.hit_count = nullptr; sp.hit_cap = 0; sp.n_compares = nullptr;
  *hA_out = hA; *nA_out = nA; *nB_out = nB;
}

// The per-guide ordering pipeline (section 3c of DESIGN.md), all on the stream, no host round trip: candidates
// `hits[0 .. min(d_stt->n_cand, cap))` (guide << tbits | database index) -> CSR rows in the output slot.  `cnt` holds the
// per-guide candidate counts (need_hist: take them here), `cursor` is zeroed scratch of the same size + the long-segment list.
static int order_grouped(ff_ctx *ctx, ff_ctx::OutSlot &os, const uint64_t *hits, PlainStatus *d_stt, size_t cap, int tbits, unsigned int *cnt,
                         unsigned int *cursor, bool need_hist, const uint64_t *d_guides, int64_t G, int max_ot, int *launches_out,
                         const uint32_t *ranks = nullptr) {
  Database &db = ctx->db;
  cudaStream_t st = ctx->stream;
  const int64_t Gp = G > 0 ? G : 1;
  size_t tmp_bytes = 0;
  int launches = 0;
  const bool bin_major = !need_hist;
  {
      FF_TRY(ctx->idx32.reserve((cap + 1) * 4));
      FF_TRY(ctx->st_targets.reserve((cap + 1) * 8));
      FF_TRY(ctx->st_mm.reserve(cap + 1));
      uint32_t *long_list = reinterpret_cast<uint32_t *>(cursor + Gp);
      const int sgrid = ctx->sm_count * 16;
      if (!bin_major) k_guide_hist<<<sgrid, 256, 0, st>>>(hits, &d_stt->n_cand, cap, tbits, cnt);
      cub::TransformInputIterator<int64_t, U32ToI64, unsigned int *> cnt64(cnt, U32ToI64());
      FF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt64, ctx->seg_start.as<int64_t>(), G + 1, st));
      FF_TRY(ctx->cub_tmp.reserve(tmp_bytes));
      FF_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, cnt64, ctx->seg_start.as<int64_t>(), G + 1, st));
      if (ranks && !need_hist) k_guide_place<<<sgrid, 256, 0, st>>>(hits, ranks, &d_stt->n_cand, cap, tbits, ctx->seg_start.as<int64_t>(), ctx->idx32.as<uint32_t>());
      else k_guide_scatter<<<sgrid, 256, 0, st>>>(hits, &d_stt->n_cand, cap, tbits, ctx->seg_start.as<int64_t>(), cursor, ctx->idx32.as<uint32_t>());
      k_mark_long<<<blocks_for(G, 256), 256, 0, st>>>(ctx->seg_start.as<int64_t>(), G, long_list, d_stt);
      k_sort_long<<<kLongCap, 512, kLongMax * 4, st>>>(ctx->idx32.as<uint32_t>(), ctx->seg_start.as<int64_t>(), long_list, d_stt);
      FF_CUDA(cudaEventRecord(ctx->ev[3], st));
      k_sort_cut<<<blocks_for(G * 32, 256), 256, 0, st>>>(ctx->idx32.as<uint32_t>(), ctx->seg_start.as<int64_t>(), G, db.d_targets, d_guides,
                                                         db.pack.cmp_mask, max_ot, ctx->st_targets.as<uint64_t>(), ctx->st_mm.as<uint8_t>(),
                                                         ctx->n_keep.as<int64_t>(), os.total_count.as<int32_t>(), os.overflowed.as<uint8_t>(), d_stt, cap);
      FF_CUDA(cudaMemsetAsync(ctx->n_keep.as<int64_t>() + G, 0, 8, st));
      FF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ctx->n_keep.as<int64_t>(), os.row_ptr.as<int64_t>(), G + 1, st));
      FF_TRY(ctx->cub_tmp.reserve(tmp_bytes));
      FF_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, ctx->n_keep.as<int64_t>(), os.row_ptr.as<int64_t>(), G + 1, st));
      k_compact_rows<<<blocks_for(G * 32, 256), 256, 0, st>>>(ctx->seg_start.as<int64_t>(), os.row_ptr.as<int64_t>(), G, ctx->st_targets.as<uint64_t>(),
                                                             ctx->st_mm.as<uint8_t>(), ctx->idx32.as<uint32_t>(), os.out_targets.as<uint64_t>(),
                                                             os.out_mm.as<uint8_t>(), os.out_tidx.as<uint32_t>(), d_stt);
      launches += 9;
      FF_CUDA(cudaEventRecord(ctx->ev[4], st));
    }
  *launches_out += launches;
  return FF_OK;
}

static int discover_plain(ff_ctx *ctx, const uint64_t *d_guides, int64_t n_guides, int max_mm, int max_ot,
                          bool want_positions, int slot, DeviceResult *res) {
  Database &db = ctx->db;
  ff_ctx::OutSlot &os = ctx->out[slot & 1];
  if (!db.resident) { set_error("no database resident in this context"); return FF_ENODB; }
  if (n_guides < 0 || max_mm < 0 || max_ot < 0 || (n_guides > 0 && !d_guides)) { set_error("bad discover argument"); return FF_EINVAL; }
  if (n_guides >= (1ll << 31)) { set_error("too many guides in one call"); return FF_EINVAL; }
  cudaStream_t st = ctx->stream;
  int launches = 0;
  ff_timings tm = {};

  const int64_t G = n_guides;
  const int64_t Gp = G > 0 ? G : 1;
  FF_TRY(ctx->counters.reserve(256));
  FF_TRY(ctx->seg_start.reserve((Gp + 1) * 8));
  FF_TRY(ctx->n_keep.reserve((Gp + 1) * 8));
  FF_TRY(os.row_ptr.reserve((Gp + 1) * 8));
  FF_TRY(os.total_count.reserve(Gp * 4));
  FF_TRY(os.overflowed.reserve(Gp));

  FF_CUDA(cudaEventRecord(ctx->ev[0], st));

  // ---- plan the two passes for this k
  ScanParams sp;
  int hA = 0, nA = 1, nB = 0;
  fill_scan_params(ctx, d_guides, G, max_mm, &sp, &hA, &nA, &nB);
  const int k_eff = sp.k, tbits = sp.tbits;

  // ---- candidate buffer: expected candidates per random guide = N x P(a random P-mer is within k) (116 at k = 4 on a
  // human-sized index); start with 1.4x that (+ slack), never more than 2^28 keys up front -- the call is repeated with
  // a larger buffer if it was not enough
  if (ctx->hit_cap == 0) ctx->hit_cap = 1u << 22;
  double expected_per_guide = 0.0;
  {
    double prob = 0.0, term = 1.0;  // term = C(P, i) 3^i
    for (int i = 0; i <= k_eff; ++i) {
      prob += term;
      term = term * 3.0 * (double)(db.proto_bases - i) / (double)(i + 1);
    }
    prob /= std::pow(4.0, (double)db.proto_bases);
    expected_per_guide = (double)db.n_targets * prob;
    const double per_guide = expected_per_guide * 1.4 + 64.0;
    size_t want = (size_t)std::min((double)G * per_guide, 268435456.0);
    if (want > ctx->hit_cap) ctx->hit_cap = want;
  }
  PlainStatus *d_stt = ctx->counters.as<PlainStatus>();
  PlainStatus *h_stt = static_cast<PlainStatus *>(ctx->h_status);
  PlainStatus *h_stt_dev = static_cast<PlainStatus *>(ctx->h_status_dev);
  sp.hit_count = &d_stt->n_cand; sp.n_compares = &d_stt->n_compares;
  int scan_launches = 0;
  const long long n_items = G * (long long)sp.items_per_guide;
  const int max_grid = ctx->sm_count * 8;  // 8 CTAs of 8 warps per SM
  const int grid = (int)std::max<long long>(1, std::min<long long>(max_grid, (n_items + kScanWarps - 1) / kScanWarps));
  // Guide-major or bin-major?  Bin-major pays when buckets are re-read (many guides) and the index does not fit in L2.
  bool bin_major = false;
  {
    const double reuse = (double)G * (double)nA / (double)(1ull << (2 * db.A.key_bases));
    bin_major = reuse >= 1.0 && (double)db.n_targets * 8.0 > 96e6;  // measured: ahead from 12 500 guides on 3e8 targets
    if (ctx->opt.scan_kernel == 1) bin_major = false;
    if (ctx->opt.scan_kernel == 2) bin_major = G > 0;
    bin_major = bin_major && bin_scan_supported(db, hA, G);
  }
  BinScanPlan bpl;
  if (bin_major) FF_TRY(bin_scan_prepare(ctx, sp, hA, nB, &d_stt->n_compares_b, &bpl, &launches));
  // Order the candidates with the per-guide pipeline (no host round trip) when segments are expected to be short.
  bool grouped = G > 0 && ctx->opt.group_sort != 0 && expected_per_guide <= 160.0;
  static bool long_attr[64] = {false};
  if (grouped && !long_attr[ctx->device & 63]) {
    FF_CUDA(cudaFuncSetAttribute(k_sort_long, cudaFuncAttributeMaxDynamicSharedMemorySize, kLongMax * 4));
    long_attr[ctx->device & 63] = true;
  }
  FF_CUDA(cudaEventRecord(ctx->ev[1], st));
  int64_t n_cand = 0, n_hits = 0;
  const uint64_t *sorted = nullptr;
  size_t tmp_bytes = 0;
  for (;;) {
    const size_t cap = ctx->hit_cap;
    FF_TRY(ctx->hit_keys.reserve(cap * 8));
    FF_TRY(os.out_targets.reserve((cap + 1) * 8));
    FF_TRY(os.out_mm.reserve(cap + 1));
    FF_TRY(os.out_tidx.reserve((cap + 1) * 4));
    sp.hits = ctx->hit_keys.as<uint64_t>(); sp.hit_cap = cap;
    FF_CUDA(cudaMemsetAsync(d_stt, 0, sizeof(PlainStatus), st));
    unsigned int *cnt = nullptr, *cursor = nullptr;
    if (grouped) {  // per-guide candidate counts [G + 1], scatter cursors [G], list of long segments
      FF_TRY(ctx->running.reserve((size_t)(Gp + 1) * 4 * 2 + kLongCap * 4));
      cnt = ctx->running.as<unsigned int>(); cursor = cnt + (Gp + 1);
      FF_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(Gp + 1) * 4 * 2, st));
    }
    if (G > 0) {
      if (bin_major) {
        if (cnt) FF_TRY(ctx->hit_ranks.reserve(cap * 4));
        FF_TRY(bin_scan_launch(ctx, &bpl, sp, cnt, &launches, cnt ? ctx->hit_ranks.as<uint32_t>() : nullptr));  // (counts and ranks the candidates per guide as it emits them)
      } else {
        k_seed_scan<<<grid, kScanThreads, 0, st>>>(sp);
        launches++;
      }
      scan_launches++;
    }
    FF_CUDA(cudaEventRecord(ctx->ev[2], st));
    if (grouped) FF_TRY(order_grouped(ctx, os, sp.hits, d_stt, cap, tbits, cnt, cursor, !bin_major, d_guides, G, max_ot, &launches,
                                      bin_major ? ctx->hit_ranks.as<uint32_t>() : nullptr));
    k_publish_status<<<1, 32, 0, st>>>(d_stt, nullptr, h_stt_dev, ++ctx->status_seq);
    FF_TRY(wait_status(ctx, st, ctx->status_seq));
    n_cand = (int64_t)h_stt->n_cand;
    if ((size_t)n_cand > cap) {  // the buffer was too small: grow it and repeat the call
      ctx->hit_cap = (size_t)(n_cand + n_cand / 8 + 1024);
      FF_CUDA(cudaEventRecord(ctx->ev[1], st));  // time only the run that counted
      continue;
    }
    if (grouped && h_stt->flag) grouped = false;  // a segment beyond k_sort_long: order these candidates with the radix sort
    if (grouped) n_hits = h_stt->n_hits;
    break;
  }

  if (!grouped) {
    // ---- order hits by (guide, database index) with a radix sort over the significant key bits
    int gbits = 1;
    while ((1ll << gbits) < Gp) gbits++;
    sorted = ctx->hit_keys.as<uint64_t>();
    if (n_cand > 0) {
      FF_TRY(ctx->hit_keys_sorted.reserve(ctx->hit_cap * 8));
      FF_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, ctx->hit_keys.as<uint64_t>(), ctx->hit_keys_sorted.as<uint64_t>(), n_cand, 0, tbits + gbits, st));
      FF_TRY(ctx->cub_tmp.reserve(tmp_bytes));
      FF_CUDA(cub::DeviceRadixSort::SortKeys(ctx->cub_tmp.p, tmp_bytes, ctx->hit_keys.as<uint64_t>(), ctx->hit_keys_sorted.as<uint64_t>(), n_cand, 0, tbits + gbits, st));
      sorted = ctx->hit_keys_sorted.as<uint64_t>();
      launches += 2 + (tbits + gbits + 7) / 8;
    }
    FF_CUDA(cudaEventRecord(ctx->ev[3], st));
    // ---- overflow cut in database order
    k_segments<<<blocks_for(G + 1, 256), 256, 0, st>>>(sorted, n_cand, G, tbits, ctx->seg_start.as<int64_t>());
    launches++;
    if (G > 0) {
      k_overflow_cut<<<blocks_for(G * 32, 256), 256, 0, st>>>(sorted, ctx->seg_start.as<int64_t>(), db.d_targets, G, max_ot, tbits,
                                                             ctx->n_keep.as<int64_t>(), os.total_count.as<int32_t>(), os.overflowed.as<uint8_t>());
      launches++;
    }
    FF_CUDA(cudaMemsetAsync(ctx->n_keep.as<int64_t>() + G, 0, 8, st));
    FF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ctx->n_keep.as<int64_t>(), os.row_ptr.as<int64_t>(), G + 1, st));
    FF_TRY(ctx->cub_tmp.reserve(tmp_bytes));
    FF_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, ctx->n_keep.as<int64_t>(), os.row_ptr.as<int64_t>(), G + 1, st));
    launches += 2;
    k_publish_status<<<1, 32, 0, st>>>(d_stt, os.row_ptr.as<int64_t>() + G, h_stt_dev, ++ctx->status_seq);
    FF_TRY(wait_status(ctx, st, ctx->status_seq));
    n_hits = h_stt->n_hits;
    if (G > 0 && n_hits > 0) {
      k_gather<<<blocks_for(G * 32, 256), 256, 0, st>>>(sorted, ctx->seg_start.as<int64_t>(), os.row_ptr.as<int64_t>(), db.d_targets, d_guides,
                                                       db.pack.cmp_mask, G, tbits, os.out_targets.as<uint64_t>(), os.out_mm.as<uint8_t>(), os.out_tidx.as<uint32_t>());
      launches++;
    }
  }
  int64_t n_pos = 0;
  FF_TRY(gather_positions(ctx, os, want_positions, n_hits, res, &n_pos, &launches));
  if (!grouped || want_positions) {
    FF_CUDA(cudaEventRecord(ctx->ev[4], st));
    FF_CUDA(cudaStreamSynchronize(st));
  }
  FF_CUDA(cudaGetLastError());

  FF_CUDA(cudaEventSynchronize(ctx->ev[4]));  // (already complete: the status words were published after it)
  FF_CUDA(cudaEventElapsedTime(&tm.prep_ms, ctx->ev[0], ctx->ev[1]));
  FF_CUDA(cudaEventElapsedTime(&tm.scan_ms, ctx->ev[1], ctx->ev[2]));
  FF_CUDA(cudaEventElapsedTime(&tm.order_ms, ctx->ev[2], ctx->ev[3]));
  FF_CUDA(cudaEventElapsedTime(&tm.cut_ms, ctx->ev[3], ctx->ev[4]));
  FF_CUDA(cudaEventElapsedTime(&tm.total_ms, ctx->ev[0], ctx->ev[4]));
  tm.score_ms = 0.f;
  tm.scan_launches = scan_launches;
  tm.kernel_launches = launches;
  // bytes the scan kernels request (NOT the roofline's algorithmic bytes, see bench.py): per (guide, seed) the two index
  // entries, per streamed entry its bit-sliced planes (or its 4-byte word in the guide-major kernel), per guide its
  // long, per candidate hit one 8-byte key
  const uint64_t ent_a = h_stt->n_compares, ent_b = h_stt->n_compares_b;
  const double bpe_a = bin_major ? db.A.plane_stride * 4.0 / 32.0 : 4.0, bpe_b = bin_major ? db.B.plane_stride * 4.0 / 32.0 : 4.0;
  tm.scan_bytes_read = (uint64_t)G * (uint64_t)(nA + nB) * 8ull + (uint64_t)(ent_a * bpe_a + ent_b * bpe_b) + (uint64_t)G * 8ull + (uint64_t)n_cand * 8ull;
  tm.entries_part1 = ent_a; tm.entries_part2 = ent_b;
  tm.scan_part1_ms = tm.scan_ms; tm.scan_part2_ms = 0.f;
  if (bin_major && nB > 0 && G > 0) {
    FF_CUDA(cudaEventElapsedTime(&tm.scan_part1_ms, ctx->ev[1], ctx->ev[7]));
    FF_CUDA(cudaEventElapsedTime(&tm.scan_part2_ms, ctx->ev[7], ctx->ev[2]));
  }
  ctx->last = tm;

  res->n_guides = G; res->n_hits = n_hits; res->n_positions = n_pos;
  res->n_candidate_hits = (uint64_t)n_cand; res->n_compares = ent_a + ent_b;
  res->d_row_ptr = os.row_ptr.as<int64_t>(); res->d_targets = os.out_targets.as<uint64_t>();
  res->d_mismatches = os.out_mm.as<uint8_t>(); res->d_total_count = os.total_count.as<int32_t>();
  res->d_overflowed = os.overflowed.as<uint8_t>(); res->d_bulge = nullptr;
  res->d_tidx = os.out_tidx.as<uint32_t>();
  return FF_OK;
}

#include "ff_shard.inl"
#include "ff_general.inl"

}  // namespace ff
