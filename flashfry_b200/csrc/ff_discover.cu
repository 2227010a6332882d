// Off-target discovery on the GPU: guide bucketing, the prefix-pruned bin-scan kernel, hit ordering and the
// database-order overflow cut.
//
// What it replaces (FlashFry, src/main/scala/...):
//   OrderedBinTraversalFactory precompute   reference/traversal/OrderedBinTraversalFactory.scala:146-173
//   SeekTraverser / LinearTraverser scan    reference/traverser/SeekTraverser.scala:78-102
//   BlockManager.compareIndexedBlock        reference/binary/blocks/BlockManager.scala:143-201
//   BlockManager.compareLinearBlock         reference/binary/blocks/BlockManager.scala:212-254
//   BitEncoding.mismatches                  bitcoding/BitEncoding.scala:127-132
//   ResultsAggregator.updateOT / addOT      crispr/ResultsAggregator.scala:61-69, crispr/CRISPRSiteOT.scala:39-46
//
// Not a translation.  The reference filters a guide list per 7-mer bin and per 11-mer sub-bin by *testing every
// guide against every prefix*; here prefixes are *enumerated*: a substitution of a base is an XOR of its 2-bit code
// with 1, 2 or 3, so the prefixes within d mismatches of a guide's prefix are {prefix ^ m : m in M_d} for a fixed
// mask table M sorted by distance.  A CTA owns one 7-mer bin of the database; it finds its guides by looking the
// bin's neighbours up in the guides' own 7-mer histogram, then for each (guide, remaining budget) enumerates the
// neighbouring (7+s)-mer sub-bins, looks their target range up in the resident sub-bin index and compares only
// the low word (the bases below the prefix) of those few targets.  Every (guide, target) pair within k mismatches
// is reached through exactly one (bin mask, sub mask) pair, so the hit set equals the reference's brute-force set.
// Hits are emitted as (guide, target index) keys, radix-sorted, and cut per guide in database order.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ff_common.cuh"
#include "ff_kernels.cuh"

namespace ff {

// ------------------------------------------------------------------------------------------------------------
// guide preparation
__global__ void k_guide_keys(const uint64_t *__restrict__ guides, int64_t n, int key_shift, uint32_t *__restrict__ keys,
                             uint64_t *__restrict__ entry) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t g = guides[i];
  keys[i] = (uint32_t)(g >> key_shift) & (kNumBins - 1);
  entry[i] = ((uint64_t)i << 32) | (uint32_t)g;  // guide index | low word (bases below the 7-mer + PAM)
}

// goff[b] = first sorted guide whose 7-mer key >= b  (b in [0, 4^7])
__global__ void k_guide_offsets(const uint32_t *__restrict__ sorted_keys, int64_t n, uint32_t *__restrict__ goff) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > kNumBins) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted_keys[mid] < (uint32_t)b) lo = mid + 1; else hi = mid;
  }
  goff[b] = (uint32_t)lo;
}

// ------------------------------------------------------------------------------------------------------------
// the bin-scan kernel
struct ScanParams {
  const uint32_t *tlow;
  const uint32_t *sub_off;
  const uint16_t *mask7;
  const uint16_t *submask;
  const uint64_t *gentry;   // guides sorted by 7-mer: (index << 32) | low word
  const uint32_t *goff;     // [4^7 + 1]
  uint64_t *hits;
  unsigned long long *hit_count;
  unsigned long long hit_cap;
  unsigned long long *n_compares;
  unsigned int *bin_cursor;
  int m7off[kPrefixBases + 2];
  int nsub[kMaxSubBases + 2];
  int s;                // sub-index bases
  int k;                // max mismatches
  int sub_shift;        // bit offset of the sub key inside the low word
  uint32_t rem_mask;    // compared bits below the sub key (low word)
  // staged kernel only
  const uint32_t *submask32;  // sub masks sorted by distance: mask | distance << 16
  int cap;                    // targets per staged segment
  int n_tiles;                // 32-mask tiles over the light distance classes heavy_classes..min(k,7)
  int heavy_classes;          // classes 0..heavy_classes-1 have >= 32 sub masks per guide
  int tile_start[kPrefixBases + 3];
};

constexpr int kScanThreads = 256;
constexpr int kListCap = 2048;
constexpr int kHitCap = 1024;

__device__ __forceinline__ int base_dist16(uint32_t m) {  // # non-zero 2-bit digits
  return __popc((m | (m >> 1)) & 0x55555555u);
}

__global__ void __launch_bounds__(kScanThreads) k_scan_direct(ScanParams p) {
  using BlockScan = cub::BlockScan<uint32_t, kScanThreads>;
  __shared__ typename BlockScan::TempStorage scan_tmp;
  __shared__ uint64_t s_list[kListCap];
  __shared__ uint64_t s_hits[kHitCap];
  __shared__ unsigned int s_hit_n;
  __shared__ unsigned long long s_hit_base;
  __shared__ int s_bin;

  const int tid = threadIdx.x;
  const uint32_t sub_key_mask = (1u << (2 * p.s)) - 1u;
  unsigned long long my_compares = 0;
  if (tid == 0) s_hit_n = 0;
  __syncthreads();

  for (;;) {
    if (tid == 0) s_bin = (int)atomicAdd(p.bin_cursor, 1u);
    __syncthreads();
    const int b = s_bin;
    if (b >= kNumBins) break;
    const uint32_t sub_base = (uint32_t)b << (2 * p.s);
    const bool bin_empty = p.sub_off[sub_base] == p.sub_off[sub_base + sub_key_mask + 1];
    const int dmax = bin_empty ? -1 : min(p.k, kPrefixBases);

    for (int d = 0; d <= dmax; ++d) {
      const int r = p.k - d;
      const int N = p.nsub[min(r, p.s)];
      const int m_end = p.m7off[d + 1];
      for (int m_base = p.m7off[d]; m_base < m_end; m_base += kScanThreads) {
        const int j = m_base + tid;
        uint32_t lo = 0, cnt = 0;
        if (j < m_end) {
          const uint32_t nb = (uint32_t)b ^ p.mask7[j];
          lo = p.goff[nb];
          cnt = p.goff[nb + 1] - lo;
        }
        uint32_t offs, total;
        BlockScan(scan_tmp).ExclusiveSum(cnt, offs, total);
        __syncthreads();
        for (uint32_t c = 0; c < total; c += kListCap) {
          // expand this tile's guide ranges into the shared list
          uint32_t q0 = max(offs, c), q1 = min(offs + cnt, c + (uint32_t)kListCap);
          for (uint32_t q = q0; q < q1; ++q) s_list[q - c] = p.gentry[lo + (q - offs)];
          __syncthreads();
          const uint32_t n = min((uint32_t)kListCap, total - c);
          const uint32_t items = n * (uint32_t)N;
          for (uint32_t item = tid; item < items; item += kScanThreads) {
            uint32_t e, i;
            if (N == 1) { e = item; i = 0; } else { e = item / (uint32_t)N; i = item - e * (uint32_t)N; }
            const uint64_t entry = s_list[e];
            const uint32_t glow = (uint32_t)entry;
            const uint32_t m = p.submask[i];
            const int rem = r - base_dist16(m);
            const uint32_t sub = ((glow >> p.sub_shift) & sub_key_mask) ^ m;
            const uint32_t t0 = p.sub_off[sub_base + sub], t1 = p.sub_off[sub_base + sub + 1];
            my_compares += t1 - t0;
            for (uint32_t t = t0; t < t1; ++t) {
              const uint32_t x = (p.tlow[t] ^ glow) & p.rem_mask;
              if (__popc((x | (x >> 1)) & 0x55555555u) <= rem) {
                const uint64_t key = (entry & 0xFFFFFFFF00000000ull) | t;
                const unsigned int slot = atomicAdd(&s_hit_n, 1u);
                if (slot < kHitCap) {
                  s_hits[slot] = key;
                } else {  // staging buffer full: straight to global
                  const unsigned long long gslot = atomicAdd(p.hit_count, 1ull);
                  if (gslot < p.hit_cap) p.hits[gslot] = key;
                }
              }
            }
          }
          __syncthreads();
          // flush staged hits when the buffer is at least half full
          if (s_hit_n >= kHitCap / 2) {
            const unsigned int nh = min(s_hit_n, (unsigned int)kHitCap);
            if (tid == 0) s_hit_base = atomicAdd(p.hit_count, (unsigned long long)nh);
            __syncthreads();
            for (unsigned int h = tid; h < nh; h += kScanThreads)
              if (s_hit_base + h < p.hit_cap) p.hits[s_hit_base + h] = s_hits[h];
            __syncthreads();
            if (tid == 0) s_hit_n = 0;
            __syncthreads();
          }
        }
      }
    }
    __syncthreads();  // s_bin is rewritten at the top of the loop
  }
  // final flush
  __syncthreads();
  {
    const unsigned int nh = min(s_hit_n, (unsigned int)kHitCap);
    if (nh > 0) {
      if (tid == 0) s_hit_base = atomicAdd(p.hit_count, (unsigned long long)nh);
      __syncthreads();
      for (unsigned int h = tid; h < nh; h += kScanThreads)
        if (s_hit_base + h < p.hit_cap) p.hits[s_hit_base + h] = s_hits[h];
    }
  }
  // one atomic per warp for the comparison counter
  for (int o = 16; o > 0; o >>= 1) my_compares += __shfl_down_sync(0xffffffffu, my_compares, o);
  if ((tid & 31) == 0 && my_compares) atomicAdd(p.n_compares, my_compares);
}

// ------------------------------------------------------------------------------------------------------------
// v2: the same enumeration, with the bin's low words and its slice of the sub-bin index staged in shared memory.
//
// One CTA owns one 7-mer bin at a time (dynamic bin cursor).  Thread 0 arms an mbarrier and issues ONE bulk async copy
// (cp.async.bulk global -> shared, the 1-D TMA path; SASS: UBLKCP) for the bin's low words while all threads stage the
// bin's sub-bin offsets as 16-bit segment-relative values.  Bins larger than the staging buffer are walked in several
// passes cut at sub-bin boundaries.  After that every lookup and every compare of the hot loop is an LDS.
// Work inside the bin is split into tiles of 32 first-level masks handed to WARPS through a shared counter; a warp
// expands its tile's guide ranges into a private list and enumerates (guide, sub mask) items on its own, so the main
// loop has no block-wide barrier.  Hits are staged per warp and flushed with one global atomic per flush.
constexpr int kStThreads = 512;
constexpr int kStWarps = kStThreads / 32;
constexpr int kLW = 32;  // list entries per warp
constexpr int kHW = 32;  // staged hits per warp

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct WarpHits {
  uint64_t *buf;          // this warp's staging slots (shared)
  unsigned int *count;    // this warp's counter (shared)
};

__device__ __forceinline__ void emit_hit(const ScanParams &p, const WarpHits &wh, uint64_t key) {
  const unsigned int slot = atomicAdd(wh.count, 1u);
  if (slot < kHW) {
    wh.buf[slot] = key;
  } else {
    const unsigned long long g = atomicAdd(p.hit_count, 1ull);
    if (g < p.hit_cap) p.hits[g] = key;
  }
}

__device__ __forceinline__ void flush_warp_hits(const ScanParams &p, const WarpHits &wh, int lane) {
  __syncwarp();
  const unsigned int nh = min(*wh.count, (unsigned int)kHW);
  if (nh == 0) return;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(p.hit_count, (unsigned long long)nh);
  base = __shfl_sync(0xffffffffu, base, 0);
  if ((unsigned int)lane < nh && base + lane < p.hit_cap) p.hits[base + lane] = wh.buf[lane];
  __syncwarp();
  if (lane == 0) *wh.count = 0;
  __syncwarp();
}

// One (guide entry, sub mask) item.  STAGED: look up and compare in shared memory; otherwise straight from global.
template <bool STAGED>
__device__ __forceinline__ void scan_item(const ScanParams &p, const WarpHits &wh, uint64_t entry, uint32_t sm, int r,
                                          uint32_t sub_key_mask, uint32_t sub_base, uint32_t sb, uint32_t nsb_pass,
                                          uint32_t t0a, const uint32_t *__restrict__ s_tlow, const uint16_t *__restrict__ s_sub,
                                          unsigned long long &compares) {
  const uint32_t glow = (uint32_t)entry;
  const int rem = r - (int)(sm >> 16);
  const uint32_t sub = ((glow >> p.sub_shift) & sub_key_mask) ^ (sm & 0xFFFFu);
  const uint32_t rel = sub - sb;
  if (rel >= nsb_pass) return;  // sub-bin handled by another pass of this bin
  uint32_t t0, t1;
  if (STAGED) {
    t0 = s_sub[rel];
    t1 = s_sub[rel + 1];
  } else {
    t0 = p.sub_off[sub_base + sub];
    t1 = p.sub_off[sub_base + sub + 1];
  }
  compares += t1 - t0;
  for (uint32_t t = t0; t < t1; ++t) {
    const uint32_t tl = STAGED ? s_tlow[t] : p.tlow[t];
    const uint32_t x = (tl ^ glow) & p.rem_mask;
    if (__popc((x | (x >> 1)) & 0x55555555u) <= rem)
      emit_hit(p, wh, (entry & 0xFFFFFFFF00000000ull) | (STAGED ? t0a + t : t));
  }
}

template <bool STAGED>
__device__ __forceinline__ void scan_pass_tiles(const ScanParams &p, const WarpHits &wh, int *s_tile, uint64_t *my_list, int b,
                                                uint32_t sub_key_mask, uint32_t sub_base, uint32_t sb, uint32_t nsb_pass,
                                                uint32_t t0a, const uint32_t *s_tlow, const uint16_t *s_sub, int lane, int warp,
                                                unsigned long long &compares) {
  // ---- phase A: "heavy" distance classes (>= 32 sub masks per guide).  Few first-level masks, a lot of work per
  // guide: every warp walks all of these masks and takes every kStWarps-th guide entry (static round-robin), the
  // lanes stride over the guide's sub masks.
  uint32_t pos = 0;  // running entry ordinal, identical in every warp
  for (int d = 0; d < p.heavy_classes; ++d) {
    const int r = p.k - d;
    const uint32_t N = (uint32_t)p.nsub[min(r, p.s)];
    const int m_end = p.m7off[d + 1];
    for (int jb = p.m7off[d]; jb < m_end; jb += 32) {
      uint32_t lo = 0, cnt = 0;
      if (jb + lane < m_end) {
        const uint32_t nb = (uint32_t)b ^ p.mask7[jb + lane];
        lo = p.goff[nb];
        cnt = p.goff[nb + 1] - lo;
      }
      const int nl = min(32, m_end - jb);
      for (int l = 0; l < nl; ++l) {
        const uint32_t lo_l = __shfl_sync(0xffffffffu, lo, l), cnt_l = __shfl_sync(0xffffffffu, cnt, l);
        for (uint32_t e = ((uint32_t)warp - pos) & (kStWarps - 1); e < cnt_l; e += kStWarps) {
          const uint64_t entry = p.gentry[lo_l + e];
          for (uint32_t i = lane; i < N; i += 32)
            scan_item<STAGED>(p, wh, entry, p.submask32[i], r, sub_key_mask, sub_base, sb, nsb_pass, t0a, s_tlow, s_sub, compares);
          __syncwarp();
          if (*wh.count >= kHW / 2) flush_warp_hits(p, wh, lane);
        }
        pos += cnt_l;
      }
    }
  }
  // ---- phase B: "light" classes (< 32 sub masks per guide, many first-level masks): tiles of 32 masks handed out
  // through a shared counter; a warp expands its tile's guide ranges into a private list and flattens guide x mask.
  for (;;) {
    int tile = 0;
    if (lane == 0) tile = atomicAdd(s_tile, 1);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= p.n_tiles) break;
    int d = p.heavy_classes;
    while (tile >= p.tile_start[d + 1]) ++d;
    const int j = p.m7off[d] + (tile - p.tile_start[d]) * 32 + lane;
    uint32_t lo = 0, cnt = 0;
    if (j < p.m7off[d + 1]) {
      const uint32_t nb = (uint32_t)b ^ p.mask7[j];
      lo = p.goff[nb];
      cnt = p.goff[nb + 1] - lo;
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;
    const uint32_t offs = incl - cnt;
    const int r = p.k - d;
    const uint32_t N = (uint32_t)p.nsub[min(r, p.s)];
    const float inv_n = 1.0f / (float)N;
    for (uint32_t c = 0; c < total; c += kLW) {
      const uint32_t q0 = max(offs, c), q1 = min(offs + cnt, c + (uint32_t)kLW);
      for (uint32_t q = q0; q < q1; ++q) my_list[q - c] = p.gentry[lo + (q - offs)];
      __syncwarp();
      const uint32_t n = min((uint32_t)kLW, total - c);
      const uint32_t items = n * N;
      for (uint32_t item = lane; item < items; item += 32) {
        const uint32_t e = (uint32_t)(((float)item + 0.5f) * inv_n);
        const uint32_t i = item - e * N;
        scan_item<STAGED>(p, wh, my_list[e], p.submask32[i], r, sub_key_mask, sub_base, sb, nsb_pass, t0a, s_tlow, s_sub, compares);
      }
      __syncwarp();
      if (*wh.count >= kHW / 2) flush_warp_hits(p, wh, lane);
    }
  }
}

__global__ void __launch_bounds__(kStThreads, 2) k_scan_staged(ScanParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_tile, s_bin;
  __shared__ unsigned int s_hitn[kStWarps];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t nsb = 1u << (2 * p.s);
  const uint32_t sub_key_mask = nsb - 1u;
  uint32_t *s_tlow = reinterpret_cast<uint32_t *>(smem);
  uint16_t *s_sub = reinterpret_cast<uint16_t *>(smem + (size_t)p.cap * 4);
  const size_t sub_bytes = ((size_t)(nsb + 1) * 2 + 15) & ~(size_t)15;
  uint64_t *s_list = reinterpret_cast<uint64_t *>(smem + (size_t)p.cap * 4 + sub_bytes);
  uint64_t *s_hit = s_list + kStWarps * kLW;
  WarpHits wh{s_hit + warp * kHW, &s_hitn[warp]};
  uint64_t *my_list = s_list + warp * kLW;

  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (lane == 0) s_hitn[warp] = 0;
  __syncthreads();

  uint32_t parity = 0;
  unsigned long long compares = 0;
  for (;;) {
    if (tid == 0) s_bin = (int)atomicAdd(p.bin_cursor, 1u);
    __syncthreads();
    const int b = s_bin;
    if (b >= kNumBins) break;
    const uint32_t sub_base = (uint32_t)b << (2 * p.s);
    uint32_t sb = 0;
    const uint32_t bin_t1 = p.sub_off[sub_base + nsb];
    while (sb < nsb) {
      // ---- choose the pass [sb, sb_end): as many whole sub-bins as fit the staging buffer
      const uint32_t seg_t0 = p.sub_off[sub_base + sb];
      if (seg_t0 == bin_t1) break;  // nothing left in this bin
      const uint32_t t0a = seg_t0 & ~3u;  // 16-byte aligned source for the bulk copy
      uint32_t sb_end = nsb;
      if (bin_t1 - t0a > (uint32_t)p.cap) {
        uint32_t lo = sb, hi = nsb;  // largest e with sub_off[e] - t0a <= cap
        while (lo < hi) {
          const uint32_t mid = (lo + hi + 1) >> 1;
          if (p.sub_off[sub_base + mid] - t0a <= (uint32_t)p.cap) lo = mid; else hi = mid - 1;
        }
        sb_end = lo;
      }
      const bool staged = sb_end > sb;
      if (!staged) sb_end = sb + 1;  // a single sub-bin larger than the buffer: compare it straight from global
      const uint32_t seg_t1 = p.sub_off[sub_base + sb_end];
      const uint32_t nsb_pass = sb_end - sb;
      __syncthreads();  // every warp is done with the previous pass's buffers (and has read s_bin)
      if (staged) {
        if (tid == 0) {
          const uint32_t bytes = (((seg_t1 - t0a) * 4u) + 15u) & ~15u;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of the old pass before the async write
          mbar_expect_tx(&s_bar, bytes);
          bulk_g2s(s_tlow, p.tlow + t0a, bytes, &s_bar);
        }
        for (uint32_t q = tid; q <= nsb_pass; q += kStThreads) s_sub[q] = (uint16_t)(p.sub_off[sub_base + sb + q] - t0a);
      }
      if (tid == 0) s_tile = 0;
      __syncthreads();
      if (staged) {
        mbar_wait(&s_bar, parity);
        parity ^= 1u;
        scan_pass_tiles<true>(p, wh, &s_tile, my_list, b, sub_key_mask, sub_base, sb, nsb_pass, t0a, s_tlow, s_sub, lane, warp, compares);
      } else {
        scan_pass_tiles<false>(p, wh, &s_tile, my_list, b, sub_key_mask, sub_base, sb, nsb_pass, 0u, nullptr, nullptr, lane, warp, compares);
      }
      sb = sb_end;
    }
    __syncthreads();
  }
  flush_warp_hits(p, wh, lane);
  for (int o = 16; o > 0; o >>= 1) compares += __shfl_down_sync(0xffffffffu, compares, o);
  if (lane == 0 && compares) atomicAdd(p.n_compares, compares);
}

// ------------------------------------------------------------------------------------------------------------
// ordering + overflow cut
// seg_start[g] = first sorted key whose guide index >= g   (g in [0, n_guides])
__global__ void k_segments(const uint64_t *__restrict__ keys, int64_t n_hits, int64_t n_guides, int64_t *__restrict__ seg_start) {
  int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g > n_guides) return;
  const uint64_t want = (uint64_t)g << 32;
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
                               const uint64_t *__restrict__ targets, int64_t n_guides, int max_ot,
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
    if (i < s1) c = (long long)(targets[(uint32_t)keys[i]] >> 48);
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
                         const uint64_t *__restrict__ guides, uint64_t cmp_mask, int64_t n_guides,
                         uint64_t *__restrict__ out_targets, uint8_t *__restrict__ out_mm, uint32_t *__restrict__ out_tidx) {
  const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_guides) return;
  const int64_t s0 = seg_start[g];
  const int64_t r0 = row_ptr[g], r1 = row_ptr[g + 1];
  const uint64_t guide = guides[g];
  for (int64_t i = lane; i < r1 - r0; i += 32) {
    const uint32_t t = (uint32_t)keys[s0 + i];
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
static inline unsigned int blocks_for(int64_t n, int threads) { return (unsigned int)((n + threads - 1) / threads); }

int discover_on_device(ff_ctx *ctx, const uint64_t *d_guides, int64_t n_guides, int max_mm, int max_ot,
                       bool want_positions, DeviceResult *res) {
  Database &db = ctx->db;
  if (!db.resident) { set_error("no database resident in this context"); return FF_ENODB; }
  if (n_guides < 0 || max_mm < 0 || max_ot < 0 || (n_guides > 0 && !d_guides)) { set_error("bad discover argument"); return FF_EINVAL; }
  if (db.pack.five_prime) { set_error("5'-PAM (Cpf1) databases are not supported by the GPU scan yet"); return FF_EUNSUPPORTED; }
  if (n_guides >= (1ll << 31)) { set_error("too many guides in one call"); return FF_EINVAL; }
  cudaStream_t st = ctx->stream;
  int launches = 0;
  ff_timings tm = {};

  const int64_t G = n_guides;
  const int64_t Gp = G > 0 ? G : 1;
  FF_TRY(ctx->gkeys.reserve(Gp * 4));
  FF_TRY(ctx->gkeys_sorted.reserve(Gp * 4));
  FF_TRY(ctx->gentry.reserve(Gp * 8));
  FF_TRY(ctx->gentry_sorted.reserve(Gp * 8));
  FF_TRY(ctx->goff.reserve((kNumBins + 2) * 4));
  FF_TRY(ctx->counters.reserve(64));
  FF_TRY(ctx->seg_start.reserve((Gp + 1) * 8));
  FF_TRY(ctx->n_keep.reserve((Gp + 1) * 8));
  FF_TRY(ctx->row_ptr.reserve((Gp + 1) * 8));
  FF_TRY(ctx->total_count.reserve(Gp * 4));
  FF_TRY(ctx->overflowed.reserve(Gp));

  FF_CUDA(cudaEventRecord(ctx->ev[0], st));
  // ---- guide bucketing by 7-mer prefix
  const int key_shift = 2 * (db.pack.scan_len - kPrefixBases);
  size_t tmp_bytes = 0;
  if (G > 0) {
    k_guide_keys<<<blocks_for(G, 256), 256, 0, st>>>(d_guides, G, key_shift, ctx->gkeys.as<uint32_t>(), ctx->gentry.as<uint64_t>());
    launches++;
    FF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx->gkeys.as<uint32_t>(), ctx->gkeys_sorted.as<uint32_t>(),
                                            ctx->gentry.as<uint64_t>(), ctx->gentry_sorted.as<uint64_t>(), G, 0, 2 * kPrefixBases, st));
    FF_TRY(ctx->cub_tmp.reserve(tmp_bytes));
    FF_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp_bytes, ctx->gkeys.as<uint32_t>(), ctx->gkeys_sorted.as<uint32_t>(),
                                            ctx->gentry.as<uint64_t>(), ctx->gentry_sorted.as<uint64_t>(), G, 0, 2 * kPrefixBases, st));
    launches += 3;
  }
  k_guide_offsets<<<blocks_for(kNumBins + 1, 256), 256, 0, st>>>(ctx->gkeys_sorted.as<uint32_t>(), G, ctx->goff.as<uint32_t>());
  launches++;
  FF_CUDA(cudaEventRecord(ctx->ev[1], st));

  // ---- scan (repeated once with a larger buffer if the hit buffer overflowed)
  if (ctx->hit_cap == 0) ctx->hit_cap = 1u << 22;
  {
    size_t want = (size_t)G * 192;  // ~116 expected hits per random guide at k=4 on a human-sized index
    if (want > ctx->hit_cap) ctx->hit_cap = want;
  }
  ScanParams sp;
  sp.tlow = db.d_tlow; sp.sub_off = db.d_sub_off; sp.mask7 = db.d_mask7; sp.submask = db.d_submask;
  sp.gentry = ctx->gentry_sorted.as<uint64_t>(); sp.goff = ctx->goff.as<uint32_t>();
  for (int i = 0; i < kPrefixBases + 2; ++i) sp.m7off[i] = db.m7off[i];
  for (int i = 0; i < kMaxSubBases + 2; ++i) sp.nsub[i] = db.nsub[i];
  sp.s = db.sub_bases; sp.k = max_mm;
  sp.sub_shift = 2 * (db.pack.scan_len - kPrefixBases - db.sub_bases);
  sp.rem_mask = (uint32_t)(db.pack.cmp_mask & ((1ull << sp.sub_shift) - 1ull));
  // staged kernel: tiles of 32 first-level masks per distance class, staging capacity from the shared-memory budget
  sp.submask32 = db.d_submask32;
  {
    const int dmax = max_mm < kPrefixBases ? max_mm : kPrefixBases;
    int dh = 0;
    while (dh <= dmax && db.nsub[std::min(max_mm - dh, db.sub_bases)] >= 32) ++dh;
    sp.heavy_classes = dh;
    for (int d = 0; d <= kPrefixBases + 2; ++d) sp.tile_start[d] = 0;
    for (int d = dh; d <= kPrefixBases + 1; ++d) {
      const int n_masks = d <= dmax ? db.m7off[d + 1] - db.m7off[d] : 0;
      sp.tile_start[d + 1] = sp.tile_start[d] + (n_masks + 31) / 32;
    }
    sp.n_tiles = sp.tile_start[dmax + 1];
  }
  const size_t sub_bytes = ((((size_t)1 << (2 * db.sub_bases)) + 1) * 2 + 15) & ~(size_t)15;
  const size_t fixed_bytes = sub_bytes + (size_t)kStWarps * (kLW + kHW) * 8;
  const size_t smem_budget = 113 * 1024;  // two CTAs per SM
  sp.cap = (int)(((smem_budget - fixed_bytes) / 4) & ~(size_t)3);
  const size_t staged_smem = (size_t)sp.cap * 4 + fixed_bytes;
  bool staged_mode = true;
  if (const char *e = getenv("FF_SCAN_MODE")) staged_mode = strcmp(e, "direct") != 0;
  if (staged_mode) FF_CUDA(cudaFuncSetAttribute(k_scan_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem));
  unsigned long long *d_cnt = ctx->counters.as<unsigned long long>();  // [0] hits [1] compares [2] bin cursor
  sp.hit_count = d_cnt; sp.n_compares = d_cnt + 1; sp.bin_cursor = reinterpret_cast<unsigned int *>(d_cnt + 2);
  unsigned long long h_cnt[3] = {0, 0, 0};
  int scan_launches = 0;
  const int grid = ctx->sm_count * 4;
  for (;;) {
    FF_TRY(ctx->hit_keys.reserve(ctx->hit_cap * 8));
    FF_TRY(ctx->hit_keys_sorted.reserve(ctx->hit_cap * 8));
    sp.hits = ctx->hit_keys.as<uint64_t>(); sp.hit_cap = ctx->hit_cap;
    FF_CUDA(cudaMemsetAsync(d_cnt, 0, 32, st));
    if (G > 0) {
      if (staged_mode) k_scan_staged<<<2 * ctx->sm_count, kStThreads, staged_smem, st>>>(sp);
      else k_scan_direct<<<grid, kScanThreads, 0, st>>>(sp);
      launches++;
      scan_launches++;
    }
    FF_CUDA(cudaEventRecord(ctx->ev[2], st));
    FF_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, 24, cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    if (h_cnt[0] <= ctx->hit_cap) break;
    ctx->hit_cap = (size_t)(h_cnt[0] + h_cnt[0] / 8 + 1024);
    FF_CUDA(cudaEventRecord(ctx->ev[1], st));  // time only the run that counted
  }
  const int64_t n_cand = (int64_t)h_cnt[0];

  // ---- order hits by (guide, database index)
  int gbits = 1;
  while ((1ll << gbits) < Gp) gbits++;
  const uint64_t *sorted = ctx->hit_keys.as<uint64_t>();
  if (n_cand > 0) {
    FF_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, ctx->hit_keys.as<uint64_t>(), ctx->hit_keys_sorted.as<uint64_t>(), n_cand, 0, 32 + gbits, st));
    FF_TRY(ctx->cub_tmp.reserve(tmp_bytes));
    FF_CUDA(cub::DeviceRadixSort::SortKeys(ctx->cub_tmp.p, tmp_bytes, ctx->hit_keys.as<uint64_t>(), ctx->hit_keys_sorted.as<uint64_t>(), n_cand, 0, 32 + gbits, st));
    sorted = ctx->hit_keys_sorted.as<uint64_t>();
    launches += 2 + (32 + gbits + 7) / 8;
  }
  FF_CUDA(cudaEventRecord(ctx->ev[3], st));

  // ---- overflow cut in database order
  k_segments<<<blocks_for(G + 1, 256), 256, 0, st>>>(sorted, n_cand, G, ctx->seg_start.as<int64_t>());
  launches++;
  if (G > 0) {
    k_overflow_cut<<<blocks_for(G * 32, 256), 256, 0, st>>>(sorted, ctx->seg_start.as<int64_t>(), db.d_targets, G, max_ot,
                                                           ctx->n_keep.as<int64_t>(), ctx->total_count.as<int32_t>(), ctx->overflowed.as<uint8_t>());
    launches++;
  }
  FF_CUDA(cudaMemsetAsync(ctx->n_keep.as<int64_t>() + G, 0, 8, st));
  FF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ctx->n_keep.as<int64_t>(), ctx->row_ptr.as<int64_t>(), G + 1, st));
  FF_TRY(ctx->cub_tmp.reserve(tmp_bytes));
  FF_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, ctx->n_keep.as<int64_t>(), ctx->row_ptr.as<int64_t>(), G + 1, st));
  launches += 2;
  int64_t n_hits = 0;
  FF_CUDA(cudaMemcpyAsync(&n_hits, ctx->row_ptr.as<int64_t>() + G, 8, cudaMemcpyDeviceToHost, st));
  FF_CUDA(cudaStreamSynchronize(st));
  const int64_t Hp = n_hits > 0 ? n_hits : 1;
  FF_TRY(ctx->out_targets.reserve(Hp * 8));
  FF_TRY(ctx->out_mm.reserve(Hp));
  FF_TRY(ctx->out_tidx.reserve(Hp * 4));
  if (G > 0 && n_hits > 0) {
    k_gather<<<blocks_for(G * 32, 256), 256, 0, st>>>(sorted, ctx->seg_start.as<int64_t>(), ctx->row_ptr.as<int64_t>(), db.d_targets, d_guides,
                                                     db.pack.cmp_mask, G, ctx->out_targets.as<uint64_t>(), ctx->out_mm.as<uint8_t>(), ctx->out_tidx.as<uint32_t>());
    launches++;
  }
  int64_t n_pos = 0;
  res->d_pos_ptr = nullptr; res->d_positions = nullptr;
  if (want_positions && db.d_positions) {
    FF_TRY(ctx->pos_cnt.reserve((Hp + 1) * 8));
    FF_TRY(ctx->pos_ptr.reserve((Hp + 1) * 8));
    FF_CUDA(cudaMemsetAsync(ctx->pos_cnt.p, 0, (Hp + 1) * 8, st));
    if (n_hits > 0) {
      k_pos_counts<<<blocks_for(n_hits, 256), 256, 0, st>>>(ctx->out_targets.as<uint64_t>(), n_hits, ctx->pos_cnt.as<int64_t>());
      launches++;
    }
    FF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ctx->pos_cnt.as<int64_t>(), ctx->pos_ptr.as<int64_t>(), n_hits + 1, st));
    FF_TRY(ctx->cub_tmp.reserve(tmp_bytes));
    FF_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, ctx->pos_cnt.as<int64_t>(), ctx->pos_ptr.as<int64_t>(), n_hits + 1, st));
    launches += 2;
    FF_CUDA(cudaMemcpyAsync(&n_pos, ctx->pos_ptr.as<int64_t>() + n_hits, 8, cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    FF_TRY(ctx->out_positions.reserve((n_pos > 0 ? n_pos : 1) * 8));
    if (n_hits > 0) {
      k_gather_positions<<<blocks_for(n_hits * 32, 256), 256, 0, st>>>(ctx->out_tidx.as<uint32_t>(), ctx->pos_ptr.as<int64_t>(), db.d_pos_off,
                                                                      db.d_positions, n_hits, ctx->out_positions.as<uint64_t>());
      launches++;
    }
    res->d_pos_ptr = ctx->pos_ptr.as<int64_t>();
    res->d_positions = ctx->out_positions.as<uint64_t>();
  }
  FF_CUDA(cudaEventRecord(ctx->ev[4], st));
  FF_CUDA(cudaStreamSynchronize(st));
  FF_CUDA(cudaGetLastError());

  FF_CUDA(cudaEventElapsedTime(&tm.prep_ms, ctx->ev[0], ctx->ev[1]));
  FF_CUDA(cudaEventElapsedTime(&tm.scan_ms, ctx->ev[1], ctx->ev[2]));
  FF_CUDA(cudaEventElapsedTime(&tm.order_ms, ctx->ev[2], ctx->ev[3]));
  FF_CUDA(cudaEventElapsedTime(&tm.cut_ms, ctx->ev[3], ctx->ev[4]));
  FF_CUDA(cudaEventElapsedTime(&tm.total_ms, ctx->ev[0], ctx->ev[4]));
  tm.score_ms = 0.f;
  tm.scan_launches = scan_launches;
  tm.kernel_launches = launches;
  // algorithmic bytes of the scan: every low word + every sub-index entry once, the guide
  // entries once, one 8-byte key per candidate hit
  tm.scan_bytes_read = db.n_targets * 4ull + ((1ull << (2 * (kPrefixBases + db.sub_bases))) + 1) * 4ull + (uint64_t)G * 8ull + (uint64_t)n_cand * 8ull;
  ctx->last = tm;

  res->n_guides = G; res->n_hits = n_hits; res->n_positions = n_pos;
  res->n_candidate_hits = (uint64_t)n_cand; res->n_compares = h_cnt[1];
  res->d_row_ptr = ctx->row_ptr.as<int64_t>(); res->d_targets = ctx->out_targets.as<uint64_t>();
  res->d_mismatches = ctx->out_mm.as<uint8_t>(); res->d_total_count = ctx->total_count.as<int32_t>();
  res->d_overflowed = ctx->overflowed.as<uint8_t>();
  return FF_OK;
}

}  // namespace ff
