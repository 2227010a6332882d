// Bin-major, bit-sliced seed scan (included by ff_discover.cu, inside namespace ff) -- the large-batch scan.
//
// What it replaces in the reference: the loop nest bin -> sub-bin -> target x guide of
// reference/binary/blocks/BlockManager.scala:143-254 driven by OrderedBinTraversalFactory.scala:146-177 (same order of
// the outer loop -- bins of the first seven bases in AAAAAAA..TTTTTTT order -- nothing else in common).
//
// The same (guide, seed) pairs as k_seed_scan, organised so that neither HBM nor the instruction issue slots are spent
// on anything but the compares:
//   * COMPARE.  `other` (the protospacer part a seed does not key on) is stored BIT-SLICED: entries in groups of 32,
//     one 32-bit word per bit plane.  One lane compares its pair's probe against 32 entries at once: per base two LOP3
//     (mismatch plane = (lo ^ Gl) | (hi ^ Gh), Gl / Gh = the probe's bits spread to full words, per lane), then a
//     carry-save adder tree over the 9..11 mismatch planes and a 4-bit carry chain for "count <= budget": ~40 LOP3 per
//     32 entries instead of XOR + fold + POPC + min per entry (5 x 32) -- the POPC pipe (16 lanes / clk / SM), which
//     bounded the previous kernel's ceiling, is not used at all.
//   * PART ONE (index A, short buckets: k_bin_scan).  A CTA claims one BIN -- all keys that share their first
//     key_bases - 4 bases (7 bases = FlashFry's own bin width for the 20-mers) -- and stages the bin's slice of the
//     bit-sliced array (~41 KB on a human-sized index) in shared memory with one TMA bulk copy (cp.async.bulk +
//     mbarrier).  The pairs that land in the bin are ENUMERATED, not sorted: guides are listed by the bin of their own
//     key; a bin F is reached from the guide classes F ^ m, m over the bin-part masks within the seed budget, and each
//     such guide contributes the masks over the last four key bases within the remaining budget.  Lanes take one pair
//     each and stream its bucket (2..5 groups) out of shared memory.  Every byte of the index is read from HBM once.
//   * PART TWO (index B, long buckets: k_pair_scan).  The few pairs (28 per guide at k = 4) are counting-sorted by
//     bucket; lanes take consecutive sorted pairs, so the ~10 pairs of a bucket sit in neighbouring lanes, read the
//     same addresses in the same instruction (one L1 wavefront) and every bucket comes from HBM once.
//   * HITS.  A lane that finds hits pushes (hit word, group, guide) into a per-warp queue in shared memory; the warp
//     drains the queue with all lanes when it is half full: one global atomic per ~32 hit words.

#define FF_FA(a, b, c, s, cy) { const uint32_t _x = (a) ^ (b); s = _x ^ (c); cy = ((a) & (b)) | (_x & (c)); }

// bit-sliced population count of NB one-bit planes -> 4-bit count (c3 c2 c1 c0)
template <int NB> struct PlaneCount;
template <> struct PlaneCount<9> {
  static __device__ __forceinline__ void run(const uint32_t (&m)[9], uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3) {
    uint32_t s1, k1, s2, k2, s3, k3, k4, t, k5;
    FF_FA(m[0], m[1], m[2], s1, k1) FF_FA(m[3], m[4], m[5], s2, k2) FF_FA(m[6], m[7], m[8], s3, k3)
    FF_FA(s1, s2, s3, c0, k4)
    FF_FA(k1, k2, k3, t, k5)
    c1 = t ^ k4;
    const uint32_t k6 = t & k4;
    c2 = k5 ^ k6; c3 = k5 & k6;
  }
};
template <> struct PlaneCount<10> {
  static __device__ __forceinline__ void run(const uint32_t (&m)[10], uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3) {
    uint32_t s1, k1, s2, k2, s3, k3, u, k4, t, k5, k6;
    FF_FA(m[0], m[1], m[2], s1, k1) FF_FA(m[3], m[4], m[5], s2, k2) FF_FA(m[6], m[7], m[8], s3, k3)
    FF_FA(s1, s2, s3, u, k4)
    c0 = u ^ m[9];
    const uint32_t k7 = u & m[9];
    FF_FA(k1, k2, k3, t, k5)
    FF_FA(t, k4, k7, c1, k6)
    c2 = k5 ^ k6; c3 = k5 & k6;
  }
};
template <> struct PlaneCount<11> {
  static __device__ __forceinline__ void run(const uint32_t (&m)[11], uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3) {
    uint32_t s1, k1, s2, k2, s3, k3, u, k4, k7, t, k5, k6;
    FF_FA(m[0], m[1], m[2], s1, k1) FF_FA(m[3], m[4], m[5], s2, k2) FF_FA(m[6], m[7], m[8], s3, k3)
    FF_FA(s1, s2, s3, u, k4)
    FF_FA(u, m[9], m[10], c0, k7)
    FF_FA(k1, k2, k3, t, k5)
    FF_FA(t, k4, k7, c1, k6)
    c2 = k5 ^ k6; c3 = k5 & k6;
  }
};

// What a lane holds for its pair: the probe's bits spread to words, and 15 - budget for the carry chain.
template <int NB> struct LaneProbe {
  uint32_t gl[NB], gh[NB], kb[4];
  // Bit k of x spread to a word: shift it to the top of its byte (eight shifted copies of x serve every k) and let PRMT
  // replicate that byte's sign -- 8 + one instruction per bit instead of two per bit.
  __device__ __forceinline__ void set(uint32_t probe, int budget) {
    const uint32_t kk = 15u - (uint32_t)min(max(budget, 0), 15);
    const uint32_t x = probe | (kk << (2 * NB));  // 2 NB probe bits, then the four bits of 15 - budget
    uint32_t y[8];
#pragma unroll
    for (int sft = 0; sft < 8; ++sft) y[sft] = x << sft;
    auto spread = [&](int k) {  // (prmt.b32 directly: __byte_perm ignores the sign-replication bit of the selector nibbles)
      uint32_t d;
      asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(y[7 - (k & 7)]), "r"(0u), "r"(0x8888u | (0x1111u * (uint32_t)(k >> 3))));
      return d;
    };
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      gl[i] = spread(2 * i);
      gh[i] = spread(2 * i + 1);
      asm volatile("" : "+r"(gl[i]), "+r"(gh[i]));  // keep them in registers: ptxas otherwise recomputes them inside the group loop
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      kb[b] = spread(2 * NB + b);
      asm volatile("" : "+r"(kb[b]));
    }
  }
  // bit e set <=> entry e of the group is within budget.  w = the group's plane words (w[2i] low bit, w[2i+1] high bit of base i).
  // lo_d >= 0 (part two, warp-uniform): and has MORE than lo_d mismatches -- the pairs with d1 <= hA belong to part one.
  __device__ __forceinline__ uint32_t match(const uint32_t (&w)[2 * NB], int lo_d = -1) const {
    uint32_t m[NB];
#pragma unroll
    // (tried: w ^ gl as w * (gl | 1) + gl, an IMAD on the FMA pipe, to take a quarter of the LOP3s off the ALU pipe: 2-5 % slower)
    for (int i = 0; i < NB; ++i) m[i] = (w[2 * i] ^ gl[i]) | (w[2 * i + 1] ^ gh[i]);
    uint32_t c0, c1, c2, c3;
    PlaneCount<NB>::run(m, c0, c1, c2, c3);
    // count + (15 - budget) carries out of four bits <=> count > budget
    uint32_t cy = c0 & kb[0];
    cy = (c1 & kb[1]) | (cy & (c1 | kb[1]));
    cy = (c2 & kb[2]) | (cy & (c2 | kb[2]));
    cy = (c3 & kb[3]) | (cy & (c3 | kb[3]));
    if (lo_d < 0) return ~cy;
    const uint32_t kl = 15u - (uint32_t)min(lo_d, 15);  // the same chain against lo_d
    const uint32_t l0 = 0u - (kl & 1u), l1 = 0u - ((kl >> 1) & 1u), l2 = 0u - ((kl >> 2) & 1u), l3 = 0u - ((kl >> 3) & 1u);
    uint32_t cl = c0 & l0;
    cl = (c1 & l1) | (cl & (c1 | l1));
    cl = (c2 & l2) | (cl & (c2 | l2));
    cl = (c3 & l3) | (cl & (c3 | l3));
    return ~cy & cl;
  }
};

// Where candidates go.  world == 1: this context's buffer.  world > 1 (database-sharded discover, section 7 of DESIGN.md):
// every rank scans ITS part of the index for ALL guides and pushes each candidate straight into the exchange block of
// the rank that OWNS the guide -- peer memory over NVLink.  The owner's block has one REGION per source rank, so a
// position needs no remote round trip: the source counts what it sent to every owner with LOCAL atomics (`sent`), the
// keys travel as fire-and-forget P2P stores, and one remote store per owner at the end of the scan tells it how many keys
// its region holds (k_peer_counts).  The hand-over of the candidates is fused into the scan.
struct HitSink {
  uint64_t *hits;
  unsigned long long *hit_count;
  unsigned long long hit_cap;         // world > 1: keys per REGION (block capacity / world)
  int tbits;
  unsigned int *gcnt;  // per-guide candidate counts for the ordering that follows (nullptr: not wanted)
  uint32_t *ranks;     // with gcnt: the rank of every candidate among its guide's (the value gcnt had), so that the ordering
                       // places candidates without a second round of atomics (nullptr: not wanted)
  int world, rank;                    // ranks of the exchange (world <= 1: no exchange)
  float owner_scale;                  // world / all guides
  unsigned int first[kMaxPeers + 1];  // first guide owned by every rank (ff_shard_range), first[world] = all guides
  uint8_t *peer[kMaxPeers];          // exchange block of every rank: PeerCtr, candidate keys (world regions), totals
};

struct PeerCtr {  // head of an exchange block (kPeerHeadBytes)
  unsigned int arrive;               // barrier arrivals, monotonically increasing (remote atomics)
  unsigned int pad;
  unsigned int sent[kMaxPeers];      // LOCAL: keys this rank has pushed to every owner in the current step
  unsigned int recv[kMaxPeers];      // written by the sources: keys in each region of this block
};
constexpr size_t kPeerHead = kPeerHeadBytes;

__device__ __forceinline__ int sink_owner(const HitSink &hs, uint32_t g) {
  int o = min((int)((float)g * hs.owner_scale), hs.world - 1);  // first[r] = n r / world: the estimate is off by one at most,
  while (g >= hs.first[o + 1]) ++o;                             // more only with fewer guides than ranks
  while (g < hs.first[o]) --o;
  return o;
}

// All shared memory of the two kernels is addressed as byte offsets into this one array: a generic pointer to shared
// memory costs an S2R + LEA (the shared window base) at every use inside the group loop.
extern __shared__ __align__(128) uint8_t ff_smem[];

constexpr int kQCap = 64;  // hit-queue entries per warp; drained at >= 32, and one iteration adds at most 32

// Queue entries (uint4): x = hit word (already cut to the bucket's range), y = index of the entry bit 0 stands for, z = guide index, w = probe (part
// two: for the exact d1 > hA test).  The number of queued entries lives in a warp-uniform register.
template <bool PASS_B>
__device__ __forceinline__ void drain_queue(const HitSink &hs, uint32_t q_off, uint32_t n, int lane, const uint32_t *__restrict__ canon,
                                            const uint32_t *__restrict__ other, int lo_d) {
  __syncwarp();
  const uint4 *q = reinterpret_cast<const uint4 *>(ff_smem + q_off);
  for (uint32_t i0 = 0; i0 < n; i0 += 32) {
    uint4 e = make_uint4(0u, 0u, 0u, 0u);
    if (i0 + lane < n) e = q[i0 + lane];
    uint32_t vm = e.x;  // (part two: the compare itself has dropped the entries with d1 <= hA)
    const int c = __popc(vm);
    if (hs.world > 1) {  // push to the owners' exchange blocks (region hs.rank of each): no remote round trip
      const int own = c ? sink_owner(hs, e.z) : -1;
      unsigned int my_total = 0;  // lane o: what this warp sends to owner o
      for (int o = 0; o < hs.world; ++o) {
        const unsigned int t = __reduce_add_sync(0xffffffffu, own == o ? (unsigned int)c : 0u);
        if (lane == o) my_total = t;
      }
      unsigned int base_l = 0;
      if (my_total) base_l = atomicAdd(&reinterpret_cast<PeerCtr *>(hs.peer[hs.rank])->sent[lane], my_total);  // (local memory)
      unsigned int pref = 0;
      for (int o = 0; o < hs.world; ++o) {
        if (!__ballot_sync(0xffffffffu, own == o)) continue;
        const unsigned int mine = own == o ? (unsigned int)c : 0u;
        unsigned int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned int v = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += v;
        }
        if (own == o) pref = incl - mine;
      }
      unsigned long long pos = (unsigned long long)__shfl_sync(0xffffffffu, base_l, own >= 0 ? own : 0) + pref;
      if (c) {
        uint64_t *dst = reinterpret_cast<uint64_t *>(hs.peer[own] + kPeerHead) + (size_t)hs.rank * hs.hit_cap;
        const uint64_t gk = (uint64_t)(e.z - hs.first[own]) << hs.tbits;  // the owner numbers its guides from 0
        while (vm) {
          const int b = __ffs((int)vm) - 1;
          vm &= vm - 1u;
          const uint32_t idx = e.y + (uint32_t)b;
          if (pos < hs.hit_cap) dst[pos] = gk | (canon ? canon[idx] : idx);
          ++pos;
        }
      }
      continue;
    }
    unsigned int r0 = 0;
    if (c && hs.gcnt) {
      if (hs.ranks) r0 = atomicAdd(hs.gcnt + e.z, (unsigned int)c);  // (in flight together with the position atomic below)
      else atomicAdd(hs.gcnt + e.z, (unsigned int)c);
    }
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;
    unsigned long long base = 0;
    if (lane == 31) base = atomicAdd(hs.hit_count, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    unsigned long long pos = base + (unsigned long long)(incl - c);
    const uint64_t gk = (uint64_t)e.z << hs.tbits;
    while (vm) {
      const int b = __ffs((int)vm) - 1;
      vm &= vm - 1u;
      const uint32_t idx = e.y + (uint32_t)b;
      if (pos < hs.hit_cap) {
        hs.hits[pos] = gk | (canon ? canon[idx] : idx);
        if (hs.ranks) hs.ranks[pos] = r0;
      }
      ++pos; ++r0;
    }
  }
  __syncwarp();
}

// The hit word of the lane's group `it` (of n + 1): cut it to the bucket's range -- the first group's entries below
// lo_bit and the last group's entries from hi_bit on belong to the neighbours (or are padding) --, queue it with the index
// of its entry 0, drain the queue when half full.
template <bool PASS_B>
__device__ __forceinline__ void queue_hits(uint32_t hm, int it, int n, uint32_t lo_bit, uint32_t hi_bit, uint32_t idx0, uint32_t gid,
                                           uint32_t probe, const HitSink &hs, uint32_t q_off, uint32_t &qcount, int lane, uint32_t lt_mask,
                                           const uint32_t *canon, const uint32_t *other, int lo_d) {
  if (it > n) hm = 0u;
  if (hm) {
    if (it == 0) hm &= 0xFFFFFFFFu << lo_bit;
    if (it == n) hm &= 0xFFFFFFFFu >> (32u - hi_bit);
  }
  const uint32_t hb = __ballot_sync(0xffffffffu, hm != 0u);
  if (hb) {
    if (hm) *reinterpret_cast<uint4 *>(ff_smem + q_off + 16u * (qcount + (uint32_t)__popc(hb & lt_mask))) = make_uint4(hm, idx0 + 32u * (uint32_t)it, gid, probe);
    qcount += (uint32_t)__popc(hb);
    if (qcount >= 32u) {
      drain_queue<PASS_B>(hs, q_off, qcount, lane, canon, other, lo_d);
      qcount = 0;
    }
  }
}

// Stream the lane's n + 1 groups starting at group g0 (n = -1: idle lane).  Entry 0 of group g0 has index idx0; the
// bucket owns the first group's entries from lo_bit on and the last group's entries below hi_bit.  SMEM: group `g_base`
// sits at byte offset `base_off` of ff_smem; else at gbase[0].  The loop is warp-uniform (longest bucket of the warp);
// lanes past their own bucket load nothing and drop their result.
template <int NB, int STRIDE, bool PASS_B, bool SMEM>
__device__ __forceinline__ void stream_groups(uint32_t base_off, const uint32_t *gbase, uint32_t g_base, uint32_t g0, int n, uint32_t idx0,
                                              uint32_t lo_bit, uint32_t hi_bit, const LaneProbe<NB> &lp, uint32_t gid, uint32_t probe,
                                              const HitSink &hs, uint32_t q_off, uint32_t &qcount, int lane, const uint32_t *canon,
                                              const uint32_t *other, int lo_d) {
  const int T = __reduce_max_sync(0xffffffffu, n);
  uint32_t soff = base_off + (g0 - g_base) * (uint32_t)(STRIDE * 4);
  const uint32_t *pg = SMEM ? nullptr : gbase + (size_t)(n >= 0 ? g0 - g_base : 0u) * STRIDE;
  uint32_t w[2 * NB];
#pragma unroll
  for (int j = 0; j < 2 * NB; ++j) w[j] = 0u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int it = 0; it <= T; ++it) {
    if (it <= n) {  // lanes past their bucket keep the old words: no loads, no bank conflicts
      if (SMEM) {
        if (STRIDE % 4 == 0) {
#pragma unroll
          for (int j = 0; j < (2 * NB + 3) / 4; ++j) {
            const uint4 v = *reinterpret_cast<const uint4 *>(ff_smem + soff + 16 * j);
            w[4 * j] = v.x; w[4 * j + 1] = v.y;
            if (4 * j + 2 < 2 * NB) { w[4 * j + 2] = v.z; w[4 * j + 3] = v.w; }
          }
        } else {
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const uint2 v = *reinterpret_cast<const uint2 *>(ff_smem + soff + 8 * j);
            w[2 * j] = v.x; w[2 * j + 1] = v.y;
          }
        }
        soff += STRIDE * 4;
      } else {
        if (it + 2 <= n) {  // long part-two buckets: pull the group after next towards the SM
          asm volatile("prefetch.global.L1 [%0];" ::"l"(pg + 2 * STRIDE));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(pg + 2 * STRIDE + STRIDE - 1));
        }
        if (STRIDE % 4 == 0) {
#pragma unroll
          for (int j = 0; j < (2 * NB + 3) / 4; ++j) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(pg + 4 * j));
            w[4 * j] = v.x; w[4 * j + 1] = v.y;
            if (4 * j + 2 < 2 * NB) { w[4 * j + 2] = v.z; w[4 * j + 3] = v.w; }
          }
        } else {
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(pg + 2 * j));
            w[2 * j] = v.x; w[2 * j + 1] = v.y;
          }
        }
        pg += STRIDE;
      }
    }
    queue_hits<PASS_B>(lp.match(w, PASS_B ? lo_d : -1), it, n, lo_bit, hi_bit, idx0, gid, probe, hs, q_off, qcount, lane, lt_mask, canon, other, lo_d);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Shared memory by 32-bit shared-window address.  The kernels below use cluster-scoped instructions (TMA bulk copies),
// for which ptxas rebuilds the window base (S2UR SR_CgaCtaId + UMOV + ULEA) at EVERY generic access to ff_smem -- three
// issue slots per group in the inner loop.  Addresses formed once per round and used through ld/st.shared avoid that.
__device__ __forceinline__ uint32_t atom_add_shared(uint32_t addr, uint32_t v) {  // (inline PTX also keeps nvcc from wrapping a lane-0
  uint32_t r;                                                                   //  atomic in its 15-instruction warp aggregation)
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(addr), "r"(v) : "memory");
  return r;
}

// Load the 2 NB plane words of the group at shared address `a` if r >= 0 (else keep the old words): predicated loads,
// no branch.  STRIDE % 4 == 0: 16-byte loads (the group is 16-byte aligned), else 8-byte loads.
template <int NB, int STRIDE> struct GroupLoad;
template <> struct GroupLoad<9, 18> {
  static __device__ __forceinline__ void run(uint32_t (&w)[18], uint32_t a, int r) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %19, 0;\n\t"
        "@p ld.shared.v2.u32 {%0, %1}, [%18];\n\t@p ld.shared.v2.u32 {%2, %3}, [%18+8];\n\t@p ld.shared.v2.u32 {%4, %5}, [%18+16];\n\t"
        "@p ld.shared.v2.u32 {%6, %7}, [%18+24];\n\t@p ld.shared.v2.u32 {%8, %9}, [%18+32];\n\t@p ld.shared.v2.u32 {%10, %11}, [%18+40];\n\t"
        "@p ld.shared.v2.u32 {%12, %13}, [%18+48];\n\t@p ld.shared.v2.u32 {%14, %15}, [%18+56];\n\t@p ld.shared.v2.u32 {%16, %17}, [%18+64];\n\t}"
        : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]), "+r"(w[9]), "+r"(w[10]),
          "+r"(w[11]), "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15]), "+r"(w[16]), "+r"(w[17])
        : "r"(a), "r"(r));
  }
};
template <> struct GroupLoad<10, 20> {
  static __device__ __forceinline__ void run(uint32_t (&w)[20], uint32_t a, int r) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %21, 0;\n\t"
        "@p ld.shared.v4.u32 {%0, %1, %2, %3}, [%20];\n\t@p ld.shared.v4.u32 {%4, %5, %6, %7}, [%20+16];\n\t"
        "@p ld.shared.v4.u32 {%8, %9, %10, %11}, [%20+32];\n\t@p ld.shared.v4.u32 {%12, %13, %14, %15}, [%20+48];\n\t"
        "@p ld.shared.v4.u32 {%16, %17, %18, %19}, [%20+64];\n\t}"
        : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]), "+r"(w[9]), "+r"(w[10]),
          "+r"(w[11]), "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15]), "+r"(w[16]), "+r"(w[17]), "+r"(w[18]), "+r"(w[19])
        : "r"(a), "r"(r));
  }
};
template <> struct GroupLoad<11, 24> {
  static __device__ __forceinline__ void run(uint32_t (&w)[22], uint32_t a, int r) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %23, 0;\n\t"
        "@p ld.shared.v4.u32 {%0, %1, %2, %3}, [%22];\n\t@p ld.shared.v4.u32 {%4, %5, %6, %7}, [%22+16];\n\t"
        "@p ld.shared.v4.u32 {%8, %9, %10, %11}, [%22+32];\n\t@p ld.shared.v4.u32 {%12, %13, %14, %15}, [%22+48];\n\t"
        "@p ld.shared.v4.u32 {%16, %17, %18, %19}, [%22+64];\n\t@p ld.shared.v2.u32 {%20, %21}, [%22+80];\n\t}"
        : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]), "+r"(w[9]), "+r"(w[10]),
          "+r"(w[11]), "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15]), "+r"(w[16]), "+r"(w[17]), "+r"(w[18]), "+r"(w[19]), "+r"(w[20]),
          "+r"(w[21])
        : "r"(a), "r"(r));
  }
};

// The shared-memory inner loop of both scan kernels.  The lane streams its n + 1 groups (n = -1: idle lane, and then
// last_mask must be 0) starting at shared address `sa`; entry 0 of the first group has index idx0; first_mask / last_mask
// cut the first / last group to the bucket's range (FIRST_CUT = false: buckets start at a group boundary).  The loop is
// warp-uniform (longest bucket of the warp).  A lane with a non-zero hit word queues (word, index of entry 0, guide,
// probe) at the warp's queue (shared address qa, generic offset q_off); the queue is drained when >= 32 entries wait.
template <int NB, int STRIDE, bool PASS_B, bool FIRST_CUT>
__device__ __forceinline__ void stream_smem(uint32_t sa, int n, uint32_t idx0, uint32_t first_mask, uint32_t last_mask, const LaneProbe<NB> &lp,
                                            uint32_t gid, uint32_t probe, const HitSink &hs, uint32_t qa, uint32_t q_off, uint32_t &qcount,
                                            int lane, const uint32_t *canon, const uint32_t *other, int lo_d) {
  const int T = __reduce_max_sync(0xffffffffu, n);
  uint32_t w[2 * NB];
#pragma unroll
  for (int j = 0; j < 2 * NB; ++j) w[j] = 0u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  int r = n;  // groups left after this one
  for (int it = 0; it <= T; ++it) {
    GroupLoad<NB, STRIDE>::run(w, sa, r);
    uint32_t cut = r > 0 ? 0xFFFFFFFFu : last_mask;  // r < 0: last_mask has been cleared
    if (r <= 0) last_mask = 0u;
    if (FIRST_CUT) { cut &= first_mask; first_mask = 0xFFFFFFFFu; }
    const uint32_t hm = lp.match(w, PASS_B ? lo_d : -1) & cut;
    const uint32_t hb = __ballot_sync(0xffffffffu, hm != 0u);
    if (hb) {
      if (hm)
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(qa + 16u * (qcount + (uint32_t)__popc(hb & lt_mask))), "r"(hm), "r"(idx0),
                     "r"(gid), "r"(probe) : "memory");
      qcount += (uint32_t)__popc(hb);
      if (qcount >= 32u) {
        drain_queue<PASS_B>(hs, q_off, qcount, lane, canon, other, lo_d);
        qcount = 0;
      }
    }
    sa += STRIDE * 4; idx0 += 32u; --r;
  }
}

// ------------------------------------------------------------------------------------------------------------
// part one: bins of index A through shared memory
#ifndef FF_BIN_THREADS
#define FF_BIN_THREADS 384
#endif
constexpr int kBinThreads = FF_BIN_THREADS;
constexpr int kBinWarps = kBinThreads / 32;
constexpr int kNbrCap = 1160;       // bin-part masks within the seed budget (7 bases, distance <= 3: 1156)
#ifndef FF_BIN_SLICE
#define FF_BIN_SLICE 800
#endif
#ifndef FF_BIN_PAIRCAP
#define FF_BIN_PAIRCAP 2016
#endif
#ifndef FF_BIN_VISITS
#define FF_BIN_VISITS 1536
#endif
constexpr int kSliceGroups = FF_BIN_SLICE;   // groups of 32 entries a CTA can stage (x 72 B = 56 KB; a human-sized bin is ~717 bucket-aligned groups)
constexpr int kPairCap = FF_BIN_PAIRCAP;     // pairs sorted by bucket at a time (a human-sized bin with 100 000 guides has ~3200: two chunks)
constexpr int kVisitCap = FF_BIN_VISITS;     // (class, guide) visits of a bin listed in shared memory (~1300); more: binary search

struct BinParams {
  const uint32_t *planes, *off, *goff, *canon, *himasks, *lomasks;  // goff: first group of every bucket (bucket-aligned planes)
  int n_hi;            // bin-part masks within the budget
  int cum_hi[5];       // cum_hi[d] = # bin-part masks at distance <= d (d <= hA <= 3)
  int nm[4];           // nm[d] = # last-four-bases masks a guide reached at bin distance d contributes
  uint32_t rcp[4];     // floor(2^32 / nm[d])
  int hA, k;
  uint32_t n_bins;     // scan bins [bin_base, n_bins): this rank's part of the index (all of it unless database-sharded)
  uint32_t bin_base;
  const uint2 *sg;     // guides listed by the bin of their own key: x = probe | (last four key bases) << 24, y = guide index
  const int *cls_off;  // [n_bins + 1]
  HitSink hs;
  unsigned long long *n_compares;
  unsigned int *next_bin;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct BinShared {
  unsigned long long mbar;
  uint32_t bin, b0, b1, glob, g_base, bytes, use_vis, next_slice, early, pad0;
  uint32_t R[6], PB[6];
  uint32_t wtot[kBinWarps];
  uint32_t off[260], goff[260];  // entries / groups: first of every bucket of the bin
  uint32_t lom[256];
  uint32_t pre[kNbrCap + 4];
  uint32_t start[kNbrCap + 4];
  uint32_t vis[kVisitCap];
  uint4 q[kBinWarps][kQCap];
  // the chunk of pairs being processed, counting-sorted by bucket so that neighbouring lanes stream the same groups
  uint32_t cnt[256], bstart[260];
  uint2 rec[kPairCap];      // x = probe | bucket << 24, y = guide index | budget << 28
  uint16_t rank[kPairCap], perm[kPairCap];
};

template <int NB>
__global__ void __launch_bounds__(kBinThreads, 2) k_bin_scan(BinParams bp) {
  constexpr int STRIDE = 2 * NB;
  constexpr uint32_t kShOff = (uint32_t)kSliceGroups * STRIDE * 4;
  BinShared &sh = *reinterpret_cast<BinShared *>(ff_smem + kShOff);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&sh.mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 256) sh.lom[tid] = bp.lomasks[tid];
  uint32_t phase = 0;
  unsigned long long compares = 0;
  const uint32_t q_off = kShOff + (uint32_t)offsetof(BinShared, q) + (uint32_t)warp * kQCap * 16u;
  const uint32_t sbase = smem_addr(ff_smem), claim_addr = smem_addr(&sh.next_slice);
  uint32_t qcount = 0;
  uint32_t claimed = 0;  // thread 0: the bin after the current one, claimed a whole bin ahead (its atomic is never waited for)
  if (tid == 0) claimed = bp.bin_base + atomicAdd(bp.next_bin, 1u);
  for (;;) {
    __syncthreads();  // the previous bin is finished: its slice and tables may be overwritten
    if (tid == 0) sh.bin = claimed;
    __syncthreads();
    const uint32_t bin = sh.bin;
    if (bin >= bp.n_bins) break;
    if (tid == 0) claimed = bp.bin_base + atomicAdd(bp.next_bin, 1u);
    if (tid == kBinThreads - 1) {  // (a thread that has no class to size below) start the copy of the bin's slice NOW when it fits the
      const uint32_t gs = bp.goff[bin << 8] & ~1u, ge = bp.goff[(bin << 8) + 256];  // buffer: it lands during the set-up phases
      const bool fits = ge - gs <= (uint32_t)kSliceGroups;
      sh.early = fits ? 1u : 0u;
      if (fits) {
        const uint32_t bytes = ((ge - gs) * STRIDE * 4u + 15u) & ~15u;
        sh.b1 = 256; sh.g_base = gs; sh.glob = 0u; sh.bytes = bytes;
        if (bytes) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&sh.mbar)), "r"(bytes) : "memory");
          const uint8_t *src = reinterpret_cast<const uint8_t *>(bp.planes + (size_t)gs * STRIDE);
          for (uint32_t o = 0; o < bytes; o += 32768u) {
            const uint32_t nb = min(32768u, bytes - o);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_addr(ff_smem + o)), "l"(src + o), "r"(nb), "r"(smem_addr(&sh.mbar)) : "memory");
          }
        }
      }
    }
    for (int i = tid; i <= 256; i += kBinThreads) { sh.off[i] = bp.off[(bin << 8) + i]; sh.goff[i] = bp.goff[(bin << 8) + i]; }
    // guide classes that reach this bin, and the exclusive prefix of their sizes ("visits" of the bin)
    uint32_t carry = 0;
    for (int j0 = 0; j0 < bp.n_hi; j0 += kBinThreads) {
      const int j = j0 + tid;
      uint32_t cnt = 0;
      if (j < bp.n_hi) {
        const uint32_t t = bin ^ (bp.himasks[j] & 0xFFFFFFu);
        const int a0 = bp.cls_off[t];
        cnt = (uint32_t)(bp.cls_off[t + 1] - a0);
        sh.start[j] = (uint32_t)a0;
      }
      uint32_t incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) sh.wtot[warp] = incl;
      __syncthreads();
      uint32_t before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < kBinWarps; ++w) {
        const uint32_t v = sh.wtot[w];
        if (w < warp) before += v;
        total += v;
      }
      if (j < bp.n_hi) sh.pre[j] = carry + before + incl - cnt;
      carry += total;
      __syncthreads();
    }
    if (tid == 0) {  // visits and pairs by bin distance
      sh.pre[bp.n_hi] = carry;
      sh.b0 = 0;
      sh.use_vis = carry <= (uint32_t)kVisitCap ? 1u : 0u;
      uint32_t pb = 0, prev = 0;
      sh.PB[0] = 0; sh.R[0] = 0;
      for (int d = 0; d <= bp.hA; ++d) {
        const uint32_t r = sh.pre[bp.cum_hi[d]];
        pb += (r - prev) * (uint32_t)bp.nm[d];
        sh.R[d + 1] = r; sh.PB[d + 1] = pb;
        prev = r;
      }
    }
    __syncthreads();
    const bool use_vis = sh.use_vis != 0;
    if (use_vis)  // list the visits: the guide (position in sg) of every (class, guide) combination, in class order
      for (int j = tid; j < bp.n_hi; j += kBinThreads) {
        const uint32_t p0 = sh.pre[j], c = sh.pre[j + 1] - p0, a0 = sh.start[j];
        for (uint32_t i = 0; i < c; ++i) sh.vis[p0 + i] = a0 + i;
      }
    // the bin's buckets in one pass when its slice fits the staging buffer, else in runs of buckets; a single bucket
    // larger than the buffer is streamed from global memory
    for (;;) {
      __syncthreads();
      if (tid == 0 && sh.early) {
        sh.early = 0u;  // the whole bin is one run and its copy is already under way
      } else if (tid == 0) {
        const uint32_t b0 = sh.b0;
        const uint32_t gs = sh.goff[b0] & ~1u;  // even group: 16-byte aligned source for any stride that is a multiple of 2 words
        uint32_t b1 = 256;
        if (sh.goff[256] - gs > (uint32_t)kSliceGroups) {
          b1 = b0 + 1;
          while (b1 < 256 && sh.goff[b1 + 1] - gs <= (uint32_t)kSliceGroups) ++b1;
        }
        const uint32_t ng = sh.goff[b1] - gs;
        sh.b1 = b1; sh.g_base = gs;
        sh.glob = ng > (uint32_t)kSliceGroups ? 1u : 0u;
        const uint32_t bytes = sh.glob ? 0u : ((ng * STRIDE * 4u + 15u) & ~15u);
        sh.bytes = bytes;
        if (bytes) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&sh.mbar)), "r"(bytes) : "memory");
          const uint8_t *src = reinterpret_cast<const uint8_t *>(bp.planes + (size_t)gs * STRIDE);
          for (uint32_t o = 0; o < bytes; o += 32768u) {
            const uint32_t nb = min(32768u, bytes - o);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_addr(ff_smem + o)), "l"(src + o), "r"(nb), "r"(smem_addr(&sh.mbar)) : "memory");
          }
        }
      }
      __syncthreads();
      const uint32_t b0 = sh.b0, b1 = sh.b1, g_base = sh.g_base;
      const bool glob = sh.glob != 0;
      bool waited = sh.bytes == 0;
      const uint32_t n_pairs = sh.PB[bp.hA + 1];
      // chunks of equal size (a human-sized bin with 100 000 guides: 3216 pairs -> 2 x 1608, not 2016 + 1200)
      const uint32_t n_chunks = (n_pairs + kPairCap - 1) / kPairCap, chunk = n_chunks ? (n_pairs + n_chunks - 1) / n_chunks : 0;
      for (uint32_t c0 = 0; c0 < n_pairs; c0 += chunk) {
        const uint32_t cn = min(chunk, n_pairs - c0);
        if (tid < 256) sh.cnt[tid] = 0;
        __syncthreads();
        if (tid == 0) sh.next_slice = 0;  // (every warp has left the previous chunk's claim loop)
        // enumerate the chunk's pairs; rank every pair inside its bucket
        for (uint32_t i = tid; i < cn; i += kBinThreads) {
          const uint32_t p = c0 + i;
          int d = 0;
          while (p >= sh.PB[d + 1]) ++d;
          const uint32_t qd = p - sh.PB[d], nmd = (uint32_t)bp.nm[d];
          uint32_t vq = __umulhi(qd, bp.rcp[d]), s = qd - vq * nmd;  // rcp rounds down: the quotient is exact or one short
          if (s >= nmd) { ++vq; s -= nmd; }
          const uint32_t v = sh.R[d] + vq;
          uint32_t sidx;
          if (use_vis) {
            sidx = sh.vis[v];
          } else {
            int jl = d ? bp.cum_hi[d - 1] : 0, jh = bp.cum_hi[d];
            while (jh - jl > 1) {
              const int mid = (jl + jh) >> 1;
              if (sh.pre[mid] <= v) jl = mid; else jh = mid;
            }
            sidx = sh.start[jl] + (v - sh.pre[jl]);
          }
          const uint2 rec = bp.sg[sidx];
          const uint32_t lm = sh.lom[s];
          const uint32_t bl = (rec.x >> 24) ^ (lm & 0xFFu);
          uint16_t rk = 0xFFFFu;
          if (bl >= b0 && bl < b1 && sh.off[bl + 1] > sh.off[bl]) {
            rk = (uint16_t)atomicAdd(&sh.cnt[bl], 1u);
            const int budget = min(bp.k - d - (int)(lm >> 8), 15);
            sh.rec[i] = make_uint2((rec.x & 0xFFFFFFu) | (bl << 24), rec.y | ((uint32_t)budget << 28));
          }
          sh.rank[i] = rk;
        }
        __syncthreads();
        if (warp == 0) {  // exclusive prefix over the 256 bucket counts
          uint32_t c[8], sum = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) { c[j] = sh.cnt[lane * 8 + j]; sum += c[j]; }
          uint32_t incl = sum;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          uint32_t acc = incl - sum;
#pragma unroll
          for (int j = 0; j < 8; ++j) { sh.bstart[lane * 8 + j] = acc; acc += c[j]; }
          if (lane == 31) sh.bstart[256] = acc;
        }
        __syncthreads();
        for (uint32_t i = tid; i < cn; i += kBinThreads) {
          const uint32_t rk = sh.rank[i];
          if (rk != 0xFFFFu) sh.perm[sh.bstart[sh.rec[i].x >> 24] + rk] = (uint16_t)i;
        }
        if (!waited) {  // the slice has had the whole enumeration to arrive
          uint32_t done = 0;
          while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_addr(&sh.mbar)), "r"(phase) : "memory");
          }
          phase ^= 1u;
          waited = true;
        }
        __syncthreads();
        const uint32_t nv = sh.bstart[256];
        for (;;) {  // warps take 32 sorted pairs at a time
          uint32_t j0 = 0;
          if (lane == 0) j0 = atom_add_shared(claim_addr, 32u);
          j0 = __shfl_sync(0xffffffffu, j0, 0);
          if (j0 >= nv) break;
          const uint32_t j = j0 + lane;
          uint32_t lo = 0, len = 0, g0 = g_base, gid = 0, probe = 0;
          int budget = -1;
          if (j < nv) {
            const uint2 r = sh.rec[sh.perm[j]];
            const uint32_t bl = r.x >> 24;
            lo = sh.off[bl]; len = sh.off[bl + 1] - lo; g0 = sh.goff[bl];
            probe = r.x & 0xFFFFFFu; gid = r.y & 0x0FFFFFFFu; budget = (int)(r.y >> 28);
          }
          compares += len;
          LaneProbe<NB> lp;
          lp.set(probe, budget);
          const int n = len ? (int)((len - 1u) >> 5) : -1;           // every bucket starts a group: ceil(len / 32) groups,
          const uint32_t hi_bit = ((len - 1u) & 31u) + 1u;           // only the last one is cut
          if (glob) stream_groups<NB, STRIDE, false, false>(0u, bp.planes + (size_t)g_base * STRIDE, g_base, g0, n, lo, 0u, hi_bit, lp, gid, probe, bp.hs, q_off, qcount, lane, bp.canon, nullptr, -1);
          else stream_smem<NB, STRIDE, false, false>(sbase + (g0 - g_base) * (uint32_t)(STRIDE * 4), n, lo, 0xFFFFFFFFu, len ? 0xFFFFFFFFu >> (32u - hi_bit) : 0u, lp,
                                                     gid, probe, bp.hs, sbase + q_off, q_off, qcount, lane, bp.canon, nullptr, -1);
        }
      }
      if (!waited) {  // no pairs at all: still consume the copy before the buffer is reused
        uint32_t done = 0;
        while (!done) {
          asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                       : "=r"(done) : "r"(smem_addr(&sh.mbar)), "r"(phase) : "memory");
        }
        phase ^= 1u;
      }
      if (b1 >= 256) break;
      __syncthreads();
      if (tid == 0) sh.b0 = b1;
    }
  }
  drain_queue<false>(bp.hs, q_off, qcount, lane, bp.canon, nullptr, -1);
  if (bp.hs.world > 1) __threadfence_system();  // candidates pushed to peers are visible before this rank reaches the barrier
  for (int o = 16; o > 0; o >>= 1) compares += __shfl_down_sync(0xffffffffu, compares, o);
  if (lane == 0 && compares) atomicAdd(bp.n_compares, compares);
}

// ------------------------------------------------------------------------------------------------------------
// part two: (guide, seed) pairs of index B, counting-sorted by bucket
struct PairParams {
  const uint32_t *planes, *off, *other, *canon;
  const uint4 *recs;   // x, y = the bucket's entry range [lo, hi), z = probe | budget << 24, w = guide index; sorted by bucket
  const uint32_t *start;  // [n_keys + 1] first sorted pair of every bucket
  int rb_shift;        // k_pair_scan2: a "B-bin" = 2^rb_shift consecutive buckets
  uint32_t n_bbins;
  int seg_shift;       // k_pair_scan2: a round of 32 pairs is cut into 2^seg_shift work items (consecutive group ranges)
  uint32_t bbin_base;  // k_pair_scan2 scans B-bins [bbin_base, n_bbins)
  const unsigned int *n_pairs_dev;  // database-sharded: only this rank's buckets have pairs; their number is start[n_keys]
  long long n_pairs;
  int lo_d;
  HitSink hs;
  unsigned long long *n_compares;
  unsigned long long *next_item;
};

// one thread per (guide, seed): the bucket the pair lands in
__device__ __forceinline__ void b_pair(const uint64_t *guides, const uint32_t *masks, int n_seeds, int proto_shift, uint64_t proto_mask, int b_bits,
                                       int k, long long idx, uint32_t *kk, uint32_t *pb, uint32_t *gid) {
  const long long g = idx / n_seeds;
  const int j = (int)(idx - g * n_seeds);
  const uint64_t proto = (guides[g] >> proto_shift) & proto_mask;
  const uint32_t m = masks[j];
  *kk = (uint32_t)(proto & ((1ull << b_bits) - 1ull)) ^ (m & 0xFFFFFFu);
  *pb = (uint32_t)(proto >> b_bits) | ((uint32_t)(k - (int)(m >> 24)) << 24);
  *gid = (uint32_t)g;
}

__global__ void k_bpairs_hist(const uint64_t *__restrict__ guides, long long n_pairs, const uint32_t *__restrict__ masks, int n_seeds,
                              int proto_shift, uint64_t proto_mask, int b_bits, int k, uint32_t b_lo, uint32_t b_hi, unsigned int *__restrict__ cnt) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n_pairs) return;
  uint32_t kk, pb, gid;
  b_pair(guides, masks, n_seeds, proto_shift, proto_mask, b_bits, k, idx, &kk, &pb, &gid);
  if (kk - b_lo < b_hi - b_lo) atomicAdd(cnt + kk, 1u);  // (buckets outside [b_lo, b_hi) belong to other ranks)
}

__global__ void k_bpairs_scatter(const uint64_t *__restrict__ guides, long long n_pairs, const uint32_t *__restrict__ masks, int n_seeds,
                                 int proto_shift, uint64_t proto_mask, int b_bits, int k, uint32_t b_lo, uint32_t b_hi,
                                 const unsigned int *__restrict__ start, unsigned int *__restrict__ cursor, const uint32_t *__restrict__ off,
                                 uint4 *__restrict__ recs) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n_pairs) return;
  uint32_t kk, pb, gid;
  b_pair(guides, masks, n_seeds, proto_shift, proto_mask, b_bits, k, idx, &kk, &pb, &gid);
  if (kk - b_lo < b_hi - b_lo) recs[start[kk] + atomicAdd(cursor + kk, 1u)] = make_uint4(off[kk], off[kk + 1], pb, gid);
}

// 3 CTAs per SM (80 registers).  Tried on the GPU and dropped: a register double buffer of the next group (0.99 ms), two
// groups per iteration with their loads issued together (0.90 ms; both cost a third of the resident warps), a per-lane
// cp.async ring in shared memory (2.2 ms: LDGSTS does not merge the lanes that read the same address), the item's ~4
// buckets staged in a per-warp shared-memory region by one TMA bulk copy each (0.88 ms: 15 KB per warp leave 12 warps
// per SM, and items with a fifth bucket fall back to global loads) -- against 0.73 ms.
#ifndef FF_PAIR_MIN_BLOCKS
#define FF_PAIR_MIN_BLOCKS 3
#endif
#ifndef FF_PAIR_THREADS
#define FF_PAIR_THREADS 256
#endif
constexpr int kPairThreads = FF_PAIR_THREADS;
constexpr int kPairWarps = kPairThreads / 32;

template <int NB>
__global__ void __launch_bounds__(kPairThreads, FF_PAIR_MIN_BLOCKS) k_pair_scan(PairParams pp) {
  constexpr int STRIDE = (2 * NB + 3) & ~3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t q_off = (uint32_t)warp * kQCap * 16u;  // dynamic shared memory: one hit queue per warp
  uint32_t qcount = 0;
  unsigned long long compares = 0;
  const long long n_pairs = pp.n_pairs_dev ? (long long)*pp.n_pairs_dev : pp.n_pairs;
  const long long n_items = (n_pairs + 31) / 32;
  for (;;) {  // items are claimed in bucket order: the whole grid walks the index front to back
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(pp.next_item, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if ((long long)item >= n_items) break;
    const long long pi = (long long)item * 32 + lane;
    uint32_t lo = 0, hi = 0, gid = 0, probe = 0;
    int budget = -1;
    if (pi < n_pairs) {
      const uint4 r = pp.recs[pi];
      lo = r.x; hi = r.y;
      probe = r.z & 0xFFFFFFu; budget = (int)(r.z >> 24); gid = r.w;
    }
    compares += hi - lo;
    LaneProbe<NB> lp;
    lp.set(probe, budget);
    const int n = hi > lo ? (int)(((hi - 1u) >> 5) - (lo >> 5)) : -1;  // groups are cut at multiples of 32 entries
    stream_groups<NB, STRIDE, true, false>(0u, pp.planes, 0u, lo >> 5, n, lo & ~31u, lo & 31u, ((hi - 1u) & 31u) + 1u, lp, gid, probe, pp.hs,
                                           q_off, qcount, lane, pp.canon, pp.other, pp.lo_d);
  }
  drain_queue<true>(pp.hs, q_off, qcount, lane, pp.canon, pp.other, pp.lo_d);
  if (pp.hs.world > 1) __threadfence_system();
  for (int o = 16; o > 0; o >>= 1) compares += __shfl_down_sync(0xffffffffu, compares, o);
  if (lane == 0 && compares) atomicAdd(pp.n_compares, compares);
}

// Part two through shared memory.  With many guides every bucket of index B is visited (10.7 pairs per bucket at 100 000
// guides), so the pass is a join of the sorted pair list with the whole index.  "B-bin" = 2^rb_shift consecutive buckets,
// in index order; its run of bit-sliced groups and its slice of the sorted pair list are staged with TMA bulk copies.
// A CTA owns a ring of TWO staging buffers and NO CTA-wide barrier: every warp walks the CTA's bins in sequence, waits
// for the bin's "full" mbarrier, claims work items -- 32 consecutive sorted pairs x one of 2^seg_shift consecutive
// group ranges of their buckets (a bucket is ~36 groups) -- and, when it finds none left, signs off the bin with an atomic
// on its "left" counter and moves on to the next bin in the other buffer.  The LAST warp to leave bin i refills that
// buffer with bin i + 2 (claims it from the global counter, reads its offsets, starts the copies): a split barrier.  Warps
// never wait for each other unless they are a whole bin apart.  Index B comes from HBM exactly once; the group loop
// reads shared memory (broadcast: neighbouring lanes hold pairs of the same bucket).  A bin whose run does not fit the
// buffer streams from global memory like k_pair_scan.
#ifndef FF_P2_GROUPS
#define FF_P2_GROUPS 304
#endif
#ifndef FF_P2_BLOCKS
#define FF_P2_BLOCKS 3
#endif
#ifndef FF_P2_THREADS
#define FF_P2_THREADS 256
#endif
#ifndef FF_P2_RECS
#define FF_P2_RECS 192
#endif
constexpr int kP2Threads = FF_P2_THREADS;
constexpr int kP2Warps = kP2Threads / 32;
constexpr int kP2Groups = FF_P2_GROUPS;  // groups per staging buffer (x 96 B = 29 KB)
constexpr int kP2Recs = FF_P2_RECS;      // sorted pairs staged per buffer (more: read from global memory)

struct Pair2Meta { uint32_t end, p0, p1, g_base, glob, recs_staged, pad0, pad1; };
struct Pair2Shared {
  unsigned long long full[2];
  uint32_t left[2], next_item[2];
  Pair2Meta meta[2];
  uint4 q[kP2Warps][kQCap];
};

template <int NB> struct Pair2Layout {
  static constexpr int STRIDE = (2 * NB + 3) & ~3;
  static constexpr uint32_t kSlice = (uint32_t)kP2Groups * STRIDE * 4;
  static constexpr uint32_t kRecs = 2u * kSlice;                       // uint4 recs[2][kP2Recs]
  static constexpr uint32_t kShOff = kRecs + 2u * (uint32_t)kP2Recs * 16u;
  static constexpr size_t kBytes = (size_t)kShOff + sizeof(Pair2Shared);
};

// One lane: find the CTA's next non-empty bin and start its copies into buffer b (or publish the end of the work).
template <int NB>
__device__ __forceinline__ void p2_refill(const PairParams &pp, Pair2Shared &sh, uint32_t sbase, int b) {
  using L = Pair2Layout<NB>;
  const uint32_t mbar = smem_addr(&sh.full[b]);
  const uint32_t rb = (uint32_t)pp.rb_shift;
  for (;;) {
    const uint32_t bin = pp.bbin_base + (uint32_t)atomicAdd(pp.next_item, 1ull);
    if (bin >= pp.n_bbins) {
      sh.meta[b].end = 1u;
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
      return;
    }
    const uint32_t e0 = __ldg(pp.off + ((size_t)bin << rb)), e1 = __ldg(pp.off + ((size_t)(bin + 1u) << rb));
    const uint32_t p0 = __ldg(pp.start + ((size_t)bin << rb)), p1 = __ldg(pp.start + ((size_t)(bin + 1u) << rb));
    if (p1 == p0 || e1 == e0) continue;  // no pairs or no entries: nothing to do in this bin
    const uint32_t g_base = e0 >> 5, ng = ((e1 - 1u) >> 5) - g_base + 1u;
    const bool glob = ng > (uint32_t)kP2Groups;
    const bool recs_staged = p1 - p0 <= (uint32_t)kP2Recs;
    Pair2Meta &m = sh.meta[b];
    m.end = 0u; m.p0 = p0; m.p1 = p1; m.g_base = g_base; m.glob = glob ? 1u : 0u; m.recs_staged = recs_staged ? 1u : 0u;
    sh.left[b] = 0u; sh.next_item[b] = 0u;
    const uint32_t bytes_g = glob ? 0u : ng * (uint32_t)(L::STRIDE * 4), bytes_r = recs_staged ? (p1 - p0) * 16u : 0u;
    if (bytes_g + bytes_r == 0u) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
      return;
    }
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes_g + bytes_r) : "memory");
    if (bytes_g) {
      const uint8_t *src = reinterpret_cast<const uint8_t *>(pp.planes + (size_t)g_base * L::STRIDE);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(sbase + (uint32_t)b * L::kSlice), "l"(src), "r"(bytes_g), "r"(mbar) : "memory");
    }
    if (bytes_r)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(sbase + L::kRecs + (uint32_t)b * (uint32_t)kP2Recs * 16u), "l"(pp.recs + p0), "r"(bytes_r), "r"(mbar) : "memory");
    return;
  }
}

template <int NB>
__global__ void __launch_bounds__(kP2Threads, FF_P2_BLOCKS) k_pair_scan2(PairParams pp) {
  using L = Pair2Layout<NB>;
  constexpr int STRIDE = L::STRIDE;
  Pair2Shared &sh = *reinterpret_cast<Pair2Shared *>(ff_smem + L::kShOff);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t sbase = smem_addr(ff_smem);
  const uint32_t q_off = L::kShOff + (uint32_t)offsetof(Pair2Shared, q) + (uint32_t)warp * kQCap * 16u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&sh.full[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&sh.full[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    p2_refill<NB>(pp, sh, sbase, 0);
    p2_refill<NB>(pp, sh, sbase, 1);
  }
  __syncthreads();  // (the only CTA-wide barrier: the mbarriers exist)
  uint32_t qcount = 0;
  unsigned long long compares = 0;
  const uint32_t seg_mask = (1u << pp.seg_shift) - 1u;
  for (uint32_t i = 0;; ++i) {
    const int b = (int)(i & 1u);
    {  // the bin's copies have landed (or the work has ended)
      const uint32_t mbar = smem_addr(&sh.full[b]), parity = (i >> 1) & 1u;
      uint32_t done = 0;
      while (!done)  // (tried: __nanosleep between polls -- no change, try_wait suspends the warp itself)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    }
    const Pair2Meta m = sh.meta[b];
    if (m.end) break;
    const uint32_t p0 = m.p0, p1 = m.p1, g_base = m.g_base;
    const uint32_t claim_addr = smem_addr(&sh.next_item[b]);
    const uint32_t slice = sbase + (uint32_t)b * L::kSlice, recs_sm = sbase + L::kRecs + (uint32_t)b * (uint32_t)kP2Recs * 16u;
    const uint32_t n_items = ((p1 - p0 + 31u) >> 5) << pp.seg_shift;
    for (;;) {
      uint32_t item = 0;
      if (lane == 0) item = atom_add_shared(claim_addr, 1u);
      item = __shfl_sync(0xffffffffu, item, 0);
      if (item >= n_items) break;
      const uint32_t seg = item & seg_mask;
      const uint32_t pl = ((item >> pp.seg_shift) << 5) + (uint32_t)lane;  // pair, relative to the bin's first
      uint32_t lo = 0, hi = 0, gid = 0, probe = 0;
      int budget = -1;
      if (p0 + pl < p1) {
        uint4 r;
        if (m.recs_staged) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(recs_sm + pl * 16u));
        else r = __ldg(pp.recs + p0 + pl);
        lo = r.x; hi = r.y;
        probe = r.z & 0xFFFFFFu; budget = (int)(r.z >> 24); gid = r.w;
      }
      if (seg == 0) compares += hi - lo;
      LaneProbe<NB> lp;
      lp.set(probe, budget);
      // the lane's bucket spans groups g_first .. g_first + n_all - 1; this item takes the seg-th part of them
      const uint32_t g_first = lo >> 5;
      const uint32_t n_all = hi > lo ? ((hi - 1u) >> 5) - g_first + 1u : 0u;
      const uint32_t ga = n_all * seg >> pp.seg_shift, gb = n_all * (seg + 1u) >> pp.seg_shift;
      const int n = (int)gb - (int)ga - 1;  // -1: nothing for this lane
      const uint32_t g0 = g_first + ga;
      if (m.glob) {  // (rare: a run longer than the staging buffer) the same groups from global memory; the cuts as bit positions
        const uint32_t lo_bit = ga == 0u ? (lo & 31u) : 0u, hi_bit = gb == n_all ? ((hi - 1u) & 31u) + 1u : 32u;
        stream_groups<NB, STRIDE, true, false>(0u, pp.planes, 0u, g0, n, g0 << 5, lo_bit, hi_bit, lp, gid, probe, pp.hs, q_off, qcount, lane, pp.canon,
                                               pp.other, pp.lo_d);
      } else {
        const uint32_t first_mask = ga == 0u ? 0xFFFFFFFFu << (lo & 31u) : 0xFFFFFFFFu;
        const uint32_t last_mask = n < 0 ? 0u : (gb == n_all ? 0xFFFFFFFFu >> (31u - ((hi - 1u) & 31u)) : 0xFFFFFFFFu);
        stream_smem<NB, STRIDE, true, true>(slice + (g0 - g_base) * (uint32_t)(STRIDE * 4), n, g0 << 5, first_mask, last_mask, lp, gid, probe, pp.hs,
                                            sbase + q_off, q_off, qcount, lane, pp.canon, pp.other, pp.lo_d);
      }
    }
    // sign off: this warp reads nothing of bin i any more; the last one out refills the buffer with bin i + 2
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();  // release: this warp's reads of the bin's words and slice are done ...
      if (atom_add_shared(smem_addr(&sh.left[b]), 1u) == (uint32_t)kP2Warps - 1u) {
        __threadfence_block();  // ... acquire: so are those of every other warp; the buffer may be overwritten
        p2_refill<NB>(pp, sh, sbase, b);
      }
    }
    __syncwarp();
  }
  drain_queue<true>(pp.hs, q_off, qcount, lane, pp.canon, pp.other, pp.lo_d);
  if (pp.hs.world > 1) __threadfence_system();
  for (int o = 16; o > 0; o >>= 1) compares += __shfl_down_sync(0xffffffffu, compares, o);
  if (lane == 0 && compares) atomicAdd(pp.n_compares, compares);
}

// ------------------------------------------------------------------------------------------------------------
// per-call set-up: guides listed by the bin of their key (part one), pairs sorted by bucket (part two)
__global__ void k_bin_guide_hist(const uint64_t *__restrict__ guides, int64_t n, int proto_shift, uint64_t proto_mask, int b_bits,
                                 unsigned int *__restrict__ cnt) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= n) return;
  const uint64_t proto = (guides[g] >> proto_shift) & proto_mask;
  atomicAdd(cnt + ((uint32_t)(proto >> b_bits) >> 8), 1u);
}

__global__ void k_bin_guide_scatter(const uint64_t *__restrict__ guides, int64_t n, int proto_shift, uint64_t proto_mask, int b_bits,
                                    const int *__restrict__ cls_off, unsigned int *__restrict__ cursor, uint2 *__restrict__ sg) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= n) return;
  const uint64_t proto = (guides[g] >> proto_shift) & proto_mask;
  const uint32_t ka = (uint32_t)(proto >> b_bits), kb = (uint32_t)(proto & ((1ull << b_bits) - 1ull));
  const uint32_t cls = ka >> 8;
  sg[(uint32_t)cls_off[cls] + atomicAdd(cursor + cls, 1u)] = make_uint2(kb | ((ka & 0xFFu) << 24), (uint32_t)g);
}

struct BinScanPlan {
  BinParams bp;
  PairParams pp;
  bool part_two;
  bool staged_b;   // part two through shared memory (k_pair_scan2) or lanes reading global memory (k_pair_scan)
  bool shard;      // database-sharded: candidates go to `sink` (the owners' exchange blocks)
  HitSink sink;
  int nb_a, nb_b;  // other bases of the two halves
  size_t smem_a;
};

static bool bin_scan_supported(const Database &db, int hA, int64_t G) {
  if (!db.A.d_planes || !db.A.d_goff || !db.B.d_planes || db.A.n_planes != 18) return false;
  if (db.B.n_planes != 20 && db.B.n_planes != 22) return false;
  if (hA > 3 || db.A.cum_hi[hA] > kNbrCap) return false;
  return G < (1ll << 28);
}

// Launch the set-up kernels (no host synchronisation) and fill the plan.
// Database-sharded discover: this rank scans bins / buckets [rank / world, (rank + 1) / world) of both index halves for all
// the guides; `sink` (world, first[], peer[], hit_cap, tbits) says where the candidates go.
struct ScanShard {
  int rank = 0, world = 1;
  HitSink sink = {};
};

static int bin_scan_prepare(ff_ctx *ctx, const ScanParams &sp, int hA, int nB, unsigned long long *n_compares_b, BinScanPlan *pl,
                            int *launches, const ScanShard *shard = nullptr, bool reserve_only = false) {
  Database &db = ctx->db;
  cudaStream_t st = ctx->stream;
  const int64_t G = sp.n_guides;
  const uint32_t n_bins = 1u << (2 * (db.A.key_bases - 4));
  const uint32_t n_keys_b = 1u << (2 * db.B.key_bases);
  const long long n_pairs = (long long)G * nB;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_cls_cnt = 0;                                        // u32[n_bins + 1]  (zeroed)
  const size_t o_cls_cur = o_cls_cnt + up(((size_t)n_bins + 1) * 4);  // u32[n_bins]      (zeroed)
  const size_t o_b_cnt = o_cls_cur + up((size_t)n_bins * 4);         // u32[n_keys_b + 1] (zeroed)
  const size_t o_b_cur = o_b_cnt + up(((size_t)n_keys_b + 1) * 4);    // u32[n_keys_b]     (zeroed)
  const size_t o_ctr = o_b_cur + up((size_t)n_keys_b * 4);           // counters          (zeroed)
  const size_t zero_bytes = o_ctr + 256;
  const size_t o_cls_off = zero_bytes;                               // int[n_bins + 1]
  const size_t o_b_start = o_cls_off + up(((size_t)n_bins + 1) * 4);  // u32[n_keys_b + 1]
  const size_t o_sg = o_b_start + up(((size_t)n_keys_b + 1) * 4);     // uint2[G]
  const size_t o_recs = o_sg + up((size_t)(G > 0 ? G : 1) * 8);       // uint4[n_pairs]
  const size_t total = o_recs + up((size_t)(n_pairs > 0 ? n_pairs : 1) * 16);
  FF_TRY(ctx->cell_ws.reserve(total));
  uint8_t *w = ctx->cell_ws.as<uint8_t>();
  FF_CUDA(cudaMemsetAsync(w, 0, zero_bytes, st));
  unsigned int *cls_cnt = (unsigned int *)(w + o_cls_cnt), *cls_cur = (unsigned int *)(w + o_cls_cur);
  unsigned int *b_cnt = (unsigned int *)(w + o_b_cnt), *b_cur = (unsigned int *)(w + o_b_cur);
  int *cls_off = (int *)(w + o_cls_off);
  unsigned int *b_start = (unsigned int *)(w + o_b_start);
  uint2 *sg = (uint2 *)(w + o_sg);
  uint4 *recs = (uint4 *)(w + o_recs);

  size_t tmp = 0, tmp2 = 0;
  FF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, cls_cnt, cls_off, (int)(n_bins + 1), st));
  FF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, b_cnt, b_start, (int)(n_keys_b + 1), st));
  FF_TRY(ctx->cub_tmp.reserve(std::max(tmp, tmp2)));
  if (reserve_only) return FF_OK;  // (the workspaces exist now; nothing was launched)
  if (G > 0) {
    k_bin_guide_hist<<<blocks_for(G, 256), 256, 0, st>>>(sp.guides, G, sp.proto_shift, sp.proto_mask, sp.b_bits, cls_cnt);
    FF_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, cls_cnt, cls_off, (int)(n_bins + 1), st));
    k_bin_guide_scatter<<<blocks_for(G, 256), 256, 0, st>>>(sp.guides, G, sp.proto_shift, sp.proto_mask, sp.b_bits, cls_off, cls_cur, sg);
    *launches += 3;
  }
  // B-bins of k_pair_scan2: as many buckets as fit the staging buffer on average, but enough bins to balance the grid
  PairParams &pp = pl->pp;
  {
    int kb_bits = 2 * db.B.key_bases;
    const double groups_per_bucket = (double)db.n_targets / (double)n_keys_b / 32.0;
    int rbs = 0;
    auto fits = [&](int r) {  // mean run of 2^r buckets + three standard deviations (Poisson bucket sizes) within the buffer
      const double entries = groups_per_bucket * 32.0 * (double)(1u << r);
      return entries / 32.0 + 1.0 + 3.0 * std::sqrt(entries) / 32.0 <= (double)kP2Groups;
    };
    while (rbs + 1 <= kb_bits && fits(rbs + 1)) ++rbs;
    rbs = std::min(rbs, std::max(0, kb_bits - 11));
    pp.rb_shift = rbs; pp.n_bbins = n_keys_b >> rbs; pp.bbin_base = 0;
    const double rounds = (double)n_pairs * (double)(1u << rbs) / (double)n_keys_b / 32.0;
    int ss = 0;
    while (ss < 3 && rounds * (double)(1 << ss) < 0.6 * kP2Warps && groups_per_bucket / (double)(2 << ss) >= 4.0) ++ss;
    if (ctx->opt.pair_segs > 0) { ss = 0; while ((1 << (ss + 1)) <= ctx->opt.pair_segs) ++ss; }
    pp.seg_shift = ss;
  }
  uint32_t bin_lo = 0, bin_hi = n_bins, b_lo = 0, b_hi = n_keys_b;
  if (ctx->opt.debug_bin_div > 1) bin_hi = n_bins / (uint32_t)ctx->opt.debug_bin_div;  // (diagnostic: part one over a fraction of the bins; rows incomplete)
  if (shard && shard->world > 1) {
    bin_lo = (uint32_t)((uint64_t)n_bins * shard->rank / shard->world); bin_hi = (uint32_t)((uint64_t)n_bins * (shard->rank + 1) / shard->world);
    const uint32_t bb_lo = (uint32_t)((uint64_t)pp.n_bbins * shard->rank / shard->world), bb_hi = (uint32_t)((uint64_t)pp.n_bbins * (shard->rank + 1) / shard->world);
    pp.bbin_base = bb_lo; pp.n_bbins = bb_hi;
    b_lo = bb_lo << pp.rb_shift; b_hi = bb_hi << pp.rb_shift;
  }
  if (n_pairs > 0) {
    k_bpairs_hist<<<blocks_for(n_pairs, 256), 256, 0, st>>>(sp.guides, n_pairs, db.B.d_masks, nB, sp.proto_shift, sp.proto_mask, sp.b_bits, sp.k, b_lo, b_hi, b_cnt);
    FF_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp2, b_cnt, b_start, (int)(n_keys_b + 1), st));
    k_bpairs_scatter<<<blocks_for(n_pairs, 256), 256, 0, st>>>(sp.guides, n_pairs, db.B.d_masks, nB, sp.proto_shift, sp.proto_mask, sp.b_bits, sp.k,
                                                               b_lo, b_hi, b_start, b_cur, db.B.d_off, recs);
    *launches += 3;
  }
  FF_CUDA(cudaGetLastError());

  BinParams &bp = pl->bp;
  bp.planes = db.A.d_planes; bp.off = db.A.d_off; bp.goff = db.A.d_goff; bp.canon = db.A.d_canon; bp.himasks = db.A.d_himasks; bp.lomasks = db.A.d_lomasks;
  bp.n_hi = db.A.cum_hi[hA];
  for (int d = 0; d < 5; ++d) bp.cum_hi[d] = db.A.cum_hi[std::min(d, hA)];
  for (int d = 0; d < 4; ++d) {
    bp.nm[d] = d <= hA ? db.A.cum_lo[std::min(hA - d, 4)] : 1;
    bp.rcp[d] = bp.nm[d] > 1 ? (uint32_t)(0x100000000ull / (uint64_t)bp.nm[d]) : 0xFFFFFFFFu;
  }
  bp.hA = hA; bp.k = sp.k; bp.n_bins = bin_hi; bp.bin_base = bin_lo; bp.sg = sg; bp.cls_off = cls_off;
  bp.hs = HitSink{sp.hits, sp.hit_count, sp.hit_cap, sp.tbits, nullptr};
  pl->shard = shard && shard->world > 1;
  if (pl->shard) pl->sink = shard->sink;
  bp.n_compares = sp.n_compares;
  bp.next_bin = (unsigned int *)(w + o_ctr);
  pp.planes = db.B.d_planes; pp.off = db.B.d_off; pp.other = db.B.d_other; pp.canon = db.B.d_canon; pp.recs = recs;
  pp.n_pairs = n_pairs; pp.lo_d = hA; pp.hs = bp.hs; pp.n_compares = n_compares_b;
  pp.next_item = (unsigned long long *)(w + o_ctr + 64);
  pp.start = b_start;
  pp.n_pairs_dev = pl->shard ? b_start + n_keys_b : nullptr;
  // the ring stages the WHOLE index once per call; lanes reading global memory touch only the buckets that have pairs:
  // measured on B200 (3 x 10^8 targets) the two meet at ~11 pairs per bucket (100 000 guides); the ring wins beyond
  pl->staged_b = ctx->opt.pair_kernel == 2 || (ctx->opt.pair_kernel == 0 && (double)n_pairs >= 12.0 * (double)n_keys_b);
  pl->part_two = n_pairs > 0;
  pl->nb_a = db.A.n_planes / 2; pl->nb_b = db.B.n_planes / 2;
  pl->smem_a = (size_t)kSliceGroups * db.A.n_planes * 4 + sizeof(BinShared);
  return FF_OK;
}

// (re)launch the two scan kernels of a prepared plan; hit buffer and counters come from `sp`
static int bin_scan_launch(ff_ctx *ctx, BinScanPlan *pl, const ScanParams &sp, unsigned int *gcnt, int *launches, uint32_t *ranks = nullptr) {
  cudaStream_t st = ctx->stream;
  pl->bp.hs = HitSink{sp.hits, sp.hit_count, sp.hit_cap, sp.tbits, gcnt, gcnt ? ranks : nullptr};
  if (pl->shard) { pl->bp.hs = pl->sink; pl->bp.hs.tbits = sp.tbits; pl->bp.hs.gcnt = nullptr; }
  pl->pp.hs = pl->bp.hs;
  FF_CUDA(cudaMemsetAsync(pl->bp.next_bin, 0, 128, st));
  static bool attr_set[64] = {false};
  if (!attr_set[ctx->device & 63]) {
    FF_CUDA(cudaFuncSetAttribute(k_bin_scan<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem_a));
    attr_set[ctx->device & 63] = true;
  }
  k_bin_scan<9><<<ctx->sm_count * 2, kBinThreads, pl->smem_a, st>>>(pl->bp);
  (*launches)++;
  if (pl->part_two) {
    FF_CUDA(cudaEventRecord(ctx->ev[7], st));
    if (pl->staged_b) {
      const size_t sm2 = pl->nb_b == 11 ? Pair2Layout<11>::kBytes : Pair2Layout<10>::kBytes;
      static bool attr2[64] = {false};
      if (!attr2[ctx->device & 63]) {
        FF_CUDA(cudaFuncSetAttribute(k_pair_scan2<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Pair2Layout<11>::kBytes));
        FF_CUDA(cudaFuncSetAttribute(k_pair_scan2<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Pair2Layout<10>::kBytes));
        attr2[ctx->device & 63] = true;
      }
      const int grid = (int)std::min<uint32_t>((uint32_t)(ctx->sm_count * FF_P2_BLOCKS), pl->pp.n_bbins);
      if (pl->nb_b == 11) k_pair_scan2<11><<<grid, kP2Threads, sm2, st>>>(pl->pp);
      else k_pair_scan2<10><<<grid, kP2Threads, sm2, st>>>(pl->pp);
    } else {
      const int grid = ctx->sm_count * FF_PAIR_MIN_BLOCKS;
      const size_t qsm = (size_t)kPairWarps * kQCap * 16;
      if (pl->nb_b == 11) k_pair_scan<11><<<grid, kPairThreads, qsm, st>>>(pl->pp);
      else k_pair_scan<10><<<grid, kPairThreads, qsm, st>>>(pl->pp);
    }
    (*launches)++;
  }
  FF_CUDA(cudaGetLastError());
  return FF_OK;
}
