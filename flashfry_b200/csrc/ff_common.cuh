// Shared declarations of libflashfry_b200: context layout, device buffers, error plumbing.
// Product code: nothing under oracle/ is ever included or linked here.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/flashfry_b200.h"

namespace ff {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define FF_CUDA(expr)                                                          \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return ff::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define FF_TRY(expr)          \
  do {                        \
    int _rc = (expr);         \
    if (_rc != FF_OK) return _rc; \
  } while (0)

// A grow-only device allocation (steady-state calls never hit cudaMalloc).
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);  // contents are NOT preserved when it grows
  void release();
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Pinned host allocation, pooled by the context and lent to ff_hits.
struct HostBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// standards/StandardScanParameters.scala:28-48 -- the fields the device path needs
struct Pack {
  int enzyme_index = 0, scan_len = 0, pam_len = 0, five_prime = 0;
  uint64_t cmp_mask = 0;
};
int pack_from_index(int enzyme_index, Pack *out);  // StandardScanParameters.scala:61-70

constexpr int kCells = 64;  // cells of a key space: the value of the first three key bases (4^3)

// One half of the seed index: targets ordered by `key` (a run of consecutive protospacer bases), the complementary
// part of the protospacer stored next to it.
struct SeedIndex {
  int key_bases = 0;            // w: bases in the key
  uint32_t *d_off = nullptr;    // [4^w + 1] first entry of every key
  uint32_t *d_other = nullptr;  // [n + pad] the other part of the protospacer, right-aligned
  uint32_t *d_canon = nullptr;  // [n] database-order index of every entry; nullptr when the order IS database order
  uint32_t *d_masks = nullptr;  // [4^w] XOR masks sorted by Hamming distance (in bases): mask | distance << 24
  int cum[16] = {0};            // cum[h] = # masks at distance <= h (h = 0..w)
  uint32_t *d_masks_w1 = nullptr;  // the same table over w-1 bases (bulge patterns: one key base is a wildcard)
  int cum_w1[16] = {0};
  // bit-sliced copy of `other` for the bin scan (ff_binscan.inl): entries in groups of 32, one 32-bit word per bit plane
  // (bit e of plane word j = bit j of other[32 g + e]); plane_stride words per group
  uint32_t *d_planes = nullptr;
  int n_planes = 0, plane_stride = 0;
  uint64_t n_groups = 0;
  // index A only: every bucket starts a group of its own (d_goff[key] = first group of the bucket; the last group of a
  // bucket is padded).  A bucket of ~72 entries then spans ceil(72 / 32) = 3 groups instead of 3.2 .. 4 at a random
  // alignment, and the first group needs no range mask.  nullptr: groups are cut at multiples of 32 entries (index B).
  uint32_t *d_goff = nullptr;
  // masks over the first key_bases - 4 key bases (the "bin" of a key), sorted by distance; the masks over the last four
  // key bases live in constant memory (the same 256 for every key width)
  uint32_t *d_himasks = nullptr;   // [4^(key_bases-4)] mask | distance << 24
  uint32_t *d_lomasks = nullptr;   // [256] mask | distance << 8
  int cum_hi[16] = {0}, cum_lo[16] = {0};
  void release();
};

struct Database {
  bool resident = false;
  Pack pack;
  int bin_width = 7;            // of the file header (the device index does not use the file's bins)
  uint64_t n_targets = 0, n_positions = 0;
  int proto_bases = 0;          // P: compared bases (popcount(cmp_mask) / 2)
  int proto_shift = 0;          // bit offset of the protospacer inside the target long
  uint64_t *d_targets = nullptr;   // [n_targets] target longs, database order
  uint64_t *d_pos_off = nullptr;   // [n_targets + 1] exclusive scan of counts (only with positions)
  uint64_t *d_positions = nullptr; // [n_positions]
  SeedIndex A;                  // keyed by the first a protospacer bases, other = the last P-a bases
  SeedIndex B;                  // keyed by the last P-a bases, other = the first a bases
  // database-order windows (early termination of overflowed guides): the index-A key space is cut into kCells equal
  // cells; d_cell_off[key_b * (kCells + 1) + c] = first entry of B bucket key_b whose database index lies in cell >= c.
  // Built on first use (3'-PAM databases only: their index-A order is database order).
  uint32_t *d_cell_off = nullptr;
  uint64_t device_bytes = 0;
  std::vector<std::string> contigs;
  void release();
};

struct Hits;  // host-side result owner (ff_api.cu)

// Tuning and test knobs, set through ff_set_option (no environment variables are read by the library).
struct Options {
  int scan_kernel = 0;     // 0 = choose by batch and database size, 1 = guide-major k_seed_scan, 2 = bin-major k_bin_scan / k_pair_scan
  int force_general = 0;   // 1 = take the windowed general path even for the mismatch-only search
  int window_cells = 0;    // > 0: fixed window size (cells) of the general path
  int subbatch_min = 45000;           // ff_discover cuts a guide set into up to three sub-batches of at least this size
                                      // (measured on B200: two sub-batches 60 / 40 % beat one and three for 100 000 guides)
  int subbatch_c1 = 65, subbatch_c2 = 90;  // cumulative % of the first two of three sub-batches
  int subbatch_two = 70;                   // % of the guides in the first of two sub-batches
  int group_sort = 1;      // 0 = always order hits with the radix sort
  int b_spi = 0;           // > 0: seeds per work item of the part-two pass of k_seed_scan / k_pattern_scan
  int split_a = 0;         // > 0: bases in the part-one key of the next database build
  int trace = 0;           // 1 = ff_discover prints host-side timestamps of its sub-batches to stderr
  int compact_hits = 0;    // ff_discover: 1 = ship database indices instead of target longs (ff_hits.target_index)
  int pair_kernel = 0;     // part two of the bin-major scan: 0 = by pairs per bucket, 1 = k_pair_scan (lanes read global memory),
                           // 2 = k_pair_scan2 (B-bins staged in shared memory by a two-buffer TMA ring)
  int debug_bin_div = 0;   // diagnostic: > 1 = part one of the bin-major scan covers only the first 1/n of the bins (timing; rows incomplete)
  int peer_local_only = 0; // diagnostic: the sharded scan keeps every candidate in its own block (timing without NVLink stores; rows are wrong)
  int pair_segs = 0;       // > 0: work items per round of 32 pairs in k_pair_scan2 (1, 2, 4, 8); 0 = by batch size
};

// Database-sharded discover (ff_shard.inl): this rank's exchange block and the mapped blocks of its peers.
constexpr int kMaxPeers = 8;
constexpr size_t kPeerHeadBytes = 256;  // PeerCtr at the head of an exchange block
struct PeerLink {
  bool ready = false, fresh = false;   // fresh: exported and not attached yet
  int rank = 0, world = 1;
  DevBuf block;                         // own block: PeerCtr (256 B) | candidate keys u64[hit_cap] in `world` regions | per-guide totals i32[g_cap]
  size_t hit_cap = 0;
  int64_t g_cap = 0;
  uint8_t *base[kMaxPeers] = {nullptr};  // every rank's block (base[rank] = own); peers mapped by CUDA IPC or peer access
  bool ipc_opened[kMaxPeers] = {false};
  unsigned int epoch = 0;               // barriers passed (all ranks call them in the same sequence)
};

}  // namespace ff

struct ff_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  ff::Database db;
  ff::Options opt;

  // ---- per-call workspaces (grow-only) ----
  ff::DevBuf cub_tmp;
  ff::DevBuf hit_keys, hit_keys_sorted, counters, hit_ranks;
  ff::DevBuf seg_start, n_keep;
  ff::DevBuf idx32, st_targets, st_mm;  // per-guide ordering: scattered database indices, rows staged at their segment
  ff::HostBuf host_targets;             // host mirror of db.d_targets (ff_db_host_targets), made on first use
  uint64_t host_targets_n = 0;
  // Status words of a call (candidate count, flags, hit count), written by the last kernel straight into MAPPED pinned
  // host memory: a cudaMemcpy of these few bytes would queue behind the previous sub-batch's hit-list D2H in the copy engine.
  void *h_status = nullptr, *h_status_dev = nullptr;
  unsigned int status_seq = 0;
  ff::DevBuf pos_cnt, pos_ptr, out_positions;
  ff::DevBuf cfd_per_ot, hsu_per_ot;
  ff::DevBuf scratch_guides;  // H2D staging target for ff_discover
  // windowed / bulge discover: per-guide running totals, the active guide list (two copies), kept keys, scratch
  ff::DevBuf running, active, active2, act_flags, seg_end, kept_keys, kept_sorted, n_sel;
  ff::DevBuf cell_ws;  // bin scan: guides listed by bin, (guide, seed) pairs sorted by bucket, counters
  // results of a discover call; two sets so that the D2H of one guide sub-batch overlaps the scan of the next
  struct OutSlot {
    ff::DevBuf row_ptr, total_count, overflowed, out_targets, out_mm, out_tidx, out_bulge, cfd_max, cfd_spec, hsu;
  } out[2];
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t slot_copied[2] = {nullptr, nullptr};  // D2H of the slot's previous contents has finished
  size_t hit_cap = 0;

  ff::PeerLink peer;
  cudaEvent_t ev[8] = {nullptr};
  ff_timings last = {};
  std::vector<ff::HostBuf *> host_pool;
};
