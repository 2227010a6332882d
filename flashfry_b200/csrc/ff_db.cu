// Database residency: FlashFry's on-disk format -> SoA image in HBM + the device-side prefix index.
//
// Replaces (FlashFry, src/main/scala/...):
//   BinaryHeader.readHeader                 reference/binary/BinaryHeader.scala:115-160
//   SeekTraverser.fillBlock (BGZF seek+read) reference/traverser/SeekTraverser.scala:113-121  (htsjdk 2.8.1
//                                           BlockCompressedInputStream, build.sbt:17 -- BGZF per SAM spec 4.1)
//   Utils.byteArrayToLong                   utils/Utils.scala:167-186 (native = little-endian longs)
//   block decoders                          reference/binary/blocks/BlockManager.scala:266-351
//
// HBM layout:
//   targets  u64[N_t]            the reference's target longs in DATABASE ORDER (2-bit bases in bits 0..47, occurrence
//                                count in 48..63) -- read only for emitted hits
//   seed index A                 entries ordered by the first a protospacer bases (for 3'-PAM enzymes that IS database
//                                order): off u32[4^a+1], other u32[N_t] = the last P-a bases
//   seed index B                 entries ordered by the last b = P-a bases: off u32[4^b+1], other u32[N_t] = the first a
//                                bases, canon u32[N_t] = database-order index of every entry
//   mask tables                  all XOR masks over a (resp. b) bases sorted by Hamming distance, mask | distance << 24
//   pos_off  u64[N_t+1], positions u64[N_p]   only touched for emitted hits when positions are requested
#include <cub/cub.cuh>
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "ff_common.cuh"
#include "ff_kernels.cuh"

namespace ff {

// standards/StandardScanParameters.scala:61-70 + the constants of :90-215
int pack_from_index(int idx, Pack *o) {
  o->enzyme_index = idx;
  switch (idx) {
    case 1: o->scan_len = 24; o->pam_len = 4; o->five_prime = 1; o->cmp_mask = 0x00FFFFFFFFFFull; return FF_OK;
    case 2: case 3: case 4: o->scan_len = 23; o->pam_len = 3; o->five_prime = 0; o->cmp_mask = 0x3FFFFFFFFFC0ull; return FF_OK;
    case 5: case 6: o->scan_len = 22; o->pam_len = 3; o->five_prime = 0; o->cmp_mask = 0x0FFFFFFFFFC0ull; return FF_OK;
    default: set_error("Unable to find the correct parameter pack for enzyme: %d", idx); return FF_EINVAL;
  }
}

void SeedIndex::release() {
  cudaFree(d_off); cudaFree(d_other); cudaFree(d_canon); cudaFree(d_masks); cudaFree(d_masks_w1);
  cudaFree(d_planes); cudaFree(d_goff); cudaFree(d_himasks); cudaFree(d_lomasks);
  *this = SeedIndex();
}

void Database::release() {
  cudaFree(d_targets); cudaFree(d_pos_off); cudaFree(d_positions); cudaFree(d_cell_off);
  A.release(); B.release();
  *this = Database();
}

// ------------------------------------------------------------------------------------------------------------
// strictly increasing over the 48 sequence bits?  (what makes database order == index-A order for 3'-PAM enzymes)
__global__ void k_check_sorted(const uint64_t *__restrict__ t, uint64_t n, int check_order, unsigned int *__restrict__ bad) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t m = 0xFFFFFFFFFFFFull;
  if (check_order && i > 0 && (t[i - 1] & m) >= (t[i] & m)) atomicAdd(bad, 1u);
  if ((t[i] >> 48) == 0 || (t[i] >> 63)) atomicAdd(bad + 1, 1u);
}

// split every protospacer into its first a bases (keyA) and its last b bases (keyB)
__global__ void k_split_proto(const uint64_t *__restrict__ t, uint64_t n, int proto_shift, int b_bits, uint64_t proto_mask,
                              uint32_t *__restrict__ key_a, uint32_t *__restrict__ key_b, uint32_t *__restrict__ iota) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t proto = (t[i] >> proto_shift) & proto_mask;
  key_a[i] = (uint32_t)(proto >> b_bits);
  key_b[i] = (uint32_t)(proto & ((1ull << b_bits) - 1ull));
  if (iota) iota[i] = (uint32_t)i;
}

__global__ void k_gather_u32(const uint32_t *__restrict__ src, const uint32_t *__restrict__ idx, uint64_t n, uint32_t *__restrict__ dst) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

// off[j] = first entry whose key >= j, for j in [0, n_keys]; one thread per key (binary search over the sorted keys)
__global__ void k_key_offsets(const uint32_t *__restrict__ sorted_keys, uint64_t n, uint32_t n_keys, uint32_t *__restrict__ off) {
  uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (j > n_keys) return;
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (sorted_keys[mid] < j) lo = mid + 1; else hi = mid;
  }
  off[j] = (uint32_t)lo;
}

__global__ void k_counts(const uint64_t *__restrict__ t, uint64_t n, uint64_t *__restrict__ c) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) c[i] = t[i] >> 48;
  if (i == n) c[i] = 0;
}

// groups per bucket (bucket-aligned planes of index A)
__global__ void k_bucket_groups(const uint32_t *__restrict__ off, uint32_t n_keys, uint32_t *__restrict__ ng) {
  const uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (k < n_keys) ng[k] = (off[k + 1] - off[k] + 31u) >> 5;
  if (k == n_keys) ng[k] = 0;
}

// Bit-slice `other` with every bucket starting its own group: one warp per group; the group's bucket is found by a
// binary search over the group offsets, entries past the bucket's end read as all-ones (the scan masks them away).
__global__ void k_slice_planes_aligned(const uint32_t *__restrict__ other, const uint32_t *__restrict__ off, const uint32_t *__restrict__ goff,
                                       uint32_t n_keys, uint64_t n_groups, int n_planes, int stride, uint32_t *__restrict__ planes) {
  const uint64_t g = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_groups) return;
  uint32_t lo = 0, hi = n_keys;  // last key with goff[key] <= g (keys with no entries share their successor's offset)
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo + 1) >> 1);
    if (goff[mid] <= g) lo = mid; else hi = mid - 1;
  }
  uint32_t v = 0xFFFFFFFFu;
  if (lo < n_keys && g < goff[lo + 1]) {
    const uint64_t e = (uint64_t)off[lo] + (g - goff[lo]) * 32ull + (uint64_t)lane;
    if (e < off[lo + 1]) v = other[e];
  }
  uint32_t mine = 0;
  for (int j = 0; j < n_planes; ++j) {
    const uint32_t w = __ballot_sync(0xffffffffu, (v >> j) & 1u);
    if (lane == j) mine = w;
  }
  if (lane < stride) planes[g * stride + lane] = lane < n_planes ? mine : 0u;
}

static int base_distance(uint32_t m) { return __builtin_popcount((m | (m >> 1)) & 0x55555555u); }

// all XOR masks over `bases` bases sorted by (distance, value), packed mask | distance << 24; cum[h] = # masks with distance <= h
static void make_masks(int bases, std::vector<uint32_t> *out, int *cum) {
  const uint32_t n = 1u << (2 * bases);
  std::vector<uint32_t> count(bases + 2, 0);
  for (uint32_t m = 0; m < n; ++m) count[base_distance(m) + 1]++;
  for (int d = 0; d <= bases; ++d) count[d + 1] += count[d];
  for (int h = 0; h <= bases; ++h) cum[h] = (int)count[h + 1];
  for (int h = bases + 1; h < 16; ++h) cum[h] = (int)n;
  out->assign(n, 0);
  std::vector<uint32_t> cursor(count.begin(), count.end() - 1);
  for (uint32_t m = 0; m < n; ++m) {
    const int d = base_distance(m);
    (*out)[cursor[d]++] = m | ((uint32_t)d << 24);
  }
}

// Mask tables of one key width: all masks by distance, and the same over width - 1 bases (bulge wildcards).
struct MaskTables {
  std::vector<uint32_t> masks, masks_w1;
  int cum[16], cum_w1[16];
};

static const MaskTables &mask_tables(int key_bases) {
  static std::mutex mu;
  static std::map<int, MaskTables> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key_bases);
  if (it != cache.end()) return it->second;
  MaskTables &t = cache[key_bases];
  make_masks(key_bases, &t.masks, t.cum);
  make_masks(key_bases - 1, &t.masks_w1, t.cum_w1);
  return t;
}

static unsigned int nblk(uint64_t n) { return (unsigned int)((n + 255) / 256); }

// Bit-slice `other`: one warp per group of 32 entries, plane word j = ballot of bit j.  Entries past n read the 0xFF
// padding of d_other; the scan rejects them by their index.
__global__ void k_slice_planes(const uint32_t *__restrict__ other, uint64_t n_groups, int n_planes, int stride, uint32_t *__restrict__ planes) {
  const uint64_t g = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_groups) return;
  const uint32_t v = other[g * 32 + lane];
  uint32_t mine = 0;
  for (int j = 0; j < n_planes; ++j) {
    const uint32_t w = __ballot_sync(0xffffffffu, (v >> j) & 1u);
    if (lane == j) mine = w;
  }
  if (lane < stride) planes[g * stride + lane] = lane < n_planes ? mine : 0u;
}

// Build one half of the seed index from per-target keys.  identity: entries stay in database order (keys must already be sorted).
static int build_seed_index(ff_ctx *ctx, SeedIndex *ix, int key_bases, int other_bases, bool wide_planes, const uint32_t *d_key,
                            const uint32_t *d_other_src, const uint32_t *d_iota, uint64_t n, bool identity) {
  cudaStream_t st = ctx->stream;
  ix->key_bases = key_bases;
  const uint32_t n_keys = 1u << (2 * key_bases);
  FF_CUDA(cudaMalloc(&ix->d_off, ((size_t)n_keys + 1) * 4));
  FF_CUDA(cudaMalloc(&ix->d_other, (n + 64) * 4));
  FF_CUDA(cudaMemsetAsync(ix->d_other, 0xFF, (n + 64) * 4, st));
  const uint32_t *sorted_keys = d_key;
  uint32_t *d_sorted = nullptr;
  if (identity) {
    if (n) FF_CUDA(cudaMemcpyAsync(ix->d_other, d_other_src, n * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    FF_CUDA(cudaMalloc(&d_sorted, (n + 1) * 4));
    FF_CUDA(cudaMalloc(&ix->d_canon, (n + 1) * 4));
    size_t tmp = 0;
    FF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, d_key, d_sorted, d_iota, ix->d_canon, n, 0, 2 * key_bases, st));
    FF_TRY(ctx->cub_tmp.reserve(tmp));
    FF_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, d_key, d_sorted, d_iota, ix->d_canon, n, 0, 2 * key_bases, st));
    if (n) k_gather_u32<<<nblk(n), 256, 0, st>>>(d_other_src, ix->d_canon, n, ix->d_other);
    sorted_keys = d_sorted;
  }
  k_key_offsets<<<nblk((uint64_t)n_keys + 1), 256, 0, st>>>(sorted_keys, n, n_keys, ix->d_off);
  // the mask tables depend on the key width only: built once per process, re-uploaded per database
  const MaskTables &mt = mask_tables(key_bases);
  const std::vector<uint32_t> &masks = mt.masks, &masks_w1 = mt.masks_w1;
  memcpy(ix->cum, mt.cum, sizeof ix->cum);
  memcpy(ix->cum_w1, mt.cum_w1, sizeof ix->cum_w1);
  FF_CUDA(cudaMalloc(&ix->d_masks, masks.size() * 4));
  FF_CUDA(cudaMemcpyAsync(ix->d_masks, masks.data(), masks.size() * 4, cudaMemcpyHostToDevice, st));
  FF_CUDA(cudaMalloc(&ix->d_masks_w1, masks_w1.size() * 4));
  FF_CUDA(cudaMemcpyAsync(ix->d_masks_w1, masks_w1.data(), masks_w1.size() * 4, cudaMemcpyHostToDevice, st));
  // bin scan (ff_binscan.inl): bit-sliced `other` + the masks of a key split into its bin (first key_bases - 4 bases)
  // and its last four bases.  Built for the splits the compare circuits exist for (9, 10 or 11 other bases).
  if (key_bases >= 6 && other_bases >= 9 && other_bases <= 11) {
    ix->n_planes = 2 * other_bases;
    ix->plane_stride = wide_planes ? (ix->n_planes + 3) & ~3 : ix->n_planes;
    if (wide_planes) {  // index B: long buckets, groups at multiples of 32 entries
      ix->n_groups = (n + 31) / 32 + 1;
      FF_CUDA(cudaMalloc(&ix->d_planes, ix->n_groups * ix->plane_stride * 4 + 64));
      k_slice_planes<<<nblk(ix->n_groups * 32), 256, 0, st>>>(ix->d_other, ix->n_groups, ix->n_planes, ix->plane_stride, ix->d_planes);
    } else {            // index A: short buckets, every bucket starts a group
      uint32_t *d_ng = nullptr;
      FF_CUDA(cudaMalloc(&d_ng, ((size_t)n_keys + 1) * 4));
      FF_CUDA(cudaMalloc(&ix->d_goff, ((size_t)n_keys + 1) * 4));
      k_bucket_groups<<<nblk((uint64_t)n_keys + 1), 256, 0, st>>>(ix->d_off, n_keys, d_ng);
      size_t tmp = 0;
      FF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_ng, ix->d_goff, (int)(n_keys + 1), st));
      FF_TRY(ctx->cub_tmp.reserve(tmp));
      FF_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, d_ng, ix->d_goff, (int)(n_keys + 1), st));
      uint32_t total = 0;
      FF_CUDA(cudaMemcpyAsync(&total, ix->d_goff + n_keys, 4, cudaMemcpyDeviceToHost, st));
      FF_CUDA(cudaStreamSynchronize(st));
      cudaFree(d_ng);
      ix->n_groups = (uint64_t)total + 2;  // + padding: a bin's slice is copied from an even group, in 16-byte units
      FF_CUDA(cudaMalloc(&ix->d_planes, ix->n_groups * ix->plane_stride * 4 + 64));
      k_slice_planes_aligned<<<nblk(ix->n_groups * 32), 256, 0, st>>>(ix->d_other, ix->d_off, ix->d_goff, n_keys, ix->n_groups, ix->n_planes,
                                                                   ix->plane_stride, ix->d_planes);
      ctx->db.device_bytes += ((size_t)n_keys + 1) * 4;
    }
    std::vector<uint32_t> hi, lo;
    make_masks(key_bases - 4, &hi, ix->cum_hi);
    make_masks(4, &lo, ix->cum_lo);
    for (uint32_t &m : lo) m = (m & 0xFFu) | ((m >> 24) << 8);
    FF_CUDA(cudaMalloc(&ix->d_himasks, hi.size() * 4));
    FF_CUDA(cudaMemcpyAsync(ix->d_himasks, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice, st));
    FF_CUDA(cudaMalloc(&ix->d_lomasks, lo.size() * 4));
    FF_CUDA(cudaMemcpyAsync(ix->d_lomasks, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice, st));
    FF_CUDA(cudaStreamSynchronize(st));  // hi / lo go out of scope
    ctx->db.device_bytes += ix->n_groups * ix->plane_stride * 4 + hi.size() * 4 + lo.size() * 4;
  }
  FF_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_sorted);
  FF_CUDA(cudaGetLastError());
  ctx->db.device_bytes += ((size_t)n_keys + 1) * 4 + (n + 64) * 4 + (identity ? 0 : (n + 1) * 4) + masks.size() * 4 + masks_w1.size() * 4;
  return FF_OK;
}

int db_build_index(ff_ctx *ctx) {
  Database &db = ctx->db;
  cudaStream_t st = ctx->stream;
  const uint64_t n = db.n_targets;
  if (n >= 0xFFFF0000ull) { set_error("database too large for 32-bit target indices"); return FF_EUNSUPPORTED; }
  const bool sorted_db = !db.pack.five_prime;  // 3'-PAM: database order == lexicographic order of the target string
  {
    unsigned int *d_bad = nullptr, h_bad[2] = {0, 0};
    FF_CUDA(cudaMalloc(&d_bad, 8));
    FF_CUDA(cudaMemsetAsync(d_bad, 0, 8, st));
    if (n > 0) k_check_sorted<<<nblk(n), 256, 0, st>>>(db.d_targets, n, sorted_db ? 1 : 0, d_bad);
    FF_CUDA(cudaMemcpyAsync(h_bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_bad);
    if (h_bad[0] || h_bad[1]) {
      set_error("database targets are not strictly sorted / have invalid counts (%u order, %u count violations)", h_bad[0], h_bad[1]);
      return FF_EFORMAT;
    }
  }
  // the compared bases: a contiguous run of P bases at bit proto_shift (every pack of StandardScanParameters.scala)
  db.proto_bases = __builtin_popcountll(db.pack.cmp_mask) / 2;
  db.proto_shift = __builtin_ctzll(db.pack.cmp_mask);
  const int P = db.proto_bases;
  int a = (P + 2) / 2;  // 11 | 9 for the 20-mers, 10 | 9 for the 19-mers
  { const int v = ctx->opt.split_a; if (v >= 4 && v <= 12 && P - v >= 4 && P - v <= 12) a = v; }
  const int b = P - a;
  db.device_bytes = n * 8;

  uint32_t *d_ka = nullptr, *d_kb = nullptr, *d_iota = nullptr;
  FF_CUDA(cudaMalloc(&d_ka, (n + 1) * 4));
  FF_CUDA(cudaMalloc(&d_kb, (n + 1) * 4));
  FF_CUDA(cudaMalloc(&d_iota, (n + 1) * 4));
  if (n) k_split_proto<<<nblk(n), 256, 0, st>>>(db.d_targets, n, db.proto_shift, 2 * b, (1ull << (2 * P)) - 1ull, d_ka, d_kb, d_iota);
  int rc = build_seed_index(ctx, &db.A, a, b, /*wide_planes=*/false, d_ka, d_kb, d_iota, n, /*identity=*/sorted_db);
  if (rc == FF_OK) rc = build_seed_index(ctx, &db.B, b, a, /*wide_planes=*/true, d_kb, d_ka, d_iota, n, /*identity=*/false);
  cudaFree(d_ka); cudaFree(d_kb); cudaFree(d_iota);
  FF_TRY(rc);

  if (db.d_positions) {
    FF_CUDA(cudaMalloc(&db.d_pos_off, (n + 1) * 8));
    uint64_t *d_c = nullptr;
    FF_CUDA(cudaMalloc(&d_c, (n + 1) * 8));
    k_counts<<<nblk(n + 1), 256, 0, st>>>(db.d_targets, n, d_c);
    size_t tmp = 0;
    FF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_c, db.d_pos_off, n + 1, st));
    FF_TRY(ctx->cub_tmp.reserve(tmp));
    FF_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, d_c, db.d_pos_off, n + 1, st));
    uint64_t total = 0;
    FF_CUDA(cudaMemcpyAsync(&total, db.d_pos_off + n, 8, cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_c);
    if (total != db.n_positions) {
      set_error("position count %llu does not match the sum of target counts %llu", (unsigned long long)db.n_positions, (unsigned long long)total);
      return FF_EFORMAT;
    }
    db.device_bytes += (n + 1) * 8 + db.n_positions * 8;
  }
  FF_CUDA(cudaGetLastError());
  db.resident = true;
  return FF_OK;
}

// d_cell_off[key_b * (kCells + 1) + c] = first entry of B bucket key_b whose database index is >= the first database
// index of cell c (cell c = index-A keys [c, c + 1) * 4^a / kCells); entries of a B bucket are in database order
// (stable sort), so a window of cells is a contiguous run of every bucket.
__global__ void k_cell_offsets(const uint32_t *__restrict__ off_a, const uint32_t *__restrict__ off_b, const uint32_t *__restrict__ canon_b,
                               uint32_t n_keys_b, uint32_t keys_per_cell, uint32_t *__restrict__ out) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= (uint64_t)n_keys_b * (kCells + 1)) return;
  const uint32_t kb = (uint32_t)(i / (kCells + 1)), c = (uint32_t)(i % (kCells + 1));
  const uint32_t want = off_a[c * keys_per_cell];  // first database index of cell c (c == kCells: n_targets)
  uint32_t lo = off_b[kb], hi = off_b[kb + 1];
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (canon_b[mid] < want) lo = mid + 1; else hi = mid;
  }
  out[i] = lo;
}

int db_build_cell_offsets(ff_ctx *ctx) {
  Database &db = ctx->db;
  if (db.d_cell_off) return FF_OK;
  if (db.pack.five_prime || db.A.d_canon || db.A.key_bases < 3) { set_error("database-order windows need a 3'-PAM database"); return FF_EUNSUPPORTED; }
  const uint32_t n_keys_b = 1u << (2 * db.B.key_bases);
  const uint32_t keys_per_cell = (1u << (2 * db.A.key_bases)) / kCells;
  const uint64_t n = (uint64_t)n_keys_b * (kCells + 1);
  FF_CUDA(cudaMalloc(&db.d_cell_off, n * 4));
  k_cell_offsets<<<nblk(n), 256, 0, ctx->stream>>>(db.A.d_off, db.B.d_off, db.B.d_canon, n_keys_b, keys_per_cell, db.d_cell_off);
  FF_CUDA(cudaStreamSynchronize(ctx->stream));
  FF_CUDA(cudaGetLastError());
  db.device_bytes += n * 4;
  return FF_OK;
}

int db_from_host_arrays(ff_ctx *ctx, const Pack &pack, int bin_width, const uint64_t *targets, uint64_t n_targets,
                        const uint64_t *positions, uint64_t n_positions, const std::vector<std::string> &contigs) {
  ctx->db.release();
  ctx->host_targets_n = 0;  // the host mirror of the previous database is stale
  Database &db = ctx->db;
  db.pack = pack; db.bin_width = bin_width; db.n_targets = n_targets; db.n_positions = positions ? n_positions : 0;
  db.contigs = contigs;
  FF_CUDA(cudaMalloc(&db.d_targets, (n_targets + 1) * 8));
  FF_CUDA(cudaMemcpyAsync(db.d_targets, targets, n_targets * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (positions) {
    FF_CUDA(cudaMalloc(&db.d_positions, (n_positions + 1) * 8));
    FF_CUDA(cudaMemcpyAsync(db.d_positions, positions, n_positions * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  FF_CUDA(cudaStreamSynchronize(ctx->stream));
  int rc = db_build_index(ctx);
  if (rc != FF_OK) db.release();
  return rc;
}

// ------------------------------------------------------------------------------------------------------------
// FlashFry files
struct HeaderInfo {
  int enzyme = 0, bin_width = 0;
  std::vector<uint64_t> vptr;
  std::vector<int64_t> nbytes, ntargets;
  std::vector<std::string> contigs;
};

// BinaryHeader.readHeader :115-160 (text: magic, version, enzyme index, bin count, BIN=vptr,bytes,targets lines, contigs)
static int read_header(const char *path, HeaderInfo *h) {
  std::ifstream in(path);
  if (!in) { set_error("cannot open header %s", path); return FF_EIO; }
  std::string line;
  auto next = [&](std::string *s) { return (bool)std::getline(in, *s); };
  if (!next(&line) || strtoull(line.c_str(), nullptr, 10) != 0x1234ABCDE123890ull) {
    set_error("Binary file %s doesn't have the magic number expected at the top of the file", path);
    return FF_EFORMAT;
  }
  if (!next(&line) || strtoll(line.c_str(), nullptr, 10) != 1) {
    set_error("Binary file %s doesn't have the correct version, expecting 1", path);
    return FF_EFORMAT;
  }
  if (!next(&line)) { set_error("truncated header %s", path); return FF_EFORMAT; }
  h->enzyme = atoi(line.c_str());
  if (!next(&line)) { set_error("truncated header %s", path); return FF_EFORMAT; }
  const long long bin_count = strtoll(line.c_str(), nullptr, 10);
  h->bin_width = (int)(std::log((double)bin_count) / std::log(4.0));  // :131-132
  if (bin_count <= 0 || (1ll << (2 * h->bin_width)) != bin_count) { set_error("bad bin count %lld in %s", bin_count, path); return FF_EFORMAT; }
  h->vptr.resize(bin_count); h->nbytes.resize(bin_count); h->ntargets.resize(bin_count);
  for (long long b = 0; b < bin_count; ++b) {
    if (!next(&line)) { set_error("Missing line for bin %lld in %s", b, path); return FF_EFORMAT; }
    const size_t eq = line.find('=');
    if (eq == std::string::npos || (int)eq != h->bin_width) { set_error("Failed to verify bin name in header line: %s", line.c_str()); return FF_EFORMAT; }
    long long code = 0;
    for (size_t i = 0; i < eq; ++i) {
      const char *p = strchr("ACGT", line[i]);
      if (!p) { set_error("bad bin name in header line: %s", line.c_str()); return FF_EFORMAT; }
      code = code * 4 + (p - "ACGT");
    }
    if (code != b) { set_error("Failed to verify bin name, line %s is not bin %lld", line.c_str(), b); return FF_EFORMAT; }
    unsigned long long v = 0; long long nb = 0, nt = 0;
    if (sscanf(line.c_str() + eq + 1, "%llu,%lld,%lld", &v, &nb, &nt) != 3) { set_error("bad header line: %s", line.c_str()); return FF_EFORMAT; }
    h->vptr[b] = v; h->nbytes[b] = nb; h->ntargets[b] = nt;
  }
  while (next(&line)) {
    if (line.empty()) continue;
    h->contigs.push_back(line.substr(0, line.find('=')));
  }
  return FF_OK;
}

// The BGZF body is mapped, not read: a human-sized database is ~4 GB on disk, the inflate threads page it in as they
// go and nothing is copied.
struct MappedFile {
  const uint8_t *p = nullptr;
  size_t n = 0;
  int fd = -1;
  int open_ro(const char *path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) { set_error("cannot open %s", path); return FF_EIO; }
    struct stat st;
    if (fstat(fd, &st) != 0) { set_error("cannot stat %s", path); return FF_EIO; }
    n = (size_t)st.st_size;
    if (n == 0) return FF_OK;
    void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { set_error("cannot map %s", path); return FF_EIO; }
    madvise(m, n, MADV_SEQUENTIAL);
    p = static_cast<const uint8_t *>(m);
    return FF_OK;
  }
  size_t size() const { return n; }
  const uint8_t *data() const { return p; }
  uint8_t operator[](size_t i) const { return p[i]; }
  ~MappedFile() {
    if (p) munmap(const_cast<uint8_t *>(p), n);
    if (fd >= 0) ::close(fd);
  }
};

struct Member { size_t off, len, hdr, isize, out_off; };

// Walk the BGZF members through their BSIZE fields (SAM spec 4.1).
static int scan_members(const MappedFile &raw, std::vector<Member> *ms) {
  size_t off = 0, out = 0;
  while (off < raw.size()) {
    if (off + 18 > raw.size() || raw[off] != 0x1f || raw[off + 1] != 0x8b || raw[off + 2] != 8 || !(raw[off + 3] & 4)) {
      set_error("not a BGZF member at byte %zu", off);
      return FF_EFORMAT;
    }
    const size_t xlen = raw[off + 10] | (raw[off + 11] << 8);
    if (off + 12 + xlen + 8 > raw.size()) { set_error("truncated BGZF member at byte %zu", off); return FF_EFORMAT; }
    size_t p = off + 12, end = off + 12 + xlen;
    long bsize = -1;
    while (p + 4 <= end) {
      const size_t slen = raw[p + 2] | (raw[p + 3] << 8);
      if (raw[p] == 'B' && raw[p + 1] == 'C' && slen == 2) bsize = raw[p + 4] | (raw[p + 5] << 8);
      p += 4 + slen;
    }
    if (bsize < 0 || off + bsize + 1 > raw.size() || (size_t)bsize + 1 < 12 + xlen + 8) { set_error("BGZF member without a valid BC field at byte %zu", off); return FF_EFORMAT; }
    Member m;
    m.off = off; m.len = (size_t)bsize + 1; m.hdr = 12 + xlen;
    const uint8_t *tail = raw.data() + off + m.len - 4;
    m.isize = tail[0] | (tail[1] << 8) | (tail[2] << 16) | ((size_t)tail[3] << 24);
    m.out_off = out;
    out += m.isize;
    ms->push_back(m);
    off += m.len;
  }
  return FF_OK;
}

int db_load_files(ff_ctx *ctx, const char *db_path, const char *header_path) {
  HeaderInfo h;
  FF_TRY(read_header(header_path, &h));
  Pack pack;
  FF_TRY(pack_from_index(h.enzyme, &pack));
  MappedFile raw;
  FF_TRY(raw.open_ro(db_path));
  std::vector<Member> ms;
  FF_TRY(scan_members(raw, &ms));
  const size_t total = ms.empty() ? 0 : ms.back().out_off + ms.back().isize;
  std::unique_ptr<uint8_t[]> payload(new uint8_t[total + 8]);  // (uninitialised: every byte is written by an inflate)

  // inflate members on all host threads (they are independent gzip members)
  unsigned nthreads = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
  std::vector<int> err(nthreads, 0);
  {
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < nthreads; ++w) {
      pool.emplace_back([&, w]() {
        z_stream zs;
        for (size_t i = w; i < ms.size(); i += nthreads) {
          const Member &m = ms[i];
          if (m.isize == 0) continue;
          memset(&zs, 0, sizeof zs);
          if (inflateInit2(&zs, -15) != Z_OK) { err[w] = 1; return; }
          zs.next_in = const_cast<Bytef *>(raw.data() + m.off + m.hdr);
          zs.avail_in = (uInt)(m.len - m.hdr - 8);
          zs.next_out = payload.get() + m.out_off;
          zs.avail_out = (uInt)m.isize;
          const int rc = inflate(&zs, Z_FINISH);
          inflateEnd(&zs);
          if (rc != Z_STREAM_END || zs.total_out != m.isize) { err[w] = 1; return; }
        }
      });
    }
    for (auto &t : pool) t.join();
  }
  for (int e : err) if (e) { set_error("BGZF inflate failed for %s", db_path); return FF_EFORMAT; }

  // virtual pointer -> payload offset (member file offset << 16 | offset inside the member's payload)
  const size_t n_bins = h.vptr.size();
  std::vector<size_t> bin_start(n_bins);
  {
    size_t mi = 0;
    for (size_t b = 0; b < n_bins; ++b) {
      const uint64_t coff = h.vptr[b] >> 16, uoff = h.vptr[b] & 0xFFFF;
      while (mi < ms.size() && ms[mi].off < coff) ++mi;
      size_t k = mi;
      if (k >= ms.size() || ms[k].off != coff) {  // header order is normally monotone; fall back to a search
        auto it = std::lower_bound(ms.begin(), ms.end(), coff, [](const Member &m, uint64_t v) { return m.off < v; });
        if (it == ms.end() || it->off != coff) { set_error("bin %zu: virtual pointer does not address a BGZF member", b); return FF_EFORMAT; }
        k = (size_t)(it - ms.begin());
      }
      bin_start[b] = ms[k].out_off + uoff;
      if (h.nbytes[b] < 8 || (h.nbytes[b] & 7) || bin_start[b] + (size_t)h.nbytes[b] > total) {
        set_error("bin %zu: block of %lld bytes does not fit the inflated database", b, (long long)h.nbytes[b]);
        return FF_EFORMAT;
      }
    }
  }

  // pass 1: per-bin target / position counts (walk [target, pos x count]*, BlockManager.scala:225-253)
  std::vector<uint64_t> bin_t(n_bins + 1, 0), bin_p(n_bins + 1, 0);
  std::vector<int> bad(nthreads, 0);
  auto for_bins = [&](auto fn) {
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < nthreads; ++w)
      pool.emplace_back([&, w]() {
        const size_t lo = n_bins * w / nthreads, hi = n_bins * (w + 1) / nthreads;
        for (size_t b = lo; b < hi; ++b) fn(b, w);
      });
    for (auto &t : pool) t.join();
  };
  auto block_body = [&](size_t b, const uint64_t **body, size_t *n_longs) -> bool {
    const uint64_t *blk = reinterpret_cast<const uint64_t *>(payload.get() + bin_start[b]);  // little-endian host
    uint64_t first;
    memcpy(&first, blk, 8);
    const size_t nl = (size_t)h.nbytes[b] / 8;
    if (first == 1) { *body = blk + 1; *n_longs = nl - 1; return true; }
    if (first == 2 && nl >= 257) { *body = blk + 257; *n_longs = nl - 257; return true; }  // 256-entry sub-bin table skipped
    return false;  // "Invalid bin type" BlockManager.scala:85-87
  };
  for_bins([&](size_t b, unsigned w) {
    const uint64_t *body = nullptr; size_t nl = 0;
    if (!block_body(b, &body, &nl)) { bad[w] = 1; return; }
    uint64_t nt = 0, np = 0, v;
    for (size_t i = 0; i < nl;) {
      memcpy(&v, body + i, 8);
      const uint64_t c = v >> 48;
      if (c == 0 || c > 32767 || i + 1 + c > nl) { bad[w] = 2; return; }
      nt++; np += c; i += 1 + c;
    }
    bin_t[b + 1] = nt; bin_p[b + 1] = np;
  });
  for (int e : bad) if (e) { set_error(e == 1 ? "Invalid bin type, unknown value in a block header" : "Failed to correctly parse block: position entries exceed the block"); return FF_EFORMAT; }
  for (size_t b = 0; b < n_bins; ++b) {
    if ((int64_t)bin_t[b + 1] != h.ntargets[b]) { set_error("bin %zu: header says %lld targets, block holds %llu", b, (long long)h.ntargets[b], (unsigned long long)bin_t[b + 1]); return FF_EFORMAT; }
    bin_t[b + 1] += bin_t[b]; bin_p[b + 1] += bin_p[b];
  }
  const uint64_t n_targets = bin_t[n_bins], n_positions = bin_p[n_bins];
  std::unique_ptr<uint64_t[]> targets(new uint64_t[n_targets + 1]), positions(new uint64_t[n_positions + 1]);
  // pass 2: fill
  for_bins([&](size_t b, unsigned) {
    const uint64_t *body = nullptr; size_t nl = 0;
    block_body(b, &body, &nl);
    uint64_t ti = bin_t[b], pi = bin_p[b], v;
    for (size_t i = 0; i < nl;) {
      memcpy(&v, body + i, 8);
      const uint64_t c = v >> 48;
      targets[ti++] = v;
      memcpy(&positions[pi], body + i + 1, c * 8);
      pi += c; i += 1 + c;
    }
  });
  payload.reset();  // the inflated stream is no longer needed: release it before the upload
  return db_from_host_arrays(ctx, pack, h.bin_width, targets.get(), n_targets, positions.get(), n_positions, h.contigs);
}

// ------------------------------------------------------------------------------------------------------------
// SoA image side-car: the decoded database as flat little-endian arrays.  Loading it is one
// sequential read + H2D + index build, without the BGZF inflate and the [target, position x count]* block walk that
// dominate a cold start from FlashFry's own files.  Layout: 64-byte header {magic "FFB200IM", u32 version, i32 enzyme,
// i32 bin width, u32 contigs, u64 targets, u64 positions, u64 bytes of the contig table}, the contig table (NUL-
// terminated names), zero padding to a multiple of 8, targets u64[], positions u64[].
static const char kImageMagic[8] = {'F', 'F', 'B', '2', '0', '0', 'I', 'M'};
struct ImageHeader {
  char magic[8];
  uint32_t version;
  int32_t enzyme_index, bin_width;
  uint32_t n_contigs;
  uint64_t n_targets, n_positions, contig_bytes;
  uint8_t pad[16];
};
static_assert(sizeof(ImageHeader) == 64, "image header is 64 bytes");

int db_save_image(ff_ctx *ctx, const char *path) {
  Database &db = ctx->db;
  if (!db.resident) { set_error("no database resident in this context"); return FF_ENODB; }
  std::vector<uint64_t> targets(db.n_targets + 1), positions(db.n_positions + 1);
  FF_CUDA(cudaMemcpy(targets.data(), db.d_targets, db.n_targets * 8, cudaMemcpyDeviceToHost));
  if (db.d_positions && db.n_positions) FF_CUDA(cudaMemcpy(positions.data(), db.d_positions, db.n_positions * 8, cudaMemcpyDeviceToHost));
  std::string names;
  for (const std::string &c : db.contigs) { names += c; names.push_back('\0'); }
  while (names.size() % 8) names.push_back('\0');
  ImageHeader h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, kImageMagic, 8);
  h.version = 1; h.enzyme_index = db.pack.enzyme_index; h.bin_width = db.bin_width; h.n_contigs = (uint32_t)db.contigs.size();
  h.n_targets = db.n_targets; h.n_positions = db.d_positions ? db.n_positions : 0; h.contig_bytes = names.size();
  FILE *f = fopen(path, "wb");
  if (!f) { set_error("cannot create %s", path); return FF_EIO; }
  bool ok = fwrite(&h, sizeof h, 1, f) == 1;
  ok = ok && (names.empty() || fwrite(names.data(), 1, names.size(), f) == names.size());
  ok = ok && (h.n_targets == 0 || fwrite(targets.data(), 8, h.n_targets, f) == h.n_targets);
  ok = ok && (h.n_positions == 0 || fwrite(positions.data(), 8, h.n_positions, f) == h.n_positions);
  ok = (fclose(f) == 0) && ok;
  if (!ok) { set_error("short write on %s", path); return FF_EIO; }
  return FF_OK;
}

int db_load_image(ff_ctx *ctx, const char *path) {
  MappedFile f;  // the arrays are uploaded straight from the mapping: no intermediate copies on the host
  FF_TRY(f.open_ro(path));
  ImageHeader h;
  if (f.size() < sizeof h) { set_error("%s is not a flashfry_b200 database image (bad magic or version)", path); return FF_EFORMAT; }
  memcpy(&h, f.data(), sizeof h);
  if (memcmp(h.magic, kImageMagic, 8) != 0 || h.version != 1) {
    set_error("%s is not a flashfry_b200 database image (bad magic or version)", path);
    return FF_EFORMAT;
  }
  Pack pack;
  if (pack_from_index(h.enzyme_index, &pack) != FF_OK) return FF_EFORMAT;
  // sizes come from the file: check them against its length before trusting any of them
  if (h.contig_bytes > (1u << 28) || (h.contig_bytes & 7) || h.n_targets > 0xFFFF0000ull || h.n_positions > h.n_targets * 32767ull) {
    set_error("implausible sizes in image %s", path);
    return FF_EFORMAT;
  }
  if (sizeof h + h.contig_bytes + 8ull * h.n_targets + 8ull * h.n_positions > f.size()) { set_error("truncated database image %s", path); return FF_EFORMAT; }
  const char *names = reinterpret_cast<const char *>(f.data() + sizeof h);
  std::vector<std::string> contigs;
  for (size_t i = 0, k = 0; k < h.n_contigs && i < h.contig_bytes; ++k) {
    const size_t len = strnlen(names + i, h.contig_bytes - i);
    if (i + len >= h.contig_bytes) break;  // a name must end inside the table
    contigs.emplace_back(names + i, len);
    i += len + 1;
  }
  if (contigs.size() != h.n_contigs) { set_error("contig table of %s is damaged", path); return FF_EFORMAT; }
  const uint64_t *targets = reinterpret_cast<const uint64_t *>(f.data() + sizeof h + h.contig_bytes);
  const uint64_t *positions = targets + h.n_targets;
  // the index build re-validates order and counts on the device (FF_EFORMAT on a damaged image)
  return db_from_host_arrays(ctx, pack, h.bin_width, targets, h.n_targets, h.n_positions ? positions : nullptr, h.n_positions, contigs);
}

// ------------------------------------------------------------------------------------------------------------
// synthetic database for benchmarks (SURVEY.md 8(d) cfg 3): uniform random protospacer + N, fixed PAM
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void k_synth_values(uint64_t n, uint64_t seed, int random_bits, uint64_t pam_bits, uint64_t *__restrict__ out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t r = splitmix64(seed * 0x2545F4914F6CDD1Dull + i);
  out[i] = ((r >> (64 - random_bits)) << 4) | pam_bits;
}

// occurrence counts: 94 % = 1, else 1 + Geometric(0.3); ~2000 "repeat families" with count U[500, 32767]
__global__ void k_synth_counts(uint64_t n, uint64_t seed, uint64_t family_every, uint64_t *__restrict__ t) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t v = t[i];
  const uint64_t h = splitmix64(v ^ (seed << 1));
  uint64_t c = 1;
  if ((h % 100) >= 94) {
    uint64_t g = splitmix64(h);
    while (c < 64 && (g & 1023) >= 307) { c++; g = splitmix64(g); }  // continue with p = 0.7
    c += 1;
  }
  if (family_every && (splitmix64(h ^ 0xABCDEF) % family_every) == 0) c = 500 + splitmix64(h ^ 0x123457) % 32268;
  t[i] = v | (c << 48);
}

// repeat NEIGHBOURHOODS: member i of family f = the family's consensus with 0..max_subs random substitutions in the
// protospacer (sequence-level skew: hot buckets, long candidate lists), written over the tail of the uniform values
__global__ void k_synth_families(uint64_t n_fam, uint64_t fam_size, int max_subs, uint64_t seed, int random_bits, uint64_t pam_bits,
                                 uint64_t *__restrict__ out) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n_fam * fam_size) return;
  const uint64_t f = i / fam_size;
  uint64_t v = splitmix64(seed * 0x9E3779B97F4A7C15ull + 0xFA111E5ull + f) >> (64 - random_bits);
  uint64_t h = splitmix64(seed ^ (i * 0xD6E8FEB86659FD93ull));
  const int subs = (int)(h % (uint64_t)(max_subs + 1));
  const int proto_bases = random_bits / 2 - 1;  // the last random base is the PAM's N
  for (int s = 0; s < subs; ++s) {
    h = splitmix64(h);
    const int pos = (int)(h % (uint64_t)proto_bases);
    const uint64_t x = 1ull + ((h >> 32) % 3ull);
    v ^= x << (2 * (random_bits / 2 - 1 - pos));
  }
  out[i] = (v << 4) | pam_bits;
}

int db_synth(ff_ctx *ctx, const Pack &pack, uint64_t n_targets, uint64_t seed, uint64_t n_families, uint64_t family_size, int family_subs) {
  if (pack.five_prime) { set_error("synthetic databases are spCas9-family only"); return FF_EUNSUPPORTED; }
  if (n_targets == 0 || n_targets > 3000000000ull) { set_error("bad synthetic database size"); return FF_EINVAL; }
  ctx->db.release();
  ctx->host_targets_n = 0;
  Database &db = ctx->db;
  cudaStream_t st = ctx->stream;
  const int random_bits = 2 * (pack.scan_len - 2);                     // protospacer + N
  const uint64_t pam_bits = pack.enzyme_index == 4 ? 0x2ull : 0xAull;  // AG : GG
  uint64_t *d_a = nullptr, *d_b = nullptr, *d_n = nullptr;
  FF_CUDA(cudaMalloc(&d_a, (n_targets + 1) * 8));
  FF_CUDA(cudaMalloc(&d_b, (n_targets + 1) * 8));
  FF_CUDA(cudaMalloc(&d_n, 8));
  k_synth_values<<<(unsigned int)((n_targets + 255) / 256), 256, 0, st>>>(n_targets, seed, random_bits, pam_bits, d_a);
  if (n_families * family_size > 0) {
    if (n_families * family_size > n_targets / 2) { cudaFree(d_a); cudaFree(d_b); cudaFree(d_n); set_error("families larger than half the database"); return FF_EINVAL; }
    const uint64_t nf = n_families * family_size;
    k_synth_families<<<(unsigned int)((nf + 255) / 256), 256, 0, st>>>(n_families, family_size, family_subs, seed, random_bits, pam_bits,
                                                                       d_a + (n_targets - nf));
  }
  size_t tmp = 0;
  FF_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp, d_a, d_b, n_targets, 0, 2 * pack.scan_len, st));
  FF_TRY(ctx->cub_tmp.reserve(tmp));
  FF_CUDA(cub::DeviceRadixSort::SortKeys(ctx->cub_tmp.p, tmp, d_a, d_b, n_targets, 0, 2 * pack.scan_len, st));
  FF_CUDA(cub::DeviceSelect::Unique(nullptr, tmp, d_b, d_a, d_n, n_targets, st));
  FF_TRY(ctx->cub_tmp.reserve(tmp));
  FF_CUDA(cub::DeviceSelect::Unique(ctx->cub_tmp.p, tmp, d_b, d_a, d_n, n_targets, st));
  uint64_t n_unique = 0;
  FF_CUDA(cudaMemcpyAsync(&n_unique, d_n, 8, cudaMemcpyDeviceToHost, st));
  FF_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_b); cudaFree(d_n);
  const uint64_t family_every = n_unique > 4000 ? n_unique / 2000 : 0;
  k_synth_counts<<<(unsigned int)((n_unique + 255) / 256), 256, 0, st>>>(n_unique, seed, family_every, d_a);
  FF_CUDA(cudaStreamSynchronize(st));
  FF_CUDA(cudaGetLastError());
  db.pack = pack; db.bin_width = 7; db.n_targets = n_unique; db.n_positions = 0;
  db.d_targets = d_a;
  int rc = db_build_index(ctx);
  if (rc != FF_OK) db.release();
  return rc;
}

}  // namespace ff
