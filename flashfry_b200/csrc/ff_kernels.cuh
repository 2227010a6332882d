// Device helpers and cross-file entry points of libflashfry_b200.
#pragma once

#include "ff_common.cuh"

namespace ff {

// bitcoding/BitEncoding.scala:127-132 restated for the device: per-base mismatch count of two target longs under
// the enzyme's comparison mask (count bits and PAM are masked away).
__host__ __device__ __forceinline__ int mismatches64(uint64_t a, uint64_t b, uint64_t cmp_mask) {
  uint64_t x = (a ^ b) & cmp_mask & 0xFFFFFFFFFFFFull;
  x = (x | (x >> 1)) & 0x555555555555ull;
#ifdef __CUDA_ARCH__
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}

// Results of a discover call that are still in HBM (owned by the context's workspaces).
struct DeviceResult {
  int64_t n_guides = 0, n_hits = 0, n_positions = 0;
  uint64_t n_candidate_hits = 0, n_compares = 0;
  const int64_t *d_row_ptr = nullptr;
  const uint64_t *d_targets = nullptr;
  const uint8_t *d_mismatches = nullptr;
  const uint8_t *d_bulge = nullptr;  // bulge mode only
  const uint32_t *d_tidx = nullptr;  // database index of every emitted hit (plain path)
  const int32_t *d_total_count = nullptr;
  const uint8_t *d_overflowed = nullptr;
  const int64_t *d_pos_ptr = nullptr;
  const uint64_t *d_positions = nullptr;
};

// ff_discover.cu
// bulge_flags: 0 = the reference's mismatch-only search; FF_BULGE_RNA | FF_BULGE_DNA = the 1-bp bulge extension
int discover_on_device(ff_ctx *ctx, const uint64_t *d_guides, int64_t n_guides, int max_mm, int max_ot,
                       bool want_positions, int bulge_flags, int slot, DeviceResult *res);

// database-sharded discover (ff_shard.inl): all guides in, this rank's guides' rows out; ff_peer_attach must have run
int discover_sharded(ff_ctx *ctx, const uint64_t *d_guides_all, int64_t n_guides_all, int max_mm, int max_ot, int slot, DeviceResult *res);
// all workspaces of such a call; warm: every kernel of the path loaded too (ranks sharing one device: see ff_shard.inl)
int discover_sharded_reserve(ff_ctx *ctx, int64_t n_guides_all, int max_mm, bool warm);

// ff_score.cu : CFD + Hsu2013 over a CSR hit list resident in HBM.  Any output may be null.
int score_on_device(ff_ctx *ctx, const uint64_t *d_guides, int64_t n_guides, const int64_t *d_row_ptr,
                    const uint64_t *d_targets, int64_t n_hits, uint32_t metrics, double *d_cfd_max,
                    double *d_cfd_spec, double *d_hsu, double *d_per_ot_cfd);

int hit_aggregates_on_device(ff_ctx *ctx, const uint64_t *d_guides, int64_t n_guides, const int64_t *d_row_ptr, const uint64_t *d_targets,
                             uint64_t cmp_mask, int32_t *d_out);

// ff_db.cu
int db_from_host_arrays(ff_ctx *ctx, const Pack &pack, int bin_width, const uint64_t *targets, uint64_t n_targets,
                        const uint64_t *positions, uint64_t n_positions, const std::vector<std::string> &contigs);
int db_build_index(ff_ctx *ctx);  // d_targets (+ d_positions) already in HBM -> d_tlow, d_sub_off, tables, d_pos_off
int db_build_cell_offsets(ff_ctx *ctx);  // lazily: per-bucket offsets at the database-order cell boundaries (windowed scan)
int db_save_image(ff_ctx *ctx, const char *path);  // SoA side-car of the resident database
int db_load_image(ff_ctx *ctx, const char *path);
int db_load_files(ff_ctx *ctx, const char *db_path, const char *header_path);
int db_synth(ff_ctx *ctx, const Pack &pack, uint64_t n_targets, uint64_t seed, uint64_t n_families = 0, uint64_t family_size = 0,
             int family_subs = 3);

}  // namespace ff
