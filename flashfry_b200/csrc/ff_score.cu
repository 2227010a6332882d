// Fused Doench-2016 CFD + Hsu-2013 aggregate off-target scoring over a hit list that is resident in HBM.
//
// Replaces (FlashFry, src/main/scala/...):
//   Doench2016CFDScore.scoreGuide / scoreCFD      scoring/Doench2016CFDScore.scala:53-88,132-151, tables :173-214
//   CrisprMitEduOffTarget.scoreOffTarget / getScore scoring/CrisprMitEduOffTarget.scala:85-148, coefficients :43-53
//
// One warp per guide.  Phase 1: lanes take the guide's off-targets and compute the per-off-target CFD and Hsu terms
// straight from the 2-bit codes (no strings, no hash maps); products run in ascending position order with
// round-to-nearest multiplies (no FMA contraction) so every term is bit-identical to the JVM's.  Phase 2: lane 0
// folds the terms left to right -- the reference's sums are sequential (`.sum` == foldLeft from 0.0), and
// reassociating them would change the last bits.
#include "ff_common.cuh"
#include "ff_kernels.cuh"

#define FF_TABLE_QUAL static __device__ const
#include "score_tables.h"

namespace ff {

__device__ __forceinline__ int base23(uint64_t enc, int i) { return (int)((enc >> (2 * (22 - i))) & 3ull); }

struct ScoreTables {
  double mm[20][4][4];
  double cfd_pam[4][4];
  double hsu_coef[20];
  double hsu_pam[4][4];
};

constexpr int kScoreThreads = 256;

__global__ void __launch_bounds__(kScoreThreads)
k_score(const uint64_t *__restrict__ guides, int64_t n_guides, const int64_t *__restrict__ row_ptr,
        const uint64_t *__restrict__ targets, uint32_t metrics, double *__restrict__ per_cfd, double *__restrict__ per_hsu,
        double *__restrict__ cfd_max, double *__restrict__ cfd_spec, double *__restrict__ hsu_out) {
  __shared__ ScoreTables T;
  {
    double *dst = reinterpret_cast<double *>(&T);
    for (int i = threadIdx.x; i < 320; i += kScoreThreads) dst[i] = (&FF_CFD_MM[0][0][0])[i];
    for (int i = threadIdx.x; i < 16; i += kScoreThreads) (&T.cfd_pam[0][0])[i] = (&FF_CFD_PAM[0][0])[i];
    for (int i = threadIdx.x; i < 20; i += kScoreThreads) T.hsu_coef[i] = FF_HSU_COEF[i];
    for (int i = threadIdx.x; i < 16; i += kScoreThreads) (&T.hsu_pam[0][0])[i] = (&FF_HSU_PAM[0][0])[i];
  }
  __syncthreads();
  const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_guides) return;
  const uint64_t guide = guides[g];
  const int64_t r0 = row_ptr[g], r1 = row_ptr[g + 1];
  const bool do_cfd = metrics & FF_METRIC_CFD, do_hsu = metrics & FF_METRIC_HSU2013;
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);

  // ---- phase 1: per-off-target terms
  for (int64_t h = r0 + lane; h < r1; h += 32) {
    const uint64_t ot = targets[h];
    double cfd = 1.0, p1 = 1.0;
    int mm = 0, last = -1, dist_sum = 0;
#pragma unroll
    for (int i = 0; i < 20; ++i) {
      const int gb = base23(guide, i), ob = base23(ot, i);
      if (gb != ob) {
        cfd = __dmul_rn(cfd, T.mm[i][gb][ob]);                          // Doench2016CFDScore.scala:140-148
        p1 = __dmul_rn(p1, __dsub_rn(1.0, T.hsu_coef[i]));              // CrisprMitEduOffTarget.scala:120
        if (last >= 0) dist_sum += i - last;                            // :123-126
        last = i;
        ++mm;
      }
    }
    const int b21 = base23(ot, 21), b22 = base23(ot, 22);
    if (do_cfd) {
      // :67 an off-target whose 20 protospacer bases equal the guide's is skipped; :69-73 pam * score
      per_cfd[h] = mm == 0 ? qnan : __dmul_rn(T.cfd_pam[b21][b22], cfd);
    }
    if (do_hsu) {
      double v = qnan;  // :90 zero-mismatch off-targets are not scored
      if (mm != 0) {
        double p2 = 1.0;
        if (mm >= 2) {  // :131-134
          const double avg = __ddiv_rn((double)dist_sum, (double)(mm - 1));
          p2 = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(19.0, avg), 19.0), 4.0), 1.0));
        }
        const double p3 = __ddiv_rn(1.0, (double)(mm * mm));            // :137  1/pow(mm,2)
        const double total = __dmul_rn(__dmul_rn(__dmul_rn(p1, p2), p3), 100.0);  // :139
        v = __dmul_rn(total, T.hsu_pam[b21][b22]);                      // :140-147
      }
      per_hsu[h] = v;
    }
  }
  __syncwarp();
  __threadfence_block();

  // ---- phase 2: ordered folds (lane 0 CFD, lane 1 Hsu)
  if (lane == 0 && do_cfd) {
    double sum = 0.0, mx = 0.0;
    bool any = false;
    for (int64_t h = r0; h < r1; ++h) {
      const double s = per_cfd[h];
      if (s != s) continue;
      const double cnt = (double)(int)(targets[h] >> 48);
      sum = __dadd_rn(sum, __dmul_rn(s, cnt));                           // :79
      if (!any || s > mx) mx = s;                                        // :80
      any = true;
    }
    if (cfd_spec) cfd_spec[g] = any ? __ddiv_rn(1.0, __dadd_rn(1.0, sum)) : 1.0;
    if (cfd_max) cfd_max[g] = (any && mx >= 0.023) ? mx : 0.0;          // :83-87
  }
  if (lane == 1 && do_hsu) {
    double sum = 0.0;
    for (int64_t h = r0; h < r1; ++h) {
      const double s = per_hsu[h];
      if (s == s) sum = __dadd_rn(sum, s);                               // :103-105 scores.sum
    }
    if (hsu_out) hsu_out[g] = __dmul_rn(__ddiv_rn(100.0, __dadd_rn(100.0, sum)), 100.0);
  }
}

// scoring/ClosestHit.scala:43-76 and scoring/DangerousSequences.scala:61-65 as order-independent integer reductions:
// one warp per guide, lanes stride its hits.  out: [0] closest (INT32_MAX = none), [1] occurrences at that distance,
// [2..6] occurrence histogram for 0..4 mismatches, [7] occurrences with zero mismatches.
__global__ void k_hit_aggregates(const uint64_t *__restrict__ guides, int64_t n_guides, const int64_t *__restrict__ row_ptr,
                                 const uint64_t *__restrict__ targets, uint64_t cmp_mask, int32_t *__restrict__ out) {
  const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n_guides) return;
  const uint64_t guide = guides[g];
  const int64_t r0 = row_ptr[g], r1 = row_ptr[g + 1];
  int closest = 0x7fffffff;
  int hist[5] = {0, 0, 0, 0, 0};
  for (int64_t h = r0 + lane; h < r1; h += 32) {
    const uint64_t t = targets[h];
    const int mm = mismatches64(guide, t, cmp_mask), c = (int)(int16_t)(t >> 48);
    if (mm <= 4) hist[mm] += c;
    if (mm > 0 && mm < closest) closest = mm;
  }
  for (int o = 16; o > 0; o >>= 1) closest = min(closest, __shfl_xor_sync(0xffffffffu, closest, o));
  int at_closest = 0;  // second walk: occurrences at the winning distance (distances above 4 are not in the histogram)
  for (int64_t h = r0 + lane; h < r1; h += 32) {
    const uint64_t t = targets[h];
    if (mismatches64(guide, t, cmp_mask) == closest) at_closest += (int)(int16_t)(t >> 48);
  }
  for (int o = 16; o > 0; o >>= 1) {
    at_closest += __shfl_xor_sync(0xffffffffu, at_closest, o);
#pragma unroll
    for (int m = 0; m < 5; ++m) hist[m] += __shfl_xor_sync(0xffffffffu, hist[m], o);
  }
  if (lane == 0) {
    int32_t *o = out + g * 8;
    o[0] = closest; o[1] = closest == 0x7fffffff ? 0 : at_closest;
    for (int m = 0; m < 5; ++m) o[2 + m] = hist[m];
    o[7] = hist[0];
  }
}

int hit_aggregates_on_device(ff_ctx *ctx, const uint64_t *d_guides, int64_t n_guides, const int64_t *d_row_ptr, const uint64_t *d_targets,
                             uint64_t cmp_mask, int32_t *d_out) {
  if (n_guides <= 0) return FF_OK;
  const int64_t threads = n_guides * 32;
  k_hit_aggregates<<<(unsigned int)((threads + 255) / 256), 256, 0, ctx->stream>>>(d_guides, n_guides, d_row_ptr, d_targets, cmp_mask, d_out);
  FF_CUDA(cudaGetLastError());
  return FF_OK;
}

int score_on_device(ff_ctx *ctx, const uint64_t *d_guides, int64_t n_guides, const int64_t *d_row_ptr,
                    const uint64_t *d_targets, int64_t n_hits, uint32_t metrics, double *d_cfd_max,
                    double *d_cfd_spec, double *d_hsu, double *d_per_ot_cfd) {
  // validOverEnzyme: Cas9 family with a 23-base scan length (Doench2016CFDScore.scala:96-98, CrisprMitEduOffTarget.scala:156-158)
  if (ctx->db.resident && !(ctx->db.pack.scan_len == 23 && !ctx->db.pack.five_prime)) {
    set_error("CFD / Hsu2013 are only valid for 23-bp Cas9 parameter packs");
    return FF_EUNSUPPORTED;
  }
  if (n_guides <= 0 || metrics == 0) return FF_OK;
  const int64_t Hp = n_hits > 0 ? n_hits : 1;
  double *per_cfd = d_per_ot_cfd;
  if ((metrics & FF_METRIC_CFD) && !per_cfd) {
    FF_TRY(ctx->cfd_per_ot.reserve(Hp * 8));
    per_cfd = ctx->cfd_per_ot.as<double>();
  }
  double *per_hsu = nullptr;
  if (metrics & FF_METRIC_HSU2013) {
    FF_TRY(ctx->hsu_per_ot.reserve(Hp * 8));
    per_hsu = ctx->hsu_per_ot.as<double>();
  }
  const int64_t threads = n_guides * 32;
  k_score<<<(unsigned int)((threads + kScoreThreads - 1) / kScoreThreads), kScoreThreads, 0, ctx->stream>>>(
      d_guides, n_guides, d_row_ptr, d_targets, metrics, per_cfd, per_hsu, d_cfd_max, d_cfd_spec, d_hsu);
  FF_CUDA(cudaGetLastError());
  ctx->last.kernel_launches += 1;
  return FF_OK;
}

}  // namespace ff
