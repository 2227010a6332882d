// Micro-benchmark: achievable HBM read bandwidth on B200 for random runs of R bytes (what k_seed_scan's bucket
// streaming looks like to the memory system), as a function of run length, alignment and loads in flight.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw tools/gather_bw.cu && ./gather_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ldg128(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

// each warp reads `runs_per_warp` random runs of `run_words` 32-bit words starting at (random & ~align_mask) word offset;
// DEPTH runs are requested back to back before their data is consumed
__global__ void k_fill_starts(uint32_t* st, uint64_t n, uint64_t n_words, int run_words, uint32_t align_mask) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) st[i] = (uint32_t)((mix(i * 2654435761ull + run_words) % (n_words - run_words - 256)) & ~(uint64_t)align_mask);
}

template <int DEPTH>
__global__ void k_gather(const uint32_t* __restrict__ buf, const uint32_t* __restrict__ starts, uint64_t n_words, int run_words, uint32_t align_mask, int runs_per_warp,
                         unsigned long long* sink) {
  const int lane = threadIdx.x & 31;
  const uint64_t gw = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  uint32_t acc = 0;
  for (int r = 0; r < runs_per_warp; r += 32) {
   const uint32_t my_start = starts[gw * runs_per_warp + r + lane];   // one coalesced lookup per 32 runs, like the real kernel
   for (int r2 = 0; r2 < 32; r2 += DEPTH) {
    uint4 v[DEPTH];
    uint64_t start[DEPTH];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      start[d] = __shfl_sync(0xffffffffu, my_start, r2 + d);
      const uint64_t a = (start[d] & ~3ull) + 4ull * lane;
      v[d] = make_uint4(0, 0, 0, 0);
      if (a < start[d] + run_words) v[d] = ldg128(buf + a);
    }
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      acc ^= v[d].x ^ v[d].y ^ v[d].z ^ v[d].w;
      for (uint64_t a = (start[d] & ~3ull) + 4ull * lane + 128; a < start[d] + run_words; a += 128) {
        const uint4 w = ldg128(buf + a);
        acc ^= w.x ^ w.y ^ w.z ^ w.w;
      }
    }
   }
  }
  if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

template <int DEPTH>
float run(const uint32_t* buf, uint32_t* starts, uint64_t n_words, int run_words, uint32_t align_mask, int blocks, int runs_per_warp, unsigned long long* sink) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const uint64_t n_st = (uint64_t)blocks * 8 * runs_per_warp;
  k_fill_starts<<<(unsigned)((n_st + 255) / 256), 256>>>(starts, n_st, n_words, run_words, align_mask);
  k_gather<DEPTH><<<blocks, 256>>>(buf, starts, n_words, run_words, align_mask, runs_per_warp, sink);
  cudaEventRecord(e0);
  k_gather<DEPTH><<<blocks, 256>>>(buf, starts, n_words, run_words, align_mask, runs_per_warp, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  const uint64_t n_words = 600ull * 1000 * 1000;  // 2.4 GB
  uint32_t* buf; unsigned long long* sink;
  cudaMalloc(&buf, n_words * 4); cudaMalloc(&sink, 8);
  uint32_t* starts; cudaMalloc(&starts, (size_t)148 * 8 * 8 * 4096 * 4 + 1024);
  cudaMemset(buf, 1, n_words * 4); cudaMemset(sink, 0, 8);
  const int blocks = 148 * 8;
  printf("run_bytes align depth   useful_GB/s  (random runs from a 2.4 GB buffer, %d CTAs x 8 warps)\n", blocks);
  const int runs[] = {32, 72, 128, 288, 1152, 4096};  // words: 128 B, 288 B, 512 B, 1152 B, 4.6 KB, 16 KB
  for (int rw : runs) {
    for (uint32_t am : {0u, 7u, 31u}) {
      const int rpw = rw <= 128 ? 4096 : (rw <= 1152 ? 1024 : 256);
      const double bytes = (double)blocks * 8 * rpw * rw * 4;
      float m1 = run<1>(buf, starts, n_words, rw, am, blocks, rpw, sink);
      float m2 = run<2>(buf, starts, n_words, rw, am, blocks, rpw, sink);
      float m4 = run<4>(buf, starts, n_words, rw, am, blocks, rpw, sink);
      printf("%8d %5u   d1 %8.1f   d2 %8.1f   d4 %8.1f\n", rw * 4, am + 1, bytes / m1 / 1e6, bytes / m2 / 1e6, bytes / m4 / 1e6);
    }
  }
  return 0;
}
