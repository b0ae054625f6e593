// Probe (not part of the library): which TMEM cells does tcgen05.ld.16x128b hand to which thread?  Cells are filled through the
// plain 32x32b store with value = 1000 * lane + column; every warp then reads its quadrant with .16x128b.x1 / .x4 at lane base
// 0 and 16 and the decoded (lane, column) pairs of a few threads are printed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nsfnet_b200/csrc scripts/probe_ld16.cu -o scripts/_bin/probe_ld16
#include <cstdio>
#include <cstdlib>
#include "nsf_tc.cuh"
using namespace nsftc;
#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(128) probe(float* out /* [4 warps][2 halves][32 lanes][10 regs] */) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_base, 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base;
  const uint32_t qaddr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 32; c += 8) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = 1000.f * (warp * 32 + lane) + (c + i);
    tmem_st8(qaddr + c, v);
  }
  tmem_st_wait();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  for (int half = 0; half < 2; ++half) {
    const uint32_t a = qaddr + ((uint32_t)(half * 16) << 16);
    uint32_t r[10];
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a));
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(r[8]), "=r"(r[9]) : "r"(a + 16));
    tmem_ld_wait();
    for (int i = 0; i < 10; ++i) out[((warp * 2 + half) * 32 + lane) * 10 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  float* d; CK(cudaMalloc(&d, 4 * 2 * 32 * 10 * sizeof(float)));
  probe<<<1, 128>>>(d);
  CK(cudaDeviceSynchronize());
  static float h[4 * 2 * 32 * 10];
  CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  for (int w = 0; w < 2; ++w)
    for (int half = 0; half < 2; ++half) {
      printf("warp %d lane base %d: thread -> (lane, col) of regs r0..r7 (.x4 at column 0), r8 r9 (.x1 at column 16)\n", w, half * 16);
      for (int l = 0; l < 32; ++l) {
        if (!(l < 6 || l == 31)) continue;
        printf("  t%02d:", l);
        for (int i = 0; i < 10; ++i) { const int v = (int)h[((w * 2 + half) * 32 + l) * 10 + i]; printf(" (%d,%d)", v / 1000, v % 1000); }
        printf("\n");
      }
    }
  // consistency check of the expected map: reg 2u + b of thread t = (lane base + t / 4 + 8 b, column 4 u + t % 4)
  int bad = 0;
  for (int w = 0; w < 4; ++w) for (int half = 0; half < 2; ++half) for (int l = 0; l < 32; ++l) for (int i = 0; i < 10; ++i) {
    const int v = (int)h[((w * 2 + half) * 32 + l) * 10 + i];
    const int u = i < 8 ? i / 2 : 4, b = i & 1;
    const int el = w * 32 + half * 16 + l / 4 + 8 * b, ec = 4 * u + l % 4;
    if (v != 1000 * el + ec) ++bad;
  }
  printf("expected map (reg 2u+b of thread t = lane base + t/4 + 8b, column 4u + t%%4): %d mismatches of %d\n", bad, 4 * 2 * 32 * 10);
  return 0;
}
