// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, cta_group::1, M = 128, K = 16) for the operand
// configurations the attention kernels use. One CTA per SM, one issuing thread, R back-to-back MMAs + one commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -I llamax_b200/csrc -I include tools/mma_bench.cu -o tools/mma_bench
#include <cstdio>
#include <vector>

#include "common.cuh"

using namespace lx;

struct Cfg {
  int N;          // MMA N
  int a_tmem;     // A operand from TMEM (TS mode)
  int a_mn, b_mn; // MN-major smem operands
  int n_acc;      // number of distinct accumulators cycled through (1 = always the same D)
  int load;       // extra warps hammering TMEM loads (0/1)
};

__global__ void __launch_bounds__(256, 1) mma_bench_kernel(Cfg c, int R, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    stop = 0;
  }
  if (warp == 1) {
    tmem_alloc<1>(&tmem_slot, 512);
    tmem_relinquish<1>();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(1, 1, 128, c.N, c.a_mn, c.b_mn);
      constexpr uint32_t kHi = desc_hi(1024);
      // A tile at smem + 0 (32 KB: two boxes of [128 x 128 B]); B tile at smem + 64 KB (up to 64 KB)
      const uint32_t loA = desc_lo(smem_u32(smem), c.a_mn ? 16384 : 16);
      const uint32_t loB = desc_lo(smem_u32(smem + 65536), c.b_mn ? 16384 : 16);
      const long long t0 = clock64();
      for (int i = 0; i < R; ++i) {
        const int ks = i & 7;
        const uint32_t a_off = c.a_mn ? ks * 128 : ((ks >> 2) * 16384 + (ks & 3) * 32) / 16;
        const uint32_t b_off = c.b_mn ? ks * 128 : ((ks >> 2) * 16384 + (ks & 3) * 32) / 16;
        const uint32_t d = tmem + ((i / 8) % c.n_acc) * c.N;
        if (c.a_tmem)
          umma_ts_f16(d, tmem + 384 + ks * 8, desc_join(loB + b_off, kHi), idesc, ks != 0);
        else
          umma_ss<false, 1>(d, desc_join(loA + a_off, kHi), desc_join(loB + b_off, kHi), idesc, ks != 0);
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
      stop = 1;
    }
    __syncwarp();
  } else if (warp >= 4 && c.load) {
    // concurrent TMEM readers (like the softmax warps): 32 columns per load from the upper half
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    while (!stop) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + 448 + lane_off, v);
      tmem_wait_ld_regs(v);
      acc += v[0];
    }
    if (acc == 0x12345678u) out[1] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem, 512);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  const int smem_bytes = 129 * 1024 + 1024;
  cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  struct Named { const char* name; Cfg c; };
  std::vector<Named> cfgs = {
      {"SS N=128 A,B K-major, 1 acc", {128, 0, 0, 0, 1, 0}},
      {"SS N=128 A,B K-major, 2 acc", {128, 0, 0, 0, 2, 0}},
      {"SS N=64  A,B K-major", {64, 0, 0, 0, 2, 0}},
      {"SS N=256 A,B K-major", {256, 0, 0, 0, 1, 0}},
      {"SS N=128 B MN-major", {128, 0, 0, 1, 2, 0}},
      {"SS N=128 A,B MN-major", {128, 0, 1, 1, 2, 0}},
      {"SS N=64  A,B MN-major", {64, 0, 1, 1, 2, 0}},
      {"TS N=128 B K-major", {128, 1, 0, 0, 2, 0}},
      {"TS N=128 B MN-major", {128, 1, 0, 1, 2, 0}},
      {"SS N=128 K-major + TMEM readers", {128, 0, 0, 0, 2, 1}},
      {"TS N=128 B MN-major + TMEM readers", {128, 1, 0, 1, 2, 1}},
  };
  const int R = 512;
  for (auto& n : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      mma_bench_kernel<<<148, 256, smem_bytes>>>(n.c, R, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", n.name, cudaGetErrorString(e)); return 1; }
    }
    long long cyc;
    cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    const double per = double(cyc) / R;
    printf("%-40s %7.1f cycles/MMA   %6.0f MAC/clk/SM  (floor %d)\n", n.name, per, 128.0 * n.c.N * 16 / per, n.c.N / 2);
  }
  return 0;
}
