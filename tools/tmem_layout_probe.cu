// Probe of the tcgen05.ld / tcgen05.st fragment layouts used by the attention forward v6 (no PTX manual offline):
// rows are written thread-per-row (32x32b), read back as 16x256b; packed words written as 16x128b, read back 32x32b.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tmem_layout_probe tools/tmem_layout_probe.cu && ./tools/tmem_layout_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(uint32_t* out_ld, uint32_t* out_st) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  // 1. thread-per-row store: value = row * 256 + col, 64 columns
  for (int c0 = 0; c0 < 64; c0 += 4) {
    const uint32_t row = warp * 32 + lane;
    uint32_t v0 = row * 256 + c0, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((uint32_t)(warp * 32) << 16) + c0),
                 "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // 2. read back as 16x256b.x2 (16 columns): lanes [32w, 32w+16) and [32w+16, 32w+32)
  for (int half = 0; half < 2; ++half) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(base + ((uint32_t)(warp * 32 + half * 16) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out_ld[((warp * 2 + half) * 32 + lane) * 8 + i] = r[i];
  }
  __syncthreads();
  // 3. store packed words as 16x128b.x2 (8 columns) at columns 64.., value = tag(warp, half, lane, i); read back 32x32b
  for (int half = 0; half < 2; ++half) {
    uint32_t v[4];
    for (int i = 0; i < 4; ++i) v[i] = 0x10000000u | (half << 24) | (lane << 8) | i;
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((uint32_t)(warp * 32 + half * 16) << 16) + 64),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(base + ((uint32_t)(warp * 32) << 16) + 64) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out_st[(warp * 32 + lane) * 8 + i] = r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(base) : "memory");
}

int main() {
  uint32_t *d_ld, *d_st;
  cudaMalloc(&d_ld, 4 * 2 * 32 * 8 * 4);
  cudaMalloc(&d_st, 128 * 8 * 4);
  probe<<<1, 128>>>(d_ld, d_st);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  static uint32_t h_ld[4 * 2 * 32 * 8], h_st[128 * 8];
  cudaMemcpy(h_ld, d_ld, sizeof(h_ld), cudaMemcpyDeviceToHost);
  cudaMemcpy(h_st, d_st, sizeof(h_st), cudaMemcpyDeviceToHost);
  printf("== 16x256b.x2 load, warp 0: per lane, 8 regs as (row,col)\n");
  for (int half = 0; half < 2; ++half)
    for (int lane = 0; lane < 32; ++lane) {
      printf("half %d lane %2d:", half, lane);
      for (int i = 0; i < 8; ++i) { uint32_t v = h_ld[((0 * 2 + half) * 32 + lane) * 8 + i]; printf(" (%3u,%2u)", v >> 8, v & 255); }
      printf("\n");
    }
  printf("== warp 1 half 1 lane 5:");
  for (int i = 0; i < 8; ++i) { uint32_t v = h_ld[((1 * 2 + 1) * 32 + 5) * 8 + i]; printf(" (%3u,%2u)", v >> 8, v & 255); }
  printf("\n== 16x128b.x2 store read back thread-per-row (rows 0..31 of warp 0): 8 columns as (half,lane,i)\n");
  for (int row = 0; row < 32; ++row) {
    printf("row %2d:", row);
    for (int i = 0; i < 8; ++i) { uint32_t v = h_st[row * 8 + i]; printf(" (%u,%2u,%u)", (v >> 24) & 1, (v >> 8) & 255, v & 255); }
    printf("\n");
  }
  return 0;
}
