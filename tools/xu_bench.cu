// Microbenchmark: issue cost (cycles per warp instruction per SM sub-partition) of the instructions the attention
// softmax loops are made of, alone and mixed, with 1 or 2 warps per scheduler (= one or two softmax warpgroups per CTA).
// Answers, for the forward kernel's design: which pipe does F2FP (fp32 pair -> bf16x2) share, what does ex2 cost next
// to it, and what do the packed f32x2 forms save.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 tools/xu_bench.cu -o tools/xu_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { EX2, CVT, EX2_CVT, FFMA, FADD, FMNMX, FFMA2, FADD2, FMUL2, EX2_FFMA, EX2_CVT_FFMA, ALUPACK, EX2_ALUPACK,
          POLY, EX2_POLY3, TILE_NOW, TILE_ALU, TILE_ALU_F2, NUM_OPS };
static const char* kNames[NUM_OPS] = {
    "ex2.approx.ftz.f32", "cvt.rn.bf16x2.f32 (F2FP)", "2 ex2 + 1 cvt (per pair)", "fma.rn.f32", "add.f32", "max.f32",
    "fma.rn.f32x2 (2 elts)", "add.f32x2 (2 elts)", "mul.f32x2 (2 elts)", "1 ex2 + 1 ffma", "2 ex2 + 1 cvt + 2 ffma",
    "ALU pack (2 iadd + prmt, per pair)", "2 ex2 + ALU pack", "poly exp2 deg3 (1 elt, FMA/ALU only)",
    "3 ex2 + 1 poly (4 elts)", "softmax tile body as shipped (per pair: 2 ffma 2 ex2 2 fadd cvt 2 max)",
    "same, ALU pack", "same, ALU pack + f32x2 scale/sum"};

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t cvt2(float lo, float hi) {
  uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}
__device__ __forceinline__ uint32_t alupack(float lo, float hi) {
  return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
               "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
               : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm volatile("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
               "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
               : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm volatile("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
               "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
               : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
// 2^x for x <= 0 (x > -126): Cody-Waite split x = n + f, f in [0,1) by the magic-number trick, degree-3 minimax for 2^f,
// exponent add in the integer domain. FMA + ALU pipes only.
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;                 // 1.5 * 2^23: rounds x to nearest integer in the low mantissa bits
  const float n = t - 12582912.f;
  const float f = x - n;                          // [-0.5, 0.5]
  float p = fmaf(0.0555041086f, f, 0.2402265069f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

template <int OP>
__global__ void __launch_bounds__(256, 1) bench(float* out, long long* cyc, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + i + 1) * 1e-3f - 3.f;
  uint32_t u[8] = {1, 2, 3, 4, 5, 6, 7, 8};
  float mx = -1e30f, sum = 0.f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if constexpr (OP == EX2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = ex2f(a[i]) - 2.f;
    } else if constexpr (OP == CVT) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { u[i] = cvt2(a[i], __uint_as_float(u[i])); }
    } else if constexpr (OP == EX2_CVT) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) { const float p0 = ex2f(a[i]), p1 = ex2f(a[i + 1]); u[i] ^= cvt2(p0, p1); a[i] -= 1e-3f; }
    } else if constexpr (OP == FFMA) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], 0.999f, seed);
    } else if constexpr (OP == FADD) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = a[i] + seed;
    } else if constexpr (OP == FMNMX) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], __uint_as_float(u[i] + it));
    } else if constexpr (OP == FFMA2) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) { const float2 r = ffma2(make_float2(a[i], a[i + 1]), make_float2(0.999f, 0.998f), make_float2(seed, seed)); a[i] = r.x; a[i + 1] = r.y; }
    } else if constexpr (OP == FADD2) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) { const float2 r = fadd2(make_float2(a[i], a[i + 1]), make_float2(seed, seed)); a[i] = r.x; a[i + 1] = r.y; }
    } else if constexpr (OP == FMUL2) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) { const float2 r = fmul2(make_float2(a[i], a[i + 1]), make_float2(0.999f, 0.998f)); a[i] = r.x; a[i + 1] = r.y; }
    } else if constexpr (OP == EX2_FFMA) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = ex2f(fmaf(a[i], 0.5f, -1.f));
    } else if constexpr (OP == EX2_CVT_FFMA) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float p0 = ex2f(fmaf(a[i], 0.5f, -1.f)), p1 = ex2f(fmaf(a[i + 1], 0.5f, -1.f));
        u[i] ^= cvt2(p0, p1); a[i] -= 1e-3f;
      }
    } else if constexpr (OP == ALUPACK) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) { u[i] = alupack(__uint_as_float(u[i]), a[i + 1]); }
    } else if constexpr (OP == EX2_ALUPACK) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) { const float p0 = ex2f(a[i]), p1 = ex2f(a[i + 1]); u[i] ^= alupack(p0, p1); a[i] -= 1e-3f; }
    } else if constexpr (OP == POLY) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = poly_exp2(a[i]) - 2.f;
    } else if constexpr (OP == EX2_POLY3) {
#pragma unroll
      for (int i = 0; i < 8; i += 4) {
        a[i] = ex2f(a[i]) - 2.f; a[i + 1] = ex2f(a[i + 1]) - 2.f; a[i + 2] = ex2f(a[i + 2]) - 2.f;
        a[i + 3] = poly_exp2(a[i + 3]) - 2.f;
      }
    } else if constexpr (OP == TILE_NOW || OP == TILE_ALU) {
      // 4 pairs per iteration, mirroring the shipped loop: max (previous pass), scale/sub, ex2, rowsum, pack
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        mx = fmaxf(mx, fmaxf(a[i], a[i + 1]));
        const float p0 = ex2f(fmaf(a[i], 0.5f, -1.f)), p1 = ex2f(fmaf(a[i + 1], 0.5f, -1.f));
        sum += p0 + p1;
        u[i] ^= (OP == TILE_NOW) ? cvt2(p0, p1) : alupack(p0, p1);
        a[i] -= 1e-3f;
      }
    } else if constexpr (OP == TILE_ALU_F2) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        mx = fmaxf(mx, fmaxf(a[i], a[i + 1]));
        const float2 s = ffma2(make_float2(a[i], a[i + 1]), make_float2(0.5f, 0.5f), make_float2(-1.f, -1.f));
        const float p0 = ex2f(s.x), p1 = ex2f(s.y);
        const float2 acc = fadd2(make_float2(sum, mx), make_float2(p0, p1));   // two partial sums in one packed add
        sum = acc.x; mx = acc.y;
        u[i] ^= alupack(p0, p1);
        a[i] -= 1e-3f;
      }
    }
  }
  const long long t1 = clock64();
  float r = mx + sum;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += a[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
static void run(float* out, long long* cyc, int warps_per_sched, int units_per_iter) {
  const int iters = 4096;
  for (int rep = 0; rep < 2; ++rep) {
    bench<OP><<<148, 128 * warps_per_sched>>>(out, cyc, iters, 1.0f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", kNames[OP], cudaGetErrorString(e)); return; }
  }
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-75s %d warp/sched: %7.2f cycles per unit per warp, %7.2f per scheduler\n", kNames[OP], warps_per_sched,
         double(c) / iters / units_per_iter, double(c) / iters / units_per_iter / warps_per_sched);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
  for (int w = 1; w <= 2; ++w) {
    run<EX2>(out, cyc, w, 8);   run<CVT>(out, cyc, w, 8);   run<EX2_CVT>(out, cyc, w, 4);
    run<FFMA>(out, cyc, w, 8);  run<FADD>(out, cyc, w, 8);  run<FMNMX>(out, cyc, w, 8);
    run<FFMA2>(out, cyc, w, 4); run<FADD2>(out, cyc, w, 4); run<FMUL2>(out, cyc, w, 4);
    run<EX2_FFMA>(out, cyc, w, 8); run<EX2_CVT_FFMA>(out, cyc, w, 4);
    run<ALUPACK>(out, cyc, w, 4);  run<EX2_ALUPACK>(out, cyc, w, 4);
    run<POLY>(out, cyc, w, 8);     run<EX2_POLY3>(out, cyc, w, 2);
    run<TILE_NOW>(out, cyc, w, 4); run<TILE_ALU>(out, cyc, w, 4); run<TILE_ALU_F2>(out, cyc, w, 4);
    printf("\n");
  }
  return 0;
}
