// Micro-benchmark: how do MUFU.EX2 and FMA-pipe instructions share an SM sub-partition on B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/mufu_bench tools/mufu_bench.cu
// Each thread runs ITER iterations of: NM independent ex2.approx + NF independent FFMA (+ NP FFMA2).
// Reported: SM-cycles per iteration per warp-scheduler, for 1/2/4/8 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>

template <int NM, int NF, int NP>
__global__ void k(float* out, int iters, float seed) {
  float m[8], f[8];
  float2 p[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = seed * (i + 1) * -0.01f; f[i] = seed + i; }
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = make_float2(seed + i, seed - i);
  const float c1 = seed * 0.999f, c2 = seed * 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NM; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i % 8]));
#pragma unroll
    for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i % 8]) : "f"(c1), "f"(c2));
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      float2& q = p[i % 4];
      unsigned long long v = *reinterpret_cast<unsigned long long*>(&q);
      unsigned long long a = *reinterpret_cast<const unsigned long long*>(&make_float2(c1, c1));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v) : "l"(a));
      q = *reinterpret_cast<float2*>(&v);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += m[i] + f[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += p[i].x + p[i].y;
  if (s == 12345.678f) out[0] = s;
}

template <int NM, int NF, int NP>
void run(const char* name) {
  float* d;
  cudaMalloc(&d, 4);
  const int iters = 20000;
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("%-28s", name);
  for (int wps : {1, 2, 4, 8}) {            // warps per scheduler
    const int threads = 128 * wps > 1024 ? 1024 : 128 * wps;
    const int blocks_per_sm = (128 * wps) / threads;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NM, NF, NP><<<148 * blocks_per_sm, threads>>>(d, 100, 1.0f);
    cudaEventRecord(e0);
    k<NM, NF, NP><<<148 * blocks_per_sm, threads>>>(d, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // cycles per iteration per warp on its scheduler = time * clk / (iters * warps_per_scheduler)
    const double cyc = ms * 1e-3 * clk_khz * 1e3 / iters / wps;
    printf("  w%d: %6.1f clk/iter/warp", wps, cyc);
  }
  printf("\n");
  cudaFree(d);
}

int main() {
  run<8, 0, 0>("8 MUFU");
  run<0, 32, 0>("32 FFMA");
  run<0, 0, 16>("16 FFMA2");
  run<8, 8, 0>("8 MUFU + 8 FFMA");
  run<8, 32, 0>("8 MUFU + 32 FFMA");
  run<8, 64, 0>("8 MUFU + 64 FFMA");
  run<8, 0, 16>("8 MUFU + 16 FFMA2");
  run<8, 0, 32>("8 MUFU + 32 FFMA2");
  run<4, 32, 0>("4 MUFU + 32 FFMA");
  return 0;
}
