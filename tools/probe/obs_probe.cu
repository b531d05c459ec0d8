// tools/probe/obs_probe.cu -- hardware probe (not product code): ways to stream one observation row per warp
// (read W*5 fp64 of a candle window, z-score, write W*5 + 2W fp32) on B200, in graph-replayed kernels.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <cstdint>

#define W 128
#define F 5
#define D (W * F + 2 * W + 4)
#define TROWS (1 << 19)

struct PP { const double* tab; const double* stats; float* out; long long* cyc; int n; int step; };

__device__ __forceinline__ float fin(float v) { v = (v != v) ? 0.f : v; return fminf(fmaxf(v, -10.f), 10.f); }

// V0: as in the product kernel v4: stats via smem, 30-lane loop (unroll 8), then prices loop
__global__ void __launch_bounds__(128, 8) obs_v0(const __grid_constant__ PP P) {
  __shared__ double sm[4][2 * F];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  if (warp >= P.n) return;
  long long t0 = clock64();
  const long long row0 = ((long long)warp * 9973 + P.step) % (TROWS - W - 2);
  const double* __restrict__ base = P.tab + row0 * F;
  float* __restrict__ out = P.out + (long long)warp * D;
  if (lane < F) { const double* sp = P.stats + ((row0 + W - 1) * F + lane) * 2; sm[wl][lane] = sp[0]; sm[wl][F + lane] = sp[1]; }
  __syncwarp();
  if (lane < 30) {
    const int f = lane % 5; const double m = sm[wl][f], r = sm[wl][F + f];
#pragma unroll 8
    for (int j = lane; j < W * F; j += 30) { const double x = __ldg(base + j); __stcs(out + j, fin((float)((x - m) * r))); }
  }
#pragma unroll 4
  for (int w = lane; w < W; w += 32) {
    const double p = __ldg(base + w * F + 3), q = __ldg(base + (w ? w - 1 : 0) * F + 3);
    __stcs(out + W * F + w, (float)p); __stcs(out + W * F + W + w, w ? (float)(p - q) : 0.f);
  }
  if (lane == 0) P.cyc[warp] = clock64() - t0;
}

// V1: every load issued up front (one round trip), 32 lanes, stats by LDG per lane; prices derived from the same registers
__global__ void __launch_bounds__(128, 8) obs_v1(const __grid_constant__ PP P) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P.n) return;
  long long t0 = clock64();
  const long long row0 = ((long long)warp * 9973 + P.step) % (TROWS - W - 2);
  const double* __restrict__ base = P.tab + row0 * F;
  float* __restrict__ out = P.out + (long long)warp * D;
  const double* sp = P.stats + (row0 + W - 1) * F * 2;
  double x[20], m[5], r[5];
#pragma unroll
  for (int k = 0; k < 20; k++) x[k] = __ldg(base + lane + 32 * k);      // 640 = 20 * 32
  // lane's feature for element lane + 32k is (lane + 2k) % 5: period 5 in k
#pragma unroll
  for (int q = 0; q < 5; q++) { const int f = (lane + 2 * q) % 5; m[q] = __ldg(sp + 2 * f); r[q] = __ldg(sp + 2 * f + 1); }
#pragma unroll
  for (int k = 0; k < 20; k++) __stcs(out + lane + 32 * k, fin((float)((x[k] - m[k % 5]) * r[k % 5])));
  double p[4], qv[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { const int w = lane + 32 * k; p[k] = __ldg(base + w * F + 3); qv[k] = __ldg(base + (w ? w - 1 : 0) * F + 3); }
#pragma unroll
  for (int k = 0; k < 4; k++) { const int w = lane + 32 * k; __stcs(out + W * F + w, (float)p[k]); __stcs(out + W * F + W + w, w ? (float)(p[k] - qv[k]) : 0.f); }
  if (lane == 0) P.cyc[warp] = clock64() - t0;
}

// V2: TMA bulk copy (cp.async.bulk global->shared + mbarrier) of the window into per-warp smem, compute from smem
__global__ void __launch_bounds__(128, 8) obs_v2(const __grid_constant__ PP P) {
  __shared__ __align__(128) double sw[4][W * F + 2];
  __shared__ __align__(8) unsigned long long bar[4];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  if (warp >= P.n) return;
  long long t0 = clock64();
  const long long row0 = ((long long)warp * 9973 + P.step) % (TROWS - W - 2);
  const double* gsrc = P.tab + row0 * F;                 // 8-byte aligned; bulk copy needs 16 B: copy from the even element
  const long long e0 = (row0 * F) & ~1LL;                 // element index rounded down to even (16 B aligned)
  const int shift = (int)(row0 * F - e0);                 // 0 or 1
  const unsigned bytes = (W * F + 2) * 8;                 // multiple of 16
  const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar[wl]);
  const unsigned dst_a = (unsigned)__cvta_generic_to_shared(&sw[wl][0]);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_a), "l"(P.tab + e0), "r"(bytes), "r"(bar_a) : "memory");
  }
  float* __restrict__ out = P.out + (long long)warp * D;
  const double* sp = P.stats + (row0 + W - 1) * F * 2;
  const int f = lane % 5;
  double m = 0, r = 1;
  if (lane < 30) { m = __ldg(sp + 2 * f); r = __ldg(sp + 2 * f + 1); }
  __syncwarp();
  {  // wait for the bulk copy (phase 0)
    unsigned ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar_a) : "memory");
  }
  const double* s = &sw[wl][shift];
  if (lane < 30) {
#pragma unroll 8
    for (int j = lane; j < W * F; j += 30) __stcs(out + j, fin((float)((s[j] - m) * r)));
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int w = lane + 32 * k; const double p = s[w * F + 3], q = s[(w ? w - 1 : 0) * F + 3];
    __stcs(out + W * F + w, (float)p); __stcs(out + W * F + W + w, w ? (float)(p - q) : 0.f);
  }
  if (lane == 0) P.cyc[warp] = clock64() - t0;
  (void)gsrc;
}

typedef void (*KF)(const PP);
void run(const char* label, KF k, int n, int ring) {
  PP P{}; P.n = n;
  double *tab, *stats; cudaMalloc(&tab, (size_t)TROWS * F * 8); cudaMalloc(&stats, (size_t)TROWS * F * 16);
  cudaMemset(tab, 0, (size_t)TROWS * F * 8); cudaMemset(stats, 0, (size_t)TROWS * F * 16);
  P.tab = tab; P.stats = stats;
  float* out; size_t slot = (size_t)n * D; cudaMalloc(&out, slot * 4 * ring);
  cudaMalloc(&P.cyc, n * 8);
  cudaStream_t s; cudaStreamCreate(&s);
  const int K = 200; cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < K; i++) { P.out = out + slot * (i % ring); P.step = i; k<<<(n + 3) / 4, 128, 0, s>>>(P); }
  cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
  cudaEventRecord(e0, s); cudaGraphLaunch(ge, s); cudaEventRecord(e1, s); cudaStreamSynchronize(s);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(n); cudaMemcpy(h.data(), P.cyc, n * 8, cudaMemcpyDeviceToHost);
  double mean = 0; long long mx = 0; for (auto v : h) { mean += v; if (v > mx) mx = v; } mean /= n;
  const double us = ms * 1e3 / K, bytes = (double)n * (D * 4);
  printf("%-34s n=%6d: %7.2f us/kernel  (%.2f TB/s of obs stores)  warp cycles mean %7.0f max %7lld\n", label, n, us, bytes / us / 1e6, mean, mx);
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
  cudaFree(tab); cudaFree(stats); cudaFree(out); cudaFree(P.cyc);
}

int main() {
  for (int n : {4096, 16384, 65536}) {
    int ring = n == 4096 ? 17 : (n == 16384 ? 5 : 2);
    run("V0 smem-stats, 30-lane loop u8", obs_v0, n, ring);
    run("V1 all loads up front (1 RT)", obs_v1, n, ring);
    run("V2 TMA bulk -> smem", obs_v2, n, ring);
  }
  return 0;
}
