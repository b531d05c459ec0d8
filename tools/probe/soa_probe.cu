// tools/probe/soa_probe.cu -- hardware probe (not product code): latency of the per-warp state fetch at kernel start,
// 11 uniform loads from 11 SoA columns vs ONE coalesced load of a 128-byte AoS record (+ shuffles), 4096/16384 warps.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
struct P { double* base; long long* cyc; int n; int fields; };

__global__ void __launch_bounds__(128, 8) soa(const __grid_constant__ P p) {   // columns of n doubles, 256-B aligned
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= p.n) return;
  long long t0 = clock64();
  double acc = 0;
#pragma unroll
  for (int f = 0; f < 11; f++) acc += p.base[(size_t)f * p.n + warp];
  long long t1;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) : "l"(__double_as_longlong(acc)));
  if (lane == 0) { p.cyc[warp] = t1 - t0; p.base[(size_t)(warp % 11) * p.n + warp] = acc * 0.5 + 1.0; }
}

__global__ void __launch_bounds__(128, 8) aos(const __grid_constant__ P p) {   // one 128-byte record per env
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= p.n) return;
  long long t0 = clock64();
  double v = (lane < 16) ? p.base[(size_t)warp * 16 + lane] : 0.0;   // ONE coalesced 128-B request per warp
  double acc = 0;
#pragma unroll
  for (int f = 0; f < 11; f++) acc += __shfl_sync(0xffffffffu, v, f);
  long long t1;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) : "l"(__double_as_longlong(acc)));
  if (lane == 0) p.cyc[warp] = t1 - t0;
  if (lane < 16) p.base[(size_t)warp * 16 + lane] = acc * 0.5 + lane;
}

typedef void (*KF)(const P);
void run(const char* nm, KF k, int n) {
  P p{}; p.n = n;
  cudaMalloc(&p.base, (size_t)n * 16 * 8); cudaMemset(p.base, 0, (size_t)n * 16 * 8); cudaMalloc(&p.cyc, n * 8);
  cudaStream_t s; cudaStreamCreate(&s); const int K = 200; cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < K; i++) k<<<n / 4, 128, 0, s>>>(p);
  cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
  cudaEventRecord(e0, s); cudaGraphLaunch(ge, s); cudaEventRecord(e1, s); cudaStreamSynchronize(s);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(n); cudaMemcpy(h.data(), p.cyc, n * 8, cudaMemcpyDeviceToHost);
  double mean = 0; long long mx = 0; for (auto x : h) { mean += x; if (x > mx) mx = x; } mean /= n;
  printf("%-28s n=%6d: %6.2f us/kernel, state fetch %6.0f cycles mean, %6lld max\n", nm, n, ms * 1e3 / K, mean, mx);
}
int main() {
  for (int n : {4096, 16384}) { run("11 SoA columns", soa, n); run("1 AoS 128-B record", aos, n); }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
