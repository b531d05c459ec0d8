// tools/probe/lat_probe.cu -- hardware probe (not product code): per-warp latency of dependent global loads at the
// start of back-to-back graph-replayed kernels on B200, as a function of grid shape, parameter-block size and
// background streaming stores.  nvcc -arch=sm_100a -O3 -o lat_probe lat_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

struct Big { double* st; const double* tab; long long* stamps; float* sink; int n; int stream_floats; char pad[1200]; };
struct Small { double* st; const double* tab; long long* stamps; float* sink; int n; int stream_floats; };

template <typename PT>
__global__ void probe(const __grid_constant__ PT P) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P.n) return;
  long long t0 = clock64();
  double a = P.st[warp];                       // RT 1: uniform load of "state"
  long long idx = (long long)a;                // state holds an index into the table
  long long t1 = clock64();
  double b = P.tab[(idx + warp * 9973LL) % (1 << 19) * 5];   // RT 2: dependent load
  long long t2 = clock64();
  double acc = b;
  long long t2b = (long long)b & 1;
#pragma unroll
  for (int k = 0; k < 8; k++) acc += P.tab[((idx + warp * 9973LL + 64 * k + lane + t2b) % (1 << 19)) * 5];  // RT 3: 8 independent
  long long t3 = clock64();
  if (P.stream_floats) {                       // background streaming stores (observation-like)
    float* o = P.sink + (long long)warp * P.stream_floats;
    for (int j = lane; j < P.stream_floats; j += 32) __stcs(o + j, (float)acc);
  }
  long long t4 = clock64();
  if (lane == 0) {
    P.st[warp] = a + 1.0 + (acc == 12345.678 ? 1.0 : 0.0);
    long long* s = P.stamps + (long long)warp * 5;
    s[0] = t0; s[1] = t1; s[2] = t2; s[3] = t3; s[4] = t4;
  }
}

template <typename PT>
void run(const char* label, int n_warps, int block, int stream_floats, int ring_slots) {
  PT P{};
  P.n = n_warps; P.stream_floats = stream_floats;
  cudaMalloc(&P.st, n_warps * 8); cudaMemset(P.st, 0, n_warps * 8);
  double* tab; cudaMalloc(&tab, (size_t)(1 << 19) * 5 * 8); cudaMemset(tab, 0, (size_t)(1 << 19) * 5 * 8); P.tab = tab;
  cudaMalloc(&P.stamps, (size_t)n_warps * 5 * 8);
  float* sink = nullptr; size_t slot = (size_t)n_warps * (stream_floats ? stream_floats : 1) * 4;
  cudaMalloc(&sink, slot * ring_slots);
  cudaStream_t s; cudaStreamCreate(&s);
  const int K = 200;
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  int grid = (n_warps * 32 + block - 1) / block;
  for (int k = 0; k < K; k++) { P.sink = sink + (slot / 4) * (k % ring_slots); probe<PT><<<grid, block, 0, s>>>(P); }
  cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
  cudaEventRecord(e0, s); cudaGraphLaunch(ge, s); cudaEventRecord(e1, s); cudaStreamSynchronize(s);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h((size_t)n_warps * 5);
  cudaMemcpy(h.data(), P.stamps, h.size() * 8, cudaMemcpyDeviceToHost);
  double m[4] = {0, 0, 0, 0}; long long first = h[0], last = 0;
  for (int w = 0; w < n_warps; w++) {
    for (int i = 0; i < 4; i++) m[i] += (double)(h[w * 5 + i + 1] - h[w * 5 + i]);
  }
  for (int i = 0; i < 4; i++) m[i] /= n_warps;
  printf("%-46s grid %5d x %4d: %6.2f us/kernel | RT1 %6.0f  RT2(dep) %6.0f  RT3(8 indep) %6.0f  stores %6.0f cycles\n",
         label, grid, block, ms * 1e3 / K, m[0], m[1], m[2], m[3]);
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
  cudaFree(P.st); cudaFree(tab); cudaFree(P.stamps); cudaFree(sink);
}

int main() {
  run<Small>("small params, no stores", 4096, 128, 0, 1);
  run<Big>("1.3KB params, no stores", 4096, 128, 0, 1);
  run<Small>("small params, 28 warps/CTA", 4096, 896, 0, 1);
  run<Small>("small params, 8 warps/CTA", 4096, 256, 0, 1);
  run<Small>("small, +3.6KB stcs stores/warp, ring 17", 4096, 128, 900, 17);
  run<Small>("small, +3.6KB stcs stores/warp, ring 1", 4096, 128, 900, 1);
  run<Small>("small params, no stores, 16384 warps", 16384, 128, 0, 1);
  run<Small>("small, +3.6KB stores, 16384 warps, ring 5", 16384, 128, 900, 5);
  return 0;
}
