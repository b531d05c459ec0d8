// tools/probe/param_probe.cu -- hardware probe (not product code): cost of reading a large by-value kernel parameter
// block (fields spread over ~1.3 KB) at the start of graph-replayed kernels, vs staging a global copy into smem.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
struct Big { int f[320]; double* st; long long* cyc; int n; };   // 1280 B of ints + tail

__global__ void k_const(const __grid_constant__ Big P) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P.n) return;
  long long t0 = clock64();
  int acc = 0;
#pragma unroll
  for (int i = 0; i < 20; i++) acc += P.f[i * 16];          // one field per 64 B of the parameter block
  double a = P.st[(warp + (acc & 1)) % P.n];
  long long t1 = clock64();
  if (lane == 0) { P.cyc[warp] = t1 - t0 + (a == 1234.5 ? 1 : 0); }
}

__global__ void k_smem(const Big* __restrict__ Pg) {
  __shared__ __align__(16) unsigned char sp[sizeof(Big)];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  for (int i = threadIdx.x; i < (int)(sizeof(Big) / 8); i += blockDim.x)
    reinterpret_cast<double*>(sp)[i] = reinterpret_cast<const double*>(Pg)[i];
  __syncthreads();
  const Big& P = *reinterpret_cast<const Big*>(sp);
  if (warp >= P.n) return;
  int acc = 0;
#pragma unroll
  for (int i = 0; i < 20; i++) acc += P.f[i * 16];
  double a = P.st[(warp + (acc & 1)) % P.n];
  long long t1 = clock64();
  if (lane == 0) { P.cyc[warp] = t1 - t0 + (a == 1234.5 ? 1 : 0); }
}

__global__ void k_first(const __grid_constant__ Big P) {      // only the first 64 B of the block
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P.n) return;
  long long t0 = clock64();
  int acc = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) acc += P.f[i];
  double a = P.st[(warp + (acc & 1)) % P.n];
  long long t1 = clock64();
  if (lane == 0) { P.cyc[warp] = t1 - t0 + (a == 1234.5 ? 1 : 0); }
}

int main() {
  const int n = 4096, K = 200;
  Big P{}; P.n = n;
  cudaMalloc(&P.st, n * 8); cudaMemset(P.st, 0, n * 8); cudaMalloc(&P.cyc, n * 8);
  Big* Pg; cudaMalloc(&Pg, sizeof(Big)); cudaMemcpy(Pg, &P, sizeof(Big), cudaMemcpyHostToDevice);
  cudaStream_t s; cudaStreamCreate(&s);
  for (int v = 0; v < 3; v++) {
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < K; i++) {
      if (v == 0) k_first<<<n / 4, 128, 0, s>>>(P);
      else if (v == 1) k_const<<<n / 4, 128, 0, s>>>(P);
      else k_smem<<<n / 4, 128, 0, s>>>(Pg);
    }
    cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ge, g, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
    cudaEventRecord(e0, s); cudaGraphLaunch(ge, s); cudaEventRecord(e1, s); cudaStreamSynchronize(s);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(n); cudaMemcpy(h.data(), P.cyc, n * 8, cudaMemcpyDeviceToHost);
    double mean = 0; for (auto x : h) mean += x; mean /= n;
    const char* nm[3] = {"by-value, first 64 B only", "by-value, 20 fields over 1.3 KB", "global copy staged into smem"};
    printf("%-36s %6.2f us/kernel, params+1 load: %7.0f cycles per warp\n", nm[v], ms * 1e3 / K, mean);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
