// fx_policy.cu -- the policy of the closed loop `decide_action -> env.step` (reference caller loop: app/main.py:57-65,
// with a learned actor-critic in place of strategy.decide_action) as ONE fused sm_100a kernel per step:
//
//     h1 = tanh(obs . W1^T + b1)      obs bf16 [N, KP]   (KP = obs_dim padded to a multiple of 64; written by the env
//     h2 = tanh(h1  . W2^T + b2)                           step kernel next to the float32 row)
//     logits = h2 . Wa^T + ba  (3),  value = h2 . Wv + bv
//     action = argmax(logits + gumbel)          Gumbel-max sample (noise supplied, or counter-based in-kernel)
//     logp   = logits[action] - logsumexp(logits)
//
// Layer 1 ([N x 900] . [900 x 256]) and layer 2 are real contractions: 5th-generation tensor cores.  A 128-env row tile
// is owned by a CLUSTER OF TWO CTAs, each computing 128 of the 256 hidden units of both layers: the kernel is bound by
// the bytes one SM can pull from L2 (one CTA alone streamed 875 KB per tile: the whole obs tile and all of W1 and W2), and
// only 32 of 148 SMs had a tile; the pair halves the weight bytes, the MMA and the epilogue work per SM.  Per CTA:
//   warp 0     TMA producer: 128 x 64 obs tiles + 128 x 64 W1 tiles (its half of the hidden units), then for layer 2
//              the 128 x 64 h1 tiles + 128 x 64 W2 tiles, through one 4-stage shared-memory ring (cp.async.bulk.tensor,
//              128-byte swizzle, mbarrier complete_tx);
//   warp 1     MMA issuer: one elected thread issues tcgen05.mma (cta_group::1, kind::f16, M = 128, N = 128, K = 16,
//              bf16 x bf16 -> fp32), accumulators in tensor memory (2 x 128 columns); tcgen05.commit hands ring slots
//              back to the producer and accumulators to the epilogue;
//   warps 2-5  epilogue: tcgen05.ld (32 lanes x 32 columns per instruction; thread = env row), bias + tanh.
// The two halves of h1 meet in global memory (bf16 [N][256], L2-resident: 32 KB written per CTA, then read back by both
// CTAs of the pair as the K-major A operand of layer 2 through TMA) across a cluster barrier; after layer 2 each CTA
// reduces its 128 columns of h2 against the four head rows, rank 1 hands its partial sums to rank 0 (global scratch,
// second cluster barrier), and rank 0 samples and stores.  No cuBLAS / torch on this path.  Launched with the
// programmatic-dependent-launch attribute: the W1 tiles of the first ring slots are requested before griddepcontrol.wait.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fx_policy.cuh"

namespace {

constexpr int kTileM = FX_POLICY_TILE_M;   // env rows per cluster (both CTAs work on the same rows)
constexpr int kHidden = FX_POLICY_HIDDEN;  // 256
constexpr int kHalfN = kHidden / 2;        // hidden units per CTA (= UMMA N)
constexpr int kBlockK = 64;                // bf16 elements per 128-byte swizzled row
constexpr int kStages = 4;
constexpr int kUmmaK = 16;
constexpr uint32_t kABytes = kTileM * kBlockK * 2;    // 16 KB
constexpr uint32_t kBBytes = kHalfN * kBlockK * 2;    // 16 KB
constexpr uint32_t kStageBytes = kABytes + kBBytes;
constexpr int kThreads = 192;                         // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr uint32_t kTmemCols = 256;                   // D1 in columns [0, 128), D2 in [128, 256)

struct __align__(8) Barriers {
  unsigned long long full[kStages], empty[kStages], d1_full, d2_full;
  uint32_t tmem_base;
};

// shared memory map (1024-byte aligned base): [stages: A | B] x kStages | head weights | biases | barriers
constexpr uint32_t kOffHeadW = kStages * kStageBytes;             // float [4][256]: Wa[0..2], Wv
constexpr uint32_t kOffBias = kOffHeadW + 4 * kHidden * 4;        // float b1[256], b2[256], head bias[4]
constexpr uint32_t kOffBar = kOffBias + (2 * kHidden + 4) * 4;
constexpr uint32_t kSmemBytes = kOffBar + sizeof(Barriers) + 1024;  // + slack for the 1024-byte alignment

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, unsigned long long* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), leading byte offset (unused for swizzled K-major, 1) in [16,30), stride byte offset = 1024 B (8 rows x 128 B)
// >> 4 in [32,46), descriptor version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits [4,6) = 1), A = B = BF16 ([7,10) = 1,
// [10,13) = 1), both K-major (bits 15, 16 = 0), N >> 3 in [17,23), M >> 4 in [24,29).
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kHalfN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {  // implies tcgen05.fence::before_thread_sync
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// counter-based uniform in (0, 1): a 64-bit mix of (seed, step, env, action index) -- used when no noise tensor is given
__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned step, unsigned env, unsigned a) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)step * 0x100000001B3ull + ((unsigned long long)env << 2) + a + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return ((float)(unsigned)(z >> 40) + 0.5f) * (1.0f / 16777216.0f);
}

// cluster barrier with release / acquire semantics: every thread of both CTAs executes it (the same number of times)
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}

// (min 2 CTAs per SM only to cap the registers at 168: a policy CTA then fits next to the one-warp CTAs of an env step
// that is still draining, so its prologue -- barrier init, TMEM allocation, parameter loads, the first weight tiles --
// overlaps the step's tail instead of waiting for whole SMs to empty)
__global__ void __launch_bounds__(kThreads, 2)
fx_policy_kernel(const __grid_constant__ CUtensorMap map_obs, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_h1, const FxPolicyDev pol,
                 const int num_envs, const int k_blocks1, const float* __restrict__ gumbel, const unsigned long long seed,
                 const unsigned step, int32_t* __restrict__ action, float* __restrict__ logp, float* __restrict__ value,
                 const int env_begin, const int tile_sync) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Barriers* bar = reinterpret_cast<Barriers*>(smem + kOffBar);
  float* head_w = reinterpret_cast<float*>(smem + kOffHeadW);
  float* bias = reinterpret_cast<float*>(smem + kOffBias);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = blockIdx.x & 1;                         // which half of the hidden units (cluster = CTA pair)
  const int m0 = env_begin + (blockIdx.x >> 1) * kTileM;   // a launch covers the envs [env_begin, num_envs) (env groups)
  const int n0 = rank * kHalfN;
#ifdef FXENV_ENABLE_TIMING  // kernel-chain probe (tools/chain_probe.py): CTA 0 logs {kind, entry, after the wait, exit}
  long long* klog = nullptr;
  if (pol.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
    long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const unsigned long long seqno = atomicAdd(reinterpret_cast<unsigned long long*>(pol.dbg), 1ull);
    klog = pol.dbg + 8 + (seqno % 1024ull) * 4;
    klog[0] = 0; klog[1] = g0;
  }
#endif

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_obs) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_h1) : "memory");
    for (int s = 0; s < kStages; s++) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
    mbar_init(&bar->d1_full, 1); mbar_init(&bar->d2_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // one warp allocates (and later frees) the tensor memory: 256 columns = both accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bar->tmem_base)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // small fp32 parameters (weights of this launch, not produced by the previous kernel): plain loads
  for (int i = threadIdx.x; i < 4 * kHidden; i += kThreads) head_w[i] = pol.head_w[i];
  for (int i = threadIdx.x; i < 2 * kHidden + 4; i += kThreads)
    bias[i] = (i < kHidden) ? pol.b1[i] : (i < 2 * kHidden ? pol.b2[i - kHidden] : pol.head_b[i - 2 * kHidden]);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bar->tmem_base;
  asm volatile("griddepcontrol.launch_dependents;");  // the env step that consumes our actions may get scheduled early
  const int kb2 = kHidden / kBlockK;         // 4 k-blocks of layer 2
  const int total = k_blocks1 + kb2;

  // ================= phase 1: layer 1 (this CTA's 128 hidden units), h1 half -> global =================
  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      const int pre = total < kStages ? total : kStages;
      // weight tiles of the first slots do not depend on the previous kernel: request them before the dependency wait
      for (int it = 0; it < pre; it++) {
        unsigned char* st = smem + it * kStageBytes;
        const bool l1 = it < k_blocks1;
        mbar_expect_tx(&bar->full[it], kStageBytes);
        tma_load_2d(l1 ? &map_w1 : &map_w2, &bar->full[it], st + kABytes, (l1 ? it : it - k_blocks1) * kBlockK, n0);
      }
      if (tile_sync && step > 0u) {
        // closed loop: this tile's rows are complete once its envs have finished step - 1 (counted by the step kernel,
        // which is resident with us: see FxTileSync) -- no need to wait for the whole step grid to drain
        const int valid = (num_envs - m0 < kTileM) ? num_envs - m0 : kTileM;
        const int need = valid * (int)step;
        const int32_t* cnt = pol.done_cnt + m0 / kTileM;
        int polls = 0, have;
        do {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(have) : "l"(cnt) : "memory");
          if (have >= need) break;
          __nanosleep(64);
        } while (++polls <= (1 << 22) || (atomicAdd(pol.timeouts, 1), false));
        asm volatile("fence.proxy.async;" ::: "memory");  // the rows were written through the generic proxy
      } else {
        asm volatile("griddepcontrol.wait;" ::: "memory");  // the observation rows come from the kernel before us
      }
#ifdef FXENV_ENABLE_TIMING
      if (klog) { long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1)); klog[2] = g1; }
#endif
      for (int it = 0; it < total; it++) {
        const int s = it % kStages;
        unsigned char* st = smem + s * kStageBytes;
        const bool l1 = it < k_blocks1;
        if (it >= kStages) {
          mbar_wait(&bar->empty[s], ((it / kStages) - 1) & 1);
          mbar_expect_tx(&bar->full[s], kStageBytes);
          tma_load_2d(l1 ? &map_w1 : &map_w2, &bar->full[s], st + kABytes, (l1 ? it : it - k_blocks1) * kBlockK, n0);
        }
        if (l1) tma_load_2d(&map_obs, &bar->full[s], st, it * kBlockK, m0);
        // (layer 2: the W2 tile is in flight; the h1 tile of the same stage follows in phase 2)
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      for (int it = 0; it < k_blocks1; it++) {
        const int s = it % kStages;
        mbar_wait(&bar->full[s], (it / kStages) & 1);
        tc_fence_after();
        const uint32_t a = smem_u32(smem + s * kStageBytes), b = a + kABytes;
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; k++)
          umma_f16(tmem, umma_desc(a + k * kUmmaK * 2), umma_desc(b + k * kUmmaK * 2), (it | k) ? 1u : 0u);
        umma_commit(&bar->empty[s]);
      }
      umma_commit(&bar->d1_full);  // arrives when every layer-1 MMA has retired
    }
  } else {
    // ===== epilogue 1: warp w may touch TMEM lanes [32 * (w % 4), +32); thread = env row =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const float* b1 = bias + n0;
    float v[32];
    mbar_wait(&bar->d1_full, 0);
    tc_fence_after();
    // h1[:, n0 .. n0 + 128) = tanh(D1 + b1) -> bf16, row-major [row][256] in global memory (the other half comes from the
    // peer CTA); rows beyond the env count are written too (the scratch is padded to whole tiles) and never used
    uint4* dst = reinterpret_cast<uint4*>(pol.h1 + (size_t)(m0 + row) * kHidden + n0);
#pragma unroll 1
    for (int c = 0; c < kHalfN / 32; c++) {
      tmem_ld32(lane_addr + c * 32, v);
      uint32_t packed[16];
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const float x0 = fast_tanh(v[2 * i] + b1[c * 32 + 2 * i]), x1 = fast_tanh(v[2 * i + 1] + b1[c * 32 + 2 * i + 1]);
        __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
        packed[i] = *reinterpret_cast<uint32_t*>(&h);
      }
#pragma unroll
      for (int j = 0; j < 4; j++) dst[c * 4 + j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
    }
    asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy global stores -> the peer's (and our) TMA reads
  }
  cluster_sync();  // both halves of h1 are in global memory (release / acquire at cluster scope)

  // ================= phase 2: layer 2 (this CTA's 128 units of h2), head partial sums =================
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (warp == 0) {
    if (lane == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");
      for (int j = 0; j < kb2; j++) {  // the A operand of the stages whose W2 tile was requested in phase 1
        const int it = k_blocks1 + j, s = it % kStages;
        tma_load_2d(&map_h1, &bar->full[s], smem + s * kStageBytes, j * kBlockK, m0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int j = 0; j < kb2; j++) {
        const int it = k_blocks1 + j, s = it % kStages;
        mbar_wait(&bar->full[s], (it / kStages) & 1);
        tc_fence_after();
        const uint32_t a = smem_u32(smem + s * kStageBytes), b = a + kABytes;
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; k++)
          umma_f16(tmem + kHalfN, umma_desc(a + k * kUmmaK * 2), umma_desc(b + k * kUmmaK * 2), (j | k) ? 1u : 0u);
        umma_commit(&bar->empty[s]);
      }
      umma_commit(&bar->d2_full);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const float* b2 = bias + kHidden + n0;
    float v[32];
    mbar_wait(&bar->d2_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < kHalfN / 32; c++) {
      tmem_ld32(lane_addr + kHalfN + c * 32, v);
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const float h = fast_tanh(v[i] + b2[c * 32 + i]);
        const int k = n0 + c * 32 + i;
        acc[0] = fmaf(h, head_w[k], acc[0]);
        acc[1] = fmaf(h, head_w[kHidden + k], acc[1]);
        acc[2] = fmaf(h, head_w[2 * kHidden + k], acc[2]);
        acc[3] = fmaf(h, head_w[3 * kHidden + k], acc[3]);
      }
    }
    if (rank == 1) pol.head_part[m0 + row] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
  cluster_sync();  // rank 1's partial head sums are visible to rank 0
  if (warp >= 2 && rank == 0) {
    const int row = (warp & 3) * 32 + lane;
    const int env = m0 + row;
    if (env < num_envs) {
      const float4 o = pol.head_part[env];
      const float* hb = bias + 2 * kHidden;
      // (rank 0's columns first, then rank 1's: a fixed order, so the result does not depend on timing)
      const float l0 = (acc[0] + o.x) + hb[0], l1 = (acc[1] + o.y) + hb[1], l2 = (acc[2] + o.z) + hb[2];
      float g0, g1, g2;
      if (gumbel) { g0 = gumbel[(size_t)env * 3]; g1 = gumbel[(size_t)env * 3 + 1]; g2 = gumbel[(size_t)env * 3 + 2]; }
      else {
        g0 = -__logf(-__logf(hash_uniform(seed, step, env, 0)));
        g1 = -__logf(-__logf(hash_uniform(seed, step, env, 1)));
        g2 = -__logf(-__logf(hash_uniform(seed, step, env, 2)));
      }
      const float s0 = l0 + g0, s1 = l1 + g1, s2 = l2 + g2;
      int a = 0; float best = s0;  // first maximum wins, like torch.argmax
      if (s1 > best) { best = s1; a = 1; }
      if (s2 > best) { best = s2; a = 2; }
      const float mx = fmaxf(l0, fmaxf(l1, l2));
      const float lse = mx + logf(expf(l0 - mx) + expf(l1 - mx) + expf(l2 - mx));
      action[env] = a;
      logp[env] = (a == 0 ? l0 : (a == 1 ? l1 : l2)) - lse;
      value[env] = (acc[3] + o.w) + hb[3];
    }
    if (tile_sync) {  // the tile's actions are stored: the step kernel's warps of these envs may go
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
      if (threadIdx.x == 64) {
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(pol.act_flag + m0 / kTileM), "r"((int)step + 1) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef FXENV_ENABLE_TIMING
  if (klog) { long long g2; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g2)); klog[3] = g2; }
#endif
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

// fp32 nn.Linear weights [rows][cols] -> bf16 [rows][cols_pad] (zero padded), round-to-nearest-even
__global__ void fx_policy_pack_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int rows, int cols, int cols_pad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols_pad) return;
  const int r = (int)(i / cols_pad), c = (int)(i - (int64_t)r * cols_pad);
  dst[i] = c < cols ? __bfloat16_as_ushort(__float2bfloat16_rn(src[(int64_t)r * cols + c])) : (uint16_t)0;
}

}  // namespace

cudaError_t fx_policy_pack(const float* src, uint16_t* dst, int rows, int cols, int cols_pad, cudaStream_t stream) {
  const int64_t n = (int64_t)rows * cols_pad;
  fx_policy_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, dst, rows, cols, cols_pad);
  return cudaGetLastError();
}

size_t fx_policy_smem_bytes() { return kSmemBytes; }

cudaError_t fx_policy_configure() {
  return cudaFuncSetAttribute(fx_policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
}

cudaError_t fx_launch_policy(const CUtensorMap& map_obs, const CUtensorMap& map_w1, const CUtensorMap& map_w2,
                             const CUtensorMap& map_h1, const FxPolicyDev& pol, int num_envs, int k_pad, const float* gumbel,
                             unsigned long long seed, unsigned step, int32_t* action, float* logp, float* value,
                             cudaStream_t stream, int env_begin, int env_end, bool tile_sync) {
  if (env_end < 0) env_end = num_envs;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(2 * ((env_end - env_begin + kTileM - 1) / kTileM));  // one CTA pair per 128-env tile
  lc.blockDim = dim3(kThreads);
  lc.dynamicSmemBytes = kSmemBytes;
  lc.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = 2; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = 2;
  return cudaLaunchKernelEx(&lc, fx_policy_kernel, map_obs, map_w1, map_w2, map_h1, pol, env_end, k_pad / kBlockK, gumbel, seed,
                            step, action, logp, value, env_begin, (tile_sync && pol.act_flag) ? 1 : 0);
}
