// fx_policy.cuh -- device-side parameters of the fused actor-critic policy kernel (fx_policy.cu) and its launcher.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define FX_POLICY_TILE_M 128   // env rows per CTA (= UMMA M)
#define FX_POLICY_HIDDEN 256   // hidden units of both layers (= UMMA N)
#define FX_POLICY_ACTIONS 3

// fp32 parameters the epilogue reads directly (the two weight matrices travel as bf16 through TMA tensor maps)
struct FxPolicyDev {
  const float* b1;      // [256]
  const float* b2;      // [256]
  const float* head_w;  // [4][256]: rows 0..2 = actor head (one per action), row 3 = critic head
  const float* head_b;  // [4]
};

size_t fx_policy_smem_bytes();
cudaError_t fx_policy_configure();
// One policy evaluation for all envs: obs (bf16 [num_envs][k_pad], through map_obs) -> action / log-prob / value.
// gumbel: float32 [num_envs][3] Gumbel(0,1) noise, or nullptr for the in-kernel counter-based generator (seed, step).
cudaError_t fx_launch_policy(const CUtensorMap& map_obs, const CUtensorMap& map_w1, const CUtensorMap& map_w2,
                             const FxPolicyDev& pol, int num_envs, int k_pad, const float* gumbel, unsigned long long seed,
                             unsigned step, int32_t* action, float* logp, float* value, cudaStream_t stream,
                             int env_begin = 0, int env_end = -1);  // env_begin must be a multiple of FX_POLICY_TILE_M

// fp32 [rows][cols] (nn.Linear layout) -> bf16 [rows][cols_pad], zero padded
cudaError_t fx_policy_pack(const float* src, uint16_t* dst, int rows, int cols, int cols_pad, cudaStream_t stream);
