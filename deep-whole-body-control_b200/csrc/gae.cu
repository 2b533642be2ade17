// RolloutStorage.compute_returns (RS:136-150) as one cooperative kernel (K3 + K4 of SURVEY.md):
// two-channel GAE backward scan, thread per (env, channel) serial over T, then the joint
// advantage normalisation (global mean / UNBIASED std over all T*N*2 elements) after a grid
// barrier.  HBM/L2-bound and tiny: 49 B per (t, env) -> 8 MB at T=40, N=4096.
//
// Also PPO.process_env_step's reward path (PPO:130-134).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dwbc {

constexpr int GAE_BLOCK = 64;

__device__ __forceinline__ void block_accumulate(double s, double ss, double* stats) {
  __shared__ double red[2][GAE_BLOCK / 32];
  s = warp_sum(s);
  ss = warp_sum(ss);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[0][w] = s; red[1][w] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < GAE_BLOCK / 32; ++i) { a += red[0][i]; b += red[1][i]; }
    atomicAdd(stats + 1, a);
    atomicAdd(stats + 2, b);
  }
}

template <bool kFused>
__global__ void __launch_bounds__(GAE_BLOCK)
gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ dones,
           const float* __restrict__ last_values, float* __restrict__ returns, float* __restrict__ advantages,
           double* stats, int T, int N, float gamma, float lam) {
  const int C = 2 * N;
  double s = 0.0, ss = 0.0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < C; j += gridDim.x * blockDim.x) {
    const int env = j >> 1;
    float nxt = last_values[j];
    float adv = 0.0f;
#pragma unroll 8
    for (int t = T - 1; t >= 0; --t) {
      const float r = __ldg(rewards + (size_t)t * C + j);
      const float v = __ldg(values + (size_t)t * C + j);
      const float m = 1.0f - (float)__ldg(dones + (size_t)t * N + env);
      const float delta = r + m * gamma * nxt - v;   // RS:143
      adv = delta + m * gamma * lam * adv;           // RS:144
      const float ret = adv + v;                     // RS:145
      returns[(size_t)t * C + j] = ret;
      const float a = ret - v;                       // RS:148 (returns - values, not `adv` itself)
      advantages[(size_t)t * C + j] = a;
      s += (double)a;
      ss += (double)a * (double)a;
      nxt = v;
    }
  }
  block_accumulate(s, ss, stats);
  if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(stats, (double)T * (double)C);
  if constexpr (kFused) {
    __threadfence();
    cg::this_grid().sync();
    const double n = (double)T * (double)C;
    const double sum = __ldcg(stats + 1), sq = __ldcg(stats + 2);
    const double mean = sum / n;
    const double var = fmax((sq - sum * sum / n) / (n - 1.0), 0.0);
    const float fm = (float)mean, fd = (float)sqrt(var) + 1e-8f;  // RS:150
    const size_t total = (size_t)T * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
      advantages[i] = (__ldcg(advantages + i) - fm) / fd;
  }
}

__global__ void normalize_kernel(float* __restrict__ adv, const double* __restrict__ stats, size_t total) {
  const double n = stats[0], sum = stats[1], sq = stats[2];
  const double mean = sum / n;
  const double var = fmax((sq - sum * sum / n) / (n - 1.0), 0.0);
  const float fm = (float)mean, fd = (float)sqrt(var) + 1e-8f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    adv[i] = (adv[i] - fm) / fd;
}

__global__ void store_rewards_kernel(const float* __restrict__ rew, const float* __restrict__ arm_rew,
                                     const float* __restrict__ values, const uint8_t* __restrict__ time_outs,
                                     const uint8_t* __restrict__ resets, float gamma, float* __restrict__ out,
                                     uint8_t* __restrict__ dones, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float to = time_outs ? (float)time_outs[i] : 0.0f;
  out[2 * i] = rew[i] + gamma * (values[2 * i] * to);          // PPO:133-134
  out[2 * i + 1] = arm_rew[i] + gamma * (values[2 * i + 1] * to);
  if (dones) dones[i] = resets[i] ? 1 : 0;
}

}  // namespace dwbc

using namespace dwbc;

extern "C" int dwbc_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                        float* returns, float* advantages, double* stats, int32_t T, int32_t N, float gamma, float lam,
                        int32_t normalize, dwbc_stream_t stream) {
  if (!rewards || !values || !dones || !last_values || !returns || !advantages || !stats || T <= 0 || N <= 0) return DWBC_ERR_ARG;
  const int C = 2 * N;
  int grid = (C + GAE_BLOCK - 1) / GAE_BLOCK;
  cudaStream_t st = (cudaStream_t)stream;
  if (normalize && (size_t)T * C > 1) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gae_kernel<true>, GAE_BLOCK, 0);
    const int max_coop = sms * per_sm;
    if (max_coop > 0) {
      if (grid > max_coop) grid = max_coop;  // grid-stride loops cover the rest
      void* args[] = {(void*)&rewards, (void*)&values, (void*)&dones, (void*)&last_values, (void*)&returns,
                      (void*)&advantages, (void*)&stats, (void*)&T, (void*)&N, (void*)&gamma, (void*)&lam};
      cudaError_t e = cudaLaunchCooperativeKernel((const void*)gae_kernel<true>, dim3(grid), dim3(GAE_BLOCK), args, 0, st);
      if (e == cudaSuccess) return DWBC_OK;
      (void)cudaGetLastError();  // fall through to the two-kernel path
    }
    gae_kernel<false><<<grid, GAE_BLOCK, 0, st>>>(rewards, values, dones, last_values, returns, advantages, stats, T, N, gamma, lam);
    DWBC_LAUNCH_CHECK();
    return dwbc_normalize_advantages(advantages, stats, (int64_t)T * C, stream);
  }
  gae_kernel<false><<<grid, GAE_BLOCK, 0, st>>>(rewards, values, dones, last_values, returns, advantages, stats, T, N, gamma, lam);
  DWBC_LAUNCH_CHECK();
  return DWBC_OK;
}

extern "C" int dwbc_normalize_advantages(float* advantages, const double* stats, int64_t count, dwbc_stream_t stream) {
  if (!advantages || !stats || count <= 0) return DWBC_ERR_ARG;
  int grid = (int)((count + 1023) / 1024);
  if (grid > 1184) grid = 1184;
  normalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(advantages, stats, (size_t)count);
  DWBC_LAUNCH_CHECK();
  return DWBC_OK;
}

extern "C" int dwbc_store_rewards(const float* rew, const float* arm_rew, const float* values, const uint8_t* time_outs,
                                  const uint8_t* resets, float gamma, float* rewards_out, uint8_t* dones_out,
                                  int32_t num_envs, dwbc_stream_t stream) {
  if (!rew || !arm_rew || !values || !rewards_out || num_envs <= 0 || (dones_out && !resets)) return DWBC_ERR_ARG;
  store_rewards_kernel<<<(num_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rew, arm_rew, values, time_outs, resets, gamma,
                                                                                  rewards_out, dones_out, num_envs);
  DWBC_LAUNCH_CHECK();
  return DWBC_OK;
}
