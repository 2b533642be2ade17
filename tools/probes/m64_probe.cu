// Probe (round-2 preparation): where does tcgen05.mma (cta_group::1, kind::tf32) with M = 64 put the accumulator rows in TMEM?
// The chain kernel needs two 64-row tiles in flight per CTA (so that one tile's MMAs overlap the other's epilogue); its epilogue
// reads TMEM with tcgen05.ld.32x32b (warp q may touch lanes 32q .. 32q+31 only), so the row -> (lane, column) map decides the warp roles.
// D[64 x 64] = A[64 x 8] B[64 x 8]^T with A(r, k) = (k == 0 ? r + 1 : 0), B(n, k) = (k == 0 ? 1 + n / 128 : 0)  =>  D(r, n) = (r + 1)(1 + n/128):
// every TMEM word read back identifies its (row, column).  All 4 warps dump their 32 lanes x 64 columns.
// Build:  nvcc -I../../include -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o m64_probe m64_probe.cu
#include <cstdio>
#include <cmath>
#include <vector>
#include "../../deep-whole-body-control_b200/csrc/gemm_tc.cuh"
using namespace dwbc;

__device__ __forceinline__ uint32_t idesc_m(int m, int n) {        // tc_idesc with a free M
  uint32_t d = 0;
  d |= 1u << 4;                    // D = F32
  d |= 2u << 7;                    // A = TF32
  d |= 2u << 10;                   // B = TF32
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

__global__ void probe(float* out /*[128 TMEM lanes][64 cols]*/) {
  extern __shared__ __align__(1024) float sm[];
  float* sA = sm;                 // canonical K-major, kpad = 8: element (r, k) at ((r/8)*2 + k/4)*32 + (r%8)*4 + k%4 floats
  float* sB = sm + 128 * 8;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * 8; i += blockDim.x) {
    const int r = i >> 3, k = i & 7;
    const int off = ((r >> 3) * 2 + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3);
    sA[off] = (k == 0 && r < 64) ? (float)(r + 1) : 0.0f;
    sB[off] = (k == 0 && r < 64) ? 1.0f + (float)r / 128.0f : 0.0f;       // exactly representable in TF32 for r < 64 (7 fractional bits)
  }
  if (warp == 0) tc_tmem_alloc(&tmem_s, 64);
  if (tid == 0) tc_mbar_init(&bar, 1);
  tc_fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  if (tid == 0) {
    const uint64_t ad = tc_desc(tc_smem_u32(sA), 128, 256), bd = tc_desc(tc_smem_u32(sB), 128, 256);
    tc_mma_tf32(tmem, ad, bd, idesc_m(64, 64), 0);
    tc_commit(&bar);
  }
  tc_mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    float v[32];
    tc_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<float> h(128 * 64);
  float* d;
  cudaMalloc(&d, h.size() * 4);
  cudaMemset(d, 0xff, h.size() * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 128 * 8 * 4 + 1024);
  probe<<<1, 128, 2 * 128 * 8 * 4, 0>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
  printf("err=%s\n", cudaGetErrorString(e));
  // decode: value = (r + 1) * (1 + n / 128)  ->  for a TMEM word find (r, n)
  for (int lane = 0; lane < 128; ++lane) {
    int r0 = -1, n0 = -1, r63 = -1, n63 = -1, valid = 0;
    for (int c = 0; c < 64; ++c) {
      const float v = h[lane * 64 + c];
      for (int r = 0; r < 64 && std::isfinite(v); ++r) {
        const float n = (v / (float)(r + 1) - 1.0f) * 128.0f;
        if (n >= -0.001f && n < 63.5f && fabsf(n - roundf(n)) < 1e-3f) {
          ++valid;
          if (c == 0) { r0 = r; n0 = (int)roundf(n); }
          if (c == 63) { r63 = r; n63 = (int)roundf(n); }
          break;
        }
      }
    }
    printf("tmem lane %3d: %2d of 64 columns hold products; col 0 -> D(%d, %d), col 63 -> D(%d, %d)\n", lane, valid, r0, n0, r63, n63);
  }
  return 0;
}
