// Probe: does tcgen05.mma.kind::tf32 accept MN-major shared-memory operands (no swizzle)?
// D[128 x 128] = sum_k P[k][m] * Q[k][n] with P, Q stored as [rows(k) x feat] tiles in the K-major canonical layout of a
// [128 x 128] activation tile, re-read as MN-major operands.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../deep-whole-body-control_b200/csrc/gemm_tc.cuh"
using namespace dwbc;

__device__ __forceinline__ uint32_t idesc_v(int n, int amn, int bmn) { return tc_idesc(n, amn, bmn); }

// variant: 0 = A MN / B MN (LBO=kgroup stride, SBO=mn stride); 1 = swapped; 2 = A K-major, B MN-major (dgrad form); 3 = 2 swapped
__global__ void probe(const float* P, const float* Q, float* D, int variant) {
  extern __shared__ __align__(1024) float sm[];
  float* sA = sm;
  float* sB = sm + 128 * 128;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sw = variant >> 2;          // swizzle hypothesis (0..2), only for MN-major BASE32B variants (variant >= 8)
  for (int i = tid; i < 128 * 128; i += blockDim.x) {
    const int r = i >> 7, f = i & 127;
    int off;
    if (variant < 8) {
      off = ((r >> 3) * 32 + (f >> 2)) * 32 + (r & 7) * 4 + (f & 3);
    } else {
      const int fi = f & 31;
      int within;                       // byte offset inside the 128-byte row of the atom
      if (sw == 2) within = ((((fi >> 3) ^ (r & 3)) << 5) + ((fi & 7) << 2));                       // 32-byte chunks XOR k-row
      else if (sw == 3) within = ((fi >> 2) << 4) + ((((fi & 3) ^ ((fi >> 2) & 3))) << 2);        // element-in-chunk XOR chunk
      else within = fi << 2;                                                                        // no swizzle
      off = ((f >> 5) * 512 + (r >> 2) * 2048 + (r & 3) * 128 + within) >> 2;
    }
    sA[off] = P[i];
    sB[off] = Q[i];
  }
  if (warp == 0) tc_tmem_alloc(&tmem_s, 128);
  if (tid == 0) tc_mbar_init(&bar, 1);
  tc_fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  if (tid == 0) {
    const uint32_t a0 = tc_smem_u32(sA), b0 = tc_smem_u32(sB);
    for (int kk = 0; kk < 128; kk += 8) {
      uint64_t ad, bd;
      uint32_t id;
      if (variant >= 8) {
        // BASE32B MN-major: atoms 32 mn x 4 k; mn-atom stride 512 B, k-atom stride 2048 B; 2 k-atoms per K=8 MMA
        const bool swap = variant & 1;
        const uint32_t lbo = swap ? 2048 : 512, sbo = swap ? 512 : 2048;
        ad = tc_desc(a0 + (kk >> 2) * 2048, lbo, sbo) | ((uint64_t)1 << 61);
        bd = tc_desc(b0 + (kk >> 2) * 2048, lbo, sbo) | ((uint64_t)1 << 61);
        id = idesc_v(128, 1, 1);
      } else if (variant == 4) {
        ad = tc_desc(a0 + (kk >> 2) * 128, 128, 4096);
        bd = tc_desc(b0 + (kk >> 2) * 128, 128, 4096);
        id = idesc_v(128, 0, 0);
      } else if (variant == 0 || variant == 1) {
        const uint32_t lbo = variant == 0 ? 4096 : 128, sbo = variant == 0 ? 128 : 4096;
        ad = tc_desc(a0 + (kk >> 3) * 4096, lbo, sbo);
        bd = tc_desc(b0 + (kk >> 3) * 4096, lbo, sbo);
        id = idesc_v(128, 1, 1);
      } else {
        // A: K-major tile of P (m = row, k = feature): D[m][n] = sum_f P[m][f] * Q[f][n]  (Q = [k rows x n feat], MN-major)
        const uint32_t lbo = variant == 2 ? 4096 : 128, sbo = variant == 2 ? 128 : 4096;
        ad = tc_desc(a0 + (kk >> 2) * 128, 128, 4096);
        bd = tc_desc(b0 + (kk >> 3) * 4096, lbo, sbo);
        id = idesc_v(128, 0, 1);
      }
      tc_mma_tf32(tmem, ad, bd, id, kk > 0);
    }
    tc_commit(&bar);
  }
  tc_mbar_wait(&bar, 0);
  tc_fence_after();
  if (warp < 4) {
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float v[32];
      tc_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + c0 + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_tmem_dealloc(tmem, 128);
}

int main() {
  std::vector<float> P(128 * 128), Q(128 * 128), D(128 * 128);
  srand(1);
  for (auto& x : P) x = (float)((rand() % 17) - 8) / 8.0f;   // exactly representable in tf32
  for (auto& x : Q) x = (float)((rand() % 13) - 6) / 4.0f;
  float *dP, *dQ, *dD;
  cudaMalloc(&dP, 65536); cudaMalloc(&dQ, 65536); cudaMalloc(&dD, 65536);
  cudaMemcpy(dP, P.data(), 65536, cudaMemcpyHostToDevice);
  cudaMemcpy(dQ, Q.data(), 65536, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536 + 1024);
  for (int variant : {4, 8, 9, 12, 13, 16, 17}) {
    cudaMemset(dD, 0, 65536);
    probe<<<1, 128, 2 * 65536, 0>>>(dP, dQ, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, 65536, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0, sumabs = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 128; ++n) {
        double ref = 0;
        if (variant == 4) for (int k = 0; k < 128; ++k) ref += (double)P[m * 128 + k] * Q[n * 128 + k];
        else if (variant < 2 || variant >= 8) for (int k = 0; k < 128; ++k) ref += (double)P[k * 128 + m] * Q[k * 128 + n];
        else for (int k = 0; k < 128; ++k) ref += (double)P[m * 128 + k] * Q[k * 128 + n];
        maxerr = fmax(maxerr, fabs(ref - D[m * 128 + n])); maxref = fmax(maxref, fabs(ref)); sumabs += fabs(D[m * 128 + n]);
      }
    printf("variant %d: err=%s maxerr=%g maxref=%g sum|D|=%g\n", variant, cudaGetErrorString(e), maxerr, maxref, sumabs);
  }
  return 0;
}
