// Probe (round-2 preparation): TMA tensor-map loads with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B land row-major fp32 tiles directly in
// the MN-major SWIZZLE_128B_BASE32B operand layout tcgen05.mma kind::tf32 needs for D[out x in] = P^T Q (weight gradients):
//   one box = 32 features (128 B) x 64 rows -> 8 KB, 32-byte chunk c of row r at c ^ (r & 3) (the hypothesis to verify);
//   descriptor: LBO (feature-atom stride) = 8192 B (next box), SBO (4-row atom stride) = 512 B, layout type 1; K = 8 rows per MMA.
// Build:  nvcc -I../../include -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tma_mn_probe tma_mn_probe.cu
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../deep-whole-body-control_b200/csrc/gemm_tc.cuh"
using namespace dwbc;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int ROWS = 64, COLS = 128, BOX_BYTES = 32 * 4 * ROWS;     // 8 KB per box

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(tc_smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(tc_smem_u32(bar))
               : "memory");
}

__global__ void probe(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapQ, float* D, float* dump) {
  extern __shared__ __align__(1024) float sm[];
  float* sA = sm;                       // 4 boxes of P
  float* sB = sm + 4 * BOX_BYTES / 4;   // 4 boxes of Q
  __shared__ uint64_t ld_bar, mma_bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tc_mbar_init(&ld_bar, 1);
    tc_mbar_init(&mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tc_tmem_alloc(&tmem_s, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(&ld_bar)), "r"(8 * BOX_BYTES) : "memory");
    for (int b = 0; b < 4; ++b) {
      tma_load_2d(sA + b * BOX_BYTES / 4, &mapP, 32 * b, 0, &ld_bar);
      tma_load_2d(sB + b * BOX_BYTES / 4, &mapQ, 32 * b, 0, &ld_bar);
    }
  }
  tc_mbar_wait(&ld_bar, 0);
  if (dump) for (int i = tid; i < 2048; i += blockDim.x) dump[i] = sA[i];      // first box of P as it sits in shared memory
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t a0 = tc_smem_u32(sA), b0 = tc_smem_u32(sB);
    const uint32_t idesc = tc_idesc(128, true, true);
    for (int kk = 0; kk < ROWS; kk += 8) {
      const uint64_t ad = tc_desc(a0 + kk * 128, 8192, 512) | ((uint64_t)1 << 61);
      const uint64_t bd = tc_desc(b0 + kk * 128, 8192, 512) | ((uint64_t)1 << 61);
      tc_mma_tf32(tmem, ad, bd, idesc, kk > 0);
    }
    tc_commit(&mma_bar);
  }
  tc_mbar_wait(&mma_bar, 0);
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
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
    printf("cuTensorMapEncodeTiled not available\n");
    return 1;
  }
  std::vector<float> P(ROWS * COLS), Q(ROWS * COLS), D(128 * 128), dump(2048);
  srand(2);
  for (auto& x : P) x = (float)((rand() % 17) - 8) / 8.0f;
  for (auto& x : Q) x = (float)((rand() % 13) - 6) / 4.0f;
  float *dP, *dQ, *dD, *dDump;
  cudaMalloc(&dP, P.size() * 4); cudaMalloc(&dQ, Q.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dDump, dump.size() * 4);
  cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap mp, mq;
  const cuuint64_t gdim[2] = {COLS, ROWS}, gstr[1] = {COLS * 4};
  const cuuint32_t box[2] = {32, ROWS}, estr[2] = {1, 1};
  CUresult r1 = encode(&mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dP, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = encode(&mq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dQ, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d %d\n", (int)r1, (int)r2);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * BOX_BYTES + 1024);
  cudaMemset(dD, 0, D.size() * 4);
  probe<<<1, 128, 8 * BOX_BYTES, 0>>>(mp, mq, dD, dDump);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(dump.data(), dDump, dump.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double ref = 0;
      for (int k = 0; k < ROWS; ++k) ref += (double)P[k * COLS + m] * Q[k * COLS + n];
      maxerr = fmax(maxerr, fabs(ref - D[m * 128 + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("err=%s  max|D - P^T Q| = %g (max |ref| %g)\n", cudaGetErrorString(e), maxerr, maxref);
  // where did the TMA put element (row r, feature f) of the first box?  expected: r*128 B + ((f/8) ^ (r & 3))*32 B + (f%8)*4 B
  int bad = 0;
  for (int r = 0; r < 8; ++r)
    for (int f = 0; f < 32; ++f) {
      const int off = r * 32 + (((f >> 3) ^ (r & 3)) << 3) + (f & 7);
      if (dump[off] != P[r * COLS + f]) ++bad;
    }
  printf("shared-memory layout check of box 0 (rows 0-7): %d mismatches against chunk32 ^ (row & 3)\n", bad);
  return 0;
}
