// tcgen05 (5th-gen tensor core) building blocks of the ActorCritic path, TF32 inputs / FP32 accumulate in TMEM: PTX wrappers,
// shared-memory matrix descriptors and the instruction descriptor shared by gemm_tc2.cuh, wgrad_group.cuh and mlp_chain2.cuh.
// (The round-1 single-stage GEMM kernel that lived here was superseded by those and has been removed.)
//
// Same three operand modes and the same GemmArgs as the CUDA-core block (gemm_simt.cuh):
//   FWD      Y[m,n]  = act( beta*Y + sum_k X[m,k] W[n,k] + b[n] )      A = X,   B = W
//   BWD_DATA dX[m,n] = ( beta*dX + sum_k G[m,k] W[k,n] ) act'(Xact)    A = G,   B = W^T (transposed while filling smem)
//   BWD_WGT  dW[m,n] += sum_k G[k,m] X[k,n];  db[m] += sum_k G[k,m]    A = G^T, B = X^T (reduction over rows)
//
// Every layer of the widowGo1 networks has N, K <= 128 and only the row count is large, so one CTA owns a 128-row tile
// (FWD / BWD_DATA) or a slab of rows (BWD_WGT), keeps the whole weight operand in shared memory and needs ONE accumulator
// tile: D[128 x N<=128] = 128 TMEM columns.  Operands are written to shared memory by the CTA's threads in the canonical
// no-swizzle UMMA layouts (8-row x 16-byte core matrices, cute/atom/mma_traits_sm100.hpp) because the A operand is gathered
// through the mini-batch index (RS:189-201) and padded (K to 8, N to 16) on the fly; `tcgen05.mma.kind::tf32` is issued by
// one thread, completion arrives on an mbarrier through `tcgen05.commit`, the epilogue reads the accumulator with
// `tcgen05.ld` and fuses bias / ELU / tanh (forward) or the activation derivative (backward).
//
// TF32 keeps 10 mantissa bits of each input (FP32 accumulate): `precision="tf32"` mode, tolerances stated in the tests.
#pragma once
#include "gemm_simt.cuh"

namespace dwbc {

constexpr int TC_M = 128;           // rows of the accumulator tile == TMEM lanes
constexpr int TC_MAXK = 128;        // reduction chunk held in shared memory
constexpr int TC_MAXN = 128;

// ---- PTX wrappers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "TC_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TC_DONE;\n\t"
      "bra TC_WAIT;\n\t"
      "TC_DONE:\n\t}" ::"r"(tc_smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
// 32 lanes x 32 columns of FP32 accumulator -> 32 registers per thread (thread = lane = output row)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp: SmemDescriptor), offsets in bytes
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // base_offset = 0, lbo_mode = 0, layout_type = 0 (no swizzle)
}
// instruction descriptor (InstrDescriptor): D = F32, A = B = TF32, M = 128
__device__ __forceinline__ uint32_t tc_idesc(int n, bool a_mn_major, bool b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= 2u << 7;                       // a_format = TF32
  d |= 2u << 10;                      // b_format = TF32
  d |= (a_mn_major ? 1u : 0u) << 15;  // a_major
  d |= (b_mn_major ? 1u : 0u) << 16;  // b_major
  d |= (uint32_t)(n >> 3) << 17;      // n_dim
  d |= (uint32_t)(TC_M >> 4) << 24;   // m_dim
  return d;
}

// 3xTF32 split: kind::tf32 reads the 19 leading bits of a 32-bit operand element and TRUNCATES the 13 low mantissa bits (measured on
// B200 for operands in shared and in tensor memory, tools/probes/ts_probe.cu).  hi = that part (no instruction needed), lo = x - hi (exact).
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

inline bool tc_shape_ok(int mode, const GemmArgs& g) {
  if (mode == GEMM_BWD_WGT) return g.M <= TC_M && g.N <= TC_MAXN;
  return g.N <= TC_MAXN && g.K <= TC_MAXK;
}

}  // namespace dwbc
