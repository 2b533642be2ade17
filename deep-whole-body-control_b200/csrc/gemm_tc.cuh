// tcgen05 (5th-gen tensor core) GEMM building block of the ActorCritic path, TF32 inputs / FP32 accumulate in TMEM.
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

constexpr int TC_THREADS = 256;          // 8 warps: all fill operands; warps w and w+4 drain the two column halves of TMEM lane quarter w%4
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

// ---- operand fills (generic-proxy stores into the canonical no-swizzle layouts) -------------------------------------
// K-major: element (r, k) of a [rows_pad x kpad] tile at  ((r/8)*(kpad/4) + k/4)*128 + (r%8)*16 + (k%4)*4  bytes
//   -> core matrices of one 8-row group are contiguous along K:  LBO = 128 B, SBO = (kpad/4)*128 B
// Source: R.row(row0 + r)[k], valid for r < nrows and k < kvalid; everything else is zero-filled.
__device__ __forceinline__ void tc_fill_kmajor(float* smem, const RowMat& R, int64_t row0, int nrows, int rows_pad, int kvalid, int kpad,
                                               bool vec) {
  const int chunks = kpad >> 2, total = rows_pad * chunks;
  for (int base = threadIdx.x; base < total; base += 4 * TC_THREADS) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {   // four independent 16-byte loads in flight per thread
      const int i = base + u * TC_THREADS;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < total) {
        // lanes 0-7 -> 8 rows of one core matrix (conflict-free 128 B), next lanes -> next K chunk of the same rows
        const int r8 = i & 7, c = (i >> 3) % chunks, g = (i >> 3) / chunks;
        const int r = g * 8 + r8;
        if (r < nrows) {
          const float* src = R.row(row0 + r) + 4 * c;
          if (vec && 4 * c + 3 < kvalid) v[u] = ldg_stream(reinterpret_cast<const float4*>(src));
          else {
            if (4 * c + 0 < kvalid) v[u].x = src[0];
            if (4 * c + 1 < kvalid) v[u].y = src[1];
            if (4 * c + 2 < kvalid) v[u].z = src[2];
            if (4 * c + 3 < kvalid) v[u].w = src[3];
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * TC_THREADS;
      if (i < total) *reinterpret_cast<float4*>(smem + (size_t)i * 4) = v[u];   // canonical offset == i * 16 bytes
    }
  }
}
// Transposing K-major fill: the source is indexed [k][mn] (mn contiguous in global memory: W[k_out][n_in] for the data
// gradient, G[row][out] / X[row][in] for the weight gradient) and lands in the same K-major canonical layout as above with
// (r = mn, k):  ((mn/8)*(kpad/4) + k/4)*128 + (mn%8)*16 + (k%4)*4 bytes.  (The MN-major descriptor path -- a_major/b_major
// = 1 with TF32 operands -- returned zeros on B200 in round 1 and is not used; a shared-memory transpose costs 4 scalar
// stores per 16-byte load instead.)  Source valid for k < nk, mn < mnvalid; the rest is zero-filled.
__device__ __forceinline__ void tc_fill_kmajor_T(float* smem, const RowMat& R, int64_t k0, int nk, int kpad, int mnvalid, int mnpad, bool vec) {
  const int chunks = mnpad >> 2, kq = kpad >> 2, cg = (chunks + 3) >> 2, total = (kpad >> 3) * cg * 32;
  for (int base = threadIdx.x; base < total; base += 4 * TC_THREADS) {
    float4 v[4];
    int cc[4], kk[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * TC_THREADS;
      // lanes: 8 consecutive k x 4 consecutive mn-chunks (64 B per source row) -> <= 4-way bank conflicts on the scalar stores
      const int k8 = i & 7, c4 = (i >> 3) & 3, rest = i >> 5;
      const int c = (rest % cg) * 4 + c4, k = (rest / cg) * 8 + k8;
      cc[u] = (i < total && c < chunks) ? c : -1;
      kk[u] = k;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cc[u] >= 0 && k < nk) {
        const float* src = R.row(k0 + k) + 4 * c;
        if (vec && 4 * c + 3 < mnvalid) v[u] = ldg_stream(reinterpret_cast<const float4*>(src));
        else {
          if (4 * c + 0 < mnvalid) v[u].x = src[0];
          if (4 * c + 1 < mnvalid) v[u].y = src[1];
          if (4 * c + 2 < mnvalid) v[u].z = src[2];
          if (4 * c + 3 < mnvalid) v[u].w = src[3];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (cc[u] < 0) continue;
      const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      const int k = kk[u];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int mn = 4 * cc[u] + j;
        smem[((size_t)((mn >> 3) * kq + (k >> 2)) * 8 + (mn & 7)) * 4 + (k & 3)] = vv[j];
      }
    }
  }
}

struct TcShared {
  uint64_t bar;
  uint32_t tmem_base;
};

// One kernel for the three modes.  grid.x CTAs, each loops over work items (row tiles / row slabs).
template <int kMode>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const GemmArgs g, const int items, const int vecA, const int vecB) {
  extern __shared__ __align__(1024) float tc_smem[];
  __shared__ TcShared sh;
  float* sA = tc_smem;                               // [128 x 128] floats max
  float* sB = tc_smem + TC_M * TC_MAXK;              // [128 x 128] floats max
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tc_tmem_alloc(&sh.tmem_base, 128);
  if (tid == 0) {
    tc_mbar_init(&sh.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  uint32_t phase = 0;

  if (kMode != GEMM_BWD_WGT) {
    // ---------------- row-tile GEMM: D[128 rows x Npad] = A[128 x Kpad] * B^T -------------------------------------
    const int N = g.N, K = g.K;
    const int npad = (N + 15) & ~15, kpad = (K + 7) & ~7;
    // weight operand once per CTA
    if (kMode == GEMM_FWD) tc_fill_kmajor(sB, g.B, 0, N, npad, K, kpad, vecB);            // W[n][k], k contiguous
    else tc_fill_kmajor_T(sB, g.B, 0, K, kpad, N, npad, vecB);                             // W[k][n], n contiguous -> transposed
    const uint32_t idesc = tc_idesc(npad, false, false);
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int64_t m0 = (int64_t)it * TC_M;
      const int rows = (int)min((int64_t)TC_M, (int64_t)g.M - m0);
      tc_fill_kmajor(sA, g.A, m0, rows, TC_M, K, kpad, vecA);
      tc_fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t a0 = tc_smem_u32(sA), b0 = tc_smem_u32(sB);
        for (int kk = 0; kk < kpad; kk += 8) {
          // K-major operand: the two 16-byte K chunks of this MMA are 128 B apart (LBO); row groups (kpad/4)*128 B apart (SBO)
          const uint64_t ad = tc_desc(a0 + (kk >> 2) * 128, 128, (kpad >> 2) * 128);
          const uint64_t bd = tc_desc(b0 + (kk >> 2) * 128, 128, (kpad >> 2) * 128);
          tc_mma_tf32(tmem, ad, bd, idesc, kk > 0 ? 1u : 0u);
        }
        tc_commit(&sh.bar);
      }
      tc_mbar_wait(&sh.bar, phase);
      phase ^= 1;
      tc_fence_after();
      // epilogue: thread = output row (TMEM lane quarter warp%4), warps w / w+4 take alternate 32-column blocks
      const int rloc = (warp & 3) * 32 + lane;
      const int64_t m = m0 + rloc;
      float* crow = g.C + m * g.ldc;
      const bool rowok = rloc < rows;
      const float* xrow = (kMode == GEMM_BWD_DATA && g.act != ACT_NONE && rowok) ? g.Xact.row(m) : nullptr;
      const bool v4 = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
                      (!xrow || (((reinterpret_cast<uintptr_t>(xrow)) & 15) == 0));
      for (int c0 = (warp >> 2) * 32; c0 < npad; c0 += 64) {
        float v[32];
        tc_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
        if (rowok) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const int n = c0 + 4 * j4;
            if (n >= N) break;
            float x[4] = {v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]};
            const bool full = v4 && (n + 3 < N);
            float old[4] = {0.f, 0.f, 0.f, 0.f}, y[4] = {0.f, 0.f, 0.f, 0.f};
            if (g.beta) {
              if (full) { float4 t = *reinterpret_cast<const float4*>(crow + n); old[0] = t.x; old[1] = t.y; old[2] = t.z; old[3] = t.w; }
              else for (int q = 0; q < 4; ++q) if (n + q < N) old[q] = crow[n + q];
            }
            if (xrow) {
              if (full) { float4 t = *reinterpret_cast<const float4*>(xrow + n); y[0] = t.x; y[1] = t.y; y[2] = t.z; y[3] = t.w; }
              else for (int q = 0; q < 4; ++q) if (n + q < N) y[q] = xrow[n + q];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float t = x[q] + old[q];
              if (kMode == GEMM_FWD) {
                if (g.bias && n + q < N) t += g.bias[n + q];
                if (g.act == ACT_ELU) t = elu_f(t);
                else if (g.act == ACT_TANH) t = tanhf(t);
              } else if (xrow) {
                if (g.act == ACT_ELU) t *= (y[q] > 0.0f ? 1.0f : y[q] + 1.0f);
                else if (g.act == ACT_TANH) t *= (1.0f - y[q] * y[q]);
              }
              x[q] = t;
            }
            if (full) *reinterpret_cast<float4*>(crow + n) = make_float4(x[0], x[1], x[2], x[3]);
            else for (int q = 0; q < 4; ++q) if (n + q < N) crow[n + q] = x[q];
          }
        }
      }
      tc_fence_before();
      __syncthreads();   // accumulator drained and sA free before the next tile
      tc_fence_after();
    }
  } else {
    // ---------------- weight gradient: D[Nout(<=128) x Nin(<=128)] += G^T X over a slab of rows ---------------------
    const int Mo = g.M, Ni = g.N;                 // dW is Mo x Ni
    const int nipad = (Ni + 15) & ~15;
    const uint32_t idesc = tc_idesc(nipad, false, false);
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int64_t k_begin = (int64_t)it * g.k_chunk;
      const int64_t k_end = min((int64_t)g.K, k_begin + g.k_chunk);
      bool first = true;
      for (int64_t k0 = k_begin; k0 < k_end; k0 += TC_MAXK) {
        const int nk = (int)min((int64_t)TC_MAXK, k_end - k0);
        const int kpad = (nk + 7) & ~7;
        tc_fill_kmajor_T(sA, g.A, k0, nk, kpad, Mo, TC_M, vecA);      // A(m = out feature, k = row) = G[row][out]
        tc_fill_kmajor_T(sB, g.B, k0, nk, kpad, Ni, nipad, vecB);     // B(n = in feature,  k = row) = X[row][in]
        tc_fence_async_smem();
        __syncthreads();
        if (tid == 0) {
          tc_fence_after();
          const uint32_t a0 = tc_smem_u32(sA), b0 = tc_smem_u32(sB);
          for (int kk = 0; kk < kpad; kk += 8) {
            const uint64_t ad = tc_desc(a0 + (kk >> 2) * 128, 128, (kpad >> 2) * 128);
            const uint64_t bd = tc_desc(b0 + (kk >> 2) * 128, 128, (kpad >> 2) * 128);
            tc_mma_tf32(tmem, ad, bd, idesc, (first && kk == 0) ? 0u : 1u);
          }
          tc_commit(&sh.bar);
        }
        first = false;
        tc_mbar_wait(&sh.bar, phase);   // operands consumed: sA / sB may be refilled
        phase ^= 1;
        tc_fence_after();
        // bias gradient: column sums of G over this chunk (thread = out feature)
        if (g.dbias && tid < 2 * TC_M && (tid >> 1) < Mo) {   // two threads per out feature, each half of the chunk
          const int o = tid >> 1, h = tid & 1;
          float s = 0.0f;
          for (int k = h * (kpad >> 1); k < (h + 1) * (kpad >> 1) && k < nk; ++k) s += sA[((size_t)((o >> 3) * (kpad >> 2) + (k >> 2)) * 8 + (o & 7)) * 4 + (k & 3)];
          atomicAdd(g.dbias + o, s);
        }
        __syncthreads();
      }
      // epilogue: split-K partial -> global (atomics); thread = out feature row
      {
        const int o = (warp & 3) * 32 + lane;
        for (int c0 = (warp >> 2) * 32; c0 < nipad; c0 += 64) {
          float v[32];
          tc_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
          if (o < Mo) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < Ni) atomicAdd(g.C + (int64_t)o * g.ldc + c0 + j, v[j]);
          }
        }
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  __syncthreads();
  if (warp == 0) tc_tmem_dealloc(tmem, 128);
}

inline bool tc_shape_ok(int mode, const GemmArgs& g) {
  if (mode == GEMM_BWD_WGT) return g.M <= TC_M && g.N <= TC_MAXN;
  return g.N <= TC_MAXN && g.K <= TC_MAXK;
}

template <int kMode>
inline int launch_gemm_tc(const GemmArgs& g_in, cudaStream_t st) {
  GemmArgs g = g_in;
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return DWBC_ERR_ARG;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int items;
  if (kMode == GEMM_BWD_WGT) {
    int64_t chunk = (g.K + sms - 1) / sms;
    chunk = (chunk + 127) / 128 * 128;
    if (chunk < 128) chunk = 128;
    g.k_chunk = (int)chunk;
    items = (int)((g.K + chunk - 1) / chunk);
  } else {
    items = (g.M + TC_M - 1) / TC_M;
  }
  const int grid = items < sms ? items : sms;
  const size_t smem = (size_t)(TC_M * TC_MAXK + TC_MAXN * TC_MAXK) * sizeof(float);
  static bool attr[3] = {false, false, false};
  if (!attr[kMode]) {
    if (cudaFuncSetAttribute(gemm_tc_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return DWBC_ERR_LAUNCH;
    attr[kMode] = true;
  }
  gemm_tc_kernel<kMode><<<grid, TC_THREADS, smem, st>>>(g, items, rowmat_vec_ok(g.A) ? 1 : 0, rowmat_vec_ok(g.B) ? 1 : 0);
  ++dwbc_launch_counter;
  return cudaGetLastError() == cudaSuccess ? DWBC_OK : DWBC_ERR_LAUNCH;
}

}  // namespace dwbc
