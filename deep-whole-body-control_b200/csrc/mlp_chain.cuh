// Fused layer chain on the tcgen05 tensor cores: a whole sub-network of the ActorCritic (AC:86-353) per launch.
//
// The layer-wise kernels (gemm_tc2.cuh) are bounded by operand fills and epilogue stores: every hidden activation makes a
// round trip through L2/HBM between two layers.  Here one persistent CTA owns a 128-row tile and runs a small PROGRAM of
// layer ops on it; activations stay in shared memory in the K-major operand layout of the next layer:
//
//   loads : gathered rows of the observation (mini-batch index, RS:189-201) -> operand tiles, cp.async, zero-filled tails
//   op    : D[128 x N] (TMEM) = A tile[128 x K] (smem) * W^T ;  W comes pre-packed in the UMMA canonical layout
//           (pack_weights_kernel, one tiny launch per call) and is fetched with ONE bulk async copy per op, bias appended;
//           epilogue: TMEM -> registers -> bias + ELU / tanh -> (a) operand tile of the next op (same or other buffer, at
//           a column offset: this is how cat([obs_prop, z]) of AC:211 is formed without a concat buffer), (b) optionally the
//           global activation buffer the backward pass needs, written from the tile with 64-byte row segments.
//
// Warp roles (288 threads): warps 0-3 tile loads, warp 4 one elected thread: weight bulk copies + tcgen05.mma
// + tcgen05.commit, warps 0-3 and 5-8 epilogue (TMEM lane quarter = warp % 4, the two warps of a quarter split the columns).  All hand-offs are mbarriers; the ops of one tile
// are strictly sequential (each needs the previous output), the weight copy of op i+1 overlaps the epilogue of op i and
// the global stores of op i overlap the MMAs of op i+1.
//
// Tiles use ONE geometry: [128 rows x 128 k] fp32, element (r, k) at float ((r/8)*32 + k/4)*32 + (r%8)*4 + k%4
// (8-row x 16-byte core matrices, LBO = 128 B, SBO = 4096 B); an op reads k < kpad only.
#pragma once
#include "gemm_tc2.cuh"

namespace dwbc {

constexpr int CH_MAX_OPS = 12, CH_MAX_LOADS = 3, CH_MAX_PACK = 24;
constexpr int CH_TILE = 128 * 128;                 // floats per operand tile
constexpr int CH_WBUF = 128 * 128 + 2 * 128;       // packed weight image + two bias slots (op parity)
constexpr int CH_NARROW = 128 * 32;                // narrow input tile (buffer 2): [128 rows x 32 k], 8 pieces per row group (SBO 1024 B)
constexpr int CH_SMEM_FLOATS = 2 * CH_TILE + CH_WBUF + CH_NARROW;
constexpr int CH_GROUPS = 4;                       // epilogue warp groups (4 warps each, one per TMEM lane quarter)
constexpr int CH_THREADS = 32 * (5 + 4 * (CH_GROUPS - 1));   // warps 0-3 (loads + epilogue group 1), 4 (MMA), 5-8 (group 0), 9.. (groups 2, 3)

struct ChainLoad {
  RowMat src;        // rows of the source (already offset to the first column)
  int ncols;         // columns copied (multiple of 4)
  int buf, col0;     // destination tile and column (multiple of 4)
  int zero_to;       // columns [col0 + ncols, zero_to) are zero-filled (K padding of the consuming op)
};
struct ChainOp {
  const float* wp;   // packed image: canonical K-major [npad x kpad] weights, then [npad] bias
  float* y;          // global output (nullable), row-major
  int64_t ldy;
  int a_buf, kpad;   // input tile (0, 1: full tiles; 2: narrow tile), padded reduction length (multiple of 8)
  int N, npad;       // outputs (npad: multiple of 16)
  int act;
  int out_buf, out_col0;   // next operand tile (-1: none) and column offset (multiple of 4)
  // ---- backward (data-gradient) ops: D = dZ * W, epilogue D (+ add) (*) act'(xact) instead of bias + act
  int mode;          // 0 forward, 1 backward
  int a_col0;        // first column of the A operand inside its tile (multiple of 4)
  int tslot, accum;  // TMEM accumulator (0/1); accum = 1: add onto what the slot already holds
  int no_epi;        // 1: leave the result in TMEM (a later op accumulates onto it)
  const float* xact; int64_t ldx;     // activation OUTPUT [M x ldx] whose derivative multiplies (staged through buffer 1); null: none
  const float* add; int64_t ldadd;    // optional addend [M x ldadd] (narrow ops only)
};
struct ChainProg {
  int M, n_loads, n_ops;
  ChainLoad ld[CH_MAX_LOADS];
  ChainOp op[CH_MAX_OPS];
};

// transpose = 0: image(n, k) = w[n*ldw + (k - k0)];  transpose = 1: image(n, k) = w[(k - k0)*ldw + n];  zero outside 0 <= k - k0 < K, n < N
// up to two independent programs over the same rows (actor and critic) share one launch: work items (tile, program) are
// dealt round-robin to the persistent CTAs, which evens out the load (320 tiles x {9, 7} ops over 148 CTAs) and lets the two
// networks of the rollout's act() run side by side
struct ChainProg2 { int nprog; ChainProg p[2]; };

struct PackItem { const float* w; int64_t ldw; const float* bias; int N, K, npad, kpad; int64_t dst; int transpose, k0; };
struct PackList { int n; float* out; PackItem it[CH_MAX_PACK]; };

// weights [N x K] (row stride ldw) -> canonical K-major image [npad x kpad] (+ bias[npad]); pads are zero
__global__ void pack_weights_kernel(const PackList pl) {
  const PackItem& it = pl.it[blockIdx.y];
  const int wn = it.npad * it.kpad, total = wn + it.npad;
  float* dst = pl.out + it.dst;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < wn) {
      const int n = i / it.kpad, k = i - n * it.kpad, kk = k - it.k0;
      const float v = (n < it.N && kk >= 0 && kk < it.K) ? (it.transpose ? it.w[(int64_t)kk * it.ldw + n] : it.w[(int64_t)n * it.ldw + kk]) : 0.0f;
      dst[((size_t)((n >> 3) * (it.kpad >> 2) + (k >> 2)) * 8 + (n & 7)) * 4 + (k & 3)] = v;
    } else {
      const int n = i - wn;
      dst[i] = (it.bias && n < it.N) ? it.bias[n] : 0.0f;
    }
  }
}

struct ChShared {
  uint64_t w_full, ld_full, mma_done, epi_done, tile_done, x_full;
  uint32_t tmem_base;
  int64_t rowoff[CH_MAX_LOADS][128];
};

__device__ __forceinline__ void ch_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ch_bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc_smem_u32(dst_smem)), "l"(src),
               "r"(bytes), "r"(tc_smem_u32(bar))
               : "memory");
}

template <int kAct>
__device__ __forceinline__ void ch_bias_act(float* v, const float* bias, int nvalid) {
  if (kAct == ACT_TANH) {                 // narrow output heads only: skip the columns beyond N
    for (int jj = 0; jj < 32; ++jj) v[jj] = jj < nvalid ? t2_tanh(v[jj] + bias[jj]) : 0.0f;
    return;
  }
#pragma unroll
  for (int jj = 0; jj < 32; ++jj) {
    float x = v[jj] + bias[jj];
    if (kAct == ACT_ELU) x = x > 0.0f ? x : __expf(x) - 1.0f;
    v[jj] = jj < nvalid ? x : 0.0f;
  }
}

__global__ void __launch_bounds__(CH_THREADS, 1) chain_fwd_kernel(const __grid_constant__ ChainProg2 pp, const int tiles) {
  extern __shared__ __align__(1024) float ch_smem[];
  __shared__ ChShared sh;
  float* buf[3] = {ch_smem, ch_smem + CH_TILE, ch_smem + 2 * CH_TILE + CH_WBUF};
  float* wbuf = ch_smem + 2 * CH_TILE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tc_mbar_init(&sh.w_full, 1);
    tc_mbar_init(&sh.ld_full, T2_PROD);
    tc_mbar_init(&sh.mma_done, 1);
    tc_mbar_init(&sh.epi_done, 128 * CH_GROUPS);
    tc_mbar_init(&sh.tile_done, T2_EPI);
    tc_mbar_init(&sh.x_full, T2_PROD);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tc_tmem_alloc(&sh.tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  if (tid == 0) T2_STAMP(62);
  const int items = tiles * pp.nprog;                 // item it = (tile it / nprog, program it % nprog)
  const int my_tiles = blockIdx.x < items ? (items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int nprog = pp.nprog;

  if (warp == 4) {
    // ===================== weight copies + MMA issue (one thread) =====================
    if (lane == 0 && my_tiles > 0) {
      // weights of op (global index n) -> wbuf, its bias -> bias slot n & 1: the epilogue of op n still reads its bias while
      // the weights of op n+1 stream in
      auto fetch_w = [&](const ChainOp& w, uint32_t nn) {
        const uint32_t wbytes = (uint32_t)(w.npad * w.kpad) * 4u, bbytes = (uint32_t)w.npad * 4u;
        ch_expect_tx(&sh.w_full, wbytes + bbytes);
        ch_bulk_g2s(wbuf, w.wp, wbytes, &sh.w_full);
        ch_bulk_g2s(wbuf + CH_TILE + (nn & 1) * 128, w.wp + w.npad * w.kpad, bbytes, &sh.w_full);
      };
      fetch_w(pp.p[blockIdx.x % nprog].op[0], 0);
      uint32_t n = 0;
      const uint32_t b0 = tc_smem_u32(wbuf);
      for (int j = 0; j < my_tiles; ++j) {
        const int it = blockIdx.x + j * gridDim.x;
        const ChainProg& pr = pp.p[it % nprog];
        const int nops = pr.n_ops;
        for (int i = 0; i < nops; ++i, ++n) {
          const ChainOp& o = pr.op[i];
          tc_mbar_wait(&sh.w_full, n & 1);
          if (n < 10) T2_STAMP(6 * n + 0);
          if (i == 0) tc_mbar_wait(&sh.ld_full, j & 1);
          if (n > 0) tc_mbar_wait(&sh.epi_done, (n - 1) & 1);   // previous epilogue: accumulator drained, its output tile written
          if (n < 10) T2_STAMP(6 * n + 1);
          tc_fence_async_smem();                                // generic-proxy tile writes (cp.async, epilogue stores) -> async-proxy MMA reads
          tc_fence_after();
          const uint32_t idesc = tc_idesc(o.npad, false, false);
          const uint32_t a0 = tc_smem_u32(buf[o.a_buf]) + (o.a_col0 >> 2) * 128;
          const uint32_t a_sbo = o.a_buf == 2 ? 1024u : 4096u;
          for (int kk = 0; kk < o.kpad; kk += 8) {
            const uint64_t ad = tc_desc(a0 + (kk >> 2) * 128, 128, a_sbo);
            const uint64_t bd = tc_desc(b0 + (kk >> 2) * 128, 128, (o.kpad >> 2) * 128);
            tc_mma_tf32(tmem + o.tslot * 128, ad, bd, idesc, (kk > 0 || o.accum) ? 1u : 0u);
          }
          tc_commit(&sh.mma_done);
          tc_mbar_wait(&sh.mma_done, n & 1);                    // weights consumed: the buffer may be refilled while the epilogue runs
          if (n < 10) T2_STAMP(6 * n + 2);
          if (i + 1 < nops) fetch_w(pr.op[i + 1], n + 1);
          else if (j + 1 < my_tiles) fetch_w(pp.p[(it + gridDim.x) % nprog].op[0], n + 1);
        }
      }
    }
  } else {
    // ===================== LOADS (warps 0-3) + EPILOGUE (warps 0-3 and 5-8) =====================
    // CH_GROUPS warps share each TMEM lane quarter (warp % 4) and split an op's 32-column chunks round-robin: group 0 = warps 5-8,
    // group 1 = warps 0-3 (idle between two tile loads otherwise), groups 2, 3 = warps 9-12, 13-16.  A warp copies to global memory
    // exactly the chunks it wrote itself (128 contiguous bytes per row), so no cross-warp synchronisation is needed.
    const int h = warp < 4 ? 1 : (warp < 9 ? 0 : 2 + ((warp - 9) >> 2));   // epilogue group: takes the 32-column chunks ci with ci % CH_GROUPS == h
    const int ptid = tid;                   // producer thread index (group 1 only)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;            // tile row of this thread
    uint32_t n = 0, nx = 0;
    for (int j = 0; j < my_tiles; ++j) {
      const int it = blockIdx.x + j * gridDim.x;
      const ChainProg& pr = pp.p[it % nprog];
      const int nops = pr.n_ops;
      const int64_t m0 = (int64_t)(it / nprog) * TC_M;
      const int rows = (int)min((int64_t)TC_M, (int64_t)pr.M - m0);
      if (h == 1) {
        t2_pbar();                                             // row-offset table of the previous tile no longer read
        for (int l = 0; l < pr.n_loads; ++l)
          sh.rowoff[l][ptid] = ptid < rows ? (pr.ld[l].src.row(m0 + ptid) - pr.ld[l].src.p) : 0;
        if (j > 0) tc_mbar_wait(&sh.tile_done, (j - 1) & 1);   // group 0 has finished the previous tile too (all MMAs retired before that)
        t2_pbar();
        for (int l = 0; l < pr.n_loads; ++l) {
          const ChainLoad& L = pr.ld[l];
          const int cpr = L.ncols >> 2, c40 = L.col0 >> 2;
          const int ppg = L.buf == 2 ? 8 : 32;                 // 16-byte pieces per row per 8-row group
          const uint32_t d0 = tc_smem_u32(buf[L.buf]);
          const float* base = L.src.p;
          const int64_t* ro = sh.rowoff[l];
          const int total = TC_M * cpr;
          const int z0 = (L.col0 + L.ncols) >> 2, nz = (L.zero_to >> 2) - z0;      // zero pieces per row (K padding), written first
          for (int i = ptid; i < TC_M * nz; i += T2_PROD) {
            const int rr = i / nz, cz = i - rr * nz;
            *reinterpret_cast<float4*>(buf[L.buf] + ((size_t)((rr >> 3) * ppg + z0 + cz) * 8 + (rr & 7)) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          for (int i = ptid; i < total; i += T2_PROD) {
            const int r8 = i & 7, rest = i >> 3;
            const int g = rest / cpr, cc = rest - g * cpr;
            const int rr = g * 8 + r8;
            const float* src = rr < rows ? base + ro[rr] + 4 * cc : base;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d0 + (uint32_t)(((g * ppg + c40 + cc) * 8 + r8) * 16)), "l"(src),
                         "r"(rr < rows ? 16 : 0)
                         : "memory");
          }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc_smem_u32(&sh.ld_full)) : "memory");
      }
      for (int i = 0; i < nops; ++i, ++n) {
        const ChainOp& o = pr.op[i];
        const bool stage_x = o.mode == 1 && o.xact != nullptr && !o.no_epi;
        if (stage_x && h == 1) {
          // activation tile whose derivative multiplies this op's result -> buffer 1, while the MMAs run.  Everybody has
          // finished reading buffer 1 for the previous op once its epi_done phase has completed.
          if (n > 0) tc_mbar_wait(&sh.epi_done, (n - 1) & 1);
          const int cpr = (o.N + 3) >> 2;
          const uint32_t d0 = tc_smem_u32(buf[1]);
          for (int i2 = ptid; i2 < TC_M * cpr; i2 += T2_PROD) {
            const int r8 = i2 & 7, rest = i2 >> 3;
            const int g = rest / cpr, cc = rest - g * cpr;
            const int rr = g * 8 + r8;
            const float* src = rr < rows ? o.xact + (m0 + rr) * o.ldx + 4 * cc : o.xact;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d0 + (uint32_t)(((g * 32 + cc) * 8 + r8) * 16)), "l"(src),
                         "r"(rr < rows ? 16 : 0)
                         : "memory");
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc_smem_u32(&sh.x_full)) : "memory");
        }
        tc_mbar_wait(&sh.mma_done, n & 1);
        tc_fence_after();
        if (o.no_epi) {                     // result stays in TMEM for a later accumulating op
          tc_fence_before();
          t2_arrive(&sh.epi_done);
          continue;
        }
        if (stage_x) { tc_mbar_wait(&sh.x_full, nx & 1); ++nx; }
        const float* bias = wbuf + CH_TILE + (n & 1) * 128;
        float* otile = o.out_buf >= 0 ? buf[o.out_buf] + ((size_t)((r >> 3) * 32 + (o.out_col0 >> 2)) * 8 + (r & 7)) * 4 : nullptr;
        const float* xtile = buf[1] + ((size_t)((r >> 3) * 32) * 8 + (r & 7)) * 4;
        const bool direct = o.y != nullptr && o.out_buf < 0;
        float* yrow = o.y ? o.y + (m0 + r) * o.ldy : nullptr;
        const bool yal = (o.ldy & 3) == 0 && (o.N & 3) == 0 && (reinterpret_cast<uintptr_t>(o.y) & 15) == 0;
        for (int c0 = 32 * h; c0 < o.npad; c0 += 32 * CH_GROUPS) {
          float v[32];
          tc_ld32(tmem + o.tslot * 128 + ((uint32_t)(q * 32) << 16) + c0, v);
          if (o.mode == 0) {
            // the activation is selected by a warp-uniform branch OUTSIDE the element loop (an if-converted tanh would be
            // issued for every ELU element otherwise); columns >= N are forced to zero (stale pad values never propagate)
            if (o.act == ACT_ELU) ch_bias_act<ACT_ELU>(v, bias + c0, o.N - c0);
            else if (o.act == ACT_TANH) ch_bias_act<ACT_TANH>(v, bias + c0, o.N - c0);
            else ch_bias_act<ACT_NONE>(v, bias + c0, o.N - c0);
          } else {
            if (o.add != nullptr && r < rows) {
              const float* ar = o.add + (m0 + r) * o.ldadd + c0;
#pragma unroll
              for (int jj = 0; jj < 32; ++jj)
                if (c0 + jj < o.N) v[jj] += ar[jj];
            }
            if (stage_x) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 y4 = *reinterpret_cast<const float4*>(xtile + (size_t)((c0 >> 2) + j4) * 32);
                const float yy[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float d = o.act == ACT_TANH ? 1.0f - yy[e] * yy[e] : (yy[e] > 0.0f ? 1.0f : yy[e] + 1.0f);   // AC ELU / tanh derivatives from the outputs
                  v[4 * j4 + e] *= d;
                }
              }
            }
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] = c0 + jj < o.N ? v[jj] : 0.0f;
          }
          if (otile) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              *reinterpret_cast<float4*>(otile + (size_t)((c0 >> 2) + j4) * 32) = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
          }
          if (direct && r < rows) {
            if (yal) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4)
                if (c0 + 4 * j4 < o.N) *reinterpret_cast<float4*>(yrow + c0 + 4 * j4) = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj)
                if (c0 + jj < o.N) yrow[c0 + jj] = v[jj];
            }
          }
        }
        tc_fence_before();
        if (otile) tc_fence_async_smem();
        if (tid == 160 && n < 10) T2_STAMP(6 * n + 3);
        t2_arrive(&sh.epi_done);
        if (o.y != nullptr && o.out_buf >= 0) {
          // global copy of the chunks this warp just wrote, out of the tile: 8 rows x 64 contiguous bytes per instruction
          __syncwarp();
          const int r8 = lane & 7, pp = lane >> 3;
          const float* t0 = buf[o.out_buf];
          for (int c0 = 32 * h; c0 < o.N; c0 += 32 * CH_GROUPS) {
            if (yal) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const int rr = q * 32 + g * 8 + r8;
                if (rr >= rows) continue;
                const float* trow = t0 + ((size_t)((rr >> 3) * 32 + ((o.out_col0 + c0) >> 2)) * 8 + r8) * 4;
                float* yr = o.y + (m0 + rr) * o.ldy + c0;
#pragma unroll
                for (int p0 = 0; p0 < 8; p0 += 4) {
                  const int piece = p0 + pp;
                  if (c0 + 4 * piece < o.N) *reinterpret_cast<float4*>(yr + 4 * piece) = *reinterpret_cast<const float4*>(trow + (size_t)piece * 32);
                }
              }
            } else if (r < rows) {
              const float* trow = t0 + ((size_t)((r >> 3) * 32 + ((o.out_col0 + c0) >> 2)) * 8 + (r & 7)) * 4;
              for (int cc = 0; cc < 32 && c0 + cc < o.N; ++cc) yrow[c0 + cc] = trow[(size_t)(cc >> 2) * 32 + (cc & 3)];
            }
          }
        }
        if (tid == 160 && n < 10) T2_STAMP(6 * n + 4);
      }
      if (h == 0) t2_arrive(&sh.tile_done);
    }
  }
  __syncthreads();
  if (tid == 0) T2_STAMP(63);
  if (warp == 4) tc_tmem_dealloc(tmem, 256);
}

// ---- host side ------------------------------------------------------------------------------------------------------
inline bool chain_aligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline int launch_pack(const PackList& pl, cudaStream_t st) {
  if (pl.n <= 0) return DWBC_OK;
  pack_weights_kernel<<<dim3(8, pl.n), 256, 0, st>>>(pl);
  ++dwbc_launch_counter;
  return cudaGetLastError() == cudaSuccess ? DWBC_OK : DWBC_ERR_LAUNCH;
}

inline int launch_chain2(const ChainProg* a, const ChainProg* b, cudaStream_t st) {
  ChainProg2 pp{};
  pp.nprog = b ? 2 : 1;
  pp.p[0] = *a;
  if (b) pp.p[1] = *b;
  for (int k = 0; k < pp.nprog; ++k) {
    const ChainProg& pr = pp.p[k];
    if (pr.M <= 0 || pr.M != pp.p[0].M || pr.n_ops <= 0 || pr.n_ops > CH_MAX_OPS || pr.n_loads < 0 || pr.n_loads > CH_MAX_LOADS) return DWBC_ERR_ARG;
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int tiles = (pp.p[0].M + TC_M - 1) / TC_M;
  const int items = tiles * pp.nprog;
  const int grid = items < sms ? items : sms;
  const size_t smem = (size_t)CH_SMEM_FLOATS * sizeof(float);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(chain_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return DWBC_ERR_LAUNCH;
    attr = true;
  }
  chain_fwd_kernel<<<grid, CH_THREADS, smem, st>>>(pp, tiles);
  ++dwbc_launch_counter;
  return cudaGetLastError() == cudaSuccess ? DWBC_OK : DWBC_ERR_LAUNCH;
}
inline int launch_chain(const ChainProg& pr, cudaStream_t st) { return launch_chain2(&pr, nullptr, st); }

// Small builder: keeps the pack list and the program in step.
struct ChainBuilder {
  ChainProg pr{};
  PackList* pl;
  int64_t* pack_off;       // running float offset into the packed-weight buffer
  bool ok = true;
  ChainBuilder(PackList* pl_, int64_t* off, int M) : pl(pl_), pack_off(off) { pr.M = M; }
  void load(RowMat src, int ncols, int buf, int col0, int zero_to) {
    if (pr.n_loads >= CH_MAX_LOADS || (ncols & 3) || (col0 & 3) || (zero_to & 3) || zero_to < col0 + ncols || zero_to > (buf == 2 ? 32 : 128) ||
        !chain_aligned(src.p) || (src.stride_g & 3) || (src.ld & 3)) { ok = false; return; }
    pr.ld[pr.n_loads++] = ChainLoad{src, ncols, buf, col0, zero_to};
  }
  // y = act(A[:, :K] W^T + b);  W [N x K] row-major with row stride ldw
  void op(const float* W, int64_t ldw, const float* bias, int N, int K, int act, int a_buf, int out_buf, int out_col0, float* y, int64_t ldy) {
    const int npad = (N + 15) & ~15, kpad = (K + 7) & ~7;
    if (pr.n_ops >= CH_MAX_OPS || pl->n >= CH_MAX_PACK || N > 128 || K > 128 || (out_col0 & 3) || (out_buf >= 0 && out_col0 + ((npad + 31) & ~31) > 128)) { ok = false; return; }
    PackItem& it = pl->it[pl->n++];
    it = PackItem{W, ldw, bias, N, K, npad, kpad, *pack_off, 0, 0};
    ChainOp o{};
    o.wp = pl->out ? pl->out + *pack_off : nullptr; o.y = y; o.ldy = ldy; o.a_buf = a_buf; o.kpad = kpad; o.N = N; o.npad = npad; o.act = act;
    o.out_buf = out_buf; o.out_col0 = out_col0;
    pr.op[pr.n_ops++] = o;
    *pack_off += (int64_t)npad * kpad + npad;
    *pack_off = (*pack_off + 63) & ~(int64_t)63;      // 256-byte aligned images (bulk copies need 16)
  }
  // backward op: dX[:, :Nin] = (dZ[:, :Kout] W[Kout x Nin] (+ add)) (*) act'(xact);  W row-major with row stride ldw.
  // kreal / k0: the dZ operand has kreal valid columns starting at column k0 of the kpad-wide A window (critic heads share one g_v tile).
  void bwd(const float* W, int64_t ldw, int Nin, int Kout, int k0, int kwin, int act, const float* xact, int64_t ldx, const float* add, int64_t ldadd,
           int a_buf, int a_col0, int out_buf, float* y, int64_t ldy, int tslot, int accum, int no_epi) {
    const int npad = (Nin + 15) & ~15, kpad = (kwin + 7) & ~7;
    if (pr.n_ops >= CH_MAX_OPS || pl->n >= CH_MAX_PACK || Nin > 128 || kpad > 128 || (a_col0 & 3) || (xact && ((ldx & 3) || !chain_aligned(xact))) ||
        (add && Nin > 32)) { ok = false; return; }
    PackItem& it = pl->it[pl->n++];
    it = PackItem{W, ldw, nullptr, Nin, Kout, npad, kpad, *pack_off, 1, k0};
    ChainOp o{};
    o.wp = pl->out ? pl->out + *pack_off : nullptr; o.y = y; o.ldy = ldy; o.a_buf = a_buf; o.kpad = kpad; o.N = Nin; o.npad = npad; o.act = act;
    o.out_buf = out_buf; o.out_col0 = 0; o.mode = 1; o.a_col0 = a_col0; o.tslot = tslot; o.accum = accum; o.no_epi = no_epi;
    o.xact = act == ACT_NONE ? nullptr : xact; o.ldx = ldx; o.add = add; o.ldadd = ldadd;
    pr.op[pr.n_ops++] = o;
    *pack_off += (int64_t)npad * kpad + npad;
    *pack_off = (*pack_off + 63) & ~(int64_t)63;
  }
};

}  // namespace dwbc
