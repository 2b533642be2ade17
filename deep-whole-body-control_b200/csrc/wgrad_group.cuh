// All weight-gradient GEMMs of one PPO mini-batch in ONE persistent launch (tcgen05, TF32, MN-major operands).
//
//   dW_l[out x in] += dZ_l^T X_l ,  db_l[out] += colsum(dZ_l)        for every layer l, reduction over the mini-batch rows
//
// Layer by layer (gemm_tc2_kernel<GEMM_BWD_WGT>) each GEMM is a 128-CTA launch of ~20 us, a third of which is the split-K
// epilogue, plus launch gaps and tails.  Here a work item is (layer, slab of `slab` rows); the items are dealt round-robin to
// one persistent CTA per SM, so the epilogue is amortised over 4x longer slabs and there are no gaps.
// Operand tiles, the SWIZZLE_128B_BASE32B MN-major layout, the 8-warp cp.async fill with asynchronous barrier arrival, the
// bias-gradient column sums and the staged vector-reduction epilogue are those of gemm_tc2.cuh (weight-gradient branch).
#pragma once
#include <stdlib.h>

#include <algorithm>

#include "gemm_tc2.cuh"

namespace dwbc {

constexpr int WG_MAX = 20;
// warp roles: WG_PW producer warps (operand copies; 3xTF32: and the low-part split), one MMA warp, four epilogue warps (bias-gradient column
// sums of every chunk, then the split-K reduction of the accumulator).  The three roles only meet through mbarriers: the producers run up
// to a whole stage ring ahead, the accumulator is double-buffered in tensor memory and the epilogue has its own staging tile, so the
// copies of the next work item are in flight while the previous one is reduced.
constexpr int WG_PW = 16;
constexpr int WG_PROD = 32 * WG_PW;               // producer threads
constexpr int WG_FILL = WG_PROD;                  // threads that copy operand pieces
constexpr int WG_THREADS = WG_PROD + 32 + T2_EPI;
constexpr int WG_KS32 = WG_FILL / 32;             // row step of a copy thread on a 128-wide operand (32 pieces per row)
constexpr int WG_LDS = 36;                        // row stride (floats) of the epilogue staging tile [128][32 + 4]
struct WGItem {
  RowMat G, X;          // dZ [rows x Mo], X [rows x Ni]
  float* dW; int64_t lddw;
  float* db;            // nullable
  int Mo, Ni;
  int fastG, fastX;     // 16-byte aligned rows and width % 4 == 0: cp.async path
};
// snake != 0 (experiment, OFF by default): the GEMMs are sorted by operand width (host) and the work items, enumerated GEMM-major, are
// dealt to the CTAs boustrophedon (round j forwards for even j, backwards for odd j), so that every CTA gets the same mix of wide and
// narrow items -- dealt round-robin in construction order (snake == 0) the heaviest CTA carries 1.22 x the mean operand bytes of a
// widowGo1 mini-batch.  MEASURED SLOWER on B200 (update() 21.83 against 21.12 ms, 3xTF32; 17.36 against 16.92 ms, TF32): GEMM-major
// order makes all 148 CTAs reduce into the SAME 64 KB dW at the same time (red.global.add contention in L2), round-robin spreads the
// epilogues of one moment over all 17 GEMMs.
// rev != 0: slabs are taken from the LAST rows to the first.  The backward chain that runs right before this kernel walks the tiles upwards, so
// its most recent writes (dZ images) and reads (activation images) -- the part of the 0.5 GB working set that is still in the 126 MB L2 --
// belong to the last rows; walking upwards too, this kernel started with the rows that had been evicted longest ago.
struct WGroup { int n, rows, slab, nslab, snake, rev; WGItem it[WG_MAX]; };

// elect.sync: true in exactly one lane of the (converged) warp
__device__ __forceinline__ bool wg_elect() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// profiling aid (tools/wgrad_profile.py): clock64 stamps of the CTA's second work item, [grid][64] (dwbc_debug_set_wg_cycle_buffer)
__device__ unsigned long long* g_wg_cycles = nullptr;
#define WG_STAMP(slot) do { if (g_wg_cycles && (slot) < 64) g_wg_cycles[blockIdx.x * 64 + (slot)] = clock64(); } while (0)

struct WGShared {
  uint64_t full[4], empty[4], lo_empty[2], tfull[2], tempty[2];
  uint32_t tmem_base;
  int64_t rowoff[2][128];      // (scalar-copied operands only) row offsets (floats) of the chunk being copied and of the next one: [slot][0..63] = G rows, [64..127] = X rows
};
// the 16-byte column piece a copy thread owns in a vector-copied operand: the thread copies rows k0, k0 + WG_KS32, ... of every chunk (the
// step is a multiple of 4, so the swizzle phase k & 3 and with it everything but a constant stride of the shared-memory address is fixed).
// Row-major operand: a warp takes one whole row per instruction (c4 = lane, idle lanes past the row's width); tile image: 8 rows x 4 pieces.
struct WGCol { int fixed, active, c4, k0; uint32_t doff; };

// X3 = error-compensated mode (3xTF32): chunks are 32 rows; four raw stages {G, X} receive the cp.async copies, two more buffers hold the
// low parts G_lo = G - trunc_tf32(G), X_lo of the chunk about to be multiplied (the tensor core truncates the 13 low mantissa bits of the
// raw tiles itself), and every chunk is three accumulating MMA groups: G^T X + G_lo^T X + G^T X_lo.  The thread that copied a 16-byte
// piece also splits it (after cp.async.wait_group), two chunks behind its copies, so ~96 KB of copies are in flight per SM.
template <bool X3>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_group_kernel(const __grid_constant__ WGroup grp) {
  extern __shared__ __align__(1024) float wg_smem[];
  __shared__ WGShared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int WCH = X3 ? 32 : T2_WCH;             // rows per chunk
  constexpr int TILE = WCH * 128;                   // floats per operand tile: WCH rows x 128 features
  constexpr int STAGE = 2 * TILE;                   // {G, X}
  constexpr int NST = X3 ? 4 : 3;                   // raw stages
  constexpr int AHEAD = 2;                          // X3: chunks the copies run ahead of the split (3 measured slower: 22.5 vs 21.0 ms per update)
  float* const lo_smem = wg_smem + NST * STAGE;     // X3: two {G_lo, X_lo} buffers
  float* const stg = wg_smem + 6 * T2_WCH * 128;    // epilogue staging tile [128][WG_LDS], behind the 192 KB of operand buffers
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      tc_mbar_init(&sh.full[i], WG_FILL);
      tc_mbar_init(&sh.empty[i], 5);                // MMA commit + one lane of each of the 4 epilogue warps (bias-gradient reads)
    }
    tc_mbar_init(&sh.lo_empty[0], 1);
    tc_mbar_init(&sh.lo_empty[1], 1);
    for (int i = 0; i < 2; ++i) { tc_mbar_init(&sh.tfull[i], 1); tc_mbar_init(&sh.tempty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WG_PW) tc_tmem_alloc(&sh.tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  if (tid == 0) WG_STAMP(62);
  const int items = grp.n * grp.nslab;
  const int nb = (int)gridDim.x, bid = (int)blockIdx.x;
  int my_items;
  if (grp.snake) {
    const int full = items / nb, rem = items - full * nb;        // the last, partial round runs forwards or backwards like any other
    my_items = full + ((((full & 1) ? nb - 1 - bid : bid) < rem) ? 1 : 0);
  } else {
    my_items = bid < items ? (items - bid + nb - 1) / nb : 0;
  }

  // byte offset of the 16-byte piece (row k, piece c4) inside an operand tile (SWIZZLE_128B_BASE32B, MN-major)
  auto piece_off = [](int k, int c4) -> uint32_t {
    return (uint32_t)((c4 >> 3) * 512 + (k >> 2) * 2048 + (k & 3) * 128 + ((((c4 >> 1) & 3) ^ (k & 3)) << 5) + ((c4 & 1) << 4));
  };
  constexpr int NR = WCH / WG_KS32;                 // rows of a chunk per copy thread and operand
  constexpr uint32_t DSTEP = (WG_KS32 >> 2) * 2048; // shared-memory bytes between them
  auto make_col = [&](int ncol, bool fast, bool image, int pt) -> WGCol {
    WGCol f{};
    if (!fast) return f;                              // scalar copies through the row-offset table
    f.fixed = 1;
    if (image) {
      // tile image: eight consecutive rows share each 128-byte line (row r at byte 16 * (r & 7) of piece c4's line), so a warp takes
      // 8 rows x 4 pieces = four whole lines per copy instruction instead of 32 half-used sectors
      static_assert(WG_FILL % 256 == 0 && 8 * (WG_FILL / 256) == WG_KS32, "image mapping");
      f.k0 = (pt & 7) + 8 * (pt >> 8); f.c4 = (pt >> 3) & 31; f.active = 1;
    } else {
      f.k0 = pt >> 5; f.c4 = pt & 31; f.active = f.c4 < (ncol >> 2);
    }
    f.doff = piece_off(f.k0, f.c4);
    return f;
  };
  // byte offsets of this thread's rows of the chunk starting at row r0 (-1 = past the end of the slab: zero-filled).  For a gathered operand
  // this is the load of the mini-batch index; it is issued one chunk ahead of its use.
  auto rows_of = [&](const WGItem& g, const WGCol (&col)[2], int64_t r0, int64_t k_end, int64_t (&out)[2][NR]) {
#pragma unroll
    for (int op = 0; op < 2; ++op) {
      const RowMat& R = op == 0 ? g.G : g.X;
      const bool on = col[op].fixed && col[op].active;
      const int64_t rb = r0 + col[op].k0;
      if (R.rpg == 0) {                                // tile image (vector-copied operands have rpg <= 1: WGroupBuilder)
#pragma unroll
        for (int n = 0; n < NR; ++n) {
          const int64_t r = rb + n * WG_KS32;
          out[op][n] = (on && r < k_end) ? 4 * ((r >> 7) * 16384 + ((r & 127) >> 3) * 1024 + (r & 7) * 4) : -1;
        }
      } else {
        const int64_t* idx = R.idx;
        const int64_t sb = 4 * R.stride_g;
#pragma unroll
        for (int n = 0; n < NR; ++n) {
          const int64_t r = rb + n * WG_KS32;
          out[op][n] = (on && r < k_end) ? (idx ? idx[r] : r) * sb : -1;
        }
      }
    }
  };
  // row-offset table (scalar-copied operands only) of the chunk starting at k0; `pt` < 64 -> G rows, 64 <= pt < 128 -> X rows
  auto row_offset = [&](const WGItem& g, int64_t k0, int nk, int pt) -> int64_t {
    if (pt < WCH) return pt < nk ? (g.G.row(k0 + pt) - g.G.p) : 0;
    if (pt >= 64 && pt < 64 + WCH) return (pt - 64) < nk ? (g.X.row(k0 + pt - 64) - g.X.p) : 0;
    return 0;
  };
  // copies of one chunk; `cc` is the CTA-wide running chunk counter (stage cc % NST, use cc / NST): identical in every thread
  auto fill_chunk = [&](const WGItem& g, const WGCol (&col)[2], const int64_t (&rows)[2][NR], int nk, uint32_t cc, int slot, int pt) {
    const int s = cc % NST;
    tc_mbar_wait(&sh.empty[s], ((cc / NST) & 1) ^ 1);
#pragma unroll
    for (int op = 0; op < 2; ++op) {
      float* dst = wg_smem + s * STAGE + op * TILE;
      const RowMat& R = op == 0 ? g.G : g.X;
      const float* base = R.p;
      const WGCol& f = col[op];
      if (f.fixed) {
        if (f.active) {
          const char* cb = reinterpret_cast<const char*>(base + (R.image() ? 32 : 4) * f.c4);     // piece stride: 4 floats (row-major) or 32 (tile image)
          const uint32_t d = tc_smem_u32(dst) + f.doff;
#pragma unroll
          for (int n = 0; n < NR; ++n) {
            const int64_t o = rows[op][n];
            const bool v = o >= 0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d + n * DSTEP), "l"(v ? cb + o : reinterpret_cast<const char*>(base)), "r"(v ? 16 : 0) : "memory");
          }
        }
      } else {
        const int64_t* ro = sh.rowoff[slot] + 64 * op;
        const int ncol = op == 0 ? g.Mo : g.Ni;
        for (int i = pt; i < WCH * ncol; i += WG_FILL) {
          const int k = i / ncol, f = i - k * ncol;
          const int off = ((f >> 5) * 512 + (k >> 2) * 2048 + (k & 3) * 128 + (((((f & 31) >> 3)) ^ (k & 3)) << 5) + ((f & 7) << 2)) >> 2;
          dst[off] = k < nk ? base[ro[k] + f] : 0.0f;
        }
      }
    }
    if (!X3) asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc_smem_u32(&sh.full[s])) : "memory");
    else asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // X3: low parts of the pieces this thread copied into stage cc % NST (its copies have landed: cp.async.wait_group before the call)
  auto split_chunk = [&](const WGItem& g, const WGCol (&col)[2], uint32_t cc, int pt) {
    const int s = cc % NST, l = cc & 1;
    tc_mbar_wait(&sh.lo_empty[l], ((cc >> 1) & 1) ^ 1);
#pragma unroll
    for (int op = 0; op < 2; ++op) {
      const float* src = wg_smem + s * STAGE + op * TILE;
      float* dst = lo_smem + l * STAGE + op * TILE;
      const WGCol& f = col[op];
      if (f.fixed) {
        if (f.active) {
          const float* sp = src + (f.doff >> 2);
          float* dp = dst + (f.doff >> 2);
#pragma unroll
          for (int n = 0; n < NR; ++n) {
            const float4 v = *reinterpret_cast<const float4*>(sp + n * (DSTEP >> 2));
            *reinterpret_cast<float4*>(dp + n * (DSTEP >> 2)) = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
          }
        }
      } else {
        const int ncol = op == 0 ? g.Mo : g.Ni;
        for (int i = pt; i < WCH * ncol; i += WG_FILL) {
          const int k = i / ncol, f = i - k * ncol;
          const int off = ((f >> 5) * 512 + (k >> 2) * 2048 + (k & 3) * 128 + (((((f & 31) >> 3)) ^ (k & 3)) << 5) + ((f & 7) << 2)) >> 2;
          dst[off] = tf32_lo(src[off]);
        }
      }
    }
    tc_fence_async_smem();
    t2_arrive(&sh.full[s]);
  };
  // the copy schedule of one item for a producer thread.  Vector-copied operands need no synchronisation among the producers: every thread
  // addresses its own rows (their offsets fetched one chunk ahead).  Only an item with a scalar-copied operand (a width that is not a
  // multiple of 4, or unaligned rows) goes through the shared row-offset table, double-buffered, one producer-wide barrier per chunk.
  auto fill_item = [&](const WGItem& g, int64_t k_begin, int64_t k_end, int nch, uint32_t cc, int pt, bool stamp) {
    const WGCol col[2] = {make_col(g.Mo, g.fastG, g.G.image(), pt), make_col(g.Ni, g.fastX, g.X.image(), pt)};
    const bool table = !g.fastG || !g.fastX;
    int64_t cur[2][NR], nx[2][NR];
    rows_of(g, col, k_begin, k_end, cur);
    if (table) {
      // (bar.sync is the warp-aligned form: reconverge first -- lane 0 may still be behind its mbarrier arrive)
      __syncwarp();
      asm volatile("bar.sync 4, %0;" ::"n"(WG_FILL) : "memory");   // tables of the previous item no longer read
      if (pt < 128) sh.rowoff[0][pt] = row_offset(g, k_begin, (int)min((int64_t)WCH, k_end - k_begin), pt);
    }
    for (int c = 0; c < nch; ++c) {
      const int64_t k0 = k_begin + (int64_t)c * WCH;
      int64_t tnext = 0;
      const bool more = table && c + 1 < nch && pt < 128;
      if (table) {
        __syncwarp();
        asm volatile("bar.sync 4, %0;" ::"n"(WG_FILL) : "memory");   // table c complete; table c-1 no longer read
        if (more) tnext = row_offset(g, k0 + WCH, (int)min((int64_t)WCH, k_end - k0 - WCH), pt);
      }
      if (c + 1 < nch) rows_of(g, col, k0 + WCH, k_end, nx);
      fill_chunk(g, col, cur, (int)min((int64_t)WCH, k_end - k0), cc + c, c & 1, pt);
      if (more) sh.rowoff[(c + 1) & 1][pt] = tnext;
      if (stamp && c < 12) WG_STAMP(c);
      if (X3 && c >= AHEAD) {
        asm volatile("cp.async.wait_group %0;" ::"n"(AHEAD) : "memory");
        split_chunk(g, col, cc + c - AHEAD, pt);
        if (stamp && c - AHEAD < 12) WG_STAMP(12 + c - AHEAD);
      }
#pragma unroll
      for (int op = 0; op < 2; ++op)
#pragma unroll
        for (int n = 0; n < NR; ++n) cur[op][n] = nx[op][n];
    }
    if (X3) {
      for (int c = max(nch - AHEAD, 0); c < nch; ++c) {
        const int left = nch - 1 - c;                    // copy groups issued after chunk c
        if (left >= 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else if (left == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        split_chunk(g, col, cc + c, pt);
      }
    }
  };

  uint32_t cc = 0;                                     // running chunk counter
  for (int j = 0; j < my_items; ++j) {
    int layer, slab_i;
    if (grp.snake) {
      const int e = j * nb + ((j & 1) ? nb - 1 - bid : bid);
      layer = e / grp.nslab;
      slab_i = e - layer * grp.nslab;
    } else {
      const int w = bid + j * nb;
      layer = w % grp.n;
      slab_i = w / grp.n;
    }
    const WGItem& g = grp.it[layer];
    if (grp.rev) slab_i = grp.nslab - 1 - slab_i;
    const int64_t k_begin = (int64_t)slab_i * grp.slab;
    const int64_t k_end = min((int64_t)grp.rows, k_begin + grp.slab);
    const int nch = (int)((k_end - k_begin + WCH - 1) / WCH);
    const int Mo = g.Mo, Ni = g.Ni, nipad = (Ni + 15) & ~15;
    if (warp < WG_PW) {
      if (tid == 0 && j == 1) WG_STAMP(60);
      fill_item(g, k_begin, k_end, nch, cc, tid, tid == 0 && j == 1);
    } else if (warp == WG_PW) {
      // the whole warp runs the loop (converged, warp-uniform values); the tcgen05 instructions sit under elect.sync -- issued from inside
      // `if (lane == 0)` every tcgen05.mma was wrapped in an elect / R2UR.BROADCAST / branch loop (operands not provably uniform)
      const uint32_t idesc = tc_idesc(nipad, true, true);
      const uint32_t acc = tmem + (uint32_t)(j & 1) * 128;
      if (nch > 0) {
        tc_mbar_wait(&sh.tempty[j & 1], ((j >> 1) & 1) ^ 1);      // the epilogue of item j-2 has read this accumulator
        tc_fence_after();
      }
      for (int c = 0; c < nch; ++c) {
        const uint32_t u = cc + c;
        const int s = u % NST;
        tc_mbar_wait(&sh.full[s], (u / NST) & 1);
        tc_fence_async_smem();          // generic-proxy writes of the producers (made visible by the barrier) -> async proxy reads of the MMA
        tc_fence_after();
        if (lane == 0 && j == 1 && c < 12) WG_STAMP(24 + c);
        const uint32_t a0 = tc_smem_u32(wg_smem + s * STAGE), b0 = a0 + TILE * 4;
        const uint32_t al = tc_smem_u32(lo_smem + (u & 1) * STAGE), bl = al + TILE * 4;
        if (wg_elect()) {
          uint64_t ad = tc_desc(a0, 512, 2048) | ((uint64_t)1 << 61), bd = tc_desc(b0, 512, 2048) | ((uint64_t)1 << 61);
          uint64_t adl = tc_desc(al, 512, 2048) | ((uint64_t)1 << 61), bdl = tc_desc(bl, 512, 2048) | ((uint64_t)1 << 61);
#pragma unroll
          for (int kk = 0; kk < WCH; kk += 8) {           // one K step = 8 rows = two 4-row atoms: +4096 bytes = +256 in the address field
            tc_mma_tf32(acc, ad, bd, idesc, (c > 0 || kk > 0) ? 1u : 0u);
            if (X3) {
              tc_mma_tf32(acc, adl, bd, idesc, 1u);
              tc_mma_tf32(acc, ad, bdl, idesc, 1u);
              adl += 256; bdl += 256;
            }
            ad += 256; bd += 256;
          }
          tc_commit(&sh.empty[s]);
          if (X3) tc_commit(&sh.lo_empty[u & 1]);
          if (c + 1 == nch) tc_commit(&sh.tfull[j & 1]);
          if (j == 1 && c < 12) WG_STAMP(36 + c);
        }
        __syncwarp();
      }
    } else {
      const int et = tid - (WG_PROD + 32);             // 0..127: output feature whose bias gradient / accumulator row this thread owns
      const int q = warp & 3;
      float bsum = 0.0f;
      const int fo = (et >> 5) * 128 + (et & 7);       // float offset of feature et inside row 0 of its atom (before the chunk swizzle)
      const int c32 = (et & 31) >> 3;
      auto bias_chunk = [&](uint32_t u) {
        const int s = u % NST;
        tc_mbar_wait(&sh.full[s], (u / NST) & 1);
        const float* gt = wg_smem + s * STAGE;
        if (g.db != nullptr && et < Mo) {
          // four independent partial sums (one per row of the 4-row atoms): the loads of a chunk are all in flight before the first add
          float p4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
          for (int k = 0; k < WCH; ++k) p4[k & 3] += gt[fo + (k >> 2) * 512 + (k & 3) * 32 + ((c32 ^ (k & 3)) << 3)];
          bsum += (p4[0] + p4[1]) + (p4[2] + p4[3]);
        }
        __syncwarp();
        if (lane == 0) t2_arrive(&sh.empty[s]);
      };
      for (int c = 0; c < nch; ++c) {
        bias_chunk(cc + c);
        if (et == 0 && j == 1 && c < 12) WG_STAMP(48 + c);
      }
      if (nch > 0) {
        if (g.db && et < Mo) atomicAdd(g.db + et, bsum);
        const int o = q * 32 + lane;
        tc_mbar_wait(&sh.tfull[j & 1], (j >> 1) & 1);
        tc_fence_after();
        // split-K partial: 32 accumulator columns at a time through the staging tile (warp q owns rows q*32 .. q*32+31 of it), out as
        // row-contiguous 128-byte vector reductions (eight lanes per row, four rows per instruction)
        const bool v4 = (g.lddw & 3) == 0 && (Ni & 3) == 0 && (reinterpret_cast<uintptr_t>(g.dW) & 15) == 0;
        const int rows_q = min(32, Mo - q * 32);
        for (int c0 = 0; c0 < nipad; c0 += 32) {
          float v[32];
          tc_ld32(tmem + (uint32_t)(j & 1) * 128 + ((uint32_t)(q * 32) << 16) + c0, v);
          __syncwarp();                                 // the previous column block's reads of the staging rows are done
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            *reinterpret_cast<float4*>(stg + (size_t)o * WG_LDS + 4 * j4) = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
          __syncwarp();
          if (v4) {
            const int cq = c0 + 4 * (lane & 7);
            for (int r = lane >> 3; r < rows_q; r += 4) {
              if (cq < Ni) {
                const float4 a = *reinterpret_cast<const float4*>(stg + (size_t)(q * 32 + r) * WG_LDS + 4 * (lane & 7));
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g.dW + (int64_t)(q * 32 + r) * g.lddw + cq), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
              }
            }
          } else {
            for (int r = 0; r < rows_q; ++r)
              if (c0 + lane < Ni) atomicAdd(g.dW + (int64_t)(q * 32 + r) * g.lddw + c0 + lane, stg[(size_t)(q * 32 + r) * WG_LDS + lane]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) t2_arrive(&sh.tempty[j & 1]);
      }
    }
    cc += nch;
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) WG_STAMP(63);
  if (warp == WG_PW) tc_tmem_dealloc(tmem, 256);
}

struct WGroupBuilder {
  WGroup g{};
  bool ok = true;
  void add(RowMat G, RowMat X, float* dW, int64_t lddw, float* db, int Mo, int Ni) {
    if (g.n >= WG_MAX || Mo > 128 || Ni > 128 || Mo <= 0 || Ni <= 0) { ok = false; return; }
    WGItem& it = g.it[g.n++];
    it.G = G; it.X = X; it.dW = dW; it.lddw = lddw; it.db = db; it.Mo = Mo; it.Ni = Ni;
    it.fastG = rowmat_vec_ok(G) && (Mo & 3) == 0 && G.rpg <= 1;
    it.fastX = rowmat_vec_ok(X) && (Ni & 3) == 0 && X.rpg <= 1;
    if ((G.rpg == 0 && !it.fastG) || (X.rpg == 0 && !it.fastX)) ok = false;      // tile images are always read in 16-byte pieces
  }
};

inline int wg_items_per_cta = 4;                     // tuning aid (dwbc_debug_set_wgrad_items)
inline int wg_reverse = 0;                            // tuning aid (dwbc_debug_set_wgrad_reverse): 0 = slabs from the first rows upwards
inline int wg_snake = 0;                             // tuning aid (dwbc_debug_set_wgrad_snake): 0 = round-robin deal in construction order

inline int launch_wgrad_group(WGroup& g, int rows, bool x3, cudaStream_t st) {
  if (g.n <= 0 || rows <= 0) return DWBC_ERR_ARG;
  g.snake = wg_snake;
  g.rev = wg_reverse;
  if (g.snake)         // widest operand pair first (a work item's time goes with the bytes it fetches per row)
    std::stable_sort(g.it, g.it + g.n, [](const WGItem& a, const WGItem& b) {
      return ((a.Mo + 15) & ~15) + ((a.Ni + 15) & ~15) > ((b.Mo + 15) & ~15) + ((b.Ni + 15) & ~15);
    });
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  g.rows = rows;
  // slab length: ~4 work items per CTA, at least 4 chunks so that the epilogue stays amortised
  const int per_cta = wg_items_per_cta;                // work items per CTA
  int64_t total_items = (int64_t)per_cta * sms;
  int64_t nslab = (total_items + g.n - 1) / g.n;
  int64_t slab = (rows + nslab - 1) / nslab;
  slab = (slab + T2_WCH - 1) / T2_WCH * T2_WCH;          // (a multiple of the 32-row chunks of the 3xTF32 mode too)
  if (slab < 4 * T2_WCH) slab = 4 * T2_WCH;
  g.slab = (int)slab;
  g.nslab = (int)((rows + slab - 1) / slab);
  const int items = g.n * g.nslab;
  const int grid = items < sms ? items : sms;
  const size_t smem = (size_t)(6 * T2_WCH * 128 + 128 * WG_LDS) * sizeof(float);       // 192 KB of operand buffers + the 18 KB staging tile
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(wgrad_group_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(wgrad_group_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return DWBC_ERR_LAUNCH;
    attr = true;
  }
  if (x3) wgrad_group_kernel<true><<<grid, WG_THREADS, smem, st>>>(g);
  else wgrad_group_kernel<false><<<grid, WG_THREADS, smem, st>>>(g);
  ++dwbc_launch_counter;
  return cudaGetLastError() == cudaSuccess ? DWBC_OK : DWBC_ERR_LAUNCH;
}

}  // namespace dwbc
