// corr1d_bwd_tca.cu -- tensor-core (tcgen05 + TMEM) backward of the 1 x P horizontal correlation, second generation:
// BOTH gradients build their A operand (the band matrix made of g) straight into TMEM, nothing of it goes through
// shared memory, and the warp roles are decoupled so that no warp pays more than two barrier round trips per K chunk.
//
//   mode 0: gin1[c,x0+x] = sum_j Gd[x][j] * in2[c][x0+oo0+j]      Gd[x][j] = g[p = j-delta0-x][x0+x]
//   mode 1: gin2[c,x0+x] = sum_j Gd[x][j] * in1[c][x0+oo1+j]      Gd[x][j] = g[p = x+c1-j][x0+oo1+j]
// Both are deterministic gathers (no atomics); one launch, CTAs [0, n_cta0) compute gin1, the rest gin2.
// A tile = 128 output columns x of one image row x one block of <= 128 channels:
//   D[x (M = 128 TMEM lanes)][c (N = Cbox columns)] = sum over the band (K = 32*NKC columns, 320 for P = 192).
//
// Why TMEM for both.  The first generation re-laid mode 1's A operand shared->shared (LDS + 2 STS per element, then
// the MMA read it back: 56 KB of shared-memory traffic per 32-column chunk against 24 KB here) and was bound by
// shared-memory bandwidth; mode 0 already wrote A into TMEM but its builders also split the feature band and paid
// ~6 barrier round trips (each ~100 cycles) per chunk.  tcgen05.st needs thread = TMEM lane = Gd row x, so the g values
// a warp fetches per instruction sit on a diagonal of the staged g slice:
//   mode 0: the slice is the tile's own [P][128] block of g; lane x reads column x of row p = j-delta0-x: the word
//           address is p*128 + x, consecutive lanes hit consecutive banks -> conflict-free;
//   mode 1: lane x needs g[p][x0 + x - s_p] (s_p = p - r): a different column shift per plane.  TMA cannot start a box
//           on a 4-byte boundary (an inner coordinate that is not a multiple of 16 bytes raises "illegal instruction",
//           scripts/microbench/tma_rows.cu), so the slice is loaded as [4 planes][136 columns] boxes whose start is the
//           aligned column just left of the shifted origin of the group's last plane; row p then holds the wanted
//           value at column x + 3 - (p & 3) + e0.  In natural order (all lanes the same j) the residue p & 3 differs
//           from lane to lane and the reads are 4-way bank conflicted whatever the row pitch; instead lane x walks every
//           group of 4 band columns in the rotated order j = 4g + ((u + x) & 3), u = 0..3, which makes p & 3 uniform
//           across the warp (bank = 5x - 4((u+x)&3) + const with pitch 136: a permutation, conflict-free) and then
//           rotates each group of 4 registers back by x & 3 (two selects per element).
// Warp roles (16 warps, one CTA per SM, persistent):
//   warp 0 lane 0: TMA producer of the feature band = B operand ([Cbox][32] K-major, 128-byte swizzle), deep ring
//   warp 1      MMA issuer: tcgen05.mma kind::tf32, A from TMEM (TS form), M=128, N=Cbox; 3xTF32 = A_hi x [B_hi;B_lo]
//               (N = 2*Cbox, A_hi read once) + A_lo x B_hi per k-step; two TMEM accumulators when they fit
//   warps 2-5   epilogue: tcgen05.ld -> coalesced 128-byte row stores of gin[c][x]
//   warps 6-13  A builders, two groups of 4 warps (one per TMEM lane quarter) taking the K chunks alternately:
//               wait a_empty -> LDS the diagonal -> tcgen05.st hi (raw fp32; kind::tf32 ignores the low 13 mantissa
//               bits) and lo = x - trunc_tf32(x) -> wait::st -> arrive a_built
//   warps 14-15 band split (3xTF32): lo = x - trunc_tf32(x) of each landed band chunk, written next to it
//   warp 0 lanes 16-23: TMA producer of the g slice (mode 0: [32][128] boxes, mode 1: [4][136] shifted boxes); a
//               32-plane box is recycled for the next tile as soon as the last chunk that reads it is built
#include <stdlib.h>

#include "tc_common.cuh"

namespace pmt {
namespace {

constexpr int kTM = 128;                  // output columns per tile (UMMA M)
constexpr int kKC = 32;                   // band columns per K chunk (4 k-steps of 8)
constexpr int kEpiWarps = 4;
constexpr int kGroups = 2;                // builder groups
constexpr int kBuildWarps = 4 * kGroups;
constexpr int kSplitWarps = 2;
constexpr int kWarpBuild0 = 2 + kEpiWarps;             // 6
constexpr int kWarpSplit0 = kWarpBuild0 + kBuildWarps; // 14
constexpr int kThreads = 32 * (kWarpSplit0 + kSplitWarps);   // 512 (16 warps: 128 registers per thread; a 17th warp would cap them at 96)
constexpr unsigned kRawLanes = 0x00ff0000u;              // lanes 16..23 of warp 0 = producer of the g slice
constexpr int kPitch0 = kTM;              // mode 0 slice row pitch (floats)
constexpr int kPitch1 = kTM + 8;          // mode 1: 128 + 3 (plane inside its group of 4) + 3 (alignment) -> 136
constexpr int kBoxBytes0 = 32 * kPitch0 * 4;   // 16 KB per 32 planes
constexpr int kBoxBytes1 = 32 * kPitch1 * 4;   // 17 KB per 32 planes (8 TMA boxes of 4 planes)
constexpr int kMaxBandSlots = 8, kMaxASlots = 4, kMaxGBoxes = 8;
// The slice keeps kRow0 zero rows in front of plane 0 and (at least) one after plane P-1: a builder clamps the plane of
// every element into [-1, P] with ONE unsigned min and reads unconditionally -- branch-free (predicated loads compiled
// into a divergent branch per element and made the builders, not the tensor pipe, the bottleneck).
constexpr int kRow0 = 4;                  // multiple of 4: keeps the [4][136] boxes of mode 1 128-byte aligned
constexpr int kSliceRowsExtra = 2 * kRow0;

struct TcaArgs {
  int C, H, W, P;
  int Cbox;            // channels per block rounded up to 32 (UMMA N, TMEM columns per accumulator half)
  int n_cblk;          // channel blocks (1 unless C > 128)
  int NKC;             // K chunks per tile
  int n_xtiles, n_tiles;   // tiles per gradient = B * H * n_xtiles * n_cblk
  int n_cta0;          // CTAs [0, n_cta0) compute gin1 (mode 0), the rest gin2 (mode 1)
  int n_gboxes;        // 32-plane boxes of the g slice
  int passes;          // 1 = plain TF32, 3 = 3xTF32
  int acc_cols, n_acc; // TMEM columns per accumulator (Cbox or 2*Cbox), number of accumulators (1 or 2)
  int a_base, a_slots, aslot_cols;   // TMEM A ring
  int band_off, band_slots, band_slot_bytes, band_lo_off;
  int bar_off, tmem_cols;
  int oo[2], delta[2];
  int koff0;           // mode 0: box b is last read by chunk min(NKC-1, b + koff0)
  int c1;              // mode 1: p = x + c1 - j
  int e0;              // mode 1: slice column of lane x in plane p is x + 3 - (p & 3) + e0
  int r;               // (P-1)/2
  int only_mode;       // development builds: -1 = both gradients, 0 / 1 = run that gradient only
};

__device__ __forceinline__ float lo_tf32(float x) {
  return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

// Per-role wait / section cycle counters of the first CTA of each mode (development builds with -DPMT_BWD_PROFILE):
// prof[(mode*16 + warp)*8 + slot]; slots 0-3 = cycles blocked on a barrier family, 4 = total, 5-7 = work sections.
#ifdef PMT_BWD_PROFILE
#define PW(slot, stmt)                    \
  do {                                    \
    const long long _t0 = clock64();      \
    stmt;                                 \
    pcyc[slot] += clock64() - _t0;        \
  } while (0)
#else
#define PW(slot, stmt) stmt
#endif

struct TileCoord {
  int x0, h, n, c0;
};
__device__ __forceinline__ TileCoord tile_coord(const TcaArgs& a, int cta_in_mode, int ctas_of_mode, int i) {
  int t = cta_in_mode + i * ctas_of_mode;
  TileCoord c;
  c.c0 = (t % a.n_cblk) * 128;
  t /= a.n_cblk;
  c.x0 = (t % a.n_xtiles) * kTM;
  c.h = (t / a.n_xtiles) % a.H;
  c.n = t / (a.n_xtiles * a.H);
  return c;
}

template <int kPasses>
__global__ void __launch_bounds__(kThreads, 1)
corr1d_bwd_tca_kernel(const __grid_constant__ CUtensorMap tmIn1, const __grid_constant__ CUtensorMap tmIn2,
                      const __grid_constant__ CUtensorMap tmG0, const __grid_constant__ CUtensorMap tmG1,
                      float* __restrict__ gin1, float* __restrict__ gin2, const TcaArgs a, long long* __restrict__ prof) {
  extern __shared__ __align__(1024) unsigned char smem[];
#ifdef PMT_BWD_PROFILE
  long long pcyc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long p_start = clock64();
#endif
  uint64_t* band_full = reinterpret_cast<uint64_t*>(smem + a.bar_off);   // band chunk landed (TMA)
  uint64_t* band_ready = band_full + kMaxBandSlots;                      // its lo copy written (4 split warps)
  uint64_t* band_empty = band_ready + kMaxBandSlots;                     // consumed (MMA commit)
  uint64_t* a_built = band_empty + kMaxBandSlots;                        // A slot written (4 builder warps)
  uint64_t* a_empty = a_built + kMaxASlots;                              // A slot consumed (MMA commit)
  uint64_t* raw_full = a_empty + kMaxASlots;                             // 32-plane box of the g slice landed
  uint64_t* raw_empty = raw_full + kMaxGBoxes;                           // ... no longer read by any builder warp
  uint64_t* tmem_full = raw_empty + kMaxGBoxes;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int mode = (int)blockIdx.x < a.n_cta0 ? 0 : 1;
  const int cta_in_mode = mode == 0 ? (int)blockIdx.x : (int)blockIdx.x - a.n_cta0;
  const int ctas_of_mode = mode == 0 ? a.n_cta0 : (int)gridDim.x - a.n_cta0;
  unsigned char* band_ring = smem + a.band_off;
  const CUtensorMap* tmBand = mode == 0 ? &tmIn2 : &tmIn1;
  float* __restrict__ dst = mode == 0 ? gin1 : gin2;

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int band_bytes = a.Cbox * 128;
  int n_my = (a.n_tiles - cta_in_mode + ctas_of_mode - 1) / ctas_of_mode;  // tiles of this CTA
#ifdef PMT_DEV_KNOBS
  if (a.only_mode >= 0 && a.only_mode != mode) n_my = 0;   // development builds: time one gradient alone
#endif
  constexpr bool three = kPasses == 3;

  if (tid == 0) {
    for (int s = 0; s < kMaxBandSlots; ++s) {
      mbar_init(&band_full[s], 1);
      mbar_init(&band_ready[s], kSplitWarps);
      mbar_init(&band_empty[s], 1);
    }
    for (int s = 0; s < kMaxASlots; ++s) {
      mbar_init(&a_built[s], 4);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kMaxGBoxes; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], kBuildWarps);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (wid == 1) {
    tc::tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tc::tmem_relinquish();
  }
  // zero rows of the g slice (rows before plane 0 and after the last box; TMA never writes them)
  for (int i = tid; i < a.band_off / 16; i += kThreads) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (wid == 0 && lane < 16) {
    // ===== TMA producer 1 (lane 0 of warp 0): the feature band, one swizzled [Cbox][32] box per K chunk; the ring spans
    // tile boundaries.  The two producers share a warp as divergent lane groups (independent thread scheduling keeps
    // both spin loops progressing; both sleep between polls), which keeps the CTA at 16 warps. =====
    if (lane == 0) {
      tma_prefetch_desc(tmBand);
      const int oo = mode == 0 ? a.oo[0] : a.oo[1];
      int bs = 0;
      uint32_t bph = 1;  // parity to wait for on band_empty (the first pass over the ring is free)
      for (int i = 0; i < n_my; ++i) {
        const TileCoord tc_ = tile_coord(a, cta_in_mode, ctas_of_mode, i);
        for (int k = 0; k < a.NKC; ++k) {
          PW(0, mbar_wait_relaxed(&band_empty[bs], bph));
          mbar_arrive_expect_tx(&band_full[bs], (uint32_t)band_bytes);
          tma_load_4d(band_ring + (size_t)bs * a.band_slot_bytes, tmBand, tc_.x0 + oo + kKC * k, tc_.h, tc_.c0, tc_.n,
                      &band_full[bs]);
          if (++bs == a.band_slots) bs = 0, bph ^= 1u;
        }
      }
    }
  } else if (wid == 0) {
    // ===== TMA producer 2 (lanes 16-23 of warp 0): the g slice of the tile, one 32-plane box per barrier, in the order
    // the chunks need them (mode 0: ascending planes; mode 1: descending).  A box is reloaded for the next tile as soon
    // as every builder warp has released it, so the loads run about one tile ahead of the MMAs. =====
    if (lane < 24) {
      const int rl = lane - 16;   // 0..7
      if (rl == 0) tma_prefetch_desc(mode == 0 ? &tmG0 : &tmG1);
      for (int i = 0; i < n_my; ++i) {
        const TileCoord tc_ = tile_coord(a, cta_in_mode, ctas_of_mode, i);
        const bool more = i + 1 < n_my;
        const TileCoord nx = tile_coord(a, cta_in_mode, ctas_of_mode, more ? i + 1 : i);
        for (int o = 0; o < a.n_gboxes; ++o) {
          const int b = mode == 0 ? o : a.n_gboxes - 1 - o;
          if (mode == 0) {
            if (rl == 0) {
              // the smem slice only reaches about one tile ahead: pull the next tile's box into L2 now
              if (more) tma_prefetch_l2_4d(&tmG0, nx.x0, nx.h, 32 * b, nx.n);
              PW(1, mbar_wait_relaxed(&raw_empty[b], ((uint32_t)i & 1u) ^ 1u));   // the previous tile is done with this box
              mbar_arrive_expect_tx(&raw_full[b], (uint32_t)kBoxBytes0);
              tma_load_4d(smem + kRow0 * kPitch0 * 4 + b * kBoxBytes0, &tmG0, tc_.x0, tc_.h, 32 * b, tc_.n, &raw_full[b]);
            }
          } else {
            PW(1, mbar_wait_relaxed(&raw_empty[b], ((uint32_t)i & 1u) ^ 1u));
            if (rl == 0) mbar_arrive_expect_tx(&raw_full[b], (uint32_t)kBoxBytes1);
            __syncwarp(kRawLanes);
            // planes pg..pg+3, columns from the aligned origin of plane pg+3: a = x0 - (pg+3) + r - e0 (multiple of 4)
            const int pg = 32 * b + 4 * rl;
            tma_load_4d(smem + (size_t)(pg + kRow0) * kPitch1 * 4, &tmG1, tc_.x0 - (pg + 3) + a.r - a.e0, tc_.h, pg, tc_.n, &raw_full[b]);
            if (more) tma_prefetch_l2_4d(&tmG1, nx.x0 - (pg + 3) + a.r - a.e0, nx.h, pg, nx.n);
          }
        }
      }
    }
  } else if (wid == 1) {
    // ===== MMA issuer =====
    // The whole warp runs this loop converged (every lane polls the barriers) and ONE elected lane issues: the
    // compiler keeps descriptors, slots and phases in uniform registers and emits straight-line UTCHMMA.  With the loop
    // under `if (lane == 0)` every MMA cost ~10 instructions (R2UR moves + an ELECT / BRA.U.ANY serialisation loop) and
    // this single thread, not the tensor pipe, bounded the kernel (measured: it waited < 12 % of its time).
    {
      const uint32_t idesc = tc::make_idesc(2, 0, 0, kTM, a.Cbox);
      const uint32_t idesc2 = tc::make_idesc(2, 0, 0, kTM, 2 * a.Cbox);  // B = [band_hi ; band_lo] stacked along N
      const uint64_t dB0 = tc::smem_desc(smem_u32(band_ring), 16, 1024, 2);
      const uint32_t b_step = (uint32_t)a.band_slot_bytes >> 4;
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int buf = 0;
      uint32_t eph = 1;   // parity to wait for on tmem_empty[buf]
      for (int i = 0; i < n_my; ++i) {
        PW(0, mbar_wait(&tmem_empty[buf], eph));  // the epilogue has drained this accumulator
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * a.acc_cols);
        for (int k = 0; k < a.NKC; ++k) {
          PW(1, mbar_wait(&a_built[as], aph));
          if (three) PW(2, mbar_wait(&band_ready[bs], bph));
          else PW(2, mbar_wait(&band_full[bs], bph));
          tc::fence_after_sync();
          const uint64_t dB = dB0 + (uint64_t)(b_step * (uint32_t)bs);
          const uint32_t ta = tmem_base + (uint32_t)(a.a_base + as * a.aslot_cols);
          if (tc::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kKC / 8; ++kk) {
              const uint32_t acc = (k > 0 || kk > 0) ? 1u : 0u;
              if (three) {
                // D[:, 0:C] += A_hi*B_hi + A_lo*B_hi ; D[:, C:2C] += A_hi*B_lo (the epilogue adds the two column blocks)
                tc::mma_tf32_ts(d_tmem, ta + 8 * kk, dB + 2 * kk, idesc2, acc);
                tc::mma_tf32_ts(d_tmem, ta + 32 + 8 * kk, dB + 2 * kk, idesc, 1u);
              } else {
                tc::mma_tf32_ts(d_tmem, ta + 8 * kk, dB + 2 * kk, idesc, acc);
              }
            }
            tc::mma_commit(&a_empty[as]);
            tc::mma_commit(&band_empty[bs]);
            if (k == a.NKC - 1) tc::mma_commit(&tmem_full[buf]);
          }
          __syncwarp();
          if (++as == a.a_slots) as = 0, aph ^= 1u;
          if (++bs == a.band_slots) bs = 0, bph ^= 1u;
        }
        if (++buf == a.n_acc) buf = 0, eph ^= 1u;
      }
    }
  } else if (wid < kWarpBuild0) {
    // ===== epilogue: TMEM -> coalesced global stores =====
    const int q = wid & 3;
    const int xl = 32 * q + lane;
    const int64_t pstride = (int64_t)a.H * a.W;
    int buf = 0;
    uint32_t fph = 0;
    for (int i = 0; i < n_my; ++i) {
      const TileCoord tc_ = tile_coord(a, cta_in_mode, ctas_of_mode, i);
      PW(0, mbar_wait_relaxed(&tmem_full[buf], fph));
      tc::fence_after_sync();
      const bool ok = tc_.x0 + xl < a.W;
      const int cvalid = a.C - tc_.c0;   // channels of this block that exist
      float* o = dst + (((int64_t)tc_.n * a.C + tc_.c0) * a.H + tc_.h) * (int64_t)a.W + tc_.x0 + xl;
      const uint32_t t0 = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * a.acc_cols);
      if (three) {
        for (int cb = 0; cb < a.Cbox; cb += 16) {
          float v[16], v2[16];
          tc::tmem_ld16(t0 + (uint32_t)cb, v);
          tc::tmem_ld16(t0 + (uint32_t)(a.Cbox + cb), v2);
#pragma unroll
          for (int cc = 0; cc < 16; ++cc)
            if (ok && cb + cc < cvalid) o[(int64_t)(cb + cc) * pstride] = v[cc] + v2[cc];
        }
      } else {
        for (int cb = 0; cb < a.Cbox; cb += 32) {
          float v[32];
          tc::tmem_ld32(t0 + (uint32_t)cb, v);
#pragma unroll
          for (int cc = 0; cc < 32; ++cc)
            if (ok && cb + cc < cvalid) o[(int64_t)(cb + cc) * pstride] = v[cc];
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      if (++buf == a.n_acc) buf = 0, fph ^= 1u;
    }
  } else if (wid < kWarpSplit0) {
    // ===== A builders =====
    const int bw = wid - kWarpBuild0;     // 0..7
    const int grp = bw >> 2;              // this warp handles chunks g = grp (mod kGroups)
    const int q = wid & 3;                // TMEM lane quarter
    const int xl = 32 * q + lane;         // Gd row = TMEM lane
    const unsigned Pu = (unsigned)a.P;
    // mode 1 per-lane constants: step u of a group of 4 band columns reads column j = 4g + ((u + xl) & 3)
    int bu[4], cu[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      bu[u] = xl - ((u + xl) & 3);
      cu[u] = xl + 3 - ((a.c1 - u) & 3) + a.e0;
    }
    const bool rot1 = (xl & 1) != 0, rot2 = (xl & 2) != 0;
    int as = grp % a.a_slots;
    uint32_t aeph = (((uint32_t)(grp / a.a_slots)) & 1u) ^ 1u;   // parity to wait for on a_empty[as]
    int g = 0;
    for (int i = 0; i < n_my; ++i) {
      int boxes_ready = 0;   // boxes (in load order) known to have landed for this tile
      int next_rel = 0;      // boxes (in load order) already handed back to the producer
      for (int k = 0; k < a.NKC; ++k, ++g) {
        if (g % kGroups != grp) continue;
        // --- the boxes this chunk reads (load order o: mode 0 box o, mode 1 box n_gboxes-1-o) ---
        int need;
        if (mode == 0) {
          need = k + 1 < a.n_gboxes ? k + 1 : a.n_gboxes;                   // planes <= 32k+31-delta0
        } else {
          int lowp = a.c1 - kKC * k - (kKC - 1);                             // lowest plane read (row x = 0)
          int lowb = lowp > 0 ? lowp >> 5 : 0;
          if (lowb > a.n_gboxes - 1) lowb = a.n_gboxes - 1;
          need = a.n_gboxes - lowb;
        }
        while (boxes_ready < need) {
          const int b = mode == 0 ? boxes_ready : a.n_gboxes - 1 - boxes_ready;
          PW(0, mbar_wait(&raw_full[b], (uint32_t)i & 1u));
          ++boxes_ready;
        }
        // Row of plane p in the slice = umin(p + 1, P + 1) + kRow0 - 1: planes outside [0, P) land on a zero row.
        float w[32];
        const unsigned Pp1 = Pu + 1u;
        if (mode == 0) {
          const float* col = reinterpret_cast<const float*>(smem) + (kRow0 - 1) * kPitch0 + xl;
          const unsigned pu1 = (unsigned)(kKC * k - a.delta[0] - xl + 1);      // plane + 1 of band column jj is pu1 + jj
#pragma unroll
          for (int t = 0; t < 32; ++t) w[t] = col[min(pu1 + (unsigned)t, Pp1) * kPitch0];
        } else {
          const float* S = reinterpret_cast<const float*>(smem) + (kRow0 - 1) * kPitch1;
          const int pk1 = a.c1 - kKC * k + 1;
#pragma unroll
          for (int gg = 0; gg < 8; ++gg) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = S[min((unsigned)(bu[u] + pk1 - 4 * gg), Pp1) * kPitch1 + cu[u]];
            // v[u] holds band column 4gg + ((u + xl) & 3): rotate right by xl & 3 so that register m holds column 4gg + m
            // (selects, not branches: the amount differs from lane to lane)
            const float t0 = rot1 ? v[3] : v[0], t1 = rot1 ? v[0] : v[1], t2 = rot1 ? v[1] : v[2], t3 = rot1 ? v[2] : v[3];
            w[4 * gg + 0] = rot2 ? t2 : t0;
            w[4 * gg + 1] = rot2 ? t3 : t1;
            w[4 * gg + 2] = rot2 ? t0 : t2;
            w[4 * gg + 3] = rot2 ? t1 : t3;
          }
        }
        // --- hand back every box this warp will not read again (warp-uniform loop, lane 0 arrives) ---
        __syncwarp();
        {
          const bool last_visit = k + kGroups >= a.NKC;
          while (next_rel < a.n_gboxes) {
            int kl;   // last chunk that reads the box at load position next_rel
            if (mode == 0) kl = next_rel + a.koff0;
            else kl = (a.c1 + (kTM - 1) - 32 * (a.n_gboxes - 1 - next_rel)) >> 5;
            if (kl > a.NKC - 1) kl = a.NKC - 1;
            if (!(kl <= k + kGroups - 1 || last_visit)) break;
            const int b = mode == 0 ? next_rel : a.n_gboxes - 1 - next_rel;
            // Never release a box this warp has not seen land: the wait orders this arrival after the producer's
            // reload, i.e. after EVERY warp's release of the previous tile -- otherwise a warp running a tile ahead
            // (short K loops) could complete the previous phase with its own second arrival.
            if (next_rel >= boxes_ready) {
              mbar_wait(&raw_full[b], (uint32_t)i & 1u);
              ++boxes_ready;
            }
            if (lane == 0) mbar_arrive(&raw_empty[b]);
            ++next_rel;
          }
        }
        // --- into TMEM: thread = Gd row (TMEM lane), 32 consecutive band columns; hi = raw fp32, lo = x - trunc_tf32(x) ---
        PW(1, mbar_wait(&a_empty[as], aeph));
        tc::fence_after_sync();
        const uint32_t ta = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(a.a_base + as * a.aslot_cols);
        tc::tmem_st16(ta, *reinterpret_cast<const float(*)[16]>(&w[0]));
        tc::tmem_st16(ta + 16, *reinterpret_cast<const float(*)[16]>(&w[16]));
        if (three) {
#pragma unroll
          for (int t = 0; t < 32; ++t) w[t] = lo_tf32(w[t]);
          tc::tmem_st16(ta + 32, *reinterpret_cast<const float(*)[16]>(&w[0]));
          tc::tmem_st16(ta + 48, *reinterpret_cast<const float(*)[16]>(&w[16]));
        }
        PW(5, tc::tmem_st_wait());
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_built[as]);
        as += kGroups;
        while (as >= a.a_slots) as -= a.a_slots, aeph ^= 1u;
      }
    }
  } else {
    // ===== band split (3xTF32): lo = x - trunc_tf32(x) of the landed band chunk (position-wise, layout agnostic) =====
    if (three) {
      constexpr int kSplitThreads = 32 * kSplitWarps;
      const int t = tid - 32 * kWarpSplit0;   // 0..kSplitThreads-1
      const int nch = band_bytes / 16;
      int bs = 0;
      uint32_t bph = 0;
      for (int i = 0; i < n_my; ++i) {
        for (int k = 0; k < a.NKC; ++k) {
          PW(0, mbar_wait(&band_full[bs], bph));
          unsigned char* sb = band_ring + (size_t)bs * a.band_slot_bytes;
          for (int cb = t; cb < nch; cb += 4 * kSplitThreads) {   // 4 loads in flight (C = 64: two rounds)
            float4 x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (cb + u * kSplitThreads < nch) x[u] = *reinterpret_cast<const float4*>(sb + 16 * (cb + u * kSplitThreads));
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (cb + u * kSplitThreads < nch)
                *reinterpret_cast<float4*>(sb + a.band_lo_off + 16 * (cb + u * kSplitThreads)) =
                    make_float4(lo_tf32(x[u].x), lo_tf32(x[u].y), lo_tf32(x[u].z), lo_tf32(x[u].w));
          }
          fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&band_ready[bs]);
          if (++bs == a.band_slots) bs = 0, bph ^= 1u;
        }
      }
    }
  }

#ifdef PMT_BWD_PROFILE
  if (prof != nullptr && cta_in_mode == 0 && (lane == 0 || (wid == 0 && lane == 16))) {
    long long* o = prof + ((size_t)mode * 17 + (wid == 0 && lane == 16 ? 16 : wid)) * 8;
    pcyc[4] = clock64() - p_start;
    for (int s = 0; s < 8; ++s) o[s] = pcyc[s];
  }
#endif
  tc::fence_before_sync();
  __syncthreads();
  if (wid == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

int fill_args(TcaArgs* a, int C, int H, int W, int P, int passes) {
  a->C = C, a->H = H, a->W = W, a->P = P, a->r = (P - 1) / 2, a->passes = passes;
  a->n_cblk = ceil_div(C, 128);
  a->Cbox = a->n_cblk > 1 ? 128 : round_up(C, 32);   // the epilogue reads TMEM in 16/32-column groups; UMMA N % 16 == 0
  const int oo0 = -a->r, oo1 = -(P - 1 - a->r);
  a->delta[0] = ((oo0 % 4) + 4) % 4;
  a->oo[0] = oo0 - a->delta[0];
  a->delta[1] = ((oo1 % 4) + 4) % 4;
  a->oo[1] = oo1 - a->delta[1];
  a->koff0 = (kTM + kKC - 2 + a->delta[0]) / kKC;
  a->c1 = P - 1 + a->delta[1];
  a->e0 = (((a->r - 3) % 4) + 4) % 4;
  const int dmax = a->delta[0] > a->delta[1] ? a->delta[0] : a->delta[1];
  a->NKC = ceil_div(kTM + P - 1 + dmax, kKC);
  a->n_xtiles = ceil_div(W, kTM);
  a->n_tiles = 0;  // set by the launcher (needs B)
  a->n_gboxes = ceil_div(P, 32);
  if (a->n_gboxes > kMaxGBoxes) return 1;
  const int mult = passes == 3 ? 2 : 1;
  // TMEM: accumulators first, then the A ring
  a->acc_cols = a->Cbox * mult;
  a->aslot_cols = 32 * mult;
  a->n_acc = (2 * a->acc_cols + kGroups * a->aslot_cols <= 512) ? 2 : 1;
  a->a_base = a->n_acc * a->acc_cols;
  a->a_slots = (512 - a->a_base) / a->aslot_cols;
  if (a->a_slots > kMaxASlots) a->a_slots = kMaxASlots;
  if (a->a_slots < kGroups) return 1;
  int cols = 32;
  while (cols < a->a_base + a->a_slots * a->aslot_cols) cols *= 2;
  a->tmem_cols = cols;
  // shared memory: the g slice (the larger, mode 1, layout decides), then the band ring
  a->band_off = round_up((a->n_gboxes * 32 + kSliceRowsExtra) * kPitch1 * 4, 1024);
  a->band_lo_off = a->Cbox * 128;
  a->band_slot_bytes = a->Cbox * 128 * mult;
  const int budget = 227 * 1024 - 1024;
  int bslots = (budget - a->band_off) / a->band_slot_bytes;
  if (bslots > kMaxBandSlots) bslots = kMaxBandSlots;
  if (bslots < 2) return 1;
  a->band_slots = bslots;
  a->bar_off = a->band_off + bslots * a->band_slot_bytes;
  return 0;
}

}  // namespace

extern long long* g_bwd_prof;  // corr1d_bwd_tc.cu (development builds: device buffer set through pmt_debug_set_ptr)

bool corr1d_bwd_tca_ok(const void* in1, const void* in2, int C, int H, int W, int P, int dilp, int passes) {
  if (dilp != 1 || P < 1 || C < 1 || W % 4 != 0 || !aligned16(in1) || !aligned16(in2)) return false;
  TcaArgs a;
  return fill_args(&a, C, H, W, P, passes) == 0;
}

int launch_corr1d_bwd_tca(const float* in1, const float* in2, const float* gout, float* gin1, float* gin2, int B,
                          int C, int H, int W, int P, int passes, cudaStream_t st) {
  TcaArgs a;
  PMT_CHECK_ARG(passes == 1 || passes == 3, "corr1d tc: passes must be 1 (tf32) or 3 (3xtf32)");
  PMT_CHECK_ARG(fill_args(&a, C, H, W, P, passes) == 0, "corr1d tc bwd: unsupported shape C=%d P=%d", C, P);
  CUtensorMap tm1, tm2, tmG0, tmG1;
  if (int e = make_tmap_nchw_ex(&tm1, in1, B, C, H, W, kKC, a.Cbox, 1)) return e;
  if (int e = make_tmap_nchw_ex(&tm2, in2, B, C, H, W, kKC, a.Cbox, 1)) return e;
  if (int e = make_tmap_nchw_ex(&tmG0, gout, B, P, H, W, kPitch0, 32, 0)) return e;
  if (int e = make_tmap_nchw_ex(&tmG1, gout, B, P, H, W, kPitch1, 4, 0)) return e;
  const int smem_bytes = a.bar_off + 1024;
  const int64_t tiles = (int64_t)B * H * a.n_xtiles * a.n_cblk;
  PMT_CHECK_ARG(tiles < (1ll << 31), "corr1d tc bwd: too many tiles");
  a.n_tiles = (int)tiles;
  // persistent: one CTA per SM, half of them per gradient (both modes now cost the same per tile)
  const int sms = sm_count();
  int64_t n_cta = 2 * tiles < sms ? 2 * tiles : sms;
  if (n_cta < 2) n_cta = 2;
  int n0 = (int)(n_cta / 2);
  if (const int e = PMT_ENV_INT("PMT_BWD_SPLIT", 0)) n0 = e;
  if (n0 < 1) n0 = 1;
  if (n0 > n_cta - 1) n0 = (int)n_cta - 1;
  a.n_cta0 = n0;
  a.only_mode = PMT_ENV_INT("PMT_TCA_ONLY", -1);
  if (passes == 3) {
    PMT_CUDA_OK(cudaFuncSetAttribute(corr1d_bwd_tca_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    corr1d_bwd_tca_kernel<3><<<dim3((unsigned)n_cta), kThreads, smem_bytes, st>>>(tm1, tm2, tmG0, tmG1, gin1, gin2, a, g_bwd_prof);
  } else {
    PMT_CUDA_OK(cudaFuncSetAttribute(corr1d_bwd_tca_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    corr1d_bwd_tca_kernel<1><<<dim3((unsigned)n_cta), kThreads, smem_bytes, st>>>(tm1, tm2, tmG0, tmG1, gin1, gin2, a, g_bwd_prof);
  }
  PMT_LAUNCH_OK("corr1d_bwd_tca_kernel");
  return PMT_OK;
}

}  // namespace pmt
