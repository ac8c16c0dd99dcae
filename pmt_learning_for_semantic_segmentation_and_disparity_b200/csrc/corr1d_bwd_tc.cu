// corr1d_bwd_tc.cu -- tensor-core (tcgen05 + TMEM) backward of the 1 x P horizontal correlation.
// Both gradients are deterministic gathers (no atomics), one launch; CTAs [0, n_cta0) compute gin1, the rest gin2:
//   mode 0: gin1[c,x] = sum_j Gd[x][j] * in2[c][x0+oo+j]        mode 1: gin2[c,x] = sum_j Gd[x][j] * in1[c][x0+oo+j]
// where Gd[x][j] is the band matrix made of g (mode 0: g[j-delta-x][x]; mode 1: g[x+P-1+delta-j][x0+oo+j]).
// A tile = 128 output columns x of one image row, all C channels (C <= 128):
//   D[x (M=128 TMEM lanes)][c (N=C columns)] = sum over the band (K = 32*NKC columns, 320 for P=192).
// Persistent CTAs (one per SM, half of them per gradient) walk tiles with every stage of the pipeline
// running ahead across tile boundaries:
//   * warp 0 and the warp after the builders (TMA producers): the feature band = B operand (rows c, K contiguous along
//     w -> K-major), one 128-byte-swizzled [C][32] box per K chunk into a deep ring; raw g into a second ring -- mode 0:
//     the tile's [P][128] slice as 32-row boxes (a box is recycled for the next tile as soon as its last chunk is
//     built), mode 1: one 128-byte-swizzled [160][32] block per chunk (OOB rows/columns zero-filled by TMA), 6 deep;
//   * builder warps (groups of 4 warps, one per TMEM lane quarter, that take the K chunks round-robin; 3 groups when
//     the rings fit, else 2): both modes write the A operand Gd straight into TMEM when the columns are there (C <= 64
//     for 3xTF32; tcgen05.st: thread = Gd row, registers = the 32 band columns of the chunk, hi = raw fp32 and
//     lo = x - trunc_tf32(x)) with conflict-free reads -- mode 0: column reads of the resident g slice; mode 1: the
//     diagonal of the swizzled raw block in a rotated column order, rotated back in registers.  Without TMEM room
//     (C = 128 with 3xTF32) they re-lay raw g out shared->shared into the swizzled K-major A operand instead;
//   * two teams of band-split warps (3xTF32) write lo = x - trunc_tf32(x) of every landed band chunk next to it;
//   * warp 1 issues tcgen05.mma kind::tf32 (M=128, N=C, K=8; A_hi x [B_hi;B_lo] + A_lo x B_hi for 3xTF32) into one
//     of two TMEM accumulators and commits to the mbarriers that recycle the rings;
//   * warps 2-5 (epilogue) drain the other accumulator with tcgen05.ld and store gin[c][x] directly:
//     a warp writes 32 consecutive columns of one channel per instruction (coalesced 128 bytes).
// What bounds it (cycle-stamped per-warp profile, -DPMT_BWD_PROFILE): not DRAM, not the tensor pipe (45 % active), but
// the latency chain of a builder group's chunk (barrier waits of ~200 cycles each, LDS -> tcgen05.st -> wait::st, fence,
// arrive): the rate is `groups` chunks per chain, so everything that is not the build itself was moved off the
// builders (band split -> own warps, box releases -> after the A operand is announced).
#include <stdlib.h>

#include "tc_common.cuh"

namespace pmt {
namespace {

constexpr int kTM = 128;                  // output columns per tile (UMMA M)
constexpr int kKC = 32;                   // band columns per K chunk (4 k-steps of 8)
constexpr int kGdBytes = kTM * kKC * 4;   // 16 KB: one A-operand chunk
constexpr int kEpiWarps = 4;
// Builder warps.  Every builder warp pays a fixed cost per chunk (barrier waits, tcgen05.st / fences, arrive, ring
// bookkeeping) on top of the elements it moves, so few, fat warps win over many thin ones (16 -> 8 warps: 366 -> 341 us
// at the headline shape).  The builders work in groups of 4 warps (one per TMEM lane quarter) that take chunks
// round-robin, so a group's latency chain (waits, build, fence, arrive) overlaps the other groups'.  Every ring a group
// waits on needs one slot per group (parity waits tolerate one phase of lead per waiter).  3 groups are used when the
// rings fit (both A operands in TMEM, or three shared-memory A slots for gin2 at C <= 64), else 2.
template <int kPasses, int kGroups_>
struct BwdCfg {
  static constexpr int kGroups = kGroups_;
  static constexpr int kGroupWarps = 4;
  static constexpr int kBuilders = kGroupWarps * kGroups;
  // 3xTF32: dedicated warps split the landed feature band into hi / lo.  The builders used to do it at the end of every
  // chunk; a cycle-stamped profile of one CTA showed a builder group spending ~900 of its ~3800 cycles per chunk there (a
  // barrier wait, 8 KB through LDS/STS, a proxy fence), and the rate of the kernel is kGroups chunks per such group cycle.
  // Two teams take the chunks alternately (a ring a team waits on needs one slot per team: band_slots >= 2, and a team must
  // never start on the ring's second phase -- a parity test on a fresh barrier passes at once); a team is 2 warps that
  // share a chunk (4 with 2 builder groups: wide bands, C = 128 is 16 KB per chunk).  23 warps x 80 registers fit.
  static constexpr int kSplitTeams = 2;
  static constexpr int kSplitPer = kGroups_ == 2 ? 4 : 2;
  static constexpr int kSplitWarps = kPasses == 3 ? kSplitTeams * kSplitPer : 0;
  static constexpr int kThreads = 32 * (3 + kEpiWarps + kBuilders + kSplitWarps);  // band producer, MMA, 4 epilogue, builders, raw-g producer, band split
};
constexpr int kBox0Bytes = 32 * kTM * 4;          // mode 0: 32 rows of g x 128 columns (16 KB)
constexpr int kRawRows1 = 160;                    // mode 1: rows of a raw block (>= 128+32-1)
constexpr int kRawSlot1 = kRawRows1 * kKC * 4;    // 20 KB
constexpr int kMaxRawSlots1 = 4;
constexpr int kMaxBandSlots = 8, kMaxGdSlots = 4;

struct TcBwdMode {
  int oo;      // band column j <-> image column x0 + oo + j  (multiple of 4)
  int delta;
  int koff;    // mode 0: box b of the g slice is last used by chunk min(NKC-1, b + koff)
  int a_slots;     // A-operand ring depth (smem Gd slots, or TMEM slots when tmem_a)
  int band_slots;  // band ring depth
  int band_off;    // byte offset of the band ring
  int tmem_a;      // 1: Gd chunks are written straight into TMEM (tcgen05.st) and the MMA takes A from TMEM
  int gd_off;      // byte offset of the shared-memory Gd ring (mode 1)
  int raw_slots;   // mode 1: raw-g ring depth
};

struct TcBwdArgs {
  int C, H, W, P, rW;
  int Cbox;            // channels rounded up to 32 (UMMA N, TMEM columns per accumulator)
  int NKC;             // K chunks per tile
  int n_xtiles, n_tiles;
  int n_cta0;          // CTAs [0, n_cta0) compute gin1 (mode 0), the rest gin2 (mode 1)
  int n_gboxes;        // mode 0: 32-row boxes of the resident g slice (= raw ring slots)
  int a_base, aslot_cols;  // TMEM A ring: first column, columns per slot (32 hi [+ 32 lo])
  int gd_slot_bytes, gd_lo_off;       // Gd ring slot: hi [16 KB] (+ lo [16 KB])
  int band_slot_bytes, band_lo_off;   // band ring slot: hi [Cbox*128] (+ lo)
  int bar_off;                        // byte offset of the mbarriers in dynamic shared memory (raw ring at 0)
  int tmem_cols;
  int acc_cols;        // TMEM columns per accumulator (Cbox, or 2*Cbox for 3xTF32: hi and lo column blocks)
  TcBwdMode m[2];
  int relax_ns;        // sleep between polls of the non-critical waiters (producers, epilogue); 0 = spin
  int debug;           // PMT_TC_DEBUG ablation bits: 4 skip Gd build, 32 skip band split, 16 skip MMA, 8 skip epilogue, 2048 / 4096 skip gin2 / gin1
};

// K-major, 128-byte swizzle: element (row, col) of a [rows][32 floats] chunk
__device__ __forceinline__ uint32_t kmajor_off(int row, int col) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 2) ^ (row & 7)) << 4) | ((col & 3) << 2)));
}

// 3xTF32 split.  kind::tf32 reads only the top 19 bits of an fp32 operand (verified on B200), so the raw value
// can serve as the "hi" operand and lo = x - trunc_tf32(x) (exact in fp32) carries the remaining 13 bits.
__device__ __forceinline__ float lo_tf32(float x) {
  return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

// Optional per-role wait profiling (PMT_TC_DEBUG bit 1024): cycles spent blocked on each barrier family, per warp.
#ifdef PMT_BWD_PROFILE
#define PWAIT(slot, bar, par)                      \
  do {                                             \
    const long long _t0 = clock64();               \
    mbar_wait(bar, par);                           \
    wait_cyc[slot] += clock64() - _t0;             \
  } while (0)
#define PWAIT_RELAXED(slot, bar, par) PWAIT(slot, bar, par)
#define PSEC_BEGIN() const long long _s0 = clock64()
#define PSEC_END(slot) sec_cyc[slot] += clock64() - _s0
// timeline trace of chunks [kTraceG0, kTraceG0+16) of the first CTA of each mode: prof[1024 + mode*1024 + role*128 + (g-G0)*8 + idx]
#define kTraceG0 200
#define PTRACE(role, g, idx)                                                                          \
  do {                                                                                                \
    if (prof != nullptr && cta_in_mode == 0 && lane == 0 && (g) >= kTraceG0 && (g) < kTraceG0 + 16)   \
      prof[1024 + mode * 1024 + (role) * 128 + ((g) - kTraceG0) * 8 + (idx)] = clock64();             \
  } while (0)
#else
#define PTRACE(role, g, idx)
#define PWAIT(slot, bar, par) mbar_wait(bar, par)
#define PWAIT_RELAXED(slot, bar, par)                    \
  do {                                                   \
    if (relax_ns) mbar_wait_relaxed(bar, par, relax_ns); \
    else mbar_wait(bar, par);                            \
  } while (0)
#define PSEC_BEGIN()
#define PSEC_END(slot)
#endif

struct TileCoord {
  int x0, h, n;
};
__device__ __forceinline__ TileCoord tile_coord(const TcBwdArgs& a, int i) {
  // CTAs of one mode walk that mode's tiles with a stride equal to the number of CTAs of the mode
  const bool m0 = (int)blockIdx.x < a.n_cta0;
  const int t = (m0 ? (int)blockIdx.x : (int)blockIdx.x - a.n_cta0) + i * (m0 ? a.n_cta0 : (int)gridDim.x - a.n_cta0);
  TileCoord c;
  c.x0 = (t % a.n_xtiles) * kTM;
  c.h = (t / a.n_xtiles) % a.H;
  c.n = t / (a.n_xtiles * a.H);
  return c;
}

template <int kPasses, int kGroups>
__global__ void __launch_bounds__(BwdCfg<kPasses, kGroups>::kThreads, 1)
corr1d_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmIn1, const __grid_constant__ CUtensorMap tmIn2,
                     const __grid_constant__ CUtensorMap tmG0, const __grid_constant__ CUtensorMap tmG1,
                     float* __restrict__ gin1, float* __restrict__ gin2, const TcBwdArgs a,
                     long long* __restrict__ prof) {
  constexpr int kBuilders = BwdCfg<kPasses, kGroups>::kBuilders;
  constexpr int kGroupWarps = BwdCfg<kPasses, kGroups>::kGroupWarps;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* band_full = reinterpret_cast<uint64_t*>(smem + a.bar_off);
  uint64_t* band_empty = band_full + kMaxBandSlots;
  uint64_t* gd_built = band_empty + kMaxBandSlots;
  uint64_t* gd_empty = gd_built + kMaxGdSlots;
  uint64_t* raw_full = gd_empty + kMaxGdSlots;
  uint64_t* raw_empty = raw_full + 8;
  uint64_t* tmem_full = raw_empty + 8;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* band_split = tmem_empty + 2;   // 3xTF32: lo half of a band slot written (one split warp per chunk)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(band_split + kMaxBandSlots);


  const int mode = (int)blockIdx.x < a.n_cta0 ? 0 : 1;
  const int cta_in_mode = mode == 0 ? (int)blockIdx.x : (int)blockIdx.x - a.n_cta0;
  const int ctas_of_mode = mode == 0 ? a.n_cta0 : (int)gridDim.x - a.n_cta0;
  const TcBwdMode m = a.m[mode];
  unsigned char* band_ring = smem + m.band_off;
  unsigned char* gd_ring = smem + m.gd_off;
  const CUtensorMap* tmBand = mode == 0 ? &tmIn2 : &tmIn1;
  float* __restrict__ dst = mode == 0 ? gin1 : gin2;

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const unsigned relax_ns = (unsigned)a.relax_ns;
#ifdef PMT_BWD_PROFILE
  long long wait_cyc[4] = {0, 0, 0, 0};
  long long sec_cyc[4] = {0, 0, 0, 0};
  const long long t_start = clock64();
#endif
  const int band_bytes = a.Cbox * 128;
  int n_my = (a.n_tiles - cta_in_mode + ctas_of_mode - 1) / ctas_of_mode;  // tiles of this CTA
  if (PMT_DBG(a, (mode == 0 ? 4096 : 2048))) n_my = 0;   // ablation: run one gradient only

  if (tid == 0) {
    for (int s = 0; s < kMaxBandSlots; ++s) {
      mbar_init(&band_full[s], 1);
      mbar_init(&band_empty[s], 1);
      mbar_init(&band_split[s], BwdCfg<kPasses, kGroups>::kSplitPer);
    }
    for (int s = 0; s < kMaxGdSlots; ++s) {
      mbar_init(&gd_built[s], kGroupWarps);
      mbar_init(&gd_empty[s], 1);
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], mode == 0 ? kBuilders : kGroupWarps);
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
  if (mode == 0) {
    // The resident g slice sits one row down: row 0 and the rows after plane P-1 are zero and never written by the TMA,
    // so a builder clamps the plane of an element into [-1, P] with one unsigned min and loads unconditionally
    // (the predicated loads compiled into a divergent branch per element).
    float* Z = reinterpret_cast<float*>(smem);
    for (int i = tid; i < kTM; i += blockDim.x) Z[i] = 0.f;
    for (int i = (a.P + 1) * kTM + tid; i < (a.n_gboxes * 32 + 2) * kTM; i += blockDim.x) Z[i] = 0.f;
    fence_proxy_async();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (wid == 0) {
    // ===== TMA producer 1: the feature band, one swizzled [C][32] box per K chunk, ring spans tile boundaries =====
    if (lane == 0) {
      tma_prefetch_desc(tmBand);
      int bs = 0;
      uint32_t bph = 1;  // parity to wait for on band_empty (first pass over the ring is free)
      for (int i = 0; i < n_my; ++i) {
        const TileCoord tc_ = tile_coord(a, i);
        for (int k = 0; k < a.NKC; ++k) {
          PWAIT_RELAXED(0, &band_empty[bs], bph);
          PTRACE(0, i * a.NKC + k, 0);
          if PMT_DBG(a, 128) {   // ablation: no band traffic
            mbar_arrive(&band_full[bs]);
          } else {
            mbar_arrive_expect_tx(&band_full[bs], (uint32_t)band_bytes);
            tma_load_4d(band_ring + (size_t)bs * a.band_slot_bytes, tmBand, tc_.x0 + m.oo + kKC * k, tc_.h, 0, tc_.n,
                        &band_full[bs]);
          }
          if (++bs == m.band_slots) bs = 0, bph ^= 1u;
        }
      }
    }
  } else if (wid == 2 + kEpiWarps + kBuilders) {
    // ===== TMA producer 2: raw g.  Mode 0: 32-row boxes of the tile's [P][128] slice, box k lives in slot k and is
    // reloaded for the next tile as soon as the builders release it; mode 1: one [160][32] block per chunk.  It
    // only waits on slot releases from the builders, so it runs as far ahead as the ring allows. =====
    if (lane == 0) {
      int slot = 0;
      uint32_t sph = 1;  // mode 1: parity to wait for on raw_empty[slot]
      for (int i = 0; i < n_my; ++i) {
        const TileCoord tc_ = tile_coord(a, i);
        const bool more = !PMT_DBG(a, 512) && i + 1 < n_my;
        const TileCoord nx = tile_coord(a, more ? i + 1 : i);
        for (int k = 0; k < a.NKC; ++k) {
          if (mode == 0) {
            if (k < a.n_gboxes) {
              // the smem ring only reaches about one tile ahead: pull the next tile's box into L2 now
              if (more && !PMT_DBG(a, 64)) tma_prefetch_l2_4d(&tmG0, nx.x0, nx.h, 32 * k, nx.n);
              PWAIT_RELAXED(0, &raw_empty[k], ((uint32_t)i & 1u) ^ 1u);  // previous tile is done with this box
              if PMT_DBG(a, 64) {   // ablation: no raw-g traffic
                mbar_arrive(&raw_full[k]);
              } else {
                mbar_arrive_expect_tx(&raw_full[k], (uint32_t)kBox0Bytes);
                tma_load_4d(smem + kTM * 4 + k * kBox0Bytes, &tmG0, tc_.x0, tc_.h, 32 * k, tc_.n, &raw_full[k]);
              }
            }
          } else {
            const int p0 = a.P - 1 + m.delta - kKC * k - (kKC - 1);
            if (more && !PMT_DBG(a, 64)) tma_prefetch_l2_4d(&tmG1, nx.x0 + m.oo + kKC * k, nx.h, p0, nx.n);
            PWAIT_RELAXED(0, &raw_empty[slot], sph);
            if PMT_DBG(a, 64) {
              mbar_arrive(&raw_full[slot]);
            } else {
              mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)kRawSlot1);
              tma_load_4d(smem + slot * kRawSlot1, &tmG1, tc_.x0 + m.oo + kKC * k, tc_.h, p0, tc_.n, &raw_full[slot]);
            }
            if (++slot == m.raw_slots) slot = 0, sph ^= 1u;
          }
        }
      }
    }
  } else if (wid == 1) {
    // ===== MMA issuer =====
    // The warp runs the loop converged (all lanes poll the barriers) and one elected lane issues, so that descriptors,
    // slots and phases live in uniform registers and the UTCHMMA sequence is straight-line code (tc::elect_one); with
    // the loop under `if (lane == 0)` this single thread, not the tensor pipe, was the busiest unit of the kernel.
    {
      const uint32_t idesc = tc::make_idesc(2, 0, 0, kTM, a.Cbox);
      const uint32_t idesc2 = tc::make_idesc(2, 0, 0, kTM, 2 * a.Cbox);  // B = [band_hi ; band_lo] stacked along N
      const uint64_t dA0 = tc::smem_desc(smem_u32(gd_ring), 16, 1024, 2);
      const uint64_t dB0 = tc::smem_desc(smem_u32(band_ring), 16, 1024, 2);
      const uint32_t a_step = (uint32_t)a.gd_slot_bytes >> 4, b_step = (uint32_t)a.band_slot_bytes >> 4;
      const uint32_t a_lo = (uint32_t)a.gd_lo_off >> 4;
      const bool tmem_a = m.tmem_a != 0, skip = PMT_DBG(a, 16) != 0;
      int gs = 0, bs = 0;
      uint32_t gph = 0, bph = 0;
      for (int i = 0; i < n_my; ++i) {
        const int buf = i & 1;
        PWAIT(0, &tmem_empty[buf], ((uint32_t)(i >> 1) & 1u) ^ 1u);  // epilogue drained this accumulator
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * a.acc_cols);
        for (int k = 0; k < a.NKC; ++k) {
          PWAIT(1, &gd_built[gs], gph);
          if (kPasses == 1) PWAIT(2, &band_full[bs], bph);
          else PWAIT(2, &band_split[bs], bph);   // hi landed (TMA) and lo written (split warps)
          tc::fence_after_sync();
          const uint64_t dB = dB0 + (uint64_t)(b_step * (uint32_t)bs);
          const uint32_t ta = tmem_base + (uint32_t)(a.a_base + gs * a.aslot_cols);
          const uint64_t dA = dA0 + (uint64_t)(a_step * (uint32_t)gs);
          PSEC_BEGIN();
          if (tc::elect_one()) {
            if (!skip) {
              if (tmem_a) {
#pragma unroll
                for (int kk = 0; kk < kKC / 8; ++kk) {
                  const uint32_t acc = (k > 0 || kk > 0) ? 1u : 0u;
                  if (kPasses == 3) {
                    tc::mma_tf32_ts(d_tmem, ta + 8 * kk, dB + 2 * kk, idesc2, acc);
                    tc::mma_tf32_ts(d_tmem, ta + 32 + 8 * kk, dB + 2 * kk, idesc, 1u);
                  } else {
                    tc::mma_tf32_ts(d_tmem, ta + 8 * kk, dB + 2 * kk, idesc, acc);
                  }
                }
              } else {
#pragma unroll
                for (int kk = 0; kk < kKC / 8; ++kk) {
                  const uint32_t acc = (k > 0 || kk > 0) ? 1u : 0u;
                  if (kPasses == 3) {
                    // D[:, 0:C] += A_hi*B_hi + A_lo*B_hi ; D[:, C:2C] += A_hi*B_lo  (A_hi is read once for both B
                    // halves; the epilogue adds the two column blocks)
                    tc::mma_tf32(d_tmem, dA + 2 * kk, dB + 2 * kk, idesc2, acc);
                    tc::mma_tf32(d_tmem, dA + a_lo + 2 * kk, dB + 2 * kk, idesc, 1u);
                  } else {
                    tc::mma_tf32(d_tmem, dA + 2 * kk, dB + 2 * kk, idesc, acc);
                  }
                }
              }
            }
            PSEC_END(0);   // profiling build: cycles in the MMA issue (the thread blocks while the tensor queue is full)
            tc::mma_commit(&gd_empty[gs]);
            tc::mma_commit(&band_empty[bs]);
            if (k == a.NKC - 1) tc::mma_commit(&tmem_full[buf]);
          }
          __syncwarp();
          PSEC_END(1);     // issue + commits + reconvergence
          if (++gs == m.a_slots) gs = 0, gph ^= 1u;
          if (++bs == m.band_slots) bs = 0, bph ^= 1u;
        }
      }
    }
  } else if (wid > 2 + kEpiWarps + kBuilders) {
    // ===== band split (3xTF32): lo = x - trunc_tf32(x) of each landed feature band chunk, written next to it
    // (position-wise, layout agnostic).  Team t takes chunks t, t + 2, ... =====
    constexpr int kTeams = BwdCfg<kPasses, kGroups>::kSplitTeams, kPer = BwdCfg<kPasses, kGroups>::kSplitPer;
    const int j = wid - (3 + kEpiWarps + kBuilders);
    const int team = j / kPer, member = j % kPer;
    const int G = n_my * a.NKC;
    const int nch = band_bytes / 16;
    int bs = team % m.band_slots;
    uint32_t bph = (uint32_t)(team / m.band_slots) & 1u;
    for (int g = team; g < G; g += kTeams) {
      PWAIT(2, &band_full[bs], bph);
      unsigned char* sb = band_ring + (size_t)bs * a.band_slot_bytes;
      if (!PMT_DBG(a, 32)) {
        for (int base = 32 * member + lane; base < nch; base += 8 * 32 * kPer) {   // 8 loads in flight per thread
          float4 x[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (base + 32 * kPer * u < nch) x[u] = *reinterpret_cast<const float4*>(sb + 16 * (base + 32 * kPer * u));
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (base + 32 * kPer * u < nch)
              *reinterpret_cast<float4*>(sb + a.band_lo_off + 16 * (base + 32 * kPer * u)) =
                  make_float4(lo_tf32(x[u].x), lo_tf32(x[u].y), lo_tf32(x[u].z), lo_tf32(x[u].w));
        }
      }
      fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&band_split[bs]);
      bs += kTeams;
      while (bs >= m.band_slots) bs -= m.band_slots, bph ^= 1u;
    }
  } else if (wid < 2 + kEpiWarps) {
    // ===== epilogue: TMEM -> coalesced global stores =====
    const int q = wid & 3;
    const int xl = 32 * q + lane;
    const int64_t pstride = (int64_t)a.H * a.W;
    for (int i = 0; i < n_my; ++i) {
      const int buf = i & 1;
      const TileCoord tc_ = tile_coord(a, i);
      PWAIT_RELAXED(0, &tmem_full[buf], (uint32_t)(i >> 1) & 1u);
      tc::fence_after_sync();
      const bool ok = tc_.x0 + xl < a.W;
      float* o = dst + ((int64_t)tc_.n * a.C * a.H + tc_.h) * (int64_t)a.W + tc_.x0 + xl;
      if (kPasses == 3) {
        // 16 columns at a time (keeps registers low with 16 builder warps): D = block0 + block1 (the A_hi*B_lo part)
        for (int cb = 0; cb < a.Cbox && !PMT_DBG(a, 8); cb += 16) {
          float v[16];
          const uint32_t t0 = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * a.acc_cols + cb);
          tc::tmem_ld16(t0, v);
          {
            float v2[16];
            tc::tmem_ld16(t0 + (uint32_t)a.Cbox, v2);
#pragma unroll
            for (int cc = 0; cc < 16; ++cc) v[cc] += v2[cc];
          }
#pragma unroll
          for (int cc = 0; cc < 16; ++cc)
            if (ok && cb + cc < a.C) o[(int64_t)(cb + cc) * pstride] = v[cc];
        }
      } else {
        for (int cb = 0; cb < a.Cbox && !PMT_DBG(a, 8); cb += 32) {
          float v[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * a.acc_cols + cb), v);
#pragma unroll
          for (int cc = 0; cc < 32; ++cc)
            if (ok && cb + cc < a.C) o[(int64_t)(cb + cc) * pstride] = v[cc];
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
  } else {
    // ===== builder warps =====
    const int bw = wid - 2 - kEpiWarps;  // 0..kBuilders-1; consecutive warps cycle through the 4 TMEM lane quarters
    const int grp = (bw >> 2) % kGroups;                  // this warp handles chunks g = grp (mod kGroups)
    const int gw = ((bw >> 2) / kGroups) * 4 + (bw & 3);  // index inside the group, 0..kGroupWarps-1
    // running ring positions for the chunks this warp visits (g = grp, grp+2, ...): slot, and the phase parity seen
    // by a consumer-side wait (full/built); producer-side waits (empty) use the opposite parity
    int g = grp;   // chunk index (only the profiling build reads it)
    int gs = grp % m.a_slots, rs = grp % m.raw_slots;
    uint32_t gph = (uint32_t)(grp / m.a_slots) & 1u, rph = (uint32_t)(grp / m.raw_slots) & 1u;
    auto advance2 = [](int& slot, uint32_t& ph, int n) {
      slot += kGroups;
      while (slot >= n) slot -= n, ph ^= 1u;
    };
    // The group walks its own chunks g = grp, grp + kGroups, ... directly as (tile i, chunk k) with a carry (NKC >= 4 >
    // kGroups): a loop over every k that skipped the other groups' chunks cost a quarter of the builders' time in
    // loop control and branch resolution (ncu source view, profiles/r02_ncu_corr.md).
    int boxes_ready = 0;
    int next_rel = 0;                    // mode 0: boxes are handed back to the producer in increasing order
    for (int i = 0, k = grp; i < n_my;) {
      {
        unsigned char* sa = gd_ring + (size_t)gs * a.gd_slot_bytes;
        if (mode == 0) {
          // rows p <= 32k+31-delta are needed: boxes 0..k of this tile's [P][128] slice
          const int need = (k + 1 < a.n_gboxes) ? k + 1 : a.n_gboxes;
          if (gw == 0) PTRACE(2 + grp, g, 0);
          while (boxes_ready < need) { PWAIT(0, &raw_full[boxes_ready], (uint32_t)i & 1u); ++boxes_ready; }
          if (gw == 0) PTRACE(2 + grp, g, 1);
          PWAIT(1, &gd_empty[gs], gph ^ 1u);
          if (gw == 0) PTRACE(2 + grp, g, 2);
          const float* Gt = reinterpret_cast<const float*>(smem);
          { PSEC_BEGIN();
          if (m.tmem_a) {
            // A operand straight into TMEM: thread = Gd row x (TMEM lane), kCols consecutive band columns per warp
            constexpr int kCols = 32 / (kGroupWarps / 4);
            static_assert(kCols == 16 || kCols == 32, "group size");
            const int q = wid & 3, sub = gw >> 2;
            const int xl = 32 * q + lane;
            const int pb = kKC * k + sub * kCols - m.delta - xl;  // p of column jj = sub*kCols + t is pb + t
            const uint32_t ta = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(a.a_base + gs * a.aslot_cols + sub * kCols);
            // slice row = min(plane + 1, P + 1) (planes outside [0, P) hit a zero row), computed on byte offsets into the
            // slice so that an element costs one add-min and one LDS: offset = 4 xl + 512 row; 4 xl < 512, so a row
            // below 0 wraps the unsigned offset far above the limit and is clamped like a row above P
            const uint32_t lane_off = (uint32_t)xl * 4u;
            const uint32_t lim_off = lane_off + ((uint32_t)a.P + 1u) * (kTM * 4u);
#pragma unroll
            for (int c0 = 0; c0 < kCols; c0 += 16) {
              float w[16];
              const uint32_t o0 = lane_off + (uint32_t)(pb + c0 + 1) * (kTM * 4u);
#pragma unroll
              for (int t = 0; t < 16; ++t)
                w[t] = *reinterpret_cast<const float*>(smem + min(o0 + (uint32_t)t * (kTM * 4u), lim_off));
              if PMT_DBG(a, 4) {
#pragma unroll
                for (int t = 0; t < 16; ++t) w[t] = 0.f;
              }
              tc::tmem_st16(ta + c0, w);
              if (kPasses == 3) {
#pragma unroll
                for (int t = 0; t < 16; ++t) w[t] = lo_tf32(w[t]);
                tc::tmem_st16(ta + 32 + c0, w);
              }
            }
            { PSEC_BEGIN();
            tc::tmem_st_wait();
            tc::fence_before_sync();
            PSEC_END(1); }
          } else {
#pragma unroll
            for (int task = gw; task < 32; task += kGroupWarps) {   // 32 warp tasks per chunk: (16-byte column c4, 32-row block xb)
              if PMT_DBG(a, 4) break;
              const int c4 = task & 7, xb = task >> 3;
              const int xl = 32 * xb + lane;
              const int pb = kKC * k + 4 * c4 - m.delta - xl;  // p of column jj = 4*c4 + t is pb + t
              float v[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) v[t] = Gt[min((unsigned)(pb + t + 1), (unsigned)a.P + 1u) * kTM + xl];
              const uint32_t off = kmajor_off(xl, 4 * c4);
              *reinterpret_cast<float4*>(sa + off) = make_float4(v[0], v[1], v[2], v[3]);
              if (kPasses == 3)
                *reinterpret_cast<float4*>(sa + a.gd_lo_off + off) =
                    make_float4(lo_tf32(v[0]), lo_tf32(v[1]), lo_tf32(v[2]), lo_tf32(v[3]));
            }
          }
          PSEC_END(0); }
          if (gw == 0) PTRACE(2 + grp, g, 3);
        } else {
          const int slot = rs;
          if (gw == 0) PTRACE(2 + grp, g, 0);
          PWAIT(0, &raw_full[slot], rph);
          if (gw == 0) PTRACE(2 + grp, g, 1);
          PWAIT(1, &gd_empty[gs], gph ^ 1u);
          if (gw == 0) PTRACE(2 + grp, g, 2);
          const float* raw = reinterpret_cast<const float*>(smem + slot * kRawSlot1);
          { PSEC_BEGIN();
          if (m.tmem_a) {
            // A operand straight into TMEM (thread = Gd row xl = TMEM lane, registers = the 32 band columns of the chunk):
            // Gd[xl][jj] = raw[r = xl + 31 - jj][jj].  The raw block is a 128-byte-swizzled TMA box ([160 rows][32 floats]),
            // so element (r, jj) sits at byte r*128 + (((jj >> 2) ^ (r & 7)) << 4) + 4*(jj & 3).  Read in natural order
            // (all lanes the same jj) the 32 consecutive rows of a warp hit every 16-byte slot 4 times: 4-way bank
            // conflicts (ncu: half of the kernel's shared wavefronts).  Instead lane l reads, in instruction (mm, n), the
            // column jj = 4 mm + ((n + h) & 3) with h = l >> 3: the 8 lanes of one h cover the 8 slots at one 4-byte
            // position, the four h the four positions -- conflict-free -- and each group of 4 registers is rotated back
            // by h afterwards (two selects per element).  With jj = 4 mm + c: r & 7 = ((xl + 31 - c) & 7) ^ (4 (mm & 1)),
            // so the offset is base - 512 mm + ((K_mm << 4) ^ e) with per-thread constants base, e (one pair per n) and
            // compile-time K_mm = mm ^ 4 (mm & 1).  No shared->shared copy (1 LDS + 2 STS per element before), no Gd ring
            // in shared memory (96 KB that now hold a deeper raw ring), no A-operand fetches by the MMA (32 KB / chunk).
            const int q = wid & 3;
            const int xl = 32 * q + lane;
            const int h = lane >> 3;
            const uint32_t ta = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(a.a_base + gs * a.aslot_cols);
            const unsigned char* rawb = smem + slot * kRawSlot1;
            uint32_t base_n[4], e_n[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              const int c = (n + h) & 3;                 // the column (within a group of 4) this lane reads in slot n
              base_n[n] = (uint32_t)((xl + 31 - c) * 128 + 4 * c);
              e_n[n] = (uint32_t)(((xl + 31 - c) & 7) << 4);
            }
            const bool rot1 = (h & 1) != 0, rot2 = (h & 2) != 0;
#pragma unroll
            for (int c0 = 0; c0 < kKC; c0 += 16) {
              float w[16];
#pragma unroll
              for (int t = 0; t < 16; ++t) {
                const int jj = c0 + t, mm = jj >> 2, n = jj & 3;
                const uint32_t km = (uint32_t)((mm ^ (4 * (mm & 1))) << 4);
                w[t] = *reinterpret_cast<const float*>(rawb + (base_n[n] - (uint32_t)(512 * mm) + (km ^ e_n[n])));
              }
              // w[4 g + n] holds column 4 g + ((n + h) & 3): rotate every group of 4 back by h
#pragma unroll
              for (int gq = 0; gq < 16; gq += 4) {
                const float a0 = rot1 ? w[gq + 3] : w[gq + 0], a1 = rot1 ? w[gq + 0] : w[gq + 1];
                const float a2 = rot1 ? w[gq + 1] : w[gq + 2], a3 = rot1 ? w[gq + 2] : w[gq + 3];
                w[gq + 0] = rot2 ? a2 : a0, w[gq + 1] = rot2 ? a3 : a1;
                w[gq + 2] = rot2 ? a0 : a2, w[gq + 3] = rot2 ? a1 : a3;
              }
              if PMT_DBG(a, 4) {
#pragma unroll
                for (int t = 0; t < 16; ++t) w[t] = 0.f;
              }
              tc::tmem_st16(ta + c0, w);
              if (kPasses == 3) {
#pragma unroll
                for (int t = 0; t < 16; ++t) w[t] = lo_tf32(w[t]);
                tc::tmem_st16(ta + 32 + c0, w);
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_empty[slot]);   // every value of the block is in registers / on its way to TMEM
            tc::tmem_st_wait();
            tc::fence_before_sync();
          } else if (!PMT_DBG(a, 4)) {
            // 32 warp tasks per chunk (4 Gd rows x 32 columns, lane = column); this warp takes tasks gw, gw+G, ...
            // All loads are issued before the first store so the LDS latency is paid once per chunk, not per row.
            constexpr int kTasks = 32 / kGroupWarps;
            float v[kTasks * 4];
            // Gd row xl = 4*(gw + G*i) + t reads g[p_first + xl + 31 - lane][column lane]
            const float* src = raw + (4 * gw + 31 - lane) * kKC + lane;
#pragma unroll
            for (int i = 0; i < kTasks; ++i)
#pragma unroll
              for (int t = 0; t < 4; ++t) v[4 * i + t] = src[(4 * kGroupWarps * i + t) * kKC];
            // K-major SW128 offset of (xl, lane): xl & 7 = 4*(gw & 1) + t and xl >> 3 = (gw >> 1) + (G/2)*i
            unsigned char* dst = sa + (gw >> 1) * 1024 + 4 * (gw & 1) * 128 + ((lane & 3) << 2);
#pragma unroll
            for (int i = 0; i < kTasks; ++i)
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const uint32_t off = (uint32_t)((kGroupWarps / 2) * i * 1024 + t * 128) +
                                     (uint32_t)((((lane >> 2) ^ (4 * (gw & 1) + t)) & 7) << 4);
                *reinterpret_cast<float*>(dst + off) = v[4 * i + t];
                if (kPasses == 3) *reinterpret_cast<float*>(dst + a.gd_lo_off + off) = lo_tf32(v[4 * i + t]);
              }
          }
          PSEC_END(0); }
          if (!m.tmem_a) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_empty[slot]);
          }
        }
        { PSEC_BEGIN();
        if (!m.tmem_a) fence_proxy_async();   // shared-memory Gd: generic-proxy writes -> async-proxy (MMA) reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&gd_built[gs]);
        PSEC_END(3); }
        if (mode == 0) {
          // Hand back every box of the g slice this warp will not read again -- after the A operand has been announced,
          // so the releases are off the path to the MMA.  Box b is last used by chunk min(NKC-1, b+koff): boxes
          // b <= k + kGroups - 1 - koff are dead for this warp, and on its last visit of the tile all of them are.
          const int target = (k + kGroups >= a.NKC) ? a.n_gboxes : min(a.n_gboxes, k + kGroups - m.koff);
          if (lane == 0)
            for (int b = next_rel; b < target; ++b) mbar_arrive(&raw_empty[b]);
          if (target > next_rel) next_rel = target;
        }
        if (gw == 0) PTRACE(2 + grp, g, 6);
        advance2(gs, gph, m.a_slots);
        advance2(rs, rph, m.raw_slots);
      }
      g += kGroups;
      k += kGroups;
      if (k >= a.NKC) k -= a.NKC, ++i, boxes_ready = 0, next_rel = 0;
    }
  }

#ifdef PMT_BWD_PROFILE
  if (prof != nullptr && cta_in_mode == 0 && lane == 0) {
    long long* o = prof + ((size_t)mode * 32 + wid) * 10;
    o[0] = wait_cyc[0], o[1] = wait_cyc[1], o[2] = wait_cyc[2], o[3] = wait_cyc[3];
    o[4] = clock64() - t_start;
    o[5] = sec_cyc[0], o[6] = sec_cyc[1], o[7] = sec_cyc[2], o[8] = sec_cyc[3];
  }
#endif
  tc::fence_before_sync();
  __syncthreads();
  if (wid == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

int fill_args(TcBwdArgs* a, int C, int H, int W, int P, int passes, int groups) {
  a->C = C, a->H = H, a->W = W, a->P = P, a->rW = (P - 1) / 2;
  if (C > 128) return 1;
  a->Cbox = round_up(C, 32);  // the epilogue reads TMEM in 32-column groups; UMMA N % 16 == 0
  const int oo0 = -a->rW, oo1 = -(P - 1 - a->rW);
  a->m[0].delta = ((oo0 % 4) + 4) % 4;
  a->m[0].oo = oo0 - a->m[0].delta;
  a->m[1].delta = ((oo1 % 4) + 4) % 4;
  a->m[1].oo = oo1 - a->m[1].delta;
  a->m[0].koff = (kTM + kKC - 2 + a->m[0].delta) / kKC;
  a->m[1].koff = 0;
  const int dmax = a->m[0].delta > a->m[1].delta ? a->m[0].delta : a->m[1].delta;
  a->NKC = ceil_div(kTM + P - 1 + dmax, kKC);
  a->n_xtiles = ceil_div(W, kTM);
  a->n_tiles = 0;  // set by the launcher (needs B)
  a->n_gboxes = ceil_div(P, 32);
  if (a->n_gboxes > 8) return 1;
  const int raw0 = a->n_gboxes * kBox0Bytes;
  const int mult = passes == 3 ? 2 : 1;
  a->gd_lo_off = kGdBytes;
  a->gd_slot_bytes = kGdBytes * mult;
  a->band_lo_off = a->Cbox * 128;
  a->band_slot_bytes = a->Cbox * 128 * mult;
  a->acc_cols = a->Cbox * mult;                 // 3xTF32 keeps two column blocks per accumulator
  if (2 * a->acc_cols > 512) return 1;
  const int smem_budget = 227 * 1024 - 1024;
  // A operand in TMEM (mode 0 only -- its Gd rows are conflict-free column reads of the resident g slice): needs
  // columns next to the two accumulators.
  a->aslot_cols = 32 * mult;
  a->a_base = 2 * a->acc_cols;
  int tmem_slots = (512 - a->a_base) / a->aslot_cols;
  if (tmem_slots > kMaxGdSlots) tmem_slots = kMaxGdSlots;
  const bool want_tmem_a = PMT_ENV_INT("PMT_NO_TMEM_A", 0) == 0;
  int max_end = 0;
  for (int md = 0; md < 2; ++md) {
    TcBwdMode& m = a->m[md];
    // every ring a builder group waits on needs at least one slot per group (parity waits tolerate one phase of lead)
    m.tmem_a = (want_tmem_a && tmem_slots >= groups && (md == 0 || PMT_ENV_INT("PMT_BWD_TMEM_A1", 1) != 0)) ? 1 : 0;
    int gd_bytes = 0;
    if (md == 0) {
      m.raw_slots = a->n_gboxes;
      m.gd_off = round_up(raw0 + 2 * kTM * 4, 1024);   // + the zero rows before plane 0 and after the last box
    } else {
      if (m.tmem_a) m.raw_slots = 6;   // no Gd ring in shared memory: two raw blocks in flight per builder group
      else m.raw_slots = groups < 3 ? kMaxRawSlots1 : groups;   // 3 groups: 3 raw + 3 Gd slots fit, 4 + 3 do not
      m.gd_off = round_up(m.raw_slots * kRawSlot1, 1024);
    }
    if (m.tmem_a) {
      m.a_slots = tmem_slots;
    } else {
      m.a_slots = groups;
      gd_bytes = m.a_slots * a->gd_slot_bytes;
    }
    m.band_off = m.gd_off + gd_bytes;
    int bslots = (smem_budget - m.band_off) / a->band_slot_bytes;
    if (bslots > kMaxBandSlots) bslots = kMaxBandSlots;
    if (bslots < groups) return 1;
    m.band_slots = bslots;
    const int end = m.band_off + bslots * a->band_slot_bytes;
    if (end > max_end) max_end = end;
  }
  a->bar_off = max_end;
  int cols = 32;
  const int need_cols = a->m[0].tmem_a ? a->a_base + a->m[0].a_slots * a->aslot_cols : 2 * a->acc_cols;
  while (cols < need_cols) cols *= 2;
  a->tmem_cols = cols;
  a->debug = PMT_ENV_INT("PMT_TC_DEBUG", 0);
  a->relax_ns = PMT_ENV_INT("PMT_BWD_RELAX_NS", 64);
  return 0;
}

}  // namespace

bool corr1d_bwd_tc_ok(const void* in1, const void* in2, int C, int H, int W, int P, int dilp, int passes) {
  if (dilp != 1 || P < 1 || C < 1 || W % 4 != 0 || !aligned16(in1) || !aligned16(in2)) return false;
  TcBwdArgs a;
  return fill_args(&a, C, H, W, P, passes, 2) == 0;
}

long long* g_bwd_prof = nullptr;  // device buffer for PMT_BWD_PROFILE builds (set through pmt_debug_set_ptr)

int launch_corr1d_bwd_tc(const float* in1, const float* in2, const float* gout, float* gin1, float* gin2, int B,
                         int C, int H, int W, int P, int passes, cudaStream_t st) {
  TcBwdArgs a;
  PMT_CHECK_ARG(passes == 1 || passes == 3, "corr1d tc: passes must be 1 (tf32) or 3 (3xtf32)");
  int groups = passes == 3 ? 3 : 2;   // 3xTF32: 3 builder groups when their shared-memory rings fit, else 2
  if (const int e = PMT_ENV_INT("PMT_BWD_GROUPS", 0)) groups = e == 2 ? 2 : 3;
  if (fill_args(&a, C, H, W, P, passes, groups) != 0) groups = 2;
  PMT_CHECK_ARG(fill_args(&a, C, H, W, P, passes, groups) == 0, "corr1d tc bwd: unsupported shape C=%d P=%d", C, P);
  CUtensorMap tm1, tm2, tmG0, tmG1;
  if (int e = make_tmap_nchw_ex(&tm1, in1, B, C, H, W, kKC, a.Cbox, 1)) return e;
  if (int e = make_tmap_nchw_ex(&tm2, in2, B, C, H, W, kKC, a.Cbox, 1)) return e;
  if (int e = make_tmap_nchw_ex(&tmG0, gout, B, P, H, W, kTM, 32, 0)) return e;
  if (int e = make_tmap_nchw_ex(&tmG1, gout, B, P, H, W, kKC, kRawRows1, a.m[1].tmem_a ? 1 : 0)) return e;
  const int smem_bytes = a.bar_off + 1024;
  const int64_t tiles = (int64_t)B * H * a.n_xtiles;
  PMT_CHECK_ARG(tiles < (1ll << 31), "corr1d tc bwd: too many tiles");
  a.n_tiles = (int)tiles;
  // persistent: one CTA per SM.  The two gradients cost differently per tile (gin1 builds its A operand in TMEM and
  // is cheaper), so the SMs are split unevenly to finish together.
  const int sms = sm_count();
  int64_t n_cta = 2 * tiles < sms ? 2 * tiles : sms;
  if (n_cta < 2) n_cta = 2;
  // measured optimum at the headline shape (both A operands in TMEM): 72 of 148 SMs for gin1 -- its builders read the
  // resident g slice conflict-free, gin2's read the swizzled raw blocks with 4-way bank conflicts
  int n0 = (int)(n_cta * (a.m[0].tmem_a && a.m[1].tmem_a && passes == 3 ? 0.4865 : 0.5) + 0.5);
  if (const int e = PMT_ENV_INT("PMT_BWD_SPLIT", 0)) n0 = e;  // tuning knob
  if (n0 < 1) n0 = 1;
  if (n0 > n_cta - 1) n0 = (int)n_cta - 1;
  a.n_cta0 = n0;
#define PMT_LAUNCH_BWD(PASSES, GROUPS)                                                                              \
  do {                                                                                                              \
    PMT_CUDA_OK(cudaFuncSetAttribute(corr1d_bwd_tc_kernel<PASSES, GROUPS>,                                          \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));                     \
    corr1d_bwd_tc_kernel<PASSES, GROUPS><<<dim3((unsigned)n_cta), BwdCfg<PASSES, GROUPS>::kThreads, smem_bytes, st>>>( \
        tm1, tm2, tmG0, tmG1, gin1, gin2, a, g_bwd_prof);                                                           \
  } while (0)
  if (passes == 3) {
    if (groups == 3) PMT_LAUNCH_BWD(3, 3);
    else PMT_LAUNCH_BWD(3, 2);
  } else {
    if (groups == 3) PMT_LAUNCH_BWD(1, 3);
    else PMT_LAUNCH_BWD(1, 2);
  }
#undef PMT_LAUNCH_BWD
  PMT_LAUNCH_OK("corr1d_bwd_tc_kernel");
  return PMT_OK;
}

}  // namespace pmt
