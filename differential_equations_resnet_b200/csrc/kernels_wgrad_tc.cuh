// K4: dense weight gradient G[tap][ci][o] = sum_q x_halo[q + shift(tap), ci] * dZ[q, o] on tcgen05
// tensor cores (replaces TF's Conv2DBackpropFilter, training/training.py:300).
//
// GEMM view: M = input channel, N = output channel, K = pixel position in the same padded linear
// space the forward kernel uses (pitch P = W+1).  Both operands arrive NHWC, i.e. with the M/N index
// contiguous and K strided, so they are consumed as MN-major UMMA operands straight from the TMA
// halo strips -- no transposition anywhere:
//   * tf32 modes: 128B-swizzle / 32B-atom layout (the only MN-major layout tf32 accepts), chunks of
//     32 channels, 4-position groups (SBO = 512B); C = 16 is zero-padded to 32 by TMA OOB fill.
//   * bf16 mode : canonical 32/64/128B swizzles, chunks of min(C,64) channels, 8-position groups.
// A 3x3 tap is a whole-row shift of the x strip's start address.  When all channels fit one chunk
// ("beta trick") the M-chunk stride (LBO) is set to ONE position so a single M=4*chunk MMA covers
// the three taps beta=0,1,2 of a kernel row (4th chunk is junk) -- 3 instead of 9 MMAs per k-step.
// The junk padded column of dZ is zero (TMA OOB), so junk positions add nothing.
// Each CTA owns (tap group, N range, M block, slice of positions), keeps its accumulators in TMEM for its
// whole slice and writes ONE fp32 partial; partials are reduced deterministically afterwards.
// Tiles are either position-granular (KT positions from row-granular strips) or, for 128-channel operand
// blocks, ROW-ALIGNED (R whole image rows: strips hold exactly R+2 / R rows, the k-steps run into a zeroed
// pad); the dz column sums (bias gradient) are dealt over all CTAs that stage the same dz strip.
// Strict mode = 3xTF32: x_hi*dz_hi + x_hi*dz_lo + x_lo*dz_hi with lo strips produced by converter warps.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "kernels_conv_tc.cuh"
#include "sm100_ptx.cuh"

namespace b200ode {

struct WgradParams {
  int N, H, W, C, P;
  int KT;            // positions per tile (multiple of the MMA K extent)
  int tstride;       // positions between consecutive tile starts (= KT, or rows*P for row-aligned tiles)
  int rowtiles;      // row-aligned tiles: a tile is `rows` whole image rows; strips hold exactly the rows needed and
                     // the k-steps run over a zeroed pad up to the next multiple of the MMA K extent
  int tpi;           // tiles per image
  int total_tiles, nparts;
  int RBx, RBd;      // strip rows of x (halo) and dz
  int CH;            // channels per chunk
  int RWB;           // bytes per position row within a chunk
  int xchunks;       // x chunks per stage (all input channels)
  int dchunks;       // dz chunks per stage (this CTA's N range)
  int trick;         // beta trick: one MMA per kernel row
  int TG;            // taps per CTA group (normal mode)
  int NT;            // output channels per CTA group
  int ntapgroups, nngroups;
  int mgroups;       // M blocks (of Mblk input channels) dealt to different CTAs; each CTA stages only its block's x chunks
  int MB, Mblk;      // M blocks per tap and rows per block (normal mode)
  uint32_t x_chunk_bytes, d_chunk_bytes, x_chunk_stride, d_chunk_stride;
  uint32_t x_off, d_off, x_lo_off, d_lo_off, stage_stride, bar_off;
  int stages;
  uint32_t tmem_cols;
  float* partials;   // [nparts][9][C][C]
  float* bias_partials;  // [nparts][C] column sums of dz (bias gradient), written by tap group 0
  uint32_t ent_off, bsum_off;  // smem offsets: per-entry A offsets (uint32[32]) and bias scratch (float[128][8])
  uint64_t* trace;             // nullable timeline buffer (debug)
  // layer batching (blockIdx.z = layer): layer 0 reads its input through map_x0 (image n), layer l >= 1
  // through map_x (image (l-1)*N + n: the saved outputs of the chain); dz image index is l*N + n.
  int L;
  long long part_layer_stride, bias_layer_stride;   // floats between consecutive layers' partials
  // Pixel-pair mode (tf32, C = 16, W even).  The TMA maps view the activations as pixel PAIRS
  // ([.., W/2, 32 floats]; csrc/tma_layout_probe.cu: a 64-byte inner box would be padded to a 128-byte
  // pitch), the padded pitch is P = W+2 (two zero slots = one out-of-bounds pair at the end of every
  // row) and one 128-byte MN-major operand row holds two positions: M = (kernel row alpha, parity px,
  // row holds two positions: M = (row copy c, parity px, ci), N = (parity pd, o).  With the x strip
  // starting one image row above the dz strip, tap (alpha, beta) is the position shift
  // alpha*P + beta - 1 = 2*(alpha*P/2 + j) + px - pd with j = floor((beta-1+pd)/2) in {-1,0,1}: per
  // kernel row ONE M=128 MMA (four row copies c = j+1 through LBO = one row, start row alpha*P/2 - 1)
  // covers 16 positions of all three beta taps -- three MMAs per 16 positions instead of per 8, and no
  // zero-padded channels.  The row before each x strip is a zeroed 1 KB pad (x_off = 1024).
  // The accumulators are written raw ([3][128 lanes][32 cols] per part) and gathered by fold_reduce_kernel.
  int pair;
  // Double-shift mode (16-bit operands, all channels in one chunk: C = 16 / 32 / 64).  BOTH operands carry tap shifts
  // through their MN-major chunk stride: A (x strip) = s2_M / C chunks one KERNEL ROW apart (chunk a <-> alpha = a),
  // B (dz strip) = s2_N / C chunks one POSITION apart (chunk j <-> beta = 2 - j, A starting one position after B), so
  // ONE M = 4C x N = 4C MMA per 16 positions yields all nine taps (7 of its 16 blocks are junk) instead of three
  // M = 4C x N = C MMAs -- the MMAs of these small layers are bound by a fixed issue cost (~50 cycles measured at
  // M = 64, N = 16), not by their size.  C = 64: two M = 128 (alpha 0,1 | 2,-) x N = 192 MMAs instead of nine 64 x 64.
  // Both strips use pitch P = W + 2 with the two zero columns on the LEFT (TMA box from column -2), so the positions a
  // shifted B chunk reads before / after a row are zeros and nothing outside a strip's rows is ever needed.
  int shift2, s2_M, s2_N, s2_nmma;
  // CTA pairs by TAPS (TWO with one M block, C = 128): both CTAs of the pair stage the SAME input channels, the peer one image
  // row lower, so the rows 128..255 of the pair's M = 256 MMA are kernel row alpha + 1 of the leader's alpha: six M = 256 MMAs
  // per k-step (three of them half junk: kernel row 3) replace nine M = 128 ones, and each CTA stages half of the dz strip.
  int tappair;
  int dbg_nostack;   // debug (B200ODE_WGRAD_NOSTACK): strict mode issues three MMAs per entry instead of the stacked pair (A/B runs)
  int dbg;           // debug (B200ODE_WGRAD_DBG): bit 0 = issue no MMAs, bit 1 = no bias column sums (timing experiments only)
  int PB;            // bytes per position in shared memory (64 in pair mode, else RWB)
  long long part_stride;   // floats per partial (9*C*C, or 3*128*32 in pair mode)
};

// TWO = true (bf16, C = 256 with one 128-channel M block per CTA): CTA PAIRS (cta_group::2, see kernels_conv_tc.cuh).  The
// two M blocks of an (N range, tap group, position slice) used to be two independent CTAs that each staged the whole dz
// strip; as a cluster of 2 they issue ONE M = 256 MMA per tap and k-step -- rows 0..127 = input channels 0..127 from the
// leader's x strip, rows 128..255 = channels 128..255 from the peer's -- with the dz strip split between them (NT/2 output
// channels each): per SM 4 KB + NT*16 B of operand reads per MMA instead of 4 KB + NT*32 B, and half the dz staging.
template <int MODE, bool F16 = false, bool TWO = false>
__global__ void __launch_bounds__((MODE == MODE_STRICT ? 10 : 6) * 32, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x,
                const __grid_constant__ CUtensorMap map_d, const WgradParams p) {
  constexpr bool STRICT = MODE == MODE_STRICT;
  constexpr bool BF16 = MODE == MODE_BF16;
  constexpr int UKP = BF16 ? 16 : 8;  // positions per MMA
  static_assert(!TWO || (BF16 && !F16), "CTA pairs: bf16 weight gradient");
  const uint32_t prank = TWO ? cluster_ctarank() : 0u;   // = M group of this CTA (the pair's cluster spans the M groups)

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.bar_off);  // [stages]
  uint64_t* empty = full + p.stages;
  uint64_t* conv = empty + p.stages;
  uint64_t* acc_full = conv + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Trace tr;
  tr.begin(p.trace);
  if (threadIdx.x == 0) tr.wall(0);
  // CTA pair: the cluster spans blockIdx.x (x = 2 * part + M group), blockIdx.y = (tap group, N range)
  const int group = (TWO && !p.tappair) ? (int)blockIdx.y * p.mgroups + (int)(blockIdx.x & 1) : (int)blockIdx.y;
  const int mgroup = group % p.mgroups;
  const int tapgroup = (group / p.mgroups) / p.nngroups, ngroup = (group / p.mgroups) % p.nngroups;
  // The dz column sums (bias gradient) of an N range are dealt over the CTAs that stage the same dz strip (all tap
  // groups and M blocks): each sums its share of the strip's 16-byte column units, so no role is slower than the others.
  const int b_upr = p.RWB >> 4;                    // 16-byte units per operand row
  const int b_units = p.dchunks * b_upr;           // column units of this CTA's dz strip (power of two, <= 64)
  // (CTA pair: each CTA stages only ITS half of the strip, shared with the other tap groups of the same pair position)
  const int b_roles = p.trick ? 1 : TWO ? p.ntapgroups : p.ntapgroups * p.mgroups, b_rid = p.trick ? 0 : TWO ? tapgroup : tapgroup * p.mgroups + mgroup;
  const int b_lo = b_rid * b_units / b_roles, b_n = (b_rid + 1) * b_units / b_roles - b_lo;
  const bool do_bias = b_n > 0 && !(p.dbg & 2);
  const int part = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int layer = blockIdx.z;
  const int ukp = p.pair ? (BF16 ? 32 : 16) : UKP;    // positions per k-step (pixel pairs: two per operand row)
  const CUtensorMap* mx = layer == 0 ? &map_x0 : &map_x;
  const int img_x0 = layer == 0 ? 0 : (layer - 1) * p.N;
  const int img_d0 = layer * p.N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(mx);
    tma_prefetch_desc(&map_d);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], do_bias ? 5 : 1); mbar_init(&conv[i], TWO ? 1 : 4); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (TWO) { tmem_alloc2(tmem_slot, p.tmem_cols); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
  }
  if (p.pair) {   // zero the 1 KB pad in front of every stage's x strip (row -1 of the first kernel row); strict: of the lo strip too
    for (int i = threadIdx.x; i < p.stages * 64; i += blockDim.x) {
      reinterpret_cast<uint4*>(smem + (i >> 6) * p.stage_stride)[i & 63] = make_uint4(0u, 0u, 0u, 0u);
      if (STRICT) reinterpret_cast<uint4*>(smem + (i >> 6) * p.stage_stride + (p.x_lo_off - p.x_off))[i & 63] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  if (p.rowtiles) {   // zero what TMA never writes between a chunk's rows and its stride: the k-steps past the last row read it
    const uint32_t xpad = (p.x_chunk_stride - p.x_chunk_bytes) >> 4, dpad = (p.d_chunk_stride - p.d_chunk_bytes) >> 4;
    const uint32_t per_stage = p.xchunks * xpad + p.dchunks * dpad;
    for (uint32_t i = threadIdx.x; i < p.stages * per_stage; i += blockDim.x) {
      const uint32_t s = i / per_stage, r = i % per_stage;
      uint8_t* dst;
      if (r < p.xchunks * xpad) dst = smem + s * p.stage_stride + p.x_off + (r / xpad) * p.x_chunk_stride + p.x_chunk_bytes + ((r % xpad) << 4);
      else { const uint32_t r2 = r - p.xchunks * xpad; dst = smem + s * p.stage_stride + p.d_off + (r2 / dpad) * p.d_chunk_stride + p.d_chunk_bytes + ((r2 % dpad) << 4); }
      *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (TWO) cluster_sync_all();     // both CTAs' barriers exist before any remote complete_tx / multicast commit
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) tr.mark(1);
  const uint32_t stage_bytes = p.xchunks * p.x_chunk_bytes + p.dchunks * p.d_chunk_bytes;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, rs = 0, rph = 0;   // ring position kept incrementally
      for (int tile = part; tile < p.total_tiles; tile += p.nparts, ++it) {
        const int n = tile / p.tpi, q0 = (tile % p.tpi) * p.tstride;
        const int row0 = q0 / p.P;
        const int xc0 = p.pair ? 0 : p.shift2 ? -2 : -1;    // pair mode: the zero slots sit at the END of every row
        const int dc0 = p.shift2 ? -2 : 0;
        const uint32_t s = rs, ph = rph;
        if (++rs == (uint32_t)p.stages) { rs = 0; rph ^= 1; }
        mbar_wait_sleep_lean(&empty[s], ph ^ 1);
        uint8_t* sb = smem + s * p.stage_stride;
        if constexpr (TWO) {
          // p.dchunks = the chunks THIS CTA stages (its half of the N range).  The loads complete on this CTA's OWN barrier
          // (its bias warps read the dz strip too); the peer's MMA warp forwards "stage full" to the leader (conv[s]).
          mbar_expect_tx(&full[s], stage_bytes);
          for (int c = 0; c < p.xchunks; ++c)
            tma_load_4d(sb + p.x_off + c * p.x_chunk_stride, mx, &full[s], mgroup * p.Mblk + c * p.CH, xc0,
                        row0 - 1 + (p.tappair ? (int)prank : 0), img_x0 + n);
          for (int c = 0; c < p.dchunks; ++c)
            tma_load_4d(sb + p.d_off + c * p.d_chunk_stride, &map_d, &full[s], ngroup * p.NT + ((int)prank * p.dchunks + c) * p.CH, dc0, row0,
                        img_d0 + n);
          continue;
        }
        mbar_expect_tx(&full[s], stage_bytes);
        for (int c = 0; c < p.xchunks; ++c)
          tma_load_4d(sb + p.x_off + c * p.x_chunk_stride, mx, &full[s], mgroup * p.Mblk * (p.mgroups > 1) + c * p.CH, xc0, row0 - 1, img_x0 + n);
        for (int c = 0; c < p.dchunks; ++c)
          tma_load_4d(sb + p.d_off + c * p.d_chunk_stride, &map_d, &full[s], ngroup * p.NT + c * p.CH, dc0, row0, img_d0 + n);
      }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform control flow, one elected lane issues (see kernels_conv_tc.cuh)
    const bool committer = elect_one() && prank == 0;       // CTA pair: the leader CTA issues and commits for both
    const bool leader = committer && !(p.dbg & 1);
    const int Mrows = TWO ? 256 : p.pair ? 128 : p.shift2 ? p.s2_M : p.trick ? 4 * p.CH : p.Mblk;
    const uint32_t idesc = make_instr_desc(BF16 ? (F16 ? FMT_F16 : FMT_BF16) : FMT_TF32, Mrows, p.shift2 ? p.s2_N : p.NT, 1, 1);
    const uint32_t lt = BF16 ? swz_layout_type(p.RWB) : 1u;   // 1 = SWIZZLE_128B_BASE32B
    const uint32_t sbo = BF16 ? 8u * p.RWB : 512u;
    const uint32_t hi32 = (sbo >> 4) | (1u << 14) | (lt << 29);
    // M-chunk stride: pair mode = one kernel row (P/2 operand rows); beta trick = one position; else one channel chunk
    const uint32_t lbo_a = (((p.trick ? (uint32_t)p.RWB : p.x_chunk_stride) >> 4) & 0x3FFF) << 16;
    const uint32_t lbo_b = ((p.d_chunk_stride >> 4) & 0x3FFF) << 16;
    auto mk = [&](uint32_t lo, uint32_t lbo) -> uint64_t { return (static_cast<uint64_t>(hi32) << 32) | (lo | lbo); };
    const uint32_t RU = (uint32_t)p.PB >> 4;           // 16-byte units per position
    const uint32_t smem_base = smem_u32(smem);
    const int nent = p.trick ? 3 : p.TG * p.MB;
    const int ACCW = STRICT ? 2 * p.NT : p.NT;   // strict: main + correction accumulators (see conv kernel)
    uint32_t* ent = reinterpret_cast<uint32_t*>(smem + p.ent_off);
    if (lane < nent) {
      int shift; uint32_t a_off;
      if (p.trick) { shift = lane * p.P; a_off = 0; }
      else {
        const int tap = tapgroup * p.TG * (TWO && p.tappair ? 2 : 1) + lane / p.MB;    // tap pairs: the LEADER's kernel row 2 * tapgroup
        shift = (tap / 3) * p.P + (tap % 3);
        a_off = (lane % p.MB) * (p.Mblk / p.CH) * p.x_chunk_stride;
      }
      // pixel-pair mode starts every kernel row one operand row early (the zero slots sit at the END of the previous row)
      ent[lane] = (uint32_t)shift * RU + (a_off >> 4) - (p.pair ? (uint32_t)p.RWB >> 4 : 0u);
    }
    __syncwarp();
    uint32_t entr[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) entr[e] = e < nent ? ent[e] : 0u;
    const int ksteps = p.KT / ukp;
    const uint32_t lo_x = (p.x_lo_off - p.x_off) >> 4, lo_d = (p.d_lo_off - p.d_off) >> 4;
    // Strict, one dz chunk per CTA (C <= 32): the hi and lo dz strips are consumed by ONE MMA with N = 2 * NT whose second
    // N chunk (LBO = distance hi strip -> lo strip) is the lo strip, so x_hi*dz_hi and x_hi*dz_lo land in the adjacent
    // [main | correction] accumulators at the operand-read cost of one MMA: two MMAs per entry and k-step instead of three.
    const bool stack = STRICT && p.dchunks == 1 && 2 * p.NT <= 256 && lo_d < 0x4000u && !p.dbg_nostack;
    const uint32_t idesc2 = make_instr_desc(FMT_TF32, Mrows, 2 * p.NT, 1, 1);
    const uint32_t lbo_b_stack = (lo_d & 0x3FFFu) << 16;
    uint32_t it = 0, rs = 0, rph = 0;
    if (TWO && prank != 0) {
      // peer CTA of a pair: its strips are consumed by the leader's MMAs; this warp tells the leader when a stage has landed
      for (int tile = part; tile < p.total_tiles; tile += p.nparts) {
        const uint32_t s = rs, ph = rph;
        if (++rs == (uint32_t)p.stages) { rs = 0; rph ^= 1; }
        mbar_wait_lean(&full[s], ph);
        if (lane == 0) mbar_arrive_cluster(&conv[s], 0);
        __syncwarp();
      }
    } else
    for (int tile = part; tile < p.total_tiles; tile += p.nparts, ++it) {
      const int q0 = (tile % p.tpi) * p.tstride;
      const uint32_t off0 = (uint32_t)(q0 - (q0 / p.P) * p.P) * RU;
      const uint32_t s = rs, ph = rph;
      if (++rs == (uint32_t)p.stages) { rs = 0; rph ^= 1; }
      mbar_wait_lean(STRICT ? &conv[s] : &full[s], ph);
      if constexpr (TWO) mbar_wait_lean(&conv[s], ph);     // the peer's half of the stage
      if (it == 0 && lane == 0) tr.mark(2);
      tc_fence_after_sync();
      uint32_t xu = ((smem_base + s * p.stage_stride + p.x_off) >> 4) + off0;
      uint32_t du = ((smem_base + s * p.stage_stride + p.d_off) >> 4) + off0;
      if (BF16 && p.shift2) {
        // double shift: chunk stride of A = one kernel row, of B = one position; A starts one position after B
        const uint32_t lbo_a2 = ((((uint32_t)p.P * (uint32_t)p.RWB) >> 4) & 0x3FFF) << 16;
        const uint32_t lbo_b2 = (((uint32_t)p.RWB >> 4) & 0x3FFF) << 16;
        const uint32_t e1 = 2u * (uint32_t)p.P * RU;          // second MMA (C = 64): kernel rows 2 (and a junk row 3)
        const uint32_t d0 = tmem_base, d1 = tmem_base + (uint32_t)p.s2_N;
        const bool two = p.s2_nmma == 2;
        xu += RU;
#pragma unroll 4
        for (int ks = 0; ks < ksteps; ++ks, xu += ukp * RU, du += ukp * RU) {
          const uint64_t dsc_b = mk(du, lbo_b2);
          const uint32_t accum = (it | ks) != 0;
          const uint64_t a0 = mk(xu, lbo_a2), a1 = mk(xu + e1, lbo_a2);
          if (leader) {
            umma_f16(d0, a0, dsc_b, idesc, accum);
            if (two) umma_f16(d1, a1, dsc_b, idesc, accum);
          }
        }
      } else if (p.trick && !STRICT) {
        // beta trick: three M = 4*CH MMAs per k-step (one per kernel row), offsets held in registers
        // start row of kernel row alpha: alpha*P positions; pair mode: alpha*P/2 - 1 operand rows (8 units each)
        const uint32_t urw = (uint32_t)p.RWB >> 4;             // 16-byte units per operand row (pair mode: 8 tf32, 4 fp16 / bf16)
        const uint32_t e0 = p.pair ? 0u - urw : 0u;
        const uint32_t e1 = p.pair ? (uint32_t)(p.P >> 1) * urw - urw : (uint32_t)p.P * RU;
        const uint32_t e2 = p.pair ? (uint32_t)p.P * urw - urw : 2u * e1;
        const uint32_t d0 = tmem_base, d1 = tmem_base + ACCW, d2 = tmem_base + 2 * ACCW;
#pragma unroll 4
        for (int ks = 0; ks < ksteps; ++ks, xu += ukp * RU, du += ukp * RU) {
          const uint64_t dsc_b = mk(du, lbo_b);
          const uint32_t accum = (it | ks) != 0;
          const uint64_t a0 = mk(xu + e0, lbo_a), a1 = mk(xu + e1, lbo_a), a2 = mk(xu + e2, lbo_a);
          if (leader) {
            if (BF16) { umma_f16(d0, a0, dsc_b, idesc, accum); umma_f16(d1, a1, dsc_b, idesc, accum); umma_f16(d2, a2, dsc_b, idesc, accum); }
            else { umma_tf32(d0, a0, dsc_b, idesc, accum); umma_tf32(d1, a1, dsc_b, idesc, accum); umma_tf32(d2, a2, dsc_b, idesc, accum); }
          }
        }
      } else if (nent <= 16) {
        // general tiling: per-entry A offsets held in registers; the entry count is made a compile-time
        // constant (switch below) so the MMAs of a k-step form one straight-line stream -- with a guarded
        // 16-way unroll the dependent IMAD/R2UR chain per MMA cost ~100 cycles (3x the hardware rate)
        auto run = [&](auto ne_c) {
          constexpr int NE = decltype(ne_c)::value;
#pragma unroll 2
          for (int ks = 0; ks < ksteps; ++ks, xu += ukp * RU, du += ukp * RU) {
            const uint64_t dsc_b = mk(du, lbo_b);
            const uint64_t dsc_b_lo = mk(du + lo_d, lbo_b);
            const uint32_t accum = (it | ks) != 0;
#pragma unroll
            for (int e = 0; e < NE; ++e) {
              const uint32_t au = xu + entr[e];
              const uint32_t d_tmem = tmem_base + e * ACCW;
              const uint64_t dsc_a = mk(au, lbo_a);
              if (leader) {
                if (BF16) { if constexpr (TWO) umma_f16_2cta(d_tmem, dsc_a, dsc_b, idesc, accum); else umma_f16(d_tmem, dsc_a, dsc_b, idesc, accum); }
                else {
                  if (STRICT && stack) {
                    umma_tf32(d_tmem, dsc_a, mk(du, lbo_b_stack), idesc2, accum);
                    umma_tf32(d_tmem + p.NT, mk(au + lo_x, lbo_a), dsc_b, idesc, 1);
                  } else {
                    umma_tf32(d_tmem, dsc_a, dsc_b, idesc, accum);
                    if (STRICT) {
                      umma_tf32(d_tmem + p.NT, dsc_a, dsc_b_lo, idesc, accum);
                      umma_tf32(d_tmem + p.NT, mk(au + lo_x, lbo_a), dsc_b, idesc, 1);
                    }
                  }
                }
              }
            }
          }
        };
        switch (nent) {
          case 1: run(std::integral_constant<int, 1>{}); break;
          case 2: run(std::integral_constant<int, 2>{}); break;
          case 3: run(std::integral_constant<int, 3>{}); break;
          case 4: run(std::integral_constant<int, 4>{}); break;
          case 5: run(std::integral_constant<int, 5>{}); break;
          case 6: run(std::integral_constant<int, 6>{}); break;
          case 9: run(std::integral_constant<int, 9>{}); break;
          case 10: run(std::integral_constant<int, 10>{}); break;
          default:   // entry counts the planner does not produce today: generic loop
            for (int ks = 0; ks < ksteps; ++ks, xu += ukp * RU, du += ukp * RU) {
              const uint64_t dsc_b = mk(du, lbo_b);
              const uint32_t accum = (it | ks) != 0;
              for (int e = 0; e < nent; ++e) {
                const uint32_t au = xu + ent[e];
                const uint64_t dsc_a = mk(au, lbo_a);
                if (leader) {
                  if (BF16) umma_f16(tmem_base + e * ACCW, dsc_a, dsc_b, idesc, accum);
                  else {
                    umma_tf32(tmem_base + e * ACCW, dsc_a, dsc_b, idesc, accum);
                    if (STRICT) {
                      umma_tf32(tmem_base + e * ACCW + p.NT, dsc_a, mk(du + lo_d, lbo_b), idesc, accum);
                      umma_tf32(tmem_base + e * ACCW + p.NT, mk(au + lo_x, lbo_a), dsc_b, idesc, 1);
                    }
                  }
                }
              }
            }
            break;
        }
      } else
      for (int ks = 0; ks < ksteps; ++ks, xu += ukp * RU, du += ukp * RU) {
        const uint64_t dsc_b = mk(du, lbo_b);
        const uint32_t accum = (it | ks) != 0;
        uint32_t d_tmem = tmem_base;
        for (int e = 0; e < nent; ++e, d_tmem += ACCW) {
          const uint32_t au = xu + ent[e];
          const uint64_t dsc_a = mk(au, lbo_a);
          if (leader) {
            if (BF16) umma_f16(d_tmem, dsc_a, dsc_b, idesc, accum);
            else {
              umma_tf32(d_tmem, dsc_a, dsc_b, idesc, accum);
              if (STRICT) {
                umma_tf32(d_tmem + p.NT, dsc_a, mk(du + lo_d, lbo_b), idesc, accum);
                umma_tf32(d_tmem + p.NT, mk(au + lo_x, lbo_a), dsc_b, idesc, 1);
              }
            }
          }
        }
      }
      if (committer) { if constexpr (TWO) umma_commit_2cta(&empty[s], 3); else umma_commit(&empty[s]); }
      if (it == 0 && lane == 0) tr.mark(3);
      __syncwarp();
    }
    if (committer) { if constexpr (TWO) umma_commit_2cta(acc_full, 3); else umma_commit(acc_full); }
    if (lane == 0) tr.mark(4);
  } else if (warp < 6) {
    // epilogue warps.  While the main loop runs they are otherwise idle, so tap group 0 uses them
    // to accumulate the bias gradient sum_q dz[q, o] from the dz strips already in shared memory
    // (no extra HBM traffic; junk / out-of-image positions are TMA zero fill).
    const int quarter = warp & 3;
    const int w4 = warp - 2;
    const int ACCW = STRICT ? 2 * p.NT : p.NT;
    if (do_bias) {
      // Column sums of the dz strips with 16-byte shared-memory loads: thread t owns one 16-byte unit (8 bf16 / 4 fp32
      // channels) of the operand rows r = k0, k0 + tpu, ... (a per-element loop with one dependent 2-byte load per
      // position made the bias CTAs 2.4x slower than the rest of the grid: ncu sm__cycles_active avg 0.5 of max).
      const int t = threadIdx.x - 64;
      const int upr = b_upr, units = b_n;
      const int tpu = 128 / units;                 // threads per unit (the last 128 - tpu*units threads idle)
      const int unit = b_lo + t % units, k0 = t / units;
      const int ck = unit / upr, u = unit % upr;
      const int ppr = p.RWB / p.PB;                // positions per operand row (2 in pixel-pair mode)
      const int nr = p.KT / ppr;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
      uint32_t it = 0, rs = 0, rph = 0;
      for (int tile = part; tile < p.total_tiles; tile += p.nparts, ++it) {
        const int q0 = (tile % p.tpi) * p.tstride;
        const int r0 = (q0 - (q0 / p.P) * p.P) / ppr;
        const uint32_t s = rs, ph = rph;
        if (++rs == (uint32_t)p.stages) { rs = 0; rph ^= 1; }
        mbar_wait_sleep_lean(&full[s], ph);
        const uint8_t* cb = smem + s * p.stage_stride + p.d_off + ck * p.d_chunk_stride;
#pragma unroll 4
        for (int r = k0 < tpu ? k0 : nr; r < nr; r += tpu) {
          uint32_t a = (uint32_t)(r0 + r) * p.RWB + (u << 4);
          a = BF16 ? swizzle_addr(a, p.RWB) : a ^ (((a >> 7) & 3u) << 5);
          const uint4 v = *reinterpret_cast<const uint4*>(cb + a);
          if (BF16 && F16) {
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
            const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&v.z)), f3 = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
            acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y;
            acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
          } else if (BF16) {
            acc[0] += __uint_as_float(v.x << 16); acc[1] += __uint_as_float(v.x & 0xFFFF0000u);
            acc[2] += __uint_as_float(v.y << 16); acc[3] += __uint_as_float(v.y & 0xFFFF0000u);
            acc[4] += __uint_as_float(v.z << 16); acc[5] += __uint_as_float(v.z & 0xFFFF0000u);
            acc[6] += __uint_as_float(v.w << 16); acc[7] += __uint_as_float(v.w & 0xFFFF0000u);
          } else {
            acc[0] += __uint_as_float(v.x); acc[1] += __uint_as_float(v.y);
            acc[2] += __uint_as_float(v.z); acc[3] += __uint_as_float(v.w);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      float* bs = reinterpret_cast<float*>(smem + p.bsum_off);   // [128 threads][8]
#pragma unroll
      for (int i = 0; i < 8; ++i) bs[t * 8 + i] = acc[i];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int epu = BF16 ? 8 : 4;                // channels per unit
      float* out = p.bias_partials + (size_t)layer * p.bias_layer_stride + (size_t)part * p.C;
      if (p.pair) {
        // operand row = [parity][16 channels]: unit = parity*(16/epu) + c/epu; fixed summation order (parity, then k)
        if (t < 16) {
          float sum = 0.0f;
          for (int par = 0; par < 2; ++par)
            for (int k = 0; k < tpu; ++k) sum += bs[(k * units + par * (16 / epu) + t / epu) * 8 + (t % epu)];
          out[t] = sum;
        }
      } else {
        for (int idx = t; idx < units * epu; idx += 128) {
          const int ul = idx / epu, jj = idx % epu, un = b_lo + ul;
          const int c = (un % upr) * epu + jj;
          const int ch = ngroup * p.NT + ((TWO ? (int)prank * p.dchunks : 0) + un / upr) * p.CH + c;
          if (c < p.CH && ch < p.C) {
            float sum = 0.0f;
            for (int k = 0; k < tpu; ++k) sum += bs[(k * units + ul) * 8 + jj];
            out[ch] = sum;
          }
        }
      }
    }
    if (threadIdx.x == 64) tr.mark(5);
    mbar_wait_sleep(acc_full, 0);
    if (threadIdx.x == 64) tr.mark(6);
    tc_fence_after_sync();
    const int Mrows = p.pair ? 128 : p.shift2 ? p.s2_M : p.trick ? 4 * p.CH : p.Mblk;
    const int nent = p.trick ? 3 : p.TG * p.MB;
    int m;  // accumulator row held by this thread's TMEM lane
    bool row_ok;
    if (Mrows == 128) { m = quarter * 32 + lane; row_ok = true; }
    else { m = quarter * 16 + lane; row_ok = lane < 16; }   // M=64: 16 lanes per quarter
    float* part_base = p.partials + (size_t)layer * p.part_layer_stride + (size_t)part * p.part_stride;
    if (p.pair) {
      // raw accumulators [alpha][lane (c,px,ci)][col (pd,o)]
      for (int e = 0; e < 3; ++e) {
        float* dst = part_base + ((size_t)e * 128 + m) * 32;
        for (int c0 = 0; c0 < 32; c0 += 16) {
          uint32_t r[16];
          tmem_ld_x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + e * ACCW + c0, r);
          if (STRICT) {      // main + correction accumulators
            uint32_t r2[16];
            tmem_ld_x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + e * ACCW + p.NT + c0, r2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
          } else {
            tmem_ld_wait();
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                    __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
      }
    } else if (p.shift2) {
      // accumulator e: rows (a, ci) <-> alpha = e * MA + a, columns (j, o) <-> beta = 2 - j
      const int MA = p.s2_M / p.CH, NB = p.s2_N / p.CH;
      const int a = m / p.CH, ci = m % p.CH;
      for (int e = 0; e < p.s2_nmma; ++e) {
        const int alpha = e * MA + a;
        const bool ok = row_ok && alpha < 3;
        for (int j = 0; j < (NB < 3 ? NB : 3); ++j) {
          float* dst = part_base + ((size_t)(alpha * 3 + (2 - j)) * p.C + ci) * p.C;
          for (int c0 = 0; c0 < p.CH; c0 += 16) {
            uint32_t r[16];
            tmem_ld_x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + e * p.s2_N + j * p.CH + c0, r);
            tmem_ld_wait();
            if (ok) {
#pragma unroll
              for (int jj = 0; jj < 16; jj += 4)
                *reinterpret_cast<float4*>(dst + c0 + jj) = make_float4(__uint_as_float(r[jj]), __uint_as_float(r[jj + 1]),
                                                                         __uint_as_float(r[jj + 2]), __uint_as_float(r[jj + 3]));
            }
          }
        }
      }
    } else
    for (int e = 0; e < nent; ++e) {
      int tap, ci;
      bool ok = row_ok;
      if (p.trick) { const int beta = m / p.CH; tap = e * 3 + beta; ci = m % p.CH; ok = ok && beta < 3; }
      else if (TWO && p.tappair) { tap = (2 * tapgroup + (int)prank) * p.TG + e; ci = m; }
      else { tap = tapgroup * p.TG + e / p.MB; ci = (e % p.MB) * p.Mblk + m + (p.mgroups > 1 ? mgroup * p.Mblk : 0); }
      ok = ok && ci < p.C && tap < 9;
      float* dst = part_base + ((size_t)tap * p.C + ci) * p.C + ngroup * p.NT;
      for (int c0 = 0; c0 < p.NT; c0 += 16) {
        uint32_t r[16];
        tmem_ld_x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + e * ACCW + c0, r);
        if (STRICT) {
          uint32_t r2[16];
          tmem_ld_x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + e * ACCW + p.NT + c0, r2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
        } else {
          tmem_ld_wait();
        }
        if (ok) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (ngroup * p.NT + c0 + j < p.C)
              *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                      __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
      }
    }
  } else if (STRICT) {
    const int ctid = threadIdx.x - 6 * 32;
    uint32_t it = 0;
    for (int tile = part; tile < p.total_tiles; tile += p.nparts, ++it) {
      const uint32_t s = it % p.stages, ph = (it / p.stages) & 1;
      mbar_wait(&full[s], ph);
      uint8_t* sb = smem + s * p.stage_stride;
      for (int which = 0; which < ((p.dbg & 4) ? 0 : 2); ++which) {      // dbg bit 2: no conversion (timing experiments only)
        const uint4* src = reinterpret_cast<const uint4*>(sb + (which ? p.d_off : p.x_off));
        uint4* dst = reinterpret_cast<uint4*>(sb + (which ? p.d_lo_off : p.x_lo_off));
        const int n16 = (which ? p.dchunks * p.d_chunk_stride : p.xchunks * p.x_chunk_stride) / 16;
        auto rem = [](uint4 u) {
          uint4 o;
          o.x = __float_as_uint(tf32_rna(__uint_as_float(u.x) - __uint_as_float(u.x & 0xFFFFE000u)));
          o.y = __float_as_uint(tf32_rna(__uint_as_float(u.y) - __uint_as_float(u.y & 0xFFFFE000u)));
          o.z = __float_as_uint(tf32_rna(__uint_as_float(u.z) - __uint_as_float(u.z & 0xFFFFE000u)));
          o.w = __float_as_uint(tf32_rna(__uint_as_float(u.w) - __uint_as_float(u.w & 0xFFFFE000u)));
          return o;
        };
        // eight independent 16-byte loads in flight per thread (the one-load-per-iteration loop made the converter warps the
        // bottleneck of the strict kernel: 685 us of the stage-1 launch's 1177 us with the MMAs switched off)
        int i = ctid;
        for (; i + 7 * 128 < n16; i += 8 * 128) {
          uint4 u[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) u[j] = src[i + j * 128];
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[i + j * 128] = rem(u[j]);
        }
        for (; i < n16; i += 128) dst[i] = rem(src[i]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&conv[s]);
    }
  }
  if (threadIdx.x == 64) tr.mark(7);
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (TWO) cluster_sync_all();     // the leader's last multicast commit has landed in the peer before either leaves
  if (warp == 1) { if constexpr (TWO) tmem_dealloc2(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols); }
  if (threadIdx.x == 0) { tr.mark(9); tr.wall(15); }
}

}  // namespace b200ode
