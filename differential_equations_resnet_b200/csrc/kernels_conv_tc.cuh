// K2/K3: 3x3 SAME stride-1 implicit-GEMM convolution on tcgen05 tensor cores with a fused
// Euler-step epilogue.  Used for the forward pass (x -> y = x + h*relu(conv_K(x)+b)) and, through
// the antisymmetry identity dX = dY - conv_K(dZ) + 2*gamma*dZ, for the data
// gradient with the SAME staged weights.
//
// Reference ops replaced: tf.nn.conv2d + bias (layers/tfkeras_layer_Conv2DAntisymmetric3By3.py:
// 159-169), Activation('relu') / Lambda(h*x) / add (models/tfkeras_resnets.py:89-92).
//
// Data flow per CTA (persistent, static round-robin over tiles):
//   * ONE TMA box load per K-block brings a halo strip of the NHWC input into shared memory:
//     rows of pitch P = W+1 pixels (column 0 = the zero column left of the image; the zero column
//     right of the image is column 0 of the next row), halo rows above/below zero-filled by TMA
//     out-of-bounds handling.  In this "padded linear" space every 3x3 tap is a constant row
//     offset (alpha*P + beta), so all nine A operands are the same strip addressed through UMMA
//     descriptors whose start address is shifted by whole rows (verified by csrc/umma_probe.cu).
//     The input is read from L2/HBM once per tile, not nine times.
//   * weights W[tap][o][ci] (K-major B operand, staged by pack_kernel) stream through a ring.
//   * tcgen05.mma accumulates 128-position segments into TMEM (double buffered across tiles).
//   * strict mode: converter warps write lo = x - trunc_tf32(x) next to each strip and every
//     (tap, k-step) issues hi*hi + hi*lo + lo*hi (3xTF32, fp32-level accuracy).
//   * epilogue warps: tcgen05.ld -> bias/relu/h/residual (+ relu bit mask, optional z) -> NHWC.
#pragma once

#include <cuda_bf16.h>

#include <type_traits>

#include "sm100_ptx.cuh"

namespace b200ode {

enum { MODE_STRICT = 0, MODE_TF32 = 1, MODE_BF16 = 2 };

struct ConvTcParams {
  int N, H, W;
  int P;            // padded row pitch in pixels (W + 1)
  int RB;           // halo rows per image in one strip (TMA box rows)
  int nimg;         // images per tile
  int spi;          // 128-position segments per image per tile
  int tpi;          // tiles per image (1 when nimg >= 1 covers whole images)
  int total_tiles;
  // shared memory plan (bytes from the 1024-aligned base)
  int sa, sw, tw;   // A stages, W stages, taps per W stage
  // general odd kernel size (GENK instantiations: Conv2DAntisymmetric with k = 5 / 7): k*k taps, `pad` = k/2 zero
  // columns left of every row (pitch P = W + pad; the right padding of a row is the left padding of the next one) and
  // `pad` halo rows above and below; one tap per weight ring stage
  int ksize, ntaps, pad;
  uint32_t a_bytes, a_lo_off, a_stride, w_bytes, w_stride, a_off, w_off, bar_off;
  uint32_t tmem_cols;
  // epilogue
  const void* in;      // conv input (centre term c_in * in[p]); nullable
  const void* skip;    // nullable
  void* out;           // nullable
  float* z_out;        // nullable, fp32 pre-activation
  uint8_t* mask;       // nullable
  const float* bias;   // nullable
  float acc_scale, c_in, h;
  int relu, scale_h;
  // fused backward tail (data-gradient launches): out2 = h2 * out * mask2, i.e. dZ_{l-1} = h * dY_{l-1} * relu mask_{l-1}
  // written in the same pass as dY_{l-1} (saves relu_scale_bwd_kernel's re-read of dY); computed from the value as
  // stored (bf16-rounded in bf16 mode) so the result is bit-identical to the two-kernel sequence
  const uint8_t* mask2;  // nullable [pixels][C/8]
  void* out2;            // nullable
  float h2;
  // BatchNorm statistics of the pre-activation from the epilogue registers (fp32 modes): every epilogue warp of
  // every CTA owns one row of per-channel partial sums, [2][gridDim.x * 8][C] floats (sum rows first, then the
  // sum-of-squares rows: the layout colsum_stage2 / bn_stats_finalize_kernel reduce in a fixed order -> deterministic)
  float* bn_part;        // nullable
  int* bn_rows_out;      // host-side only (run_conv_tc reports the number of partial rows = 8 * grid)
  uint64_t* trace;     // nullable timeline buffer (debug)
  // Thread-block cluster: the CTAs of a cluster walk their tiles in lockstep on the weight ring, so each
  // weight stage is fetched from L2 ONCE per cluster (TMA multicast by rank 0).  At C >= 128 every 256-position
  // tile re-streams all 9*C*C weights: unicast that is ~32 B/clk per SM, more than L2 delivers to 148 SMs.
  int cs;              // cluster size (1 = no cluster)
  int iters;           // tiles per CTA = ceil(total_tiles / gridDim.x); tile indices >= total_tiles are ghosts
};

template <int MODE, int C>
struct ConvTcCfg {
  static constexpr int EB = MODE == MODE_BF16 ? 2 : 4;            // operand element bytes
  static constexpr int ROWB = (C * EB >= 128) ? 128 : C * EB;      // bytes per pixel row in one K-block
  static constexpr int KB = ROWB / EB;                             // channels per K-block
  static constexpr int NKB = C / KB;
  static constexpr int KS = ROWB / 32;                             // 32-byte k-steps per K-block
  static constexpr bool STRICT = MODE == MODE_STRICT;
  // Epilogue warps per TMEM lane quarter.  At C = 256 (bf16) the two 256-column accumulators of a tile fill TMEM, so
  // the next tile's MMAs wait for the drain (~22k cycles per 256x256 tile against 37k cycles of MMAs; one 32x32
  // item takes a warp ~3000 cycles).  Measured at C = 256: four warps per quarter on 16-channel items are SLOWER
  // (348 vs 318 us); staggering odd CTAs by half a tile, or removing the drain's loads, stores or TMEM reads: no
  // change.  The drain is bound by the instruction throughput of its ~20 integer/convert/compare operations per
  // element (bf16 unpack, relu, mask bits, pack), not by memory.  Left at two.
  static constexpr int EPQ = 2;
  static constexpr int G = EPQ == 4 ? 16 : (C < 32 ? C : 32);      // channels per epilogue work item
  static constexpr int NWARPS = STRICT ? 14 : 2 + 4 * EPQ;         // TMA, MMA, epilogue (+4 converter)
  // taps per weight ring stage (same rule as taps_per_w_stage() on the host)
  static constexpr int PER_TAP = C * ROWB * (STRICT ? 2 : 1);
  static constexpr int TW = PER_TAP * 9 <= 40 * 1024 ? 9 : PER_TAP * 3 <= 56 * 1024 ? 3 : 1;
};

__device__ __forceinline__ float ld_as_float(const float* p) { return *p; }

// Column sums over the 32 lanes of a warp for G per-lane values (one pixel row per lane, G channels): recursive halving,
// G - 1 shuffles instead of 5 * G.  On return v[0] holds the sum over all lanes of channel `lane` (G == 32) or
// `lane >> 1` (G == 16, both lanes of a pair hold it).
template <int G>
__device__ __forceinline__ void warp_column_sums(float (&v)[G], int lane) {
  static_assert(G == 32 || G == 16, "32 or 16 channels per item");
  constexpr int OFF0 = 16;
#pragma unroll
  for (int n = G / 2, off = OFF0; n >= 1; n >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < n; ++j) {
      const float send = upper ? v[j] : v[j + n];
      const float keep = upper ? v[j + n] : v[j];
      v[j] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, off);
    }
  }
  if (G == 16) v[0] += __shfl_xor_sync(0xFFFFFFFFu, v[0], 1);
}

// BN = true: the variant that also emits the BatchNorm partial sums (ConvTcParams::bn_part); a separate instantiation so
// that the extra live registers do not touch the plain kernels (inline, they pushed the tf32 kernels into spills).
// TWO = true (bf16 / tf32, C >= 128): CTA PAIRS (cta_group::2).  A single-CTA M = 128 MMA reads its A tile (128 x 32 B) and the
// whole B tile (C x 32 B) from ONE SM's shared memory at 128 B/clk: 64 cycles at N = 128 and 96 at N = 256 against 32 / 64
// cycles of tensor math -- the 50 % / 67 % at which the single-CTA kernels sit.  A pair issues ONE M = 256 MMA over the
// two CTAs' A tiles with B split between their shared memories (C/2 rows each): 48 / 64 cycles of operand reads per SM.
// The two CTAs of a pair (cluster of 2) work on the same tile-in-image of two consecutive images, so one descriptor
// (same shared-memory offsets) addresses both A strips; each loads its own strip and its half of every weight stage with
// the complete_tx routed to the LEADER's barriers; the leader's MMA warp issues for both and its commits are multicast;
// each CTA drains its own TMEM half and releases the accumulator on the leader's barrier.
template <int MODE, int C, bool BN = false, bool GENK = false, bool TWO = false>
__global__ void __launch_bounds__(ConvTcCfg<MODE, C>::NWARPS * 32, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_w_lo, const ConvTcParams p) {
  static_assert(!TWO || (MODE != MODE_STRICT && !BN && !GENK && C >= 128), "CTA pairs: bf16 / tf32, C >= 128, k = 3");
  using Cfg = ConvTcCfg<MODE, C>;
  constexpr int ROWB = Cfg::ROWB, KB = Cfg::KB, NKB = Cfg::NKB, KS = Cfg::KS;
  constexpr bool STRICT = Cfg::STRICT;
  constexpr uint32_t LT = ROWB == 128 ? SWZ_128B : ROWB == 64 ? SWZ_64B : SWZ_32B;
  constexpr uint32_t SBO = 8 * ROWB;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* a_full = bars;                 // [sa]
  uint64_t* a_empty = a_full + p.sa;       // [sa]
  uint64_t* a_conv = a_empty + p.sa;       // [sa] (strict)
  uint64_t* w_full = a_conv + p.sa;        // [sw]
  uint64_t* w_empty = w_full + p.sw;       // [sw]
  uint64_t* w_empty_cl = w_empty + p.sw;   // [sw] rank 0 only: the other CTAs of the cluster released the stage
  uint64_t* acc_full = w_empty_cl + p.sw;  // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  Trace tr;
  tr.begin(p.trace);
  if (threadIdx.x == 0) tr.wall(0);
  const int mt = p.nimg * p.spi;           // segments (accumulators) per tile
  const int T = p.spi * 128;               // positions per image per tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    if (STRICT) tma_prefetch_desc(&map_w_lo);
    for (int i = 0; i < p.sa; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&a_conv[i], 4); }
    for (int i = 0; i < p.sw; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); mbar_init(&w_empty_cl[i], p.cs > 1 ? p.cs - 1 : 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], (TWO ? 2 : 1) * 4 * Cfg::EPQ); }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (TWO) { tmem_alloc2(tmem_slot, p.tmem_cols); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (p.cs > 1) cluster_sync_all();   // every CTA's barriers exist before any multicast / remote arrive
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) tr.mark(1);
  // strict mode keeps two accumulators per segment: the hi*hi sum and the small hi*lo + lo*hi
  // correction sum.  The tensor core truncates when it accumulates, so the error grows with the
  // number of accumulation steps into one register; splitting keeps the 2/3 of the MMAs that carry
  // ~2^-11-sized terms away from the main sum (they are added once, in fp32, in the epilogue).
  constexpr int ACCW = STRICT ? 2 * C : C;   // TMEM columns per segment
  const int acc_stages = (2 * mt * ACCW <= 512) ? 2 : 1;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t as_ = 0, aph = 0, ws = 0, wph = 0;   // ring positions kept incrementally (no integer division per entry)
      uint32_t iw = 0;
      const uint32_t crank = p.cs > 1 ? cluster_ctarank() : 0u;
      const uint16_t cmask = (uint16_t)((1u << p.cs) - 1u);
      if constexpr (TWO) {
        // CTA pair: own A strip + own half (C/2 output channels) of every weight stage, complete_tx on the leader's barriers
        for (int itile = 0; itile < p.iters; ++itile) {
          const int pt = (int)(blockIdx.x >> 1) + itile * (int)(gridDim.x >> 1);   // >= total_tiles: ghost pair
          const int n0 = 2 * (pt / p.tpi) + (int)crank;
          const int row0 = ((pt % p.tpi) * T) / p.P;
          for (int kb = 0; kb < NKB; ++kb) {
            const uint32_t s = as_, ph = aph;
            mbar_wait_sleep_lean(&a_empty[s], ph ^ 1);
            if (crank == 0) mbar_expect_tx(&a_full[s], 2 * p.a_bytes);
            tma_load_4d_pair(smem + p.a_off + s * p.a_stride, &map_a, mapa_u32(smem_u32(&a_full[s]), 0), kb * KB, -1, row0 - 1, n0);
            if (++as_ == (uint32_t)p.sa) { as_ = 0; aph ^= 1; }
            for (int tg = 0; tg < 9; tg += p.tw) {
              const uint32_t sw_ = ws, phw = wph;
              mbar_wait_sleep_lean(&w_empty[sw_], phw ^ 1);
              if (crank == 0) mbar_expect_tx(&w_full[sw_], 2 * p.w_bytes);
              tma_load_3d_pair(smem + p.w_off + sw_ * p.w_stride, &map_w, mapa_u32(smem_u32(&w_full[sw_]), 0), kb * KB,
                               (int)crank * (C / 2), tg);
              if (++ws == (uint32_t)p.sw) { ws = 0; wph ^= 1; }
            }
          }
        }
      } else
      for (int itile = 0; itile < p.iters; ++itile) {
        const int tile = blockIdx.x + itile * gridDim.x;   // >= total_tiles: ghost (TMA zero-fills the out-of-bounds strip)
        const int n0 = (tile / p.tpi) * p.nimg;
        const int q0 = (tile % p.tpi) * T;
        const int row0 = q0 / p.P;  // first halo row of the strip (halo row r <-> image row r-1)
        for (int kb = 0; kb < NKB; ++kb) {
          const uint32_t s = as_, ph = aph;
          mbar_wait_sleep_lean(&a_empty[s], ph ^ 1);
          mbar_expect_tx(&a_full[s], p.a_bytes);
          const int hp = GENK ? p.pad : 1;
          tma_load_4d(smem + p.a_off + s * p.a_stride, &map_a, &a_full[s], kb * KB, -hp, row0 - hp, n0);
          if (++as_ == (uint32_t)p.sa) { as_ = 0; aph ^= 1; }
          for (int tg = 0; tg < (GENK ? p.ntaps : 9); tg += p.tw) {
            const uint32_t sw_ = ws, phw = wph;
            mbar_wait_sleep_lean(&w_empty[sw_], phw ^ 1);     // this CTA's MMAs are done with the stage's previous contents
            uint8_t* wdst = smem + p.w_off + sw_ * p.w_stride;
            if (p.cs == 1) {
              mbar_expect_tx(&w_full[sw_], STRICT ? 2 * p.w_bytes : p.w_bytes);
              tma_load_3d(wdst, &map_w, &w_full[sw_], kb * KB, 0, tg);
              if (STRICT) tma_load_3d(wdst + p.w_bytes, &map_w_lo, &w_full[sw_], kb * KB, 0, tg);
            } else if (crank != 0) {
              if (iw >= (uint32_t)p.sw) mbar_arrive_cluster(&w_empty_cl[sw_], 0);   // tell rank 0 the stage is free here
              mbar_expect_tx(&w_full[sw_], STRICT ? 2 * p.w_bytes : p.w_bytes);     // rank 0's multicast completes it
            } else {
              if (iw >= (uint32_t)p.sw) mbar_wait_sleep(&w_empty_cl[sw_], phw ^ 1);
              mbar_expect_tx(&w_full[sw_], STRICT ? 2 * p.w_bytes : p.w_bytes);
              tma_load_3d_mc(wdst, &map_w, &w_full[sw_], kb * KB, 0, tg, cmask);
              if (STRICT) tma_load_3d_mc(wdst + p.w_bytes, &map_w_lo, &w_full[sw_], kb * KB, 0, tg, cmask);
            }
            ++iw;
            if (++ws == (uint32_t)p.sw) { ws = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the (uniform) control flow so descriptors live in uniform registers; one
    // elected lane issues the tcgen05 instructions.  Descriptor low words are advanced by adds of
    // precomputed 16-byte-unit offsets: a handful of integer instructions per MMA.
    const bool leader = elect_one() && (!TWO || cluster_ctarank() == 0);   // CTA pair: the leader CTA issues for both
    const uint32_t idesc = make_instr_desc(MODE == MODE_BF16 ? FMT_BF16 : FMT_TF32, TWO ? 256 : 128, C, 0, 0);
    const uint32_t desc_hi32 = (SBO >> 4) | (1u << 14) | (LT << 29);   // bits 32..63 of the smem descriptor
    constexpr uint32_t LBO_FIELD = 1u << 16;
    constexpr uint32_t RU = ROWB >> 4;                                  // 16-byte units per pixel row
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t seg_img_step = (uint32_t)(p.RB * p.P) * RU;          // next image inside the strip
    const uint32_t a_lo_units = p.a_lo_off >> 4, w_lo_units = p.w_bytes >> 4;
    const uint32_t tap_units = (uint32_t)((TWO ? C / 2 : C) * ROWB) >> 4;   // one tap's weight tile (CTA pair: this CTA's half)
    auto mk = [&](uint32_t lo) -> uint64_t { return (static_cast<uint64_t>(desc_hi32) << 32) | (lo | LBO_FIELD); };
    auto toff_of = [&](int t) -> uint32_t {
      return GENK ? (uint32_t)((t / p.ksize) * p.P + (t % p.ksize)) * RU : (uint32_t)((t / 3) * p.P + (t % 3)) * RU;
    };
    uint32_t as_ = 0, aph_ = 0, ws = 0, wph = 0, it = 0;
    if (TWO && cluster_ctarank() != 0) {
      // peer CTA of a pair: its operands are consumed by the leader's MMAs; nothing to issue
    } else
    for (int itile = 0; itile < p.iters; ++itile, ++it) {
      const int tile = TWO ? (int)(blockIdx.x >> 1) + itile * (int)(gridDim.x >> 1) : (int)(blockIdx.x + itile * gridDim.x);
      const int q0 = (tile % p.tpi) * T;
      const uint32_t off0_units = (uint32_t)(q0 - (q0 / p.P) * p.P) * RU;
      const uint32_t as = acc_stages == 2 ? (it & 1) : 0u, aph = acc_stages == 2 ? ((it >> 1) & 1) : (it & 1);
      mbar_wait_lean(&acc_empty[as], aph ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tile = tmem_base + as * mt * ACCW;
      for (int kb = 0; kb < NKB; ++kb) {
        const uint32_t s = as_, ph = aph_;
        mbar_wait_lean(STRICT ? &a_conv[s] : &a_full[s], ph);
        if (it == 0 && kb == 0 && lane == 0) tr.mark(2);
        const uint32_t a_units = ((smem_base + p.a_off + s * p.a_stride) >> 4) + off0_units;
        // Taps-per-stage and k-steps are compile-time: per segment the stage's MMAs form one straight-line
        // stream (runtime tap loops left a dependent IMAD/R2UR chain of ~130 cycles per MMA, twice the
        // hardware time of an N = 128 MMA).  Per accumulator the order (kb, tap, k-step) is unchanged.
        constexpr int TW = GENK ? 1 : Cfg::TW;
#pragma unroll 1
        for (int tg = 0; tg < (GENK ? p.ntaps : 9); tg += TW) {
          const uint32_t sw_ = ws, phw = wph;
          mbar_wait_lean(&w_full[sw_], phw);
          if (it == 0 && kb == 0 && tg == 0 && lane == 0) tr.mark(3);
          tc_fence_after_sync();
          const uint32_t b_base = (smem_base + p.w_off + sw_ * p.w_stride) >> 4;
          uint32_t a_img = a_units, d_seg = d_tile;
          for (int im = 0; im < p.nimg; ++im, a_img += seg_img_step) {
            uint32_t a_sg = a_img;
            for (int j = 0; j < p.spi; ++j, a_sg += 128 * RU, d_seg += ACCW) {
#pragma unroll
              for (int tt = 0; tt < TW; ++tt) {
                const uint32_t a_tap = a_sg + toff_of(tg + tt);
                const uint32_t b_units = b_base + tt * tap_units;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                  const uint64_t da = mk(a_tap + 2 * ks), db = mk(b_units + 2 * ks);
                  const uint32_t acc = (tg | tt | ks) ? 1u : (kb ? 1u : 0u);
                  if (leader) {
                    if (MODE == MODE_BF16) {
                      if constexpr (TWO) umma_f16_2cta(d_seg, da, db, idesc, acc);
                      else umma_f16(d_seg, da, db, idesc, acc);
                    } else {
                      if constexpr (TWO) umma_tf32_2cta(d_seg, da, db, idesc, acc);
                      else umma_tf32(d_seg, da, db, idesc, acc);
                      if (STRICT) {
                        umma_tf32(d_seg + C, da, mk(b_units + w_lo_units + 2 * ks), idesc, acc);
                        umma_tf32(d_seg + C, mk(a_tap + a_lo_units + 2 * ks), db, idesc, 1);
                      }
                    }
                  }
                }
              }
            }
          }
          if (leader) { if constexpr (TWO) umma_commit_2cta(&w_empty[sw_], 3); else umma_commit(&w_empty[sw_]); }
          if (++ws == (uint32_t)p.sw) { ws = 0; wph ^= 1; }
        }
        if (leader) { if constexpr (TWO) umma_commit_2cta(&a_empty[s], 3); else umma_commit(&a_empty[s]); }
        if (++as_ == (uint32_t)p.sa) { as_ = 0; aph_ ^= 1; }
      }
      if (leader) { if constexpr (TWO) umma_commit_2cta(&acc_full[as], 3); else umma_commit(&acc_full[as]); }
      if (it == 0 && lane == 0) tr.mark(4);
      __syncwarp();
    }
    if (lane == 0) tr.mark(5);
  } else if (warp < 2 + 4 * Cfg::EPQ) {
    // ===================== epilogue warps 2..9: two warps per TMEM lane quarter =====================
    // Work item = (segment, G-channel group); the items of a tile alternate between the two warps of a
    // quarter.  Every global load of an item (residual input, skip) is issued before the TMEM wait.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    constexpr int G = Cfg::G;                 // channels per work item
    constexpr int NGR = C / G;
    uint32_t it = 0;
    using IoT = typename std::conditional<MODE == MODE_BF16, __nv_bfloat16, float>::type;
    const IoT* in = reinterpret_cast<const IoT*>(p.in);
    const IoT* skip = reinterpret_cast<const IoT*>(p.skip);
    IoT* out = reinterpret_cast<IoT*>(p.out);
    float bn_s[BN ? NGR : 1], bn_q[BN ? NGR : 1];   // BatchNorm partial sums of this warp (p.bn_part), channel = f(lane)
#pragma unroll
    for (int j = 0; j < (BN ? NGR : 1); ++j) bn_s[j] = bn_q[j] = 0.0f;
    const int pair_rank = TWO ? (int)cluster_ctarank() : 0;
    for (int itile = 0; itile < p.iters; ++itile, ++it) {
      // ghost tiles: n0 >= N, nothing is stored.  CTA pair: the same tile-in-image of images 2m (leader) and 2m + 1 (peer)
      const int tile = TWO ? (int)(blockIdx.x >> 1) + itile * (int)(gridDim.x >> 1) : (int)(blockIdx.x + itile * gridDim.x);
      const int n0 = TWO ? 2 * (tile / p.tpi) + pair_rank : (tile / p.tpi) * p.nimg;
      const int q0 = (tile % p.tpi) * T;
      const uint32_t as = acc_stages == 2 ? (it & 1) : 0u, aph = acc_stages == 2 ? ((it >> 1) & 1) : (it & 1);
      if constexpr (MODE == MODE_BF16 && NGR >= 2) {
        // bf16, C >= 64: software-pipelined items.  The epilogue is bound by global-load latency x bytes in flight
        // (ncu: the epilogue warps' samples sit on the first use of the residual loads; 16 warps on half-size items
        // changed nothing), and at C = 256 it cannot overlap the next tile's MMAs (two 256-column accumulators fill
        // TMEM).  So: the raw operands of item k+1 are requested before item k is processed, and those of the
        // tile's first item before the wait for the accumulators.
        constexpr int IPS = NGR / Cfg::EPQ;   // items per segment for this warp: channel groups gi = EPQ*j + half
        constexpr int UV = G / 8;             // 16-byte vectors per operand per item
        const int nit = mt * IPS;
        const bool has_in = in != nullptr, has_skip = skip != nullptr, has_bias = p.bias != nullptr;
        auto geom = [&](int k, bool& valid, long long& pix, int& c0, int& sg) {
          sg = k / IPS;
          c0 = (Cfg::EPQ * (k - sg * IPS) + half) * G;
          const int n = n0 + sg / p.spi;
          const int q = q0 + (sg % p.spi) * 128 + row;
          const int y = q / p.P, xq = q - y * p.P;
          valid = (n < p.N) && (y < p.H) && (xq < p.W);
          pix = ((long long)n * p.H + y) * p.W + xq;
        };
        auto request = [&](bool valid, long long pix, int c0, uint4 (&a)[UV], uint4 (&b)[UV]) {
          if (valid) {
            if (has_in) {
              const uint4* ip = reinterpret_cast<const uint4*>(in + pix * C + c0);
#pragma unroll
              for (int j = 0; j < UV; ++j) a[j] = ip[j];
            }
            if (has_skip) {
              const uint4* sp = reinterpret_cast<const uint4*>(skip + pix * C + c0);
#pragma unroll
              for (int j = 0; j < UV; ++j) b[j] = sp[j];
            }
          }
        };
        uint4 cin[UV], csk[UV], nin[UV], nsk[UV];
#pragma unroll
        for (int j = 0; j < UV; ++j) cin[j] = csk[j] = nin[j] = nsk[j] = make_uint4(0u, 0u, 0u, 0u);
        // The per-element arithmetic is branch-free: the optional stages are folded into operands that make them
        // identities (relu floor -inf, scale 1, centre coefficient 0, zeroed operand registers).  With `if (flag)` per
        // element the compiler emitted ~2300 instructions per 32x32 item and the drain of a tile took 22k cycles of
        // pure ALU work (measured with every memory and TMEM access disabled).
        const float relu_floor = p.relu ? 0.0f : -__int_as_float(0x7f800000);
        const float hs = p.scale_h ? p.h : 1.0f;
        const float cin_c = has_in ? p.c_in : 0.0f;
        const bool want2 = p.out2 != nullptr;
        bool valid, nvalid = false; long long pix, npix = 0; int c0, nc0 = 0, sg, nsg = 0;
        geom(0, valid, pix, c0, sg);
        request(valid, pix, c0, cin, csk);
        mbar_wait_sleep_lean(&acc_full[as], aph);
        if (it == 0 && threadIdx.x == 64) tr.mark(6);
        tc_fence_after_sync();
#pragma unroll 1
        for (int k = 0; k < nit; ++k) {
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + (as * mt + sg) * ACCW + c0;
          uint32_t r[G];
          {
            uint32_t t16[16];
            tmem_ld_x16(taddr, t16);
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = t16[j];
            if (G == 32) {
              uint32_t u16[16];
              tmem_ld_x16(taddr + 16, u16);
#pragma unroll
              for (int j = 0; j < 16; ++j) r[16 + (j & 15)] = u16[j];
            }
          }
          if (k + 1 < nit) {
            geom(k + 1, nvalid, npix, nc0, nsg);
            request(nvalid, npix, nc0, nin, nsk);
          }
          float bs[G];
#pragma unroll
          for (int j = 0; j < G / 4; ++j) {
            float4 b4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (has_bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0) + j);
            bs[4 * j] = b4.x; bs[4 * j + 1] = b4.y; bs[4 * j + 2] = b4.z; bs[4 * j + 3] = b4.w;
          }
          uint32_t bits2 = 0;
          if (want2 && valid) {
            if (G == 32) bits2 = *reinterpret_cast<const uint32_t*>(p.mask2 + pix * (C / 8) + c0 / 8);
            else bits2 = *reinterpret_cast<const uint16_t*>(p.mask2 + pix * (C / 8) + c0 / 8);
          }
          tmem_ld_wait();
          if (valid) {
            float v[G];
#pragma unroll
            for (int j = 0; j < G; ++j) v[j] = fmaf(p.acc_scale, __uint_as_float(r[j]), bs[j]);
            if (p.z_out) {
              float4* zp = reinterpret_cast<float4*>(p.z_out + pix * C + c0);
#pragma unroll
              for (int j = 0; j < G / 4; ++j) zp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            if (p.mask) {
              uint32_t bits = 0;
#pragma unroll
              for (int j = 0; j < G; ++j) bits |= (v[j] > 0.0f ? 1u : 0u) << j;
              if (G == 32) *reinterpret_cast<uint32_t*>(p.mask + pix * (C / 8) + c0 / 8) = bits;
              else *reinterpret_cast<uint16_t*>(p.mask + pix * (C / 8) + c0 / 8) = static_cast<uint16_t>(bits);
            }
            if (out) {
              uint32_t pk[G / 2];
#pragma unroll
              for (int j = 0; j < UV; ++j) {
                const uint32_t wi[4] = {cin[j].x, cin[j].y, cin[j].z, cin[j].w};
                const uint32_t ws_[4] = {csk[j].x, csk[j].y, csk[j].z, csk[j].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int j0 = 8 * j + 2 * e;
                  // relu -> Lambda(h*x) (own rounding) -> + c_in * x -> + skip; disabled stages are identities
                  float a0 = __fmul_rn(hs, fmaxf(v[j0], relu_floor)), a1 = __fmul_rn(hs, fmaxf(v[j0 + 1], relu_floor));
                  a0 = fmaf(cin_c, __uint_as_float(wi[e] << 16), a0) + __uint_as_float(ws_[e] << 16);
                  a1 = fmaf(cin_c, __uint_as_float(wi[e] & 0xFFFF0000u), a1) + __uint_as_float(ws_[e] & 0xFFFF0000u);
                  __nv_bfloat162 b2 = __floats2bfloat162_rn(a0, a1);
                  pk[4 * j + e] = *reinterpret_cast<uint32_t*>(&b2);
                }
              }
              uint4* op = reinterpret_cast<uint4*>(out + pix * C + c0);
#pragma unroll
              for (int j = 0; j < UV; ++j) op[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
              if (want2) {   // dZ of the step below from the value as stored (see ConvTcParams::out2)
                uint4* op2 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + pix * C + c0);
#pragma unroll
                for (int j = 0; j < UV; ++j) {
                  uint32_t pk2[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int j0 = 8 * j + 2 * e;
                    const uint32_t w = pk[4 * j + e];
                    const float s0 = (bits2 >> j0) & 1u ? p.h2 * __uint_as_float(w << 16) : 0.0f;
                    const float s1 = (bits2 >> (j0 + 1)) & 1u ? p.h2 * __uint_as_float(w & 0xFFFF0000u) : 0.0f;
                    __nv_bfloat162 z2 = __floats2bfloat162_rn(s0, s1);
                    pk2[e] = *reinterpret_cast<uint32_t*>(&z2);
                  }
                  op2[j] = make_uint4(pk2[0], pk2[1], pk2[2], pk2[3]);
                }
              }
            }
          }
          valid = nvalid; pix = npix; c0 = nc0; sg = nsg;
#pragma unroll
          for (int j = 0; j < UV; ++j) { cin[j] = nin[j]; csk[j] = nsk[j]; }
        }
      } else {
      mbar_wait_sleep_lean(&acc_full[as], aph);
      if (it == 0 && threadIdx.x == 64) tr.mark(6);
      tc_fence_after_sync();
      for (int sg = 0; sg < mt; ++sg) {
        const int n = n0 + sg / p.spi;
        const int q = q0 + (sg % p.spi) * 128 + row;
        const int y = q / p.P;
        const int xq = q - y * p.P;
        const bool valid = (n < p.N) && (y < p.H) && (xq < p.W);
        const long long pix = ((long long)n * p.H + y) * p.W + xq;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + (as * mt + sg) * ACCW;
#pragma unroll
        for (int gi = 0; gi < NGR; ++gi) {
          if (((sg * NGR + gi) & (Cfg::EPQ - 1)) != half) continue;
          const int c0 = gi * G;
          uint32_t r[G];
          {
            uint32_t t16[16];
            tmem_ld_x16(taddr + c0, t16);
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = t16[j];
            if (G == 32) {
              uint32_t u16[16];
              tmem_ld_x16(taddr + c0 + 16, u16);
#pragma unroll
              for (int j = 0; j < 16; ++j) r[16 + (j & 15)] = u16[j];
            }
          }
          uint32_t r2[STRICT ? G : 1];
          if (STRICT) {
            uint32_t t16[16];
            tmem_ld_x16(taddr + C + c0, t16);
#pragma unroll
            for (int j = 0; j < 16; ++j) r2[j] = t16[j];
            if (G == 32) {
              uint32_t u16[16];
              tmem_ld_x16(taddr + C + c0 + 16, u16);
#pragma unroll
              for (int j = 0; j < 16; ++j) r2[16 + (j & 15)] = u16[j];
            }
          }
          // operands from global memory, all in flight before the TMEM wait
          float xin[G], xsk[G], bs[G];
          const bool has_in = in != nullptr, has_skip = skip != nullptr, has_bias = p.bias != nullptr;
          if (valid) {
            if (MODE == MODE_BF16) {
              if (has_in) {
                const uint4* ip = reinterpret_cast<const uint4*>(in + pix * C + c0);
#pragma unroll
                for (int j = 0; j < G / 8; ++j) {
                  const uint4 u = ip[j];
                  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    xin[8 * j + 2 * e] = __uint_as_float(w[e] << 16);
                    xin[8 * j + 2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
                  }
                }
              }
              if (has_skip) {
                const uint4* sp = reinterpret_cast<const uint4*>(skip + pix * C + c0);
#pragma unroll
                for (int j = 0; j < G / 8; ++j) {
                  const uint4 u = sp[j];
                  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    xsk[8 * j + 2 * e] = __uint_as_float(w[e] << 16);
                    xsk[8 * j + 2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
                  }
                }
              }
            } else {
              if (has_in) {
                const float4* ip = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) + pix * C + c0);
#pragma unroll
                for (int j = 0; j < G / 4; ++j) {
                  const float4 u = ip[j];
                  xin[4 * j] = u.x; xin[4 * j + 1] = u.y; xin[4 * j + 2] = u.z; xin[4 * j + 3] = u.w;
                }
              }
              if (has_skip) {
                const float4* sp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(skip) + pix * C + c0);
#pragma unroll
                for (int j = 0; j < G / 4; ++j) {
                  const float4 u = sp[j];
                  xsk[4 * j] = u.x; xsk[4 * j + 1] = u.y; xsk[4 * j + 2] = u.z; xsk[4 * j + 3] = u.w;
                }
              }
            }
          }
          if (has_bias) {
#pragma unroll
            for (int j = 0; j < G / 4; ++j) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0) + j);
              bs[4 * j] = b4.x; bs[4 * j + 1] = b4.y; bs[4 * j + 2] = b4.z; bs[4 * j + 3] = b4.w;
            }
          }
          tmem_ld_wait();
          if constexpr (BN && MODE != MODE_BF16) {
            float s1[G], s2[G];
#pragma unroll
            for (int j = 0; j < G; ++j) {
              float a = __uint_as_float(r[j]);
              if (STRICT) a += __uint_as_float(r2[j]);
              a = p.acc_scale * a;
              if (has_bias) a += bs[j];
              a = valid ? a : 0.0f;                           // junk rows of the strip (halo / padding positions)
              s1[j] = a; s2[j] = a * a;
            }
            warp_column_sums<G>(s1, lane);
            warp_column_sums<G>(s2, lane);
            bn_s[gi] += s1[0]; bn_q[gi] += s2[0];
          }
          if (valid) {
            float v[G];
#pragma unroll
            for (int j = 0; j < G; ++j) {
              float a = __uint_as_float(r[j]);
              if (STRICT) a += __uint_as_float(r2[j]);
              v[j] = p.acc_scale * a;
              if (has_bias) v[j] += bs[j];
            }
            if (p.z_out) {
              float4* zp = reinterpret_cast<float4*>(p.z_out + pix * C + c0);
#pragma unroll
              for (int j = 0; j < G / 4; ++j) zp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            if (p.mask) {
              uint32_t bits = 0;
#pragma unroll
              for (int j = 0; j < G; ++j) bits |= (v[j] > 0.0f ? 1u : 0u) << j;
              if (G == 32) *reinterpret_cast<uint32_t*>(p.mask + pix * (C / 8) + c0 / 8) = bits;
              else *reinterpret_cast<uint16_t*>(p.mask + pix * (C / 8) + c0 / 8) = static_cast<uint16_t>(bits);
            }
            if (out) {
#pragma unroll
              for (int j = 0; j < G; ++j) {
                if (p.relu) v[j] = fmaxf(v[j], 0.0f);
                if (p.scale_h) v[j] = __fmul_rn(p.h, v[j]);      // Lambda(h*x): own rounding (two layers in the reference)
                // c_in == 1 in the forward pass: plain add of x (the reference's add layer)
                if (has_in) v[j] = fmaf(p.c_in, xin[j], v[j]);
                if (has_skip) v[j] += xsk[j];
              }
              if (MODE == MODE_BF16) {
                uint4* op = reinterpret_cast<uint4*>(out + pix * C + c0);
#pragma unroll
                for (int j = 0; j < G / 8; ++j) {
                  uint32_t pk[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    __nv_bfloat162 b2 = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
                    pk[e] = *reinterpret_cast<uint32_t*>(&b2);
                  }
                  op[j] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
              } else {
                float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + pix * C + c0);
#pragma unroll
                for (int j = 0; j < G / 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              }
              if (p.out2) {
                uint32_t bits2;
                if (G == 32) bits2 = *reinterpret_cast<const uint32_t*>(p.mask2 + pix * (C / 8) + c0 / 8);
                else bits2 = *reinterpret_cast<const uint16_t*>(p.mask2 + pix * (C / 8) + c0 / 8);
                if (MODE == MODE_BF16) {
                  uint4* op2 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + pix * C + c0);
#pragma unroll
                  for (int j = 0; j < G / 8; ++j) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const int j0 = 8 * j + 2 * e;
                      const float a0 = __bfloat162float(__float2bfloat16_rn(v[j0])), a1 = __bfloat162float(__float2bfloat16_rn(v[j0 + 1]));
                      __nv_bfloat162 b2 = __floats2bfloat162_rn((bits2 >> j0) & 1u ? p.h2 * a0 : 0.0f,
                                                                (bits2 >> (j0 + 1)) & 1u ? p.h2 * a1 : 0.0f);
                      pk[e] = *reinterpret_cast<uint32_t*>(&b2);
                    }
                    op2[j] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                  }
                } else {
                  float4* op2 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out2) + pix * C + c0);
#pragma unroll
                  for (int j = 0; j < G / 4; ++j)
                    op2[j] = make_float4((bits2 >> (4 * j)) & 1u ? p.h2 * v[4 * j] : 0.0f, (bits2 >> (4 * j + 1)) & 1u ? p.h2 * v[4 * j + 1] : 0.0f,
                                         (bits2 >> (4 * j + 2)) & 1u ? p.h2 * v[4 * j + 2] : 0.0f, (bits2 >> (4 * j + 3)) & 1u ? p.h2 * v[4 * j + 3] : 0.0f);
                }
              }
            }
          }
        }
      }
      }   // per-item path (fp32 modes, C < 64)
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) { if constexpr (TWO) mbar_arrive_cluster(&acc_empty[as], 0); else mbar_arrive(&acc_empty[as]); }
      if (it == 0 && threadIdx.x == 64) tr.mark(7);
    }
    if constexpr (BN && MODE != MODE_BF16) {
      const size_t rows = (size_t)gridDim.x * (4 * Cfg::EPQ), rowi = (size_t)blockIdx.x * (4 * Cfg::EPQ) + (warp - 2);
      if (G == 32 || (lane & 1) == 0) {
        const int ch = G == 32 ? lane : lane >> 1;
#pragma unroll
        for (int gi = 0; gi < NGR; ++gi) {
          p.bn_part[rowi * C + gi * G + ch] = bn_s[gi];
          p.bn_part[(rows + rowi) * C + gi * G + ch] = bn_q[gi];
        }
      }
    }
    if (threadIdx.x == 64) tr.mark(8);
  } else {
    // ===================== strict-mode converter warps 10..13 =====================
    if (STRICT) {
      const int ctid = threadIdx.x - 10 * 32;  // 0..127
      uint32_t ia = 0;
      for (int itile = 0; itile < p.iters; ++itile) {
        for (int kb = 0; kb < NKB; ++kb) {
          const uint32_t s = ia % p.sa, ph = (ia / p.sa) & 1;
          mbar_wait_sleep_lean(&a_full[s], ph);
          const uint4* src = reinterpret_cast<const uint4*>(smem + p.a_off + s * p.a_stride);
          uint4* dst = reinterpret_cast<uint4*>(smem + p.a_off + s * p.a_stride + p.a_lo_off);
          const int n16 = p.a_bytes / 16;
          for (int i = ctid; i < n16; i += 128) {
            const uint4 u = src[i];
            uint4 o;
            o.x = __float_as_uint(tf32_rna(__uint_as_float(u.x) - __uint_as_float(u.x & 0xFFFFE000u)));
            o.y = __float_as_uint(tf32_rna(__uint_as_float(u.y) - __uint_as_float(u.y & 0xFFFFE000u)));
            o.z = __float_as_uint(tf32_rna(__uint_as_float(u.z) - __uint_as_float(u.z & 0xFFFFE000u)));
            o.w = __float_as_uint(tf32_rna(__uint_as_float(u.w) - __uint_as_float(u.w & 0xFFFFE000u)));
            dst[i] = o;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_conv[s]);
          ++ia;
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (p.cs > 1) cluster_sync_all();   // nobody leaves while a peer may still multicast into / signal this CTA
  if (warp == 1) { if constexpr (TWO) tmem_dealloc2(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols); }
  if (threadIdx.x == 0) { tr.mark(9); tr.wall(15); }
}

}  // namespace b200ode
