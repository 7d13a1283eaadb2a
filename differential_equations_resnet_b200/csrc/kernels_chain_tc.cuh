// Persistent Euler-step chains: ALL consecutive Euler steps of one residual stage in one launch.
//
// Reference path replaced: the stage loop of models/tfkeras_resnets.py:575-593, which stacks
// single_layer_identity_block (models/tfkeras_resnets.py:28-94: conv_K + bias -> relu -> h* -> +x)
// n times on a tensor of constant shape, and its TF-autodiff backward sweep (training/training.py:300).
//
// One CTA owns one image for the whole chain.  The image lives in shared memory as a "padded linear"
// halo strip (pitch P = W+1, see kernels_conv_tc.cuh) that the tcgen05 MMAs read as the A operand of
// all nine taps; the epilogue warps write the next step's input back INTO shared memory in exactly
// the swizzled K-major layout TMA would have produced, so step l+1 starts without touching L2/HBM.
// Per step only the small weight tile streams in (TMA ring, prefetched one step ahead); to HBM go the
// saved activation + 1-bit relu mask (forward) or dZ_l (backward), both needed by the weight gradient.
//
//   DIR 0 (forward):  x_{l+1} = x_l + h*relu(conv_{K_l}(x_l) + b_l)
//   DIR 1 (backward): dZ_l = h*dY_l*mask_l ; dY_{l-1} = dY_l - conv_{K_l}(dZ_l) + 2*gamma*dZ_l
//                     (the data gradient reuses the forward weights); dY stays in
//                     a second, thread-private shared buffer E.
//
// ST = true (strict mode, 3xTF32): fp32-grade results from tf32 MMAs.  The tensor core reads the top 19 bits of an fp32
// operand, so strip 0 (fp32, also the residual stream) IS the "hi" operand; the epilogue that produces a strip also writes
// its remainder lo = v - trunc_tf32(v) (exact in fp32, then rounded to tf32) into strip 1, and every tap of a weight ring
// stage carries 2C rows: the C output channels of W_hi followed by those of W_lo (pack_chain_kernel).  Per (tap, k-step)
// TWO MMAs: x_hi * [W_hi | W_lo] with N = 2C (one operand read of the strip feeds both products; they land in adjacent
// TMEM column ranges [main | correction]) and x_lo * W_hi with N = C onto the main range; the epilogue adds the two
// ranges.  (M = 128 x N <= 64 MMAs are bound by the A-operand read, so the stacked MMA costs what one did.)  Two strips (value + remainder) leave no room for the ping-pong pair of the fast
// mode: the step's input is updated IN PLACE, so the epilogue waits for the step's last MMA before it stores anything
// (MMAs and epilogue alternate; the next step's weights stream in under the epilogue).
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..9 = epilogue (two warps per TMEM lane quarter).  When all nine
// taps of a layer fit one ring stage (C <= 32) the MMAs are issued segment by segment with one commit
// per 128-position segment, so the epilogue of segment s overlaps the MMAs of segments > s.
#pragma once

#include "kernels_conv_tc.cuh"
#include "sm100_ptx.cuh"

namespace b200ode {

constexpr int CHAIN_MAXSEG = 9;

struct ChainParams {
  int N, H, W, P;
  int L;             // Euler steps to run
  int Lw;            // distinct weight layers; step l uses weights l % Lw
  int nseg;          // 128-position segments per image
  int tw, sw;        // taps per weight ring stage, ring depth
  int seg_outer;     // 1: all taps resident, per-segment commits
  uint32_t plane_bytes;   // one K-block plane of a strip (1024-aligned)
  uint32_t strip_stride;  // NKB * plane_bytes
  uint32_t x_bytes;       // bytes one TMA plane load of x0 delivers
  uint32_t e_off, w_off, w_stage_bytes, bar_off;
  uint32_t tmem_cols;
  float h, gamma;
  // forward
  float* acts;            // nullable [L][N,H,W,C]: output of every step
  uint8_t* masks;         // nullable [L][N,H,W,C/8]
  float* y_final;         // nullable [N,H,W,C]: output of the last step
  const float* bias;      // [Lw][C]
  // backward
  const float* dy;        // [N,H,W,C] gradient w.r.t. the chain output
  const uint8_t* masks_r; // [L][N,H,W,C/8]
  float* dz_all;          // [L][N,H,W,C]
  float* dx;              // [N,H,W,C] gradient w.r.t. the chain input
  uint64_t* trace;        // nullable per-CTA timeline (debug)
  // Thread-block cluster: all CTAs of a cluster run the same layer schedule on different images, so
  // the per-layer weight tiles are loaded ONCE per cluster by a TMA multicast of the rank-0 CTA.
  int cs;                 // cluster size (1 = no cluster)
  int dbg_skip_w;         // debug: do not load weights (timing experiments only)
  int iters;              // images per CTA = ceil(N / gridDim.x); CTAs whose image index is >= N run as ghosts
};

template <int C>
struct ChainCfg {
  static constexpr int ROWB = (C * 4 >= 128) ? 128 : C * 4;   // bytes per position row in one K-block plane
  static constexpr int KB = ROWB / 4;                         // channels per plane
  static constexpr int NKB = C / KB;
  static constexpr int KS = ROWB / 32;                        // tf32 k-steps (8 channels) per plane
  static constexpr int MW = (C + 31) / 32;                    // 32-bit mask words per pixel
  static constexpr int MAXSEG = C == 16 ? 9 : C == 32 ? 5 : 2; // segments a whole image may need (host plan agrees)
  // taps per weight ring stage (same rule as taps_per_w_stage() on the host)
  static constexpr int TW = (C * ROWB * 9 <= 40 * 1024) ? 9 : (C * ROWB * 3 <= 56 * 1024) ? 3 : 1;
};

// byte offset of 16-byte chunk `chunk` of position `pos` inside a swizzled plane
template <int ROWB>
__device__ __forceinline__ uint32_t strip_chunk_off(uint32_t pos, uint32_t chunk) {
  return swizzle_addr(pos * ROWB + chunk * 16, ROWB);
}
// E buffer (thread-private rows, C*4 bytes per pixel) with an XOR swizzle that spreads a warp's
// 16-byte accesses over all banks
template <int C>
__device__ __forceinline__ uint32_t e_chunk_off(uint32_t pix, uint32_t chunk) {
  const uint32_t x = (C == 16) ? ((pix >> 1) & 3u) : (pix & 7u);
  return pix * (C * 4) + ((chunk ^ x) << 4);
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// remainder of v after the tensor core's operand read (top 19 bits of the fp32 word), rounded to tf32
__device__ __forceinline__ float tf32_remainder(float v) {
  const float r = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  uint32_t o;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o) : "f"(r));
  return __uint_as_float(o);
}
__device__ __forceinline__ float4 tf32_remainder4(float4 v) {
  return make_float4(tf32_remainder(v.x), tf32_remainder(v.y), tf32_remainder(v.z), tf32_remainder(v.w));
}

template <int C, int DIR, bool ST = false>
__global__ void __launch_bounds__(320, 1)
chain_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const ChainParams p) {
  using Cfg = ChainCfg<C>;
  constexpr int ROWB = Cfg::ROWB, KB = Cfg::KB, NKB = Cfg::NKB, KS = Cfg::KS, MW = Cfg::MW, MAXSEG = Cfg::MAXSEG, TW = Cfg::TW;
  constexpr uint32_t LT = ROWB == 128 ? SWZ_128B : ROWB == 64 ? SWZ_64B : SWZ_32B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr uint32_t RU = ROWB >> 4;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* x_full = bars;                    // [1]
  uint64_t* layer_done = bars + 1;            // [1] count 4
  uint64_t* img_done = bars + 2;              // [1] count 4
  uint64_t* acc_full = bars + 3;              // [CHAIN_MAXSEG]
  uint64_t* w_full = acc_full + CHAIN_MAXSEG; // [sw]
  uint64_t* w_empty = w_full + p.sw;          // [sw]
  uint64_t* w_empty_cl = w_empty + p.sw;      // [sw] rank 0 only: the other CTAs of the cluster released the stage
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty_cl + p.sw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = smem_u32(smem);
  Trace tr;
  tr.begin(p.trace);
  if (threadIdx.x == 0) tr.wall(0);
  constexpr int TL = 4;   // traced (steady-state) step

  if (warp == 0 && lane == 0) {
    if (DIR == 0) tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    mbar_init(x_full, 1);
    mbar_init(layer_done, 8);
    mbar_init(img_done, 8);
    for (int i = 0; i < CHAIN_MAXSEG; ++i) mbar_init(&acc_full[i], 1);
    for (int i = 0; i < p.sw; ++i) {
      mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1);
      mbar_init(&w_empty_cl[i], p.cs > 1 ? p.cs - 1 : 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  {
    // both strips start as zeros: halo rows / the shared zero column are never written afterwards
    uint4* z = reinterpret_cast<uint4*>(smem);
    const uint32_t n16 = (2u * p.strip_stride) >> 4;
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (p.cs > 1) cluster_sync_all();   // every CTA's barriers exist before any multicast / remote arrive
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) tr.mark(1);
  const long long img_elems = (long long)p.H * p.W * C;
  const long long layer_elems = (long long)p.N * img_elems;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t iw = 0, ws = 0, wph = 0;
      const uint32_t crank = p.cs > 1 ? cluster_ctarank() : 0u;
      const uint16_t cmask = (uint16_t)((1u << p.cs) - 1u);
      for (uint32_t ic = 0; ic < (uint32_t)p.iters; ++ic) {
        const int img = blockIdx.x + ic * gridDim.x;   // >= N: ghost (TMA zero-fills the out-of-bounds box)
        if (DIR == 0) {
          if (ic > 0) mbar_wait_sleep(img_done, (ic - 1) & 1);
          mbar_expect_tx(x_full, NKB * p.x_bytes);
          for (int kb = 0; kb < NKB; ++kb) tma_load_4d(smem + kb * p.plane_bytes, &map_x, x_full, kb * KB, -1, -1, img);
        }
        for (int li = 0; li < p.L; ++li) {
          const int l = DIR ? p.L - 1 - li : li;
          const int lw = l % p.Lw;
          if (ic == 0 && li == TL + 1) tr.mark(13);
          if (ic == 0 && li == TL + 2) tr.mark(14);
          for (int kb = 0; kb < NKB; ++kb)
            for (int tg = 0; tg < 9; tg += p.tw) {
              const uint32_t s = ws, ph = wph;
              mbar_wait_sleep(&w_empty[s], ph ^ 1);           // this CTA's MMAs are done with the stage's previous contents
              if (p.dbg_skip_w && iw >= (uint32_t)p.sw) {
                mbar_arrive(&w_full[s]);
              } else if (p.cs == 1) {
                mbar_expect_tx(&w_full[s], p.w_stage_bytes);
                tma_load_3d(smem + p.w_off + s * p.w_stage_bytes, &map_w, &w_full[s], kb * KB, 0, lw * 9 + tg);
              } else if (crank != 0) {
                if (iw >= (uint32_t)p.sw) mbar_arrive_cluster(&w_empty_cl[s], 0);   // tell rank 0 the stage is free here
                mbar_expect_tx(&w_full[s], p.w_stage_bytes);                        // rank 0's multicast completes it
              } else {
                if (iw >= (uint32_t)p.sw) mbar_wait_sleep(&w_empty_cl[s], ph ^ 1);
                mbar_expect_tx(&w_full[s], p.w_stage_bytes);
                tma_load_3d_mc(smem + p.w_off + s * p.w_stage_bytes, &map_w, &w_full[s], kb * KB, 0, lw * 9 + tg, cmask);
              }
              ++iw;
              if (++ws == (uint32_t)p.sw) { ws = 0; wph ^= 1; }
            }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) ==========
    const bool leader = elect_one();
    const uint32_t idesc = make_instr_desc(FMT_TF32, 128, C, 0, 0);
    const uint32_t idesc2 = make_instr_desc(FMT_TF32, 128, 2 * C, 0, 0);     // strict: [W_hi | W_lo] rows of a tap
    constexpr uint32_t CW = ST ? 2 * C : C;                                   // TMEM columns per segment
    const uint32_t desc_hi32 = (SBO >> 4) | (1u << 14) | (LT << 29);
    constexpr uint32_t LBO_FIELD = 1u << 16;
    auto mk = [&](uint32_t lo) -> uint64_t { return (static_cast<uint64_t>(desc_hi32) << 32) | (lo | LBO_FIELD); };
    constexpr uint32_t tap_units = (uint32_t)((ST ? 2 : 1) * C * ROWB) >> 4;
    uint32_t toff[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) toff[t] = (uint32_t)((t / 3) * p.P + (t % 3)) * RU;
    const uint32_t plane_units = p.plane_bytes >> 4;
    uint32_t ws = 0, wph = 0, ld = 0;   // weight ring position kept incrementally (no integer division in the loop)
    for (uint32_t ic = 0; ic < (uint32_t)p.iters; ++ic) {
      for (int li = 0; li < p.L; ++li) {
        // The step's weight entries are awaited FIRST: they were prefetched a step ahead, and this warp is
        // idle while the previous step's epilogue runs, so the ~140 cycles per mbarrier wait are hidden there
        // instead of sitting between the strip hand-over and the first MMA.
        if (!ST) {     // (strict: a layer's entries may exceed the ring; each one is awaited where it is used)
          uint32_t s2 = ws, ph2 = wph;
          const int nent = p.seg_outer ? 1 : NKB * (9 / TW);
          for (int e = 0; e < nent; ++e) {
            mbar_wait(&w_full[s2], ph2);
            if (++s2 == (uint32_t)p.sw) { s2 = 0; ph2 ^= 1; }
          }
        }
        if (DIR == 0 && li == 0 && !ST) mbar_wait(x_full, ic & 1);
        else { mbar_wait(layer_done, ld & 1); ++ld; }       // (strict forward: the epilogue warps derive the lo strip of x0 first)
        tc_fence_after_sync();
        if (ic == 0 && lane == 0) { if (li == TL) tr.mark(2); if (li == TL + 1) tr.mark(5); }
        const uint32_t a_base = (smem_base + (ST ? 0u : (uint32_t)(li & 1) * p.strip_stride)) >> 4;
        const uint32_t a_lo = p.strip_stride >> 4;     // strict: remainder strip
        if (p.seg_outer) {
          const uint32_t s = ws;
          if (ic == 0 && li == TL && lane == 0) tr.mark(3);
          const uint32_t b_base = (smem_base + p.w_off + s * p.w_stage_bytes) >> 4;
          if (ST) { mbar_wait(&w_full[s], wph); tc_fence_after_sync(); }
          uint32_t a_sg = a_base, d = tmem_base;
          for (int sg = 0; sg < p.nseg; ++sg, a_sg += 128 * RU, d += CW) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
#pragma unroll
              for (int ks = 0; ks < KS; ++ks) {
                const uint64_t da = mk(a_sg + toff[t] + 2 * ks), db = mk(b_base + t * tap_units + 2 * ks);
                if (ST) {
                  const uint64_t dal = mk(a_sg + a_lo + toff[t] + 2 * ks);
                  if (leader) {
                    umma_tf32(d, da, db, idesc2, (t | ks) ? 1u : 0u);
                    umma_tf32(d, dal, db, idesc, 1u);
                  }
                } else {
                  if (leader) umma_tf32(d, da, db, idesc, (t | ks) ? 1u : 0u);
                }
              }
            }
            if (leader) umma_commit(&acc_full[sg]);
          }
          if (leader) umma_commit(&w_empty[s]);
          if (++ws == (uint32_t)p.sw) { ws = 0; wph ^= 1; }
        } else {
          // All ring entries of the layer are normally prefetched while the previous layer ran: wait for
          // them up front, then issue the layer's MMAs as one unrolled stream (no wait bubbles in between).
          long long wstall = 0;
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) {
            const uint32_t a_kb = a_base + kb * plane_units;
#pragma unroll
            for (int tg = 0; tg < 9; tg += TW) {
              const uint32_t s = ws;
              const uint32_t b_base = (smem_base + p.w_off + s * p.w_stage_bytes) >> 4;
              if (ST) { mbar_wait(&w_full[s], wph); tc_fence_after_sync(); }
              // taps and k-steps fully unrolled (descriptor arithmetic overlaps across MMAs); segments outermost
              uint32_t a_sg = a_kb, d = tmem_base;
              for (int sg = 0; sg < p.nseg; ++sg, a_sg += 128 * RU, d += CW) {
#pragma unroll
                for (int tt = 0; tt < TW; ++tt) {
#pragma unroll
                  for (int ks = 0; ks < KS; ++ks) {
                    const uint64_t da = mk(a_sg + toff[tg + tt] + 2 * ks), db = mk(b_base + tt * tap_units + 2 * ks);
                    if (ST) {
                      const uint64_t dal = mk(a_sg + a_lo + toff[tg + tt] + 2 * ks);
                      if (leader) {
                        umma_tf32(d, da, db, idesc2, (kb | tg | tt | ks) ? 1u : 0u);
                        umma_tf32(d, dal, db, idesc, 1u);
                      }
                    } else {
                      if (leader) umma_tf32(d, da, db, idesc, (kb | tg | tt | ks) ? 1u : 0u);
                    }
                  }
                }
              }
              if (leader) umma_commit(&w_empty[s]);
              if (++ws == (uint32_t)p.sw) { ws = 0; wph ^= 1; }
            }
          }
          for (int sg = 0; sg < p.nseg; ++sg)
            if (leader) umma_commit(&acc_full[sg]);
          if (ic == 0 && li == TL && lane == 0 && tr.buf) tr.buf[3] = (uint64_t)wstall;
        }
        if (ic == 0 && li == TL && lane == 0) tr.mark(4);
        __syncwarp();
      }
    }
    if (lane == 0) tr.mark(10);
  } else {
    // ===================== epilogue warps 2..9 =====================
    // Two warps per TMEM lane quarter (hardware: warp w reads lanes 32*(w%4)..+31); the (segment,
    // 16-channel group) work items of a layer alternate between the two.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t e_base = smem_base + p.e_off;
    constexpr int groups = C / 8;   // mask bytes per pixel
    constexpr int NG = C / 16;      // 16-channel groups per pixel
    uint32_t lc = 0;
    for (int ic = 0; ic < p.iters; ++ic) {
      const int img = blockIdx.x + ic * gridDim.x;
      const bool active = img < p.N;     // ghosts keep the barrier protocol but touch no global memory
      const long long img_off = (long long)img * img_elems;
      // Per-thread geometry of every segment (layer invariant): pixel index, strip position and validity are computed
      // ONCE per image.  They used to be recomputed per (layer, segment) between the accumulator barrier and the TMEM
      // load -- ~90 of the ~240 instructions of an item, on the critical path of the layer hand-over.
      int pixl_s[MAXSEG];
      uint32_t vmask = 0;
#pragma unroll
      for (int sg = 0; sg < MAXSEG; ++sg) {
        const int q = sg * 128 + row;
        const int yy = q / p.P, xq = q - yy * p.P;
        pixl_s[sg] = yy * p.W + xq;
        if (sg < p.nseg && active && yy < p.H && xq < p.W) vmask |= 1u << sg;
      }
      if (DIR == 1) {
        // ---- init: E = dY_L, strip0 = dZ_{L-1} = h * dY_L * mask_{L-1} ----
        const uint8_t* mk_l = p.masks_r + ((long long)(p.L - 1) * p.N + img) * (long long)p.H * p.W * groups;
        float* dz_l = p.dz_all + (long long)(p.L - 1) * layer_elems + img_off;
#pragma unroll
        for (int sg = 0; sg < MAXSEG; ++sg) {
          if ((vmask >> sg) & 1u) {
            const int pixl = pixl_s[sg];
            const uint32_t pos = (uint32_t)(sg * 128 + row + p.P + 1);
            const float* src = p.dy + img_off + (long long)pixl * C;
#pragma unroll
            for (int cg = 0; cg < NG; ++cg) {
              if (((sg * NG + cg) & 1) != half) continue;
              const int c0 = cg * 16;
              const uint32_t bits = *reinterpret_cast<const uint16_t*>(mk_l + (long long)pixl * groups + c0 / 8);
              const uint32_t pl = smem_base + (uint32_t)(c0 / KB) * p.plane_bytes;
              float4 d[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) d[j] = *reinterpret_cast<const float4*>(src + c0 + 4 * j);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                sts128(e_base + e_chunk_off<C>(pixl, c0 / 4 + j), d[j]);
                float4 z;
                z.x = (bits >> (4 * j)) & 1u ? p.h * d[j].x : 0.0f;
                z.y = (bits >> (4 * j + 1)) & 1u ? p.h * d[j].y : 0.0f;
                z.z = (bits >> (4 * j + 2)) & 1u ? p.h * d[j].z : 0.0f;
                z.w = (bits >> (4 * j + 3)) & 1u ? p.h * d[j].w : 0.0f;
                sts128(pl + strip_chunk_off<ROWB>(pos, (c0 % KB) / 4 + j), z);
                if (ST) sts128(pl + p.strip_stride + strip_chunk_off<ROWB>(pos, (c0 % KB) / 4 + j), tf32_remainder4(z));
                *reinterpret_cast<float4*>(dz_l + (long long)pixl * C + c0 + 4 * j) = z;
              }
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(layer_done);
      }
      if (ST && DIR == 0) {
        // ---- init: remainder strip of x0 (strip 0 arrives by TMA) ----
        mbar_wait_sleep(x_full, ic & 1);
#pragma unroll
        for (int sg = 0; sg < MAXSEG; ++sg) {
          if ((vmask >> sg) & 1u) {
            const uint32_t pos = (uint32_t)(sg * 128 + row + p.P + 1);
#pragma unroll
            for (int cg = 0; cg < NG; ++cg) {
              if (((sg * NG + cg) & 1) != half) continue;
              const int c0 = cg * 16;
              const uint32_t pl = smem_base + (uint32_t)(c0 / KB) * p.plane_bytes;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t so = strip_chunk_off<ROWB>(pos, (c0 % KB) / 4 + j);
                sts128(pl + p.strip_stride + so, tf32_remainder4(lds128(pl + so)));
              }
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(layer_done);
      }
      for (int li = 0; li < p.L; ++li, ++lc) {
        const int l = DIR ? p.L - 1 - li : li;
        const uint32_t cur = smem_base + (ST ? 0u : (uint32_t)(li & 1) * p.strip_stride);
        const uint32_t nxt = smem_base + (ST ? 0u : (uint32_t)((li & 1) ^ 1) * p.strip_stride);
        const uint32_t nlo = smem_base + p.strip_stride;       // strict: remainder strip
        const bool last = li == p.L - 1;
        // per-layer pointers
        const float4* bias4 = reinterpret_cast<const float4*>(p.bias + (l % p.Lw) * C);
        float* out_l = nullptr;        // global copy of this step's result
        uint8_t* mask_w = nullptr;
        uint32_t mkreg[MAXSEG][MW];
        if (DIR == 0) {
          if (p.acts) out_l = p.acts + (long long)l * layer_elems + img_off;
          else if (last) out_l = p.y_final + img_off;
          if (p.masks) mask_w = p.masks + ((long long)l * p.N + img) * (long long)p.H * p.W * groups;
        } else {
          if (!last) {
            out_l = p.dz_all + (long long)(l - 1) * layer_elems + img_off;
            // prefetch the relu masks of step l-1 for this thread's pixels (hidden behind the MMAs)
            const uint8_t* mk_l = p.masks_r + ((long long)(l - 1) * p.N + img) * (long long)p.H * p.W * groups;
#pragma unroll
            for (int sg = 0; sg < MAXSEG; ++sg) {
              {
                if ((vmask >> sg) & 1u) {
                  const uint8_t* mp = mk_l + (long long)pixl_s[sg] * groups;
                  if (C == 16) mkreg[sg][0] = *reinterpret_cast<const uint16_t*>(mp);
                  else {
#pragma unroll
                    for (int w = 0; w < MW; ++w) mkreg[sg][w] = *reinterpret_cast<const uint32_t*>(mp + 4 * w);
                  }
                }
              }
            }
          } else {
            out_l = p.dx + img_off;
          }
        }
        if (ST) {     // in-place update: no store into the strips before the step's last MMA has read them
          mbar_wait_sleep(&acc_full[p.nseg - 1], lc & 1);
          tc_fence_after_sync();
        }
#pragma unroll
        for (int sg = 0; sg < MAXSEG; ++sg) {
          // Every warp waits for the LAST segment's commit even when it owns no item of it: that gates
          // it to the MMA warp's progress, so no warp can arrive twice in one layer_done phase.
          if (sg < p.nseg && (NG > 1 || (sg & 1) == half || sg == p.nseg - 1)) {
            mbar_wait_sleep(&acc_full[sg], lc & 1);
            tc_fence_after_sync();
            if (threadIdx.x == 64 && lc == TL && sg == 0) tr.mark(6);
            if (threadIdx.x == 64 && lc == TL + 1 && sg == 0) tr.mark(9);
            if (threadIdx.x == 64 && lc == TL && sg == 2) tr.mark(7);
            const bool valid = (vmask >> sg) & 1u;
            const int pixl = pixl_s[sg];
            const uint32_t pos = (uint32_t)(sg * 128 + row + p.P + 1);
            // all TMEM loads of this warp's items of the segment are issued up front (one wait covers them)
            constexpr int NGI = NG > 1 ? NG / 2 : 1;
            constexpr int CWE = ST ? 2 * C : C;      // TMEM columns per segment (strict: [main | correction])
            uint32_t rr[NGI][16], rc[ST ? NGI : 1][16];
#pragma unroll
            for (int gj = 0; gj < NGI; ++gj) {
              const int cgw = NG > 1 ? half + 2 * gj : 0;
              if (NG > 1 || (sg & 1) == half) {
                tmem_ld_x16(tq + sg * CWE + cgw * 16, rr[gj]);
                if (ST) tmem_ld_x16(tq + sg * CWE + C + cgw * 16, rc[ST ? gj : 0]);
              }
            }
#pragma unroll
            for (int cg = 0; cg < NG; ++cg) {
              if (((sg * NG + cg) & 1) != half) continue;
              const int c0 = cg * 16;
              uint32_t (&r)[16] = rr[NG > 1 ? cg / 2 : 0];
              uint32_t (&rx)[16] = rc[ST && NG > 1 ? cg / 2 : 0];
              const uint32_t plo = (uint32_t)(c0 / KB) * p.plane_bytes;
              const uint32_t ch0 = (c0 % KB) / 4;
              uint32_t so[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) so[j] = plo + strip_chunk_off<ROWB>(pos, ch0 + j);
              if (DIR == 0) {
                float4 bv[4], xv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) bv[j] = __ldg(bias4 + c0 / 4 + j);
                if (valid) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) xv[j] = lds128(cur + so[j]);
                }
                tmem_ld_wait();
                if (ST) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(rx[i]));
                }
                if (valid) {
                  uint32_t bits = 0;
                  float4 o[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    float v[4] = {__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                  __uint_as_float(r[4 * j + 3])};
                    const float bs[4] = {bv[j].x, bv[j].y, bv[j].z, bv[j].w};
                    const float xs[4] = {xv[j].x, xv[j].y, xv[j].z, xv[j].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      float t = v[e] + bs[e];
                      bits |= (t > 0.0f ? 1u : 0u) << (4 * j + e);
                      t = fmaxf(t, 0.0f);
                      // Lambda(h*x) (only when h != 1, tfkeras_resnets.py:90) and add() are two layers in the
                      // reference: two roundings, so no FMA contraction here
                      if (p.h != 1.0f) t = __fmul_rn(p.h, t);
                      v[e] = __fadd_rn(xs[e], t);
                    }
                    o[j] = make_float4(v[0], v[1], v[2], v[3]);
                  }
                  if (!last) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) sts128(nxt + so[j], o[j]);
                    if (ST) {
#pragma unroll
                      for (int j = 0; j < 4; ++j) sts128(nlo + so[j], tf32_remainder4(o[j]));
                    }
                  }
                  if (out_l) {
                    float4* op = reinterpret_cast<float4*>(out_l + (long long)pixl * C + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) op[j] = o[j];
                  }
                  if (mask_w) *reinterpret_cast<uint16_t*>(mask_w + (long long)pixl * groups + c0 / 8) = (uint16_t)bits;
                }
              } else {
                const uint32_t bits = last ? 0u : (mkreg[sg][c0 / 32] >> (c0 % 32)) & 0xFFFFu;
                uint32_t eo[4];
                float4 dyv[4], zv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) eo[j] = e_base + e_chunk_off<C>(pixl, c0 / 4 + j);
                if (valid) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) dyv[j] = lds128(eo[j]);
                  if (p.gamma != 0.0f) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) zv[j] = lds128(cur + so[j]);
                  }
                }
                tmem_ld_wait();
                if (ST) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(rx[i]));
                }
                if (valid) {
                  float4 o[4], z[4];
                  const float g2 = 2.0f * p.gamma;
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    float v[4] = {-__uint_as_float(r[4 * j]), -__uint_as_float(r[4 * j + 1]), -__uint_as_float(r[4 * j + 2]),
                                  -__uint_as_float(r[4 * j + 3])};
                    if (p.gamma != 0.0f) {
                      v[0] = fmaf(g2, zv[j].x, v[0]); v[1] = fmaf(g2, zv[j].y, v[1]);
                      v[2] = fmaf(g2, zv[j].z, v[2]); v[3] = fmaf(g2, zv[j].w, v[3]);
                    }
                    v[0] += dyv[j].x; v[1] += dyv[j].y; v[2] += dyv[j].z; v[3] += dyv[j].w;
                    o[j] = make_float4(v[0], v[1], v[2], v[3]);
                    z[j].x = (bits >> (4 * j)) & 1u ? p.h * v[0] : 0.0f;
                    z[j].y = (bits >> (4 * j + 1)) & 1u ? p.h * v[1] : 0.0f;
                    z[j].z = (bits >> (4 * j + 2)) & 1u ? p.h * v[2] : 0.0f;
                    z[j].w = (bits >> (4 * j + 3)) & 1u ? p.h * v[3] : 0.0f;
                  }
                  float4* op = reinterpret_cast<float4*>(out_l + (long long)pixl * C + c0);
                  if (!last) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { sts128(eo[j], o[j]); sts128(nxt + so[j], z[j]); }
                    if (ST) {
#pragma unroll
                      for (int j = 0; j < 4; ++j) sts128(nlo + so[j], tf32_remainder4(z[j]));
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) op[j] = z[j];
                  } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) op[j] = o[j];
                  }
                }
              }
            }
          }
        }
        tc_fence_before_sync();
        if (threadIdx.x == 64 && lc == TL) tr.mark(8);
        if (!last) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(layer_done);
        } else if (DIR == 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive(img_done);
        }
      }
    }
  }

  if (threadIdx.x == 64) tr.mark(11);
  tc_fence_before_sync();
  __syncthreads();
  if (p.cs > 1) cluster_sync_all();   // nobody leaves while a peer may still multicast into / signal this CTA
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
  if (threadIdx.x == 0) { tr.mark(12); tr.wall(15); }
}

}  // namespace b200ode
