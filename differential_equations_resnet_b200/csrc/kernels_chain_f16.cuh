// Persistent Euler-step chains, second formulation: fp16 tensor-core operands (kind::f16, K = 16 per
// tcgen05.mma) around an fp32 residual stream.
//
// Reference path replaced: the stage loop of models/tfkeras_resnets.py:575-593 (n stacked
// single_layer_identity_block, :28-94) and its TF-autodiff backward sweep (training/training.py:300).
//
// Why a second formulation.  With N = C <= 64 accumulator columns a tcgen05.mma is bound by reading its
// 128-row A operand from shared memory (csrc/umma_rate.cu: 39 / 40 / 48 cycles per M128 x N16/32/64
// instruction, identical for tf32 and 16-bit operands), so the tf32 kernel (kernels_chain_tc.cuh, 8 channels
// per instruction) spends twice the instructions, shared-memory bytes and strip capacity of a 16-bit operand.
// fp16 has the SAME 11-bit significand as tf32; what it lacks is range, which the data provide: activations of
// a residual net are O(1), and the backward strip carries S*dZ with one power-of-two scale S per launch derived
// from max|dY| (device scalar, computed by amax_abs_kernel), undone in the fp32 epilogue / the gradient fold.
// Operands are rounded to nearest WHERE THEY ARE PRODUCED (cvt.rn), so nothing is left for the tensor core to
// truncate (the tf32 path feeds raw fp32 activations, which tcgen05 truncates: a one-sided error).
//
//   DIR 0 (forward):  R = x_l (fp32, in REGISTERS of the thread that owns the pixel / channel group for the
//                     whole chain), strip = fp16(x_l) (A operand of all nine taps)
//                     x_{l+1} = x_l + h*relu(conv_{K_l}(x_l) + b_l)            (two roundings: h*, +)
//                     to HBM: fp16(x_l) for the weight gradient (acts[l]), relu bit mask, last x in fp32
//   DIR 1 (backward): R = dY_l (fp32 registers), strip = fp16(S*dZ_l), dZ_l = h*dY_l*mask_l
//                     dY_{l-1} = dY_l - conv_{K_l}(dZ_l) + 2*gamma*dZ_l        
//                     to HBM: fp16(S*dZ_l) for the weight gradient, dX in fp32
//
// One CTA owns one image for the whole chain ("padded linear" halo strip, pitch P = W+1: every 3x3 tap is
// a row shift of the same strip, see kernels_conv_tc.cuh).  All nine taps of a layer are ONE TMA box
// (9*C*C*2 bytes), double buffered across layers and multicast to the CTAs of a cluster.  Warp roles:
// 0 = TMA producer (weights), 1 = MMA issuer, 2.. = EW epilogue warps (EW/4 per TMEM lane quarter).
// MMAs are issued segment by segment (128 positions) with one commit each, and the hand-over back to the MMA
// warp is per segment as well (a WAVEFRONT across steps): the MMAs of step l+1, segment s start as soon as the
// epilogue of step l has rewritten segments s-1, s, s+1 of the other strip, so for images of several segments
// the tensor pipe does not wait for the tail of the previous step's epilogue.
#pragma once

#include <cuda_fp16.h>

#include "kernels_chain_tc.cuh"
#include "sm100_ptx.cuh"

namespace b200ode {

struct ChainF16Params {
  int N, H, W, P;
  int L;             // Euler steps to run
  int Lw;            // distinct weight layers; step l uses weights l % Lw
  int nseg;          // 128-position segments per image
  int sw;            // weight ring depth (layers)
  uint32_t strip_stride;    // bytes of one strip (1024-aligned)
  uint32_t w_off, w_layer_bytes, bar_off;   // w_layer_bytes: ring stride (1024-aligned)
  uint32_t w_box_bytes;     // bytes one weight TMA box delivers (9*C*C*2)
  uint32_t tmem_cols;
  float h, gamma;
  // forward
  const float* x0;        // [N,H,W,C] chain input (fp32)
  __half* acts;           // nullable [L][N,H,W,C]: fp16 INPUT of every step (acts[0] = fp16(x0)), the weight-gradient operand
  uint8_t* masks;         // nullable [L][N,H,W,C/8]
  float* y_final;         // nullable [N,H,W,C]: output of the last step (fp32)
  const float* bias;      // [Lw][C]
  // backward
  const float* dy;        // [N,H,W,C] gradient w.r.t. the chain output (fp32)
  const uint8_t* masks_r; // [L][N,H,W,C/8]
  __half* dz_all;         // [L][N,H,W,C]: fp16(S * dZ_l)
  float* dx;              // [N,H,W,C] gradient w.r.t. the chain input (fp32)
  const float* amax;      // device scalar max|dy| (S = chain_grad_scale(h, *amax))
  uint64_t* trace;        // nullable per-CTA timeline (debug)
  int cs;                 // cluster size (1 = no cluster)
  int iters;              // images per CTA = ceil(N / gridDim.x); CTAs whose image index is >= N run as ghosts
};

// max |v| over a tensor -> *out (bit pattern of a non-negative float, atomicMax on the uint view; *out zeroed before)
__global__ void amax_abs_kernel(const float4* __restrict__ v, long long n4, unsigned int* __restrict__ out) {
  griddep_launch_dependents();      // the chain kernel behind this one may run its prologue meanwhile (it waits before reading *out)
  float m = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = v[i];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float wm[32];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? wm[threadIdx.x] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0 && m > 0.0f) atomicMax(out, __float_as_uint(m));
  }
}

// Chain weights as fp16 K-major B operands [L][9][o][ci] + fp32 biases [L][C]; blockIdx.y = layer.
__global__ void pack_chain_f16_kernel(LayerGeom g, const float* __restrict__ params, long long param_layer_stride,
                                      __half* __restrict__ w16, float* __restrict__ bias_out) {
  const long long total = (long long)g.k * g.k * g.C * g.C;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int l = blockIdx.y;
  params += (long long)l * param_layer_stride;
  if (i < g.C) bias_out[(long long)l * g.C + i] = g.use_bias ? params[g.bias_off + i] : 0.0f;
  if (i >= total) return;
  const int ci = (int)(i % g.C);
  const int o = (int)((i / g.C) % g.C);
  const int tap = (int)(i / ((long long)g.C * g.C));
  // rounding to nearest is sign-symmetric, so the staged kernel keeps K[a,b,ci,o] = -K[2-a,2-b,o,ci] bit for bit
  w16[(long long)l * total + i] = __float2half_rn(kernel_entry(g, params, tap / g.k, tap % g.k, ci, o));
}

template <int C>
struct ChainF16Cfg {
  static constexpr int ROWB = C * 2;                          // bytes per position row of the fp16 strip (one K-block)
  static constexpr int KS = C / 16;                           // 16-channel k-steps per tap
  static constexpr int NG = C / 16;                           // 16-channel groups per pixel (epilogue work items)
  static constexpr int MW = (C + 31) / 32;                    // 32-bit mask words per pixel
  static constexpr int MAXSEG = C == 16 ? 9 : C == 32 ? 5 : 2; // segments a whole image may need (host plan agrees)
};

// two fp32 -> packed fp16x2, round to nearest even, saturating to the largest finite value
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) {
  __half2 h = *reinterpret_cast<__half2*>(&v);
  return __half22float2(h);
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

template <int C, int DIR, int EW>
__global__ void __launch_bounds__(64 + EW * 32, 1)
chain_f16_kernel(const __grid_constant__ CUtensorMap map_w, const ChainF16Params p) {
  using Cfg = ChainF16Cfg<C>;
  constexpr int ROWB = Cfg::ROWB, KS = Cfg::KS, NG = Cfg::NG, MW = Cfg::MW, MAXSEG = Cfg::MAXSEG;
  constexpr uint32_t LT = ROWB == 128 ? SWZ_128B : ROWB == 64 ? SWZ_64B : SWZ_32B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr uint32_t RU = ROWB >> 4;     // 16-byte units per position row
  constexpr int NS = EW / 4;             // epilogue warps per TMEM lane quarter
  constexpr int NI = (MAXSEG * NG + NS - 1) / NS;   // work items (segment, 16-channel group) one warp may own
  constexpr int OWNERS = 4 * (NG < NS ? NG : NS);   // warps that own an item of a given segment
  static_assert(NS == 2 || NS == 3 || NS == 4, "2..4 epilogue warps per TMEM quarter");
  static_assert((NG <= NS && NS % NG == 0) || (NG > NS && NG % NS == 0), "channel groups must deal evenly to the warps of a quarter");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* seg_done = bars;                        // [CHAIN_MAXSEG] count OWNERS: segment sg of the next strip is written
  uint64_t* acc_full = seg_done + CHAIN_MAXSEG;     // [CHAIN_MAXSEG]
  uint64_t* w_full = acc_full + CHAIN_MAXSEG;       // [sw]
  uint64_t* w_empty = w_full + p.sw;                // [sw]
  uint64_t* w_empty_cl = w_empty + p.sw;            // [sw] rank 0 only: the other CTAs of the cluster released the stage
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty_cl + p.sw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = smem_u32(smem);
  Trace tr;
  tr.begin(p.trace);
  if (threadIdx.x == 0) tr.wall(0);
  constexpr int TL = 4;   // traced (steady-state) step

  griddep_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w);
    for (int i = 0; i < CHAIN_MAXSEG; ++i) { mbar_init(&seg_done[i], OWNERS); mbar_init(&acc_full[i], 1); }
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
  griddep_wait();                     // PDL: everything above touched shared memory / TMEM only
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) tr.mark(1);
  const long long img_elems = (long long)p.H * p.W * C;
  const long long layer_elems = (long long)p.N * img_elems;

  if (warp == 0) {
    // ===================== TMA producer: one box (all nine taps) per layer =====================
    if (lane == 0) {
      uint32_t iw = 0, ws = 0, wph = 0;
      const uint32_t crank = p.cs > 1 ? cluster_ctarank() : 0u;
      const uint16_t cmask = (uint16_t)((1u << p.cs) - 1u);
      for (uint32_t ic = 0; ic < (uint32_t)p.iters; ++ic) {
        for (int li = 0; li < p.L; ++li) {
          const int l = DIR ? p.L - 1 - li : li;
          const int lw = l % p.Lw;
          const uint32_t s = ws, ph = wph;
          mbar_wait_sleep(&w_empty[s], ph ^ 1);           // this CTA's MMAs are done with the stage's previous contents
          if (p.cs == 1) {
            mbar_expect_tx(&w_full[s], p.w_box_bytes);
            tma_load_3d(smem + p.w_off + s * p.w_layer_bytes, &map_w, &w_full[s], 0, 0, lw * 9);
          } else if (crank != 0) {
            if (iw >= (uint32_t)p.sw) mbar_arrive_cluster(&w_empty_cl[s], 0);   // tell rank 0 the stage is free here
            mbar_expect_tx(&w_full[s], p.w_box_bytes);                        // rank 0's multicast completes it
          } else {
            if (iw >= (uint32_t)p.sw) mbar_wait_sleep(&w_empty_cl[s], ph ^ 1);
            mbar_expect_tx(&w_full[s], p.w_box_bytes);
            tma_load_3d_mc(smem + p.w_off + s * p.w_layer_bytes, &map_w, &w_full[s], 0, 0, lw * 9, cmask);
          }
          ++iw;
          if (++ws == (uint32_t)p.sw) { ws = 0; wph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) ==========
    const bool leader = elect_one();
    const uint32_t idesc = make_instr_desc(FMT_F16, 128, C, 0, 0);
    const uint32_t desc_hi32 = (SBO >> 4) | (1u << 14) | (LT << 29);
    constexpr uint32_t LBO_FIELD = 1u << 16;
    auto mk = [&](uint32_t lo) -> uint64_t { return (static_cast<uint64_t>(desc_hi32) << 32) | (lo | LBO_FIELD); };
    constexpr uint32_t tap_units = (uint32_t)(C * ROWB) >> 4;
    uint32_t toff[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) toff[t] = (uint32_t)((t / 3) * p.P + (t % 3)) * RU;
    uint32_t ws = 0, wph = 0;
    uint32_t sd = 0;      // completions of seg_done[.] consumed so far (one per step: init, then every step but the last)
    for (uint32_t ic = 0; ic < (uint32_t)p.iters; ++ic) {
      for (int li = 0; li < p.L; ++li, ++sd) {
        // weights first (prefetched a layer ahead: this wait is hidden behind the previous step's epilogue)
        mbar_wait_lean(&w_full[ws], wph);
        if (ic == 0 && lane == 0) { if (li == TL) tr.mark(2); if (li == TL + 1) tr.mark(5); }
        const uint32_t a_base = (smem_base + (uint32_t)(li & 1) * p.strip_stride) >> 4;
        const uint32_t b_base = (smem_base + p.w_off + ws * p.w_layer_bytes) >> 4;
        uint32_t a_sg = a_base, d = tmem_base;
        for (int sg = 0; sg < p.nseg; ++sg, a_sg += 128 * RU, d += C) {
          // The A rows of segment sg span the strip rows the previous step's epilogue wrote for segments sg-1, sg and
          // sg+1 (sg-1 was awaited one iteration ago); its arrival also says the accumulator of sg has been read.
          // (one lane polls: 32 lanes spinning on shared memory next to running MMAs steal operand bandwidth)
          if (sg == 0) mbar_wait_lean(&seg_done[0], sd & 1);
          if (sg + 1 < p.nseg) mbar_wait_lean(&seg_done[sg + 1], sd & 1);
          tc_fence_after_sync();
#pragma unroll
          for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              const uint64_t da = mk(a_sg + toff[t] + 2 * ks), db = mk(b_base + t * tap_units + 2 * ks);
              if (leader) umma_f16(d, da, db, idesc, (t | ks) ? 1u : 0u);
            }
          }
          if (leader) umma_commit(&acc_full[sg]);
        }
        if (leader) umma_commit(&w_empty[ws]);
        if (++ws == (uint32_t)p.sw) { ws = 0; wph ^= 1; }
        if (ic == 0 && li == TL && lane == 0) tr.mark(4);
        __syncwarp();
      }
    }
    if (lane == 0) tr.mark(10);
  } else {
    // ===================== epilogue warps 2 .. 2+EW =====================
    // NS warps per TMEM lane quarter (hardware: warp w reads lanes 32*(w%4)..+31).  The (segment, 16-channel group)
    // work items of an image are dealt round-robin to the NS warps of a quarter: warp `sub` owns items
    // it = k*NS + sub, k = 0..NI-1 (segment it / NG, channel group it % NG), and a thread keeps the fp32 residual
    // (x_l / dY_l) of ITS items in registers for the whole chain.  The loops run over k, not over segments, so all
    // warps execute the SAME instruction stream (the two warps that share a scheduler differ only in `sub`).
    const int quarter = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    constexpr int groups = C / 8;   // mask bytes per pixel
    constexpr int IPS = NG > NS ? NG / NS : 1;      // items of one segment owned by one warp
    float S = 1.0f, invS = 1.0f;
    if (DIR == 1) { S = chain_grad_scale(p.h, *p.amax); invS = 1.0f / S; }
    const float hS = p.h * S;
    const float g2 = 2.0f * p.gamma * invS;
    // item k of this warp: segment and channel group
    auto seg_of = [&](int k) -> int { return NG >= NS ? k / IPS : k * (NS / NG) + sub / NG; };
    auto cg_of = [&](int k) -> int { return NG >= NS ? sub + (k % IPS) * NS : sub % NG; };
    uint32_t lc = 0;
    float res[NI][16];
    for (int ic = 0; ic < p.iters; ++ic) {
      const int img = blockIdx.x + ic * gridDim.x;
      const bool active = img < p.N;     // ghosts keep the barrier protocol but touch no global memory
      const long long img_off = (long long)img * img_elems;
      // per-thread geometry of every owned item (layer invariant)
      int pix_k[NI];
      uint32_t vmask = 0, omask = 0;     // item k: pixel is inside the image / segment exists
#pragma unroll
      for (int k = 0; k < NI; ++k) {
        const int sg = seg_of(k);
        const int q = sg * 128 + row;
        const int yy = q / p.P, xq = q - yy * p.P;
        pix_k[k] = yy * p.W + xq;
        if (sg < p.nseg) omask |= 1u << k;
        if (sg < p.nseg && active && yy < p.H && xq < p.W) vmask |= 1u << k;
      }
      // ---- init: R = chain input (x_0 / dY_L), strip 0 = its fp16 operand form (x_0 / S*dZ_{L-1}) ----
      {
        const float* src0 = (DIR == 0 ? p.x0 : p.dy) + img_off;
        const uint8_t* mk_l = DIR == 1 ? p.masks_r + ((long long)(p.L - 1) * p.N + img) * (long long)p.H * p.W * groups : nullptr;
        __half* cp = DIR == 0 ? (p.acts ? p.acts + img_off : nullptr) : p.dz_all + (long long)(p.L - 1) * layer_elems + img_off;
#pragma unroll
        for (int k = 0; k < NI; ++k) {
          if (!((omask >> k) & 1u)) continue;
          const int sg = seg_of(k), c0 = cg_of(k) * 16;
          if ((vmask >> k) & 1u) {
            const int pixl = pix_k[k];
            const uint32_t pos = (uint32_t)(sg * 128 + row + p.P + 1);
            float (&R)[16] = res[k];
            float4 d[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = *reinterpret_cast<const float4*>(src0 + (long long)pixl * C + c0 + 4 * j);
            uint32_t bits = 0xFFFFu;
            if (DIR == 1) bits = *reinterpret_cast<const uint16_t*>(mk_l + (long long)pixl * groups + c0 / 8);
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              R[4 * j] = d[j].x; R[4 * j + 1] = d[j].y; R[4 * j + 2] = d[j].z; R[4 * j + 3] = d[j].w;
              float4 z = d[j];
              if (DIR == 1) {
                z.x = (bits >> (4 * j)) & 1u ? hS * d[j].x : 0.0f;
                z.y = (bits >> (4 * j + 1)) & 1u ? hS * d[j].y : 0.0f;
                z.z = (bits >> (4 * j + 2)) & 1u ? hS * d[j].z : 0.0f;
                z.w = (bits >> (4 * j + 3)) & 1u ? hS * d[j].w : 0.0f;
              }
              pk[2 * j] = pack_f16x2(z.x, z.y);
              pk[2 * j + 1] = pack_f16x2(z.z, z.w);
            }
            const uint4 u0 = make_uint4(pk[0], pk[1], pk[2], pk[3]), u1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            sts128u(smem_base + strip_chunk_off<ROWB>(pos, c0 / 8), u0);
            sts128u(smem_base + strip_chunk_off<ROWB>(pos, c0 / 8 + 1), u1);
            if (cp) {
              uint4* gp = reinterpret_cast<uint4*>(cp + (long long)pixl * C + c0);
              gp[0] = u0; gp[1] = u1;
            }
          }
          if ((k % IPS) == IPS - 1) {     // last item of this warp in segment sg
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&seg_done[sg]);
          }
        }
      }
      for (int li = 0; li < p.L; ++li, ++lc) {
        const int l = DIR ? p.L - 1 - li : li;
        const uint32_t cur = smem_base + (uint32_t)(li & 1) * p.strip_stride;
        const uint32_t nxt = smem_base + (uint32_t)((li & 1) ^ 1) * p.strip_stride;
        const bool last = li == p.L - 1;
        __half* cp = nullptr;          // fp16 global copy of the NEXT step's operand
        uint8_t* mask_w = nullptr;
        uint32_t mk_k[NI];
        float4 bias_r[IPS][4];
        if (DIR == 0) {
          if (p.acts && !last) cp = p.acts + (long long)(l + 1) * layer_elems + img_off;
          if (p.masks) mask_w = p.masks + ((long long)l * p.N + img) * (long long)p.H * p.W * groups;
          // this step's biases of the channel groups this warp owns: in flight while the MMAs run
          const float4* bias4 = reinterpret_cast<const float4*>(p.bias + (l % p.Lw) * C);
#pragma unroll
          for (int k = 0; k < IPS; ++k) {
#pragma unroll
            for (int j = 0; j < 4; ++j) bias_r[k][j] = __ldg(bias4 + cg_of(k) * 4 + j);
          }
        } else if (!last) {
          cp = p.dz_all + (long long)(l - 1) * layer_elems + img_off;
          // prefetch the relu masks of step l-1 for this thread's items (hidden behind the MMAs)
          const uint8_t* mk_l = p.masks_r + ((long long)(l - 1) * p.N + img) * (long long)p.H * p.W * groups;
#pragma unroll
          for (int k = 0; k < NI; ++k)
            if ((vmask >> k) & 1u) mk_k[k] = *reinterpret_cast<const uint16_t*>(mk_l + (long long)pix_k[k] * groups + cg_of(k) * 2);
        }
#pragma unroll
        for (int k = 0; k < NI; ++k) {
          if (!((omask >> k) & 1u)) continue;
          const int sg = seg_of(k), c0 = cg_of(k) * 16;
          const bool valid = (vmask >> k) & 1u;
          const int pixl = pix_k[k];
          const uint32_t pos = (uint32_t)(sg * 128 + row + p.P + 1);
          if ((k % IPS) == 0) {
            mbar_wait_sleep_lean(&acc_full[sg], lc & 1);
            tc_fence_after_sync();
            if (threadIdx.x == 64 && lc == TL && k == 0) tr.mark(6);
            if (threadIdx.x == 64 && lc == TL + 1 && k == 0) tr.mark(9);
          }
          uint32_t r[16];
          tmem_ld_x16(tq + sg * C + c0, r);
          const uint32_t so0 = strip_chunk_off<ROWB>(pos, c0 / 8), so1 = strip_chunk_off<ROWB>(pos, c0 / 8 + 1);
          uint4 z0 = make_uint4(0u, 0u, 0u, 0u), z1 = z0;
          if (DIR == 1 && p.gamma != 0.0f && valid) { z0 = lds128u(cur + so0); z1 = lds128u(cur + so1); }
          tmem_ld_wait();
          float (&R)[16] = res[k];
          if (valid) {
            if (DIR == 0) {
              const float4 (&bq)[4] = bias_r[k % IPS];
              const float bs[16] = {bq[0].x, bq[0].y, bq[0].z, bq[0].w, bq[1].x, bq[1].y, bq[1].z, bq[1].w,
                                    bq[2].x, bq[2].y, bq[2].z, bq[2].w, bq[3].x, bq[3].y, bq[3].z, bq[3].w};
              uint32_t bits = 0;
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                float t = __uint_as_float(r[e]) + bs[e];
                bits |= (t > 0.0f ? 1u : 0u) << e;
                t = fmaxf(t, 0.0f);
                // Lambda(h*x) (only when h != 1, tfkeras_resnets.py:90) and add() are two layers in the
                // reference: two roundings, so no FMA contraction here
                if (p.h != 1.0f) t = __fmul_rn(p.h, t);
                R[e] = __fadd_rn(R[e], t);
              }
              if (!last) {
                const uint4 u0 = make_uint4(pack_f16x2(R[0], R[1]), pack_f16x2(R[2], R[3]), pack_f16x2(R[4], R[5]), pack_f16x2(R[6], R[7]));
                const uint4 u1 = make_uint4(pack_f16x2(R[8], R[9]), pack_f16x2(R[10], R[11]), pack_f16x2(R[12], R[13]), pack_f16x2(R[14], R[15]));
                sts128u(nxt + so0, u0);
                sts128u(nxt + so1, u1);
                if (cp) {
                  uint4* gp = reinterpret_cast<uint4*>(cp + (long long)pixl * C + c0);
                  gp[0] = u0; gp[1] = u1;
                }
              } else if (p.y_final) {
                float4* op = reinterpret_cast<float4*>(p.y_final + img_off + (long long)pixl * C + c0);
#pragma unroll
                for (int j = 0; j < 4; ++j) op[j] = make_float4(R[4 * j], R[4 * j + 1], R[4 * j + 2], R[4 * j + 3]);
              }
              if (mask_w) *reinterpret_cast<uint16_t*>(mask_w + (long long)pixl * groups + c0 / 8) = (uint16_t)bits;
            } else {
              // dY_{l-1} = dY_l - conv_K(S dZ_l)/S + 2 gamma (S dZ_l)/S
              if (p.gamma != 0.0f) {
                const uint32_t zr[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
#pragma unroll
                for (int e2 = 0; e2 < 8; ++e2) {
                  const float2 zf = unpack_f16x2(zr[e2]);
                  R[2 * e2] = fmaf(g2, zf.x, fmaf(-invS, __uint_as_float(r[2 * e2]), R[2 * e2]));
                  R[2 * e2 + 1] = fmaf(g2, zf.y, fmaf(-invS, __uint_as_float(r[2 * e2 + 1]), R[2 * e2 + 1]));
                }
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) R[e] = fmaf(-invS, __uint_as_float(r[e]), R[e]);
              }
              if (!last) {
                const uint32_t bits = mk_k[k];
                uint32_t pk[8];
#pragma unroll
                for (int e2 = 0; e2 < 8; ++e2) {
                  const float a0 = (bits >> (2 * e2)) & 1u ? hS * R[2 * e2] : 0.0f;
                  const float a1 = (bits >> (2 * e2 + 1)) & 1u ? hS * R[2 * e2 + 1] : 0.0f;
                  pk[e2] = pack_f16x2(a0, a1);
                }
                const uint4 u0 = make_uint4(pk[0], pk[1], pk[2], pk[3]), u1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                sts128u(nxt + so0, u0);
                sts128u(nxt + so1, u1);
                uint4* gp = reinterpret_cast<uint4*>(cp + (long long)pixl * C + c0);
                gp[0] = u0; gp[1] = u1;
              } else {
                float4* op = reinterpret_cast<float4*>(p.dx + img_off + (long long)pixl * C + c0);
#pragma unroll
                for (int j = 0; j < 4; ++j) op[j] = make_float4(R[4 * j], R[4 * j + 1], R[4 * j + 2], R[4 * j + 3]);
              }
            }
          }
          if (!last && (k % IPS) == IPS - 1) {
            // hand segment sg of the next strip (and its accumulator) back to the MMA warp
            tc_fence_before_sync();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&seg_done[sg]);
          }
        }
        if (threadIdx.x == 64 && lc == TL) tr.mark(8);
        if (last) {
          // Nobody leaves the image before ALL MMAs of its last step are complete: the next image's init rewrites
          // strip 0 (which they may read) and re-arms the segment barriers.
          mbar_wait_sleep_lean(&acc_full[p.nseg - 1], lc & 1);
          tc_fence_before_sync();
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
