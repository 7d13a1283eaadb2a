// Stride-2 transition blocks (reference models/tfkeras_resnets.py:204-269:
// out = relu(Conv2D 3x3 s2 (x) + bm) + Conv2D 1x1 s2 (x) + bs) as implicit GEMMs on warp-level tensor-core MMAs.
//
// Why warp-level mma.sync and not tcgen05 here: the two transitions hold 0.6 GFLOP of the step's 146 (cfg3); after the
// CUDA-core kernels of kernels_glue.cuh (10 TFLOP/s fp32, 26 % of the step time) the binding cost is staging a band of
// one image plus the 10 weight taps and the launch itself, not the MMA rate.  m16n8k8 tf32 MMAs with the 3xTF32 split
// (hi*hi + hi*lo + lo*hi, small terms first, fp32 accumulators) keep the result fp32-grade (<= 1e-5 relative against the
// fp32 oracle, tests/test_gpu_glue.py) for strict mode; the fast modes (FAST = true, the *_fast entry points) issue ONE tf32 MMA
// per product on operands rounded to nearest (2.9e-4 relative, the grade of their chains' operands); the stride-2 gather of the
// A operand is plain shared-memory addressing (a tcgen05 tile would need nine strided TMA boxes per 128 positions and a
// TMEM round trip for a K of 144 / 288).
//
// Fragment layout of mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 (PTX ISA, "Matrix Fragments for mma.m16n8k8"):
//   g = lane >> 2, t = lane & 3
//   A (16x8, row):  a0 = (g, t)   a1 = (g + 8, t)   a2 = (g, t + 4)   a3 = (g + 8, t + 4)
//   B (8x8, col):   b0 = (k = t, n = g)             b1 = (k = t + 4, n = g)
//   C/D (16x8):     c0 = (g, 2t)  c1 = (g, 2t + 1)  c2 = (g + 8, 2t)  c3 = (g + 8, 2t + 1)
// Shared-memory row strides are chosen so that the 32 lanes of every fragment load hit 32 different banks.
#pragma once

#include "kernels_glue.cuh"
#include "sm100_ptx.cuh"

namespace b200ode {

__device__ __forceinline__ void mma_tf32_m16n8k8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// v = hi + lo + O(2^-22 |v|): hi = v rounded to tf32 (nearest), lo = the remainder rounded to tf32
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  const float r = v - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], float b0, float b1) {
  uint32_t bh0, bl0, bh1, bl1;
  split_tf32(b0, bh0, bl0);
  split_tf32(b1, bh1, bl1);
  mma_tf32_m16n8k8(d, al, bh0, bh1);
  mma_tf32_m16n8k8(d, ah, bl0, bl1);
  mma_tf32_m16n8k8(d, ah, bh0, bh1);
}

// FAST = true: ONE tf32 MMA on operands rounded to nearest (10-bit significand + implicit bit, the grade of the fast modes'
// fp16 / tf32 chains); FAST = false: the 3xTF32 split above (fp32 grade).
template <bool FAST>
__device__ __forceinline__ void split_a(float v, uint32_t& hi, uint32_t& lo) {
  if constexpr (FAST) { asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v)); lo = 0u; }
  else split_tf32(v, hi, lo);
}
template <bool FAST>
__device__ __forceinline__ void mma_op(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], float b0, float b1) {
  if constexpr (FAST) {
    uint32_t bh0, bh1;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bh0) : "f"(b0));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bh1) : "f"(b1));
    mma_tf32_m16n8k8(d, ah, bh0, bh1);
  } else mma_3xtf32(d, ah, al, b0, b1);
}

// Staging: every global -> shared copy of a block is issued as cp.async (no register round trip), so ALL of a block's loads
// are in flight together and the block pays ONE memory round trip before its MMAs (the register-staged loops paid one per
// unrolled batch: ncu showed the kernels bound by long-scoreboard stalls of the staging phase, tensor pipe 18-25 % busy).
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// forward.  Block = (image, band of `orows` output rows); the band's input rows and all 10 weight taps (9 main + the 1x1
// shortcut) are staged in shared memory as fp32.  Warp task = (16 output positions, 1/NSPLIT of the output channels):
// M = positions, N = channels, K = 9 * CIN (main) and CIN (shortcut, second accumulator set).
// ---------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct TrFwdMma {
  static constexpr int PS = CIN + 2;     // pixel stride (floats): rows of a fragment are 2 pixels apart, 2 * PS = 4 (mod 32)
  static constexpr int WS = COUT + 8;    // weight row stride: k rows t = 0..3 land 8 banks apart
  static size_t smem_bytes(int orows, int W) {
    const int nir = (orows - 1) * 2 + 3;
    return ((size_t)(((nir * W + 1) * PS + 3) & ~3) + (size_t)10 * CIN * WS) * sizeof(float);
  }
};

template <int CIN, int COUT, int NSPLIT, bool FAST = false>
__global__ void __launch_bounds__(256) transition_fwd_mma_kernel(GlueConv g, const float* __restrict__ x, const float* __restrict__ Wm,
                                                                 const float* __restrict__ bm, const float* __restrict__ Ws,
                                                                 const float* __restrict__ bs, float* __restrict__ out,
                                                                 uint8_t* __restrict__ mask, int orows) {
  using Cfg = TrFwdMma<CIN, COUT>;
  constexpr int PS = Cfg::PS, WS = Cfg::WS, NT = COUT / 8 / NSPLIT, KS = CIN / 8;
  extern __shared__ float sm[];
  const int n = blockIdx.x;
  const int o0 = blockIdx.y * orows, o1 = min(g.Ho, o0 + orows);
  int i0 = o0 * 2 - g.pt; if (i0 < 0) i0 = 0;
  int i1 = (o1 - 1) * 2 + 2 - g.pt + 1; if (i1 > g.H) i1 = g.H;
  const int nir_max = (orows - 1) * 2 + 3;
  const int zp = nir_max * g.W;                       // index of the all-zero pixel (out-of-image taps)
  float* xs = sm;
  float* wsm = sm + (((nir_max * g.W + 1) * PS + 3) & ~3);
  griddep_launch_dependents();
  griddep_wait();
  {
    const float2* src = reinterpret_cast<const float2*>(x + ((long long)n * g.H + i0) * g.W * CIN);
    const int n2 = (i1 - i0) * g.W * (CIN / 2);      // pixel rows are 8-byte aligned only (PS = CIN + 2): 8-byte copies
    for (int i = threadIdx.x; i < n2; i += blockDim.x) cp_async_8(xs + (i / (CIN / 2)) * PS + (i % (CIN / 2)) * 2, src + i);
    if (threadIdx.x < PS) xs[zp * PS + threadIdx.x] = 0.0f;
    for (int i = threadIdx.x; i < 10 * CIN * (COUT / 4); i += blockDim.x) {
      const int row = i / (COUT / 4), c4 = i % (COUT / 4);
      const float* s = row < 9 * CIN ? Wm + (long long)row * COUT : Ws + (long long)(row - 9 * CIN) * COUT;
      cp_async_16(wsm + row * WS + 4 * c4, reinterpret_cast<const float4*>(s) + c4);
    }
    cp_async_wait_all();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int npos = (o1 - o0) * g.Wo;
  const int mtiles = (npos + 15) >> 4;
  for (int task = warp; task < mtiles * NSPLIT; task += nwarps) {
    const int mt = task / NSPLIT, nb = (task % NSPLIT) * NT * 8;
    const int p0 = mt * 16 + gq, p1 = p0 + 8;
    const bool ok0 = p0 < npos, ok1 = p1 < npos;
    const int oy0 = o0 + (ok0 ? p0 : 0) / g.Wo, ox0 = (ok0 ? p0 : 0) % g.Wo;
    const int oy1 = o0 + (ok1 ? p1 : 0) / g.Wo, ox1 = (ok1 ? p1 : 0) % g.Wo;
    float acc[NT][4], sc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc[j][e] = 0.0f; sc[j][e] = 0.0f; }
    // one tap (a, b) with weight rows wrow.. into the accumulator set d (two call sites: the tiles stay in registers)
    auto tap = [&](const int a, const int b, const int wrow, float (&d)[NT][4]) {
      const int iy0 = 2 * oy0 + a - g.pt, ix0 = 2 * ox0 + b - g.pl;
      const int iy1 = 2 * oy1 + a - g.pt, ix1 = 2 * ox1 + b - g.pl;
      const bool in0 = ok0 && iy0 >= 0 && iy0 < g.H && ix0 >= 0 && ix0 < g.W;
      const bool in1 = ok1 && iy1 >= 0 && iy1 < g.H && ix1 >= 0 && ix1 < g.W;
      const float* xa0 = xs + (in0 ? (iy0 - i0) * g.W + ix0 : zp) * PS + tq;
      const float* xa1 = xs + (in1 ? (iy1 - i0) * g.W + ix1 : zp) * PS + tq;
      const float* wb = wsm + (wrow + tq) * WS + nb + gq;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t ah[4], al[4];
        split_a<FAST>(xa0[ks * 8], ah[0], al[0]);
        split_a<FAST>(xa1[ks * 8], ah[1], al[1]);
        split_a<FAST>(xa0[ks * 8 + 4], ah[2], al[2]);
        split_a<FAST>(xa1[ks * 8 + 4], ah[3], al[3]);
#pragma unroll
        for (int j = 0; j < NT; ++j) mma_op<FAST>(d[j], ah, al, wb[(ks * 8) * WS + j * 8], wb[(ks * 8 + 4) * WS + j * 8]);
      }
    };
#pragma unroll 1
    for (int t = 0; t < 9; ++t) tap(t / 3, t % 3, t * CIN, acc);
    tap(g.pt, g.pl, 9 * CIN, sc);      // the 1x1 stride-2 convolution samples x[2 oy, 2 ox]
    const long long opix0 = ((long long)n * g.Ho + o0) * g.Wo + p0, opix1 = opix0 + 8;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int c = nb + j * 8 + 2 * tq;
      const float2 bmv = __ldg(reinterpret_cast<const float2*>(bm + c)), bsv = __ldg(reinterpret_cast<const float2*>(bs + c));
      // reference layer order: relu(conv + bm) and (conv1x1 + bs) are two layers' outputs, then add()
      const float m00 = acc[j][0] + bmv.x, m01 = acc[j][1] + bmv.y, m10 = acc[j][2] + bmv.x, m11 = acc[j][3] + bmv.y;
      uint32_t b0 = ((m00 > 0.0f ? 1u : 0u) | (m01 > 0.0f ? 2u : 0u)) << (2 * tq);
      uint32_t b1 = ((m10 > 0.0f ? 1u : 0u) | (m11 > 0.0f ? 2u : 0u)) << (2 * tq);
      b0 |= __shfl_xor_sync(0xffffffffu, b0, 1); b0 |= __shfl_xor_sync(0xffffffffu, b0, 2);
      b1 |= __shfl_xor_sync(0xffffffffu, b1, 1); b1 |= __shfl_xor_sync(0xffffffffu, b1, 2);
      if (ok0) {
        *reinterpret_cast<float2*>(out + opix0 * COUT + c) = make_float2(fmaxf(m00, 0.0f) + (sc[j][0] + bsv.x), fmaxf(m01, 0.0f) + (sc[j][1] + bsv.y));
        if (tq == 0) mask[opix0 * (COUT / 8) + (nb >> 3) + j] = (uint8_t)b0;
      }
      if (ok1) {
        *reinterpret_cast<float2*>(out + opix1 * COUT + c) = make_float2(fmaxf(m10, 0.0f) + (sc[j][2] + bsv.x), fmaxf(m11, 0.0f) + (sc[j][3] + bsv.y));
        if (tq == 0) mask[opix1 * (COUT / 8) + (nb >> 3) + j] = (uint8_t)b1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// data gradient.  With stride 2 an input pixel (y, x) receives the taps a = (y + pt) mod 2 (+ 2), b likewise: four
// parity classes with 4 / 2 / 2 / 1 main taps (the shortcut joins the class of (2 oy, 2 ox)).  Input pixels are grouped in
// 2x2 CELLS (cy = (y + pt) >> 1, ...): a warp task = 16 consecutive cells x 1/NSPLIT of the input channels and walks the
// four classes of its cells, so every task carries the same 9 + 1 taps.  Per class: M = 16 pixels, N = CIN slice,
// K = COUT per tap; A = dout * relu mask (main taps) or dout (shortcut) of the output pixel the tap reaches.
// Block = (image, band of `crows` cell rows); staged: the output rows the band reaches (raw and masked), the 10 taps.
// ---------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct TrDgradMma {
  static constexpr int PSO = COUT + 4;   // consecutive cells read consecutive output pixels: PSO = 4 (mod 32)
  static constexpr int WS = COUT + 4;    // B rows are input channels, k (= co) contiguous
  static size_t smem_bytes(int crows, int Wo) {
    return ((size_t)2 * ((crows + 1) * Wo + 1) * PSO + (size_t)10 * CIN * WS) * sizeof(float) + (size_t)(crows + 1) * Wo * (COUT / 8);
  }
};

template <int CIN, int COUT, int NSPLIT, bool FAST = false>
__global__ void __launch_bounds__(256) transition_dgrad_mma_kernel(GlueConv g, const float* __restrict__ dout, const uint8_t* __restrict__ mask,
                                                                   const float* __restrict__ Wm, const float* __restrict__ Ws,
                                                                   float* __restrict__ dx, int crows, unsigned int* __restrict__ dx_amax) {
  using Cfg = TrDgradMma<CIN, COUT>;
  float amx = 0.0f;       // max |dx| over what this thread stores (dx_amax != NULL: for the fp16 chain behind the transition)
  constexpr int PSO = Cfg::PSO, WS = Cfg::WS, NT = CIN / 8 / NSPLIT, KS = COUT / 8;
  extern __shared__ float sm[];
  const int n = blockIdx.x;
  const int CH = (g.H - 1 + g.pt) / 2 + 1, CW = (g.W - 1 + g.pl) / 2 + 1;
  const int c0 = blockIdx.y * crows, c1 = min(CH, c0 + crows);
  const int r0 = c0 > 0 ? c0 - 1 : 0, r1 = min(g.Ho, c1);          // output rows [r0, r1) reach the band
  const int zp = (crows + 1) * g.Wo;
  float* dO = sm;
  float* dM = dO + (zp + 1) * PSO;
  float* wsm = dM + (zp + 1) * PSO;
  uint8_t* msm = reinterpret_cast<uint8_t*>(wsm + 10 * CIN * WS);      // relu mask bytes of the staged rows
  griddep_launch_dependents();
  griddep_wait();
  {
    const long long base = ((long long)n * g.Ho + r0) * g.Wo * COUT;
    const float4* d4 = reinterpret_cast<const float4*>(dout + base);
    const int n4 = r1 > r0 ? (r1 - r0) * g.Wo * (COUT / 4) : 0;
    // the masked copy dM is staged as a RAW copy first and masked in place once everything has landed (mask bytes in
    // registers: one 4-bit group per float4, loaded up front by the thread that will mask that float4)
    for (int i4 = threadIdx.x; i4 < n4; i4 += blockDim.x) {
      const int i = 4 * i4;
      const int si = (i / COUT) * PSO + i % COUT;
      cp_async_16(dO + si, d4 + i4);
      cp_async_16(dM + si, d4 + i4);
    }
    for (int i = threadIdx.x; i < n4 / 8; i += blockDim.x) cp_async_4(msm + 4 * i, mask + (base >> 3) + 4 * i);   // n4 / 2 mask bytes
    if (threadIdx.x < PSO) { dO[zp * PSO + threadIdx.x] = 0.0f; dM[zp * PSO + threadIdx.x] = 0.0f; }
    for (int i = threadIdx.x; i < 10 * CIN * (COUT / 4); i += blockDim.x) {
      const int row = i / (COUT / 4), c4 = i % (COUT / 4);
      const float* s = row < 9 * CIN ? Wm + (long long)row * COUT : Ws + (long long)(row - 9 * CIN) * COUT;
      cp_async_16(wsm + row * WS + 4 * c4, reinterpret_cast<const float4*>(s) + c4);
    }
    cp_async_wait_all();
    __syncthreads();              // everybody's copies have landed (the mask bytes were copied by other threads)
    for (int i4 = threadIdx.x; i4 < n4; i4 += blockDim.x) {
      const int i = 4 * i4;
      const uint32_t mb = (uint32_t)msm[i >> 3] >> (i & 4);
      float4* q = reinterpret_cast<float4*>(dM + (i / COUT) * PSO + i % COUT);
      const float4 d = *q;
      *q = make_float4(mb & 1u ? d.x : 0.0f, mb & 2u ? d.y : 0.0f, mb & 4u ? d.z : 0.0f, mb & 8u ? d.w : 0.0f);
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int ncell = (c1 - c0) * CW;
  const int mtiles = (ncell + 15) >> 4;
  const int scy = g.pt & 1, scx = g.pl & 1;          // class of the pixels the 1x1 shortcut samples
  for (int task = warp; task < mtiles * NSPLIT; task += nwarps) {
    const int mt = task / NSPLIT, nb = (task % NSPLIT) * NT * 8;
    const int q0 = mt * 16 + gq, q1 = q0 + 8;
    const bool ok0 = q0 < ncell, ok1 = q1 < ncell;
    const int cy0 = c0 + (ok0 ? q0 : 0) / CW, cx0 = (ok0 ? q0 : 0) % CW;
    const int cy1 = c0 + (ok1 ? q1 : 0) / CW, cx1 = (ok1 ? q1 : 0) % CW;
#pragma unroll
    for (int cls = 0; cls < 4; ++cls) {
      const int py = cls >> 1, px = cls & 1;
      float acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.0f;
      // tap reaching output pixel (cell - (da, db)) with weight rows wrow.. and A operand from strip S
      auto tap = [&](const int da, const int db, const int wrow, const float* S) {
        const int oy0 = cy0 - da, ox0 = cx0 - db, oy1 = cy1 - da, ox1 = cx1 - db;
        const bool in0 = ok0 && oy0 >= r0 && oy0 < r1 && ox0 >= 0 && ox0 < g.Wo;
        const bool in1 = ok1 && oy1 >= r0 && oy1 < r1 && ox1 >= 0 && ox1 < g.Wo;
        const float* pa0 = S + (in0 ? (oy0 - r0) * g.Wo + ox0 : zp) * PSO + tq;
        const float* pa1 = S + (in1 ? (oy1 - r0) * g.Wo + ox1 : zp) * PSO + tq;
        const float* pb = wsm + (wrow + nb + gq) * WS + tq;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t ah[4], al[4];
          split_a<FAST>(pa0[ks * 8], ah[0], al[0]);
          split_a<FAST>(pa1[ks * 8], ah[1], al[1]);
          split_a<FAST>(pa0[ks * 8 + 4], ah[2], al[2]);
          split_a<FAST>(pa1[ks * 8 + 4], ah[3], al[3]);
#pragma unroll
          for (int j = 0; j < NT; ++j) mma_op<FAST>(acc[j], ah, al, pb[j * 8 * WS + ks * 8], pb[j * 8 * WS + ks * 8 + 4]);
        }
      };
#pragma unroll
      for (int a = py; a < 3; a += 2)
#pragma unroll
        for (int b = px; b < 3; b += 2) tap((a - py) >> 1, (b - px) >> 1, (a * 3 + b) * CIN, dM);
      if (py == scy && px == scx) tap(0, 0, 9 * CIN, dO);
      const int y0 = 2 * cy0 + py - g.pt, x0 = 2 * cx0 + px - g.pl;
      const int y1 = 2 * cy1 + py - g.pt, x1 = 2 * cx1 + px - g.pl;
      const bool st0 = ok0 && y0 >= 0 && y0 < g.H && x0 >= 0 && x0 < g.W;
      const bool st1 = ok1 && y1 >= 0 && y1 < g.H && x1 >= 0 && x1 < g.W;
      float* d0 = dx + (((long long)n * g.H + y0) * g.W + x0) * CIN + nb + 2 * tq;
      float* d1 = dx + (((long long)n * g.H + y1) * g.W + x1) * CIN + nb + 2 * tq;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        if (st0) { *reinterpret_cast<float2*>(d0 + j * 8) = make_float2(acc[j][0], acc[j][1]); amx = fmaxf(amx, fmaxf(fabsf(acc[j][0]), fabsf(acc[j][1]))); }
        if (st1) { *reinterpret_cast<float2*>(d1 + j * 8) = make_float2(acc[j][2], acc[j][3]); amx = fmaxf(amx, fmaxf(fabsf(acc[j][2]), fabsf(acc[j][3]))); }
      }
    }
  }
  if (dx_amax) {          // one atomic per block
    __shared__ float wmax[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amx = fmaxf(amx, __shfl_xor_sync(0xffffffffu, amx, o));
    if (lane == 0) wmax[warp] = amx;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = 0.0f;
      for (int w = 0; w < nwarps; ++w) m = fmaxf(m, wmax[w]);
      if (m > 0.0f) atomicMax(dx_amax, __float_as_uint(m));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient partials.  Per tap: dW[ci][co] = sum_pos x[pixel(pos, tap)][ci] * dmain[pos][co]: M = CIN, N = COUT,
// K = output positions.  Warp task = (tap, 16-channel block of ci) for the 9 main taps and the shortcut, plus two
// "ones" tasks whose A operand is all ones (exact in tf32): their accumulator rows are the column sums of dmain / dout,
// i.e. the two bias gradients.  A block walks its images band by band (stage -> sync -> MMAs), keeps its accumulators in
// registers across bands and images and writes ONE partial row [dWm 9*CIN*COUT | dbm COUT | dWs CIN*COUT | dbs COUT];
// reduce_rows sums the rows in fixed order (deterministic).
// ---------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct TrWgradMma {
  static constexpr int PS = CIN + 4;     // A columns (k = position) are 2 pixels apart: 2 * PS = 8 (mod 32)
  static constexpr int PSO = COUT + 8;   // B rows (k = position) land 8 banks apart
  static constexpr int NTASK = 10 * (CIN / 16) + 2;
  static constexpr int NWARP = 12;
  static constexpr int TPW = (NTASK + NWARP - 1) / NWARP;
  static size_t smem_bytes(int orows, int W, int Wo) {
    const int nir = (orows - 1) * 2 + 3, npad = (orows * Wo + 7) & ~7;
    return ((size_t)(nir * W + 1) * PS + (size_t)2 * npad * PSO) * sizeof(float) + (size_t)npad * (COUT / 8);
  }
};

template <int CIN, int COUT, bool FAST = false>
__global__ void __launch_bounds__(384) transition_wgrad_mma_kernel(GlueConv g, const float* __restrict__ x, const float* __restrict__ dout,
                                                                   const uint8_t* __restrict__ mask, float* __restrict__ part,
                                                                   int orows, int ipb) {
  using Cfg = TrWgradMma<CIN, COUT>;
  constexpr int PS = Cfg::PS, PSO = Cfg::PSO, NT = COUT / 8, MT = CIN / 16, NTASK = Cfg::NTASK, TPW = Cfg::TPW;
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int nir_max = (orows - 1) * 2 + 3;
  const int zp = nir_max * g.W;
  const int npad_max = (orows * g.Wo + 7) & ~7;
  float* xs = sm;                                   // [nir_max * W + 1][PS]
  float* dO = xs + (zp + 1) * PS;                   // [npad_max][PSO]
  float* dM = dO + npad_max * PSO;
  uint8_t* msm = reinterpret_cast<uint8_t*>(dM + npad_max * PSO);      // relu mask bytes of the staged band
  float acc[TPW][NT][4];
#pragma unroll
  for (int u = 0; u < TPW; ++u)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[u][j][e] = 0.0f;
  if (threadIdx.x < PS) xs[zp * PS + threadIdx.x] = 0.0f;
  const int n_end = min(g.N, (int)(blockIdx.x + 1) * ipb);
  for (int n = blockIdx.x * ipb; n < n_end; ++n) {
    for (int o0 = 0; o0 < g.Ho; o0 += orows) {
      const int o1 = min(g.Ho, o0 + orows);
      int i0 = o0 * 2 - g.pt; if (i0 < 0) i0 = 0;
      int i1 = (o1 - 1) * 2 + 2 - g.pt + 1; if (i1 > g.H) i1 = g.H;
      const int npos = (o1 - o0) * g.Wo, npad = (npos + 7) & ~7;
      __syncthreads();        // the previous band's MMAs are done with the strips
      {
        const float4* src = reinterpret_cast<const float4*>(x + ((long long)n * g.H + i0) * g.W * CIN);
        const int n4 = (i1 - i0) * g.W * (CIN / 4);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) cp_async_16(xs + (i / (CIN / 4)) * PS + (i % (CIN / 4)) * 4, src + i);
        const long long base = ((long long)n * g.Ho + o0) * g.Wo * COUT;
        const float4* d4 = reinterpret_cast<const float4*>(dout + base);
        for (int i4 = threadIdx.x; i4 < npad * (COUT / 4); i4 += blockDim.x) {
          const int i = 4 * i4;
          const int si = (i / COUT) * PSO + i % COUT;
          if (i4 < npos * (COUT / 4)) {
            cp_async_16(dO + si, d4 + i4);
            cp_async_16(dM + si, d4 + i4);
          } else {                                                     // positions past the band: zeros (0 * garbage could be NaN)
            *reinterpret_cast<float4*>(dO + si) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            *reinterpret_cast<float4*>(dM + si) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          }
        }
        for (int i = threadIdx.x; i < npos * (COUT / 32); i += blockDim.x) cp_async_4(msm + 4 * i, mask + (base >> 3) + 4 * i);
        cp_async_wait_all();
        __syncthreads();          // everybody's copies have landed (the mask bytes were copied by other threads)
        for (int i4 = threadIdx.x; i4 < npos * (COUT / 4); i4 += blockDim.x) {
          const int i = 4 * i4;
          const uint32_t mb = (uint32_t)msm[i >> 3] >> (i & 4);
          float4* q = reinterpret_cast<float4*>(dM + (i / COUT) * PSO + i % COUT);
          const float4 d = *q;
          *q = make_float4(mb & 1u ? d.x : 0.0f, mb & 2u ? d.y : 0.0f, mb & 4u ? d.z : 0.0f, mb & 8u ? d.w : 0.0f);
        }
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < TPW; ++u) {
        const int task = warp + u * nwarps;
        if (task >= NTASK) continue;                 // warp-uniform
        if (task < 10 * MT) {
          const int tp = task / MT, cb = (task % MT) * 16;
          const int a = tp == 9 ? g.pt : tp / 3, b = tp == 9 ? g.pl : tp % 3;
          const float* S = tp == 9 ? dO : dM;
          for (int k0 = 0; k0 < npad; k0 += 8) {
            const int pA = k0 + tq, pB = pA + 4;
            const int iyA = 2 * (o0 + pA / g.Wo) + a - g.pt, ixA = 2 * (pA % g.Wo) + b - g.pl;
            const int iyB = 2 * (o0 + pB / g.Wo) + a - g.pt, ixB = 2 * (pB % g.Wo) + b - g.pl;
            const bool inA = pA < npos && iyA >= 0 && iyA < g.H && ixA >= 0 && ixA < g.W;
            const bool inB = pB < npos && iyB >= 0 && iyB < g.H && ixB >= 0 && ixB < g.W;
            const float* xa = xs + (inA ? (iyA - i0) * g.W + ixA : zp) * PS + cb + gq;
            const float* xb = xs + (inB ? (iyB - i0) * g.W + ixB : zp) * PS + cb + gq;
            uint32_t ah[4], al[4];
            split_a<FAST>(xa[0], ah[0], al[0]);
            split_a<FAST>(xa[8], ah[1], al[1]);
            split_a<FAST>(xb[0], ah[2], al[2]);
            split_a<FAST>(xb[8], ah[3], al[3]);
            const float* pb = S + pA * PSO + gq;
#pragma unroll
            for (int j = 0; j < NT; ++j) mma_op<FAST>(acc[u][j], ah, al, pb[j * 8], pb[4 * PSO + j * 8]);
          }
        } else {
          const float* S = task == 10 * MT ? dM : dO;
          const uint32_t one = __float_as_uint(1.0f);
          const uint32_t ah[4] = {one, one, one, one};
          for (int k0 = 0; k0 < npad; k0 += 8) {
            const float* pb = S + (k0 + tq) * PSO + gq;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              uint32_t bh0, bl0, bh1, bl1;
              split_tf32(pb[j * 8], bh0, bl0);
              split_tf32(pb[4 * PSO + j * 8], bh1, bl1);
              mma_tf32_m16n8k8(acc[u][j], ah, bl0, bl1);
              mma_tf32_m16n8k8(acc[u][j], ah, bh0, bh1);
            }
          }
        }
      }
    }
  }
  const long long nm = 9LL * CIN * COUT;
  float* prow = part + (long long)blockIdx.x * (nm + COUT + (long long)CIN * COUT + COUT);
#pragma unroll
  for (int u = 0; u < TPW; ++u) {
    const int task = warp + u * nwarps;
    if (task >= NTASK) continue;
    if (task < 10 * MT) {
      const int tp = task / MT, cb = (task % MT) * 16;
      float* dst = (tp == 9 ? prow + nm + COUT : prow + (long long)tp * CIN * COUT) + (long long)(cb + gq) * COUT + 2 * tq;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        *reinterpret_cast<float2*>(dst + j * 8) = make_float2(acc[u][j][0], acc[u][j][1]);
        *reinterpret_cast<float2*>(dst + 8 * COUT + j * 8) = make_float2(acc[u][j][2], acc[u][j][3]);
      }
    } else if (gq == 0) {
      float* dst = (task == 10 * MT ? prow + nm : prow + nm + COUT + (long long)CIN * COUT) + 2 * tq;
#pragma unroll
      for (int j = 0; j < NT; ++j) *reinterpret_cast<float2*>(dst + j * 8) = make_float2(acc[u][j][0], acc[u][j][1]);
    }
  }
}

}  // namespace b200ode
