// Stem, transition and head layers of the single-block ResNet: the few
// non-antisymmetric layers either side of the Euler chains, as fused fp32 CUDA-core kernels.
//
// Reference semantics (paths relative to the reference repository):
//   input Lambda layers + stem   models/tfkeras_resnets.py:555-572  ((x-mean)/std -> Conv2D SAME -> relu)
//   transition block             models/tfkeras_resnets.py:204-269  (relu(Conv2D k s) + Conv2D 1x1 s; no h)
//   head                         models/tfkeras_resnets.py:595-597  (GlobalAveragePooling2D -> Dense softmax)
//   loss                         training/training.py:295           (mean K.categorical_crossentropy, eps 1e-7)
// These layers hold < 1% of the step's FLOPs; the kernels are simple direct convolutions whose job is
// to replace ~100 small library launches per step by 10.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "sm100_ptx.cuh"

namespace b200ode {

struct GlueConv {
  int N, H, W, Cin, Ho, Wo, Cout, sh, sw, pt, pl;   // 3x3 main conv geometry (TF SAME: pad_before = total/2)
};

// ---------------------------------------------------------------------------------------------
// stem: out = relu(conv3x3_SAME((img - sub) / div) + b), stride 1.  One thread per (pixel, 4 channels).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float stem_in(const void* img, int is_u8, long long idx, float sub, float div, int norm) {
  const float v = is_u8 ? (float)reinterpret_cast<const uint8_t*>(img)[idx] : reinterpret_cast<const float*>(img)[idx];
  // two Lambda layers in the reference: subtract, then divide (two roundings)
  return norm ? __fdiv_rn(__fsub_rn(v, sub), div) : v;
}

__global__ void stem_fwd_kernel(GlueConv g, const void* __restrict__ img, int is_u8, float sub, float div, int norm,
                                const float* __restrict__ Wk, const float* __restrict__ bias, float* __restrict__ out) {
  const int q4 = g.Cout / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.N * g.H * g.W * q4;
  if (idx >= total) return;
  const int c0 = (int)(idx % q4) * 4;
  long long p = idx / q4;
  const int x = (int)(p % g.W); p /= g.W;
  const int y = (int)(p % g.H);
  const int n = (int)(p / g.H);
  float4 acc = *reinterpret_cast<const float4*>(bias + c0);
  for (int a = 0; a < 3; ++a) {
    const int iy = y + a - 1;
    if (iy < 0 || iy >= g.H) continue;
    for (int b = 0; b < 3; ++b) {
      const int ix = x + b - 1;
      if (ix < 0 || ix >= g.W) continue;
      const long long ib = (((long long)n * g.H + iy) * g.W + ix) * g.Cin;
      for (int ci = 0; ci < g.Cin; ++ci) {
        const float v = stem_in(img, is_u8, ib + ci, sub, div, norm);
        const float4 w = __ldg(reinterpret_cast<const float4*>(Wk + ((a * 3 + b) * g.Cin + ci) * g.Cout + c0));
        acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y); acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
      }
    }
  }
  acc.x = fmaxf(acc.x, 0.0f); acc.y = fmaxf(acc.y, 0.0f); acc.z = fmaxf(acc.z, 0.0f); acc.w = fmaxf(acc.w, 0.0f);
  *reinterpret_cast<float4*>(out + (((long long)n * g.H + y) * g.W + x) * g.Cout + c0) = acc;
}

// Stem forward, one thread per output pixel computing all COUT channels (COUT = 8, 16 or 32): the 27 * COUT weights and,
// for uint8 images, the 256-entry table of normalised input values ((v - sub) / div: two IEEE operations per input
// element in the generic kernel, 27 divisions per thread) live in shared memory; weight reads are warp-uniform broadcasts.
// Same accumulation order as stem_fwd_kernel (bias, then taps row-major, then input channels): bit-identical results.
template <int COUT>
__global__ void __launch_bounds__(256) stem_fwd_pixel_kernel(GlueConv g, const void* __restrict__ img, int is_u8, float sub, float div,
                                                             int norm, const float* __restrict__ Wk, const float* __restrict__ bias,
                                                             float* __restrict__ out) {
  extern __shared__ float sm[];
  griddep_launch_dependents();          // (the chain kernel behind the stem waits before it reads `out`)
  float* wsm = sm;                      // [9 * Cin][COUT]
  float* bsm = wsm + 9 * g.Cin * COUT;  // [COUT]
  float* lut = bsm + COUT;              // [256] (uint8 input)
  for (int i = threadIdx.x; i < 9 * g.Cin * COUT; i += blockDim.x) wsm[i] = Wk[i];
  if (threadIdx.x < COUT) bsm[threadIdx.x] = bias[threadIdx.x];
  if (is_u8) {
    const float v = (float)threadIdx.x;
    if (threadIdx.x < 256) lut[threadIdx.x] = norm ? __fdiv_rn(__fsub_rn(v, sub), div) : v;
  }
  __syncthreads();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)g.N * g.H * g.W) return;
  const int x = (int)(pix % g.W), y = (int)((pix / g.W) % g.H);
  const long long n = pix / ((long long)g.W * g.H);
  float acc[COUT];
#pragma unroll
  for (int j = 0; j < COUT; ++j) acc[j] = bsm[j];
  const uint8_t* i8 = reinterpret_cast<const uint8_t*>(img);
  const float* i32 = reinterpret_cast<const float*>(img);
  for (int a = 0; a < 3; ++a) {
    const int iy = y + a - 1;
    if (iy < 0 || iy >= g.H) continue;
    for (int b = 0; b < 3; ++b) {
      const int ix = x + b - 1;
      if (ix < 0 || ix >= g.W) continue;
      const long long ib = ((n * g.H + iy) * g.W + ix) * g.Cin;
      const float* wt = wsm + (a * 3 + b) * g.Cin * COUT;
      for (int ci = 0; ci < g.Cin; ++ci) {
        const float v = is_u8 ? lut[i8[ib + ci]] : (norm ? __fdiv_rn(__fsub_rn(i32[ib + ci], sub), div) : i32[ib + ci]);
        const float4* w4 = reinterpret_cast<const float4*>(wt + ci * COUT);
#pragma unroll
        for (int j = 0; j < COUT / 4; ++j) {
          const float4 w = w4[j];
          acc[4 * j] = fmaf(v, w.x, acc[4 * j]); acc[4 * j + 1] = fmaf(v, w.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(v, w.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(v, w.w, acc[4 * j + 3]);
        }
      }
    }
  }
  float4* op = reinterpret_cast<float4*>(out + pix * COUT);
#pragma unroll
  for (int j = 0; j < COUT / 4; ++j)
    op[j] = make_float4(fmaxf(acc[4 * j], 0.0f), fmaxf(acc[4 * j + 1], 0.0f), fmaxf(acc[4 * j + 2], 0.0f), fmaxf(acc[4 * j + 3], 0.0f));
}

// stem weight gradient partials: block = (image, band of `rows` output rows).  The band's masked dz
// (dout * [out > 0]) and the normalised input patch (zero halo) are staged in shared memory; thread t
// owns kernel entries o = t, t + blockDim, ... (and the Cout bias entries) and walks the band's pixels.
__global__ void stem_wgrad_partial(GlueConv g, const void* __restrict__ img, int is_u8, float sub, float div, int norm,
                                   const float* __restrict__ out, const float* __restrict__ dout, float* __restrict__ part,
                                   int rows) {
  extern __shared__ float sm[];
  const int n = blockIdx.x;
  const int y0 = blockIdx.y * rows, y1 = min(g.H, y0 + rows);
  const int nr = y1 - y0;
  const int PW = g.W + 2;
  float* dz = sm;                                   // [rows*W][Cout]
  float* xin = sm + rows * g.W * g.Cout;            // [rows+2][W+2][Cin]
  {
    const float4* o4 = reinterpret_cast<const float4*>(out + ((long long)n * g.H + y0) * g.W * g.Cout);
    const float4* d4 = reinterpret_cast<const float4*>(dout + ((long long)n * g.H + y0) * g.W * g.Cout);
#pragma unroll 4
    for (int i = threadIdx.x; i < nr * g.W * g.Cout / 4; i += blockDim.x) {
      const float4 o = o4[i], d = d4[i];
      reinterpret_cast<float4*>(dz)[i] = make_float4(o.x > 0.0f ? d.x : 0.0f, o.y > 0.0f ? d.y : 0.0f,
                                                     o.z > 0.0f ? d.z : 0.0f, o.w > 0.0f ? d.w : 0.0f);
    }
  }
  for (int i = threadIdx.x; i < (nr + 2) * PW * g.Cin; i += blockDim.x) {
    const int ci = i % g.Cin, px = (i / g.Cin) % PW - 1, py = i / (g.Cin * PW) + y0 - 1;
    xin[i] = (py >= 0 && py < g.H && px >= 0 && px < g.W)
                 ? stem_in(img, is_u8, (((long long)n * g.H + py) * g.W + px) * g.Cin + ci, sub, div, norm) : 0.0f;
  }
  __syncthreads();
  const int nk = 9 * g.Cin * g.Cout;
  const int nout = nk + g.Cout;
  for (int o = threadIdx.x; o < nout; o += blockDim.x) {
    float acc = 0.0f;
    if (o < nk) {
      const int co = o % g.Cout, ci = (o / g.Cout) % g.Cin, tap = o / (g.Cout * g.Cin);
      const int a = tap / 3, b = tap % 3;
      for (int y = 0; y < nr; ++y) {
        const float* xr = xin + ((y + a) * PW + b) * g.Cin + ci;
        const float* dr = dz + y * g.W * g.Cout + co;
#pragma unroll 4
        for (int x = 0; x < g.W; ++x) acc = fmaf(xr[x * g.Cin], dr[x * g.Cout], acc);
      }
    } else {
      const int co = o - nk;
      for (int pix = 0; pix < nr * g.W; ++pix) acc += dz[pix * g.Cout + co];
    }
    part[((long long)n * gridDim.y + blockIdx.y) * nout + o] = acc;
  }
}

// out[o] = sum_r part[r*stride + o], o < n (deterministic; R small)
__global__ void reduce_rows_kernel(const float* __restrict__ part, int R, long long stride, long long n, float* __restrict__ out) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  float acc = 0.0f;
  for (int r = 0; r < R; ++r) acc += part[(long long)r * stride + o];
  out[o] = acc;
}

// ---------------------------------------------------------------------------------------------
// transition block: out = relu(conv3x3_s(x) + bm) + conv1x1_s(x) + bs; mask bit = (main > 0).
//
// All three kernels work on one band of rows of one image staged in shared memory (pixel rows padded
// by 4 floats so that lanes = pixels read conflict-free float4s) and keep the weight reads
// warp-uniform (every lane of a warp needs the same W[tap][ci][co..co+3]: one L1 transaction).
// ---------------------------------------------------------------------------------------------
// forward: block = (image, band of output rows, group of G channel slices); warp w of the block owns the
// COT-channel slice (blockIdx.z * G + w) whose weights W[10 taps][Cin][COT] sit in shared memory; lane l
// computes the PT pixels l, l+32, .. of the band (PT*COT register accumulators, ~10 FMAs per shared load).
// S > 0: both strides equal S at compile time (the address arithmetic of every tap loses its integer divisions; the
// round-1 kernels issued ~4 instructions per useful FMA, ncu: issue slots 57 % busy at 25 % FMA share); S = 0: run-time strides.
template <int COT, int PT, int S = 0>
__global__ void transition_fwd_kernel(GlueConv g, const float* __restrict__ x, const float* __restrict__ Wm,
                                      const float* __restrict__ bm, const float* __restrict__ Ws,
                                      const float* __restrict__ bs, float* __restrict__ out, uint8_t* __restrict__ mask,
                                      int orows) {
  extern __shared__ float sm[];
  const int n = blockIdx.x;
  const int o0 = blockIdx.y * orows, o1 = min(g.Ho, o0 + orows);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, G = blockDim.x >> 5;
  const int c0 = (blockIdx.z * G + warp) * COT;
  const int gsh = S ? S : g.sh, gsw = S ? S : g.sw;
  int i0 = o0 * gsh - g.pt; if (i0 < 0) i0 = 0;
  int i1 = (o1 - 1) * gsh + 2 - g.pt + 1; if (i1 > g.H) i1 = g.H;
  const int PS = g.Cin + 4;                          // padded pixel stride (floats)
  float* xs = sm;
  float* wsm = sm + ((orows - 1) * gsh + 3) * g.W * PS + warp * 10 * g.Cin * COT;    // this warp's weight slice
#pragma unroll 4
  for (int i = threadIdx.x; i < (i1 - i0) * g.W * (g.Cin / 4); i += blockDim.x) {
    const int c4 = i % (g.Cin / 4), pix = i / (g.Cin / 4);
    *reinterpret_cast<float4*>(xs + pix * PS + c4 * 4) =
        *reinterpret_cast<const float4*>(x + (((long long)n * g.H + i0) * g.W + pix) * g.Cin + c4 * 4);
  }
#pragma unroll 4
  for (int i = lane; i < 10 * g.Cin * (COT / 4); i += 32) {
    const int j = i % (COT / 4), ci = (i / (COT / 4)) % g.Cin, t = i / ((COT / 4) * g.Cin);
    const float* src = t == 9 ? Ws + (long long)ci * g.Cout : Wm + (long long)(t * g.Cin + ci) * g.Cout;
    *reinterpret_cast<float4*>(wsm + (t * g.Cin + ci) * COT + 4 * j) = __ldg(reinterpret_cast<const float4*>(src + c0) + j);
  }
  __syncthreads();
  const int npix = (o1 - o0) * g.Wo;
  for (int pb = 0; pb < npix; pb += 32 * PT) {
    float acc[PT][COT], sc[PT][COT];
    int oy[PT], ox[PT];
    bool ok[PT];
#pragma unroll
    for (int k = 0; k < PT; ++k) {
      const int p = pb + lane + 32 * k;
      ok[k] = p < npix;
      oy[k] = o0 + (ok[k] ? p : 0) / g.Wo; ox[k] = (ok[k] ? p : 0) % g.Wo;
#pragma unroll
      for (int j = 0; j < COT; ++j) { acc[k][j] = __ldg(bm + c0 + j); sc[k][j] = __ldg(bs + c0 + j); }
    }
    // one tap (a, b) of weight slice wp into the register tile d (taps 0..8 -> acc, the 1x1 shortcut -> sc: two call sites,
    // so the tile is never reached through a run-time pointer)
    auto tap = [&](const int a, const int b, const float* wp, float (&d)[PT][COT]) {
      const float* xp[PT];
      bool in[PT];
#pragma unroll
      for (int k = 0; k < PT; ++k) {
        const int iy = oy[k] * gsh + a - g.pt, ix = ox[k] * gsw + b - g.pl;
        in[k] = ok[k] && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
        xp[k] = xs + ((in[k] ? iy - i0 : 0) * g.W + (in[k] ? ix : 0)) * PS;
      }
      for (int ci = 0; ci < g.Cin; ci += 4) {
        float xv[PT][4];
#pragma unroll
        for (int k = 0; k < PT; ++k) {
          const float4 v = *reinterpret_cast<const float4*>(xp[k] + ci);
          xv[k][0] = in[k] ? v.x : 0.0f; xv[k][1] = in[k] ? v.y : 0.0f; xv[k][2] = in[k] ? v.z : 0.0f; xv[k][3] = in[k] ? v.w : 0.0f;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
          for (int j = 0; j < COT / 4; ++j) {
            const float4 w = *reinterpret_cast<const float4*>(wp + (ci + e) * COT + 4 * j);
#pragma unroll
            for (int k = 0; k < PT; ++k) {
              d[k][4 * j] = fmaf(xv[k][e], w.x, d[k][4 * j]); d[k][4 * j + 1] = fmaf(xv[k][e], w.y, d[k][4 * j + 1]);
              d[k][4 * j + 2] = fmaf(xv[k][e], w.z, d[k][4 * j + 2]); d[k][4 * j + 3] = fmaf(xv[k][e], w.w, d[k][4 * j + 3]);
            }
          }
        }
      }
    };
#pragma unroll
    for (int t = 0; t < 9; ++t) tap(t / 3, t % 3, wsm + t * g.Cin * COT, acc);   // same order as before: bit-identical sums
    tap(g.pt, g.pl, wsm + 9 * g.Cin * COT, sc);
#pragma unroll
    for (int k = 0; k < PT; ++k) {
      if (!ok[k]) continue;
      const long long opix = ((long long)n * g.Ho + oy[k]) * g.Wo + ox[k];
#pragma unroll
      for (int j = 0; j < COT / 8; ++j) {
        uint32_t bits = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) bits |= (acc[k][8 * j + e] > 0.0f ? 1u : 0u) << e;
        mask[opix * (g.Cout / 8) + c0 / 8 + j] = (uint8_t)bits;
      }
      float4* op = reinterpret_cast<float4*>(out + opix * g.Cout + c0);
#pragma unroll
      for (int j = 0; j < COT / 4; ++j)
        op[j] = make_float4(fmaxf(acc[k][4 * j], 0.0f) + sc[k][4 * j], fmaxf(acc[k][4 * j + 1], 0.0f) + sc[k][4 * j + 1],
                            fmaxf(acc[k][4 * j + 2], 0.0f) + sc[k][4 * j + 2], fmaxf(acc[k][4 * j + 3], 0.0f) + sc[k][4 * j + 3]);
    }
  }
}

// data gradient: dx[n,y,x,ci] = sum_{a,b,co} dmain[n,oy,ox,co] * Wm[a,b,ci,co] (oy*sh + a - pt == y, ...)
//                             + sum_co dout[n,oy,ox,co] * Ws[ci,co]            (oy*sh == y, ox*sw == x)
// with dmain = dout * mask.  Block = (image, band of input rows, slice of CIT input channels): the output
// rows reaching the band (dout and masked dmain) and the slice's weights W[10][CIT][Cout] are staged in
// shared memory.  Work items = (stride-parity class, chunk of 32*PT pixels of the class): within an item the
// set of contributing taps is warp-uniform; lane l computes the pixels l, l+32, .. of the chunk.
template <int CIT, int PT, int S = 0>
__global__ void transition_dgrad_kernel(GlueConv g_, const float* __restrict__ dout, const uint8_t* __restrict__ mask,
                                        const float* __restrict__ Wm, const float* __restrict__ Ws, float* __restrict__ dx,
                                        int rows) {
  extern __shared__ float sm[];
  GlueConv g = g_;
  if (S) { g.sh = S; g.sw = S; }        // compile-time strides: every "/ g.sh", "% g.sw" below folds into shifts / masks
  const int n = blockIdx.x;
  const int y0 = blockIdx.y * rows, y1 = min(g.H, y0 + rows);
  const int ci0 = blockIdx.z * CIT;
  int o0 = (y0 + g.pt - 2 + g.sh - 1) / g.sh; if (o0 < 0) o0 = 0;
  int o1 = (y1 - 1 + g.pt) / g.sh + 1; if (o1 > g.Ho) o1 = g.Ho;
  const int nor = o1 > o0 ? o1 - o0 : 0;
  const int nor_max = (rows - 1 + 2) / g.sh + 2;
  const int PS = g.Cout + 4;
  float* dO = sm;                                    // [nor*Wo][Cout+4]
  float* dM = sm + (long long)nor_max * g.Wo * PS;   // masked
  float* wsm = dM + (long long)nor_max * g.Wo * PS;  // [10][CIT][Cout]
  {
    const long long base = ((long long)n * g.Ho + o0) * g.Wo * g.Cout;       // multiple of 8 (Cout % 8 == 0)
    const float4* d4 = reinterpret_cast<const float4*>(dout + base);
#pragma unroll 4
    for (int i4 = threadIdx.x; i4 < nor * g.Wo * g.Cout / 4; i4 += blockDim.x) {
      const int i = 4 * i4;
      const float4 d = d4[i4];
      const uint32_t mb = mask[(base + i) >> 3] >> (i & 4);                  // bit index = element index
      const int si = (i / g.Cout) * PS + i % g.Cout;
      *reinterpret_cast<float4*>(dO + si) = d;
      *reinterpret_cast<float4*>(dM + si) = make_float4(mb & 1u ? d.x : 0.0f, mb & 2u ? d.y : 0.0f, mb & 4u ? d.z : 0.0f, mb & 8u ? d.w : 0.0f);
    }
  }
#pragma unroll 4
  for (int i = threadIdx.x; i < 10 * CIT * (g.Cout / 4); i += blockDim.x) {
    const int c4 = i % (g.Cout / 4), k = (i / (g.Cout / 4)) % CIT, t = i / ((g.Cout / 4) * CIT);
    const float* src = t == 9 ? Ws + (long long)(ci0 + k) * g.Cout : Wm + (long long)(t * g.Cin + ci0 + k) * g.Cout;
    reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(src) + c4);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int ncls = g.sh * g.sw;
  const int ny_max = (y1 - y0 + g.sh - 1) / g.sh + 1, nx_max = (g.W + g.sw - 1) / g.sw + 1;
  const int cpc = (ny_max * nx_max + 32 * PT - 1) / (32 * PT);       // chunks per class (upper bound)
  for (int wi = warp; wi < ncls * cpc; wi += nwarps) {
    const int cls = wi / cpc, chunk = wi % cpc;
    const int cy = cls / g.sw, cx = cls % g.sw;        // class: (y + pt) % sh == cy, (x + pl) % sw == cx
    const int ys = y0 + ((cy - (y0 + g.pt) % g.sh) + g.sh) % g.sh;
    const int xs0 = ((cx - g.pl % g.sw) + g.sw) % g.sw;
    const int ny = ys < y1 ? (y1 - ys + g.sh - 1) / g.sh : 0;
    const int nx = xs0 < g.W ? (g.W - xs0 + g.sw - 1) / g.sw : 0;
    if (chunk * 32 * PT >= ny * nx) continue;
    float acc[PT][CIT];
    int py[PT], px[PT];
    bool ok[PT];
#pragma unroll
    for (int k = 0; k < PT; ++k) {
      const int p = chunk * 32 * PT + lane + 32 * k;
      ok[k] = p < ny * nx;
      py[k] = ys + ((ok[k] ? p : 0) / nx) * g.sh; px[k] = xs0 + ((ok[k] ? p : 0) % nx) * g.sw;
#pragma unroll
      for (int c = 0; c < CIT; ++c) acc[k][c] = 0.0f;
    }
    // taps of this class: a = cy, cy + sh, ... (< 3), b likewise; tap index 9 = the 1x1 shortcut (class (pt%sh, pl%sw))
#pragma unroll
    for (int t = 0; t < 10; ++t) {
      const bool sh_ = t == 9;
      const int a = sh_ ? g.pt : t / 3, b = sh_ ? g.pl : t % 3;
      if ((a - cy) % g.sh != 0 || a < cy || (b - cx) % g.sw != 0 || b < cx) continue;    // warp-uniform
      const float* dp[PT];
      bool in[PT];
#pragma unroll
      for (int k = 0; k < PT; ++k) {
        const int ty = py[k] + g.pt - a, tx = px[k] + g.pl - b;
        const int oy = ty / g.sh, ox = tx / g.sw;
        in[k] = ok[k] && ty >= 0 && tx >= 0 && oy < g.Ho && ox < g.Wo;
        dp[k] = (sh_ ? dO : dM) + ((in[k] ? oy - o0 : 0) * g.Wo + (in[k] ? ox : 0)) * PS;
      }
      const float* wp = wsm + t * CIT * g.Cout;
      for (int co = 0; co < g.Cout; co += 4) {
        float4 d[PT];
#pragma unroll
        for (int k = 0; k < PT; ++k) {
          d[k] = *reinterpret_cast<const float4*>(dp[k] + co);
          if (!in[k]) d[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
#pragma unroll
        for (int c = 0; c < CIT; ++c) {
          const float4 w = *reinterpret_cast<const float4*>(wp + c * g.Cout + co);
#pragma unroll
          for (int k = 0; k < PT; ++k) {
            acc[k][c] = fmaf(d[k].x, w.x, acc[k][c]); acc[k][c] = fmaf(d[k].y, w.y, acc[k][c]);
            acc[k][c] = fmaf(d[k].z, w.z, acc[k][c]); acc[k][c] = fmaf(d[k].w, w.w, acc[k][c]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < PT; ++k) {
      if (!ok[k]) continue;
      float4* op = reinterpret_cast<float4*>(dx + (((long long)n * g.H + py[k]) * g.W + px[k]) * g.Cin + ci0);
#pragma unroll
      for (int c = 0; c < CIT / 4; ++c) op[c] = make_float4(acc[k][4 * c], acc[k][4 * c + 1], acc[k][4 * c + 2], acc[k][4 * c + 3]);
    }
  }
}

// weight gradient partials.  Block = (band of `orows` output rows of one image): dout / masked dmain of
// the band and the input rows it touches are staged in shared memory.  Thread = one 4(ci) x 4(co) register
// tile; the (Cin/4)*(Cout/4) tiles of a tap form a group of threads and the 10 "taps" (9 main + shortcut)
// are dealt round-robin to the blockDim / tiles groups (TPG taps per group).
// Partial layout per block: [dWm 9*Cin*Cout | dbm Cout | dWs Cin*Cout | dbs Cout].
template <int TPG>
__global__ void transition_wgrad_partial(GlueConv g, const float* __restrict__ x, const float* __restrict__ dout,
                                         const uint8_t* __restrict__ mask, float* __restrict__ part, int orows) {
  extern __shared__ float sm[];
  const int n = blockIdx.x;
  const int o0 = blockIdx.y * orows, o1 = min(g.Ho, o0 + orows);
  const int nor = o1 - o0;
  int i0 = o0 * g.sh - g.pt; if (i0 < 0) i0 = 0;
  int i1 = (o1 - 1) * g.sh + 2 - g.pt + 1; if (i1 > g.H) i1 = g.H;
  const int nir = i1 - i0;
  float* dO = sm;                                    // [nor*Wo][Cout]
  float* dM = dO + orows * g.Wo * g.Cout;
  float* xs = dM + orows * g.Wo * g.Cout;            // [nir][W][Cin]
  {
    const long long base = ((long long)n * g.Ho + o0) * g.Wo * g.Cout;
    const float4* d4 = reinterpret_cast<const float4*>(dout + base);
#pragma unroll 4
    for (int i4 = threadIdx.x; i4 < nor * g.Wo * g.Cout / 4; i4 += blockDim.x) {
      const float4 d = d4[i4];
      const uint32_t mb = mask[(base + 4 * i4) >> 3] >> ((4 * i4) & 4);
      reinterpret_cast<float4*>(dO)[i4] = d;
      reinterpret_cast<float4*>(dM)[i4] = make_float4(mb & 1u ? d.x : 0.0f, mb & 2u ? d.y : 0.0f, mb & 4u ? d.z : 0.0f, mb & 8u ? d.w : 0.0f);
    }
  }
#pragma unroll 4
  for (int i = threadIdx.x; i < nir * g.W * g.Cin / 4; i += blockDim.x)
    reinterpret_cast<float4*>(xs)[i] = reinterpret_cast<const float4*>(x + ((long long)n * g.H + i0) * g.W * g.Cin)[i];
  __syncthreads();
  const int ntile = (g.Cin / 4) * (g.Cout / 4);
  const int ngroups = blockDim.x / ntile;
  const int grp = threadIdx.x / ntile, tile = threadIdx.x % ntile;
  const int co = (tile % (g.Cout / 4)) * 4, ci = (tile / (g.Cout / 4)) * 4;
  const long long nm = 9LL * g.Cin * g.Cout;
  float* pb = part + ((long long)n * gridDim.y + blockIdx.y) * (nm + g.Cout + (long long)g.Cin * g.Cout + g.Cout);
  if (grp < ngroups) {
    float acc[TPG][4][4];
#pragma unroll
    for (int t = 0; t < TPG; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.0f;
    for (int oy = o0; oy < o1; ++oy)
      for (int ox = 0; ox < g.Wo; ++ox) {
        const int pi = ((oy - o0) * g.Wo + ox) * g.Cout + co;
        const float4 dm = *reinterpret_cast<const float4*>(dM + pi);
        const float4 ds = *reinterpret_cast<const float4*>(dO + pi);
#pragma unroll
        for (int t = 0; t < TPG; ++t) {
          const int tap = grp + t * ngroups;
          if (tap > 9) continue;
          const int a = tap == 9 ? g.pt : tap / 3, b = tap == 9 ? g.pl : tap % 3;
          const int iy = oy * g.sh + a - g.pt, ix = ox * g.sw + b - g.pl;
          if (iy < 0 || iy >= g.H || ix < 0 || ix >= g.W) continue;
          const float4 xv = *reinterpret_cast<const float4*>(xs + ((iy - i0) * g.W + ix) * g.Cin + ci);
          const float4 d = tap == 9 ? ds : dm;
          const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[t][i][0] = fmaf(xa[i], d.x, acc[t][i][0]); acc[t][i][1] = fmaf(xa[i], d.y, acc[t][i][1]);
            acc[t][i][2] = fmaf(xa[i], d.z, acc[t][i][2]); acc[t][i][3] = fmaf(xa[i], d.w, acc[t][i][3]);
          }
        }
      }
#pragma unroll
    for (int t = 0; t < TPG; ++t) {
      const int tap = grp + t * ngroups;
      if (tap > 9) continue;
      float* dst = tap == 9 ? pb + nm + g.Cout : pb + (long long)tap * g.Cin * g.Cout;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(dst + (long long)(ci + i) * g.Cout + co) = make_float4(acc[t][i][0], acc[t][i][1], acc[t][i][2], acc[t][i][3]);
    }
  }
  // bias gradients: one thread per output channel
  if (threadIdx.x < g.Cout) {
    float bmv = 0.0f, bsv = 0.0f;
    for (int p = 0; p < nor * g.Wo; ++p) { bmv += dM[p * g.Cout + threadIdx.x]; bsv += dO[p * g.Cout + threadIdx.x]; }
    pb[nm + threadIdx.x] = bmv;
    pb[nm + g.Cout + (long long)g.Cin * g.Cout + threadIdx.x] = bsv;
  }
}

// out[o] = sum_r part[r*stride + o] for many rows: LANES (8 or 32) row lanes per output, fixed-order combine
// (deterministic).  A single thread per output walked R = 128..512 rows serially (8-11 us per launch for a few
// hundred outputs).
// out_last != NULL: the last of the n outputs goes to *out_last instead of out[n - 1] (the head's loss beside its parameter gradients)
template <int LANES>
__global__ void reduce_rows_wide_kernel(const float* __restrict__ part, int R, long long stride, long long n, float* __restrict__ out,
                                        float* __restrict__ out_last = nullptr) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long o = gid / LANES;
  const int l = (int)(gid % LANES);
  if (o >= n) return;
  float acc = 0.0f;
#pragma unroll 4
  for (int r = l; r < R; r += LANES) acc += part[(long long)r * stride + o];
#pragma unroll
  for (int off = LANES >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off, LANES);
  if (l == 0) {
    if (out_last && o == n - 1) *out_last = acc;
    else out[o] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// head: GAP -> Dense softmax -> mean Keras categorical cross-entropy, forward AND backward.
// block = one image, C threads.  K <= 32 classes.
//   q = p / sum(p); c = clip(q, eps, 1-eps); loss_n = -sum_k y_k log c_k     (K.categorical_crossentropy)
// ---------------------------------------------------------------------------------------------
__global__ void head_kernel(const float* __restrict__ x, int HW, int C, int K, const float* __restrict__ Wfc,
                            const float* __restrict__ bfc, const float* __restrict__ onehot, float eps, int N,
                            float* __restrict__ probs, float* __restrict__ dx, float* __restrict__ part,
                            unsigned int* __restrict__ dx_amax = nullptr) {
  extern __shared__ float sm[];
  griddep_launch_dependents();
  griddep_wait();
  float* feat = sm;            // [C]
  float* pz = sm + C;          // [K] probabilities, then dlogits
  float* red = pz + 32;        // [1]
  const int n = blockIdx.x, c = threadIdx.x;
  float s = 0.0f;
  const float* xp = x + (long long)n * HW * C + c;
  for (int i = 0; i < HW; ++i) s += xp[(long long)i * C];
  feat[c] = s / (float)HW;
  __syncthreads();
  if (c < K) {
    float z = bfc[c];
    for (int j = 0; j < C; ++j) z = fmaf(feat[j], Wfc[j * K + c], z);
    pz[c] = z;
  }
  __syncthreads();
  if (c == 0) {
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, pz[k]);
    float se = 0.0f;
    for (int k = 0; k < K; ++k) { pz[k] = expf(pz[k] - m); se += pz[k]; }
    float sp = 0.0f;
    for (int k = 0; k < K; ++k) { pz[k] /= se; sp += pz[k]; }
    if (probs) for (int k = 0; k < K; ++k) probs[n * K + k] = pz[k];
    // loss and dL/dp through clip and renormalisation (clip passes gradient on [eps, 1-eps])
    float loss = 0.0f, gq[32], dot = 0.0f;
    for (int k = 0; k < K; ++k) {
      const float q = pz[k] / sp;
      const float cq = fminf(fmaxf(q, eps), 1.0f - eps);
      const float y = onehot[n * K + k];
      loss -= y * logf(cq);
      gq[k] = (q >= eps && q <= 1.0f - eps) ? -y / (cq * (float)N) : 0.0f;
      dot += gq[k] * q;
    }
    float gp[32], dot2 = 0.0f;
    for (int k = 0; k < K; ++k) { gp[k] = (gq[k] - dot) / sp; dot2 += gp[k] * pz[k]; }
    for (int k = 0; k < K; ++k) pz[k] = pz[k] * (gp[k] - dot2);   // dlogits
    red[0] = loss / (float)N;
  }
  __syncthreads();
  // partials: [dWfc C*K | dbfc K | loss]
  float* pb = part + (long long)n * (C * K + K + 1);
  float df = 0.0f;
  for (int k = 0; k < K; ++k) {
    const float dz = pz[k];
    pb[c * K + k] = feat[c] * dz;
    df = fmaf(dz, Wfc[c * K + k], df);
  }
  if (c < K) pb[C * K + c] = pz[c];
  if (c == 0) pb[C * K + K] = red[0];
  if (dx) {
    const float dv = df / (float)HW;
    float* dp = dx + (long long)n * HW * C + c;
    for (int i = 0; i < HW; ++i) dp[(long long)i * C] = dv;
    if (dx_amax) {      // max |dx| for the fp16 chain behind the head (bit pattern of a non-negative float orders like a uint)
      float m = fabsf(dv);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(dx_amax, __float_as_uint(m));
    }
  }
}

}  // namespace b200ode
