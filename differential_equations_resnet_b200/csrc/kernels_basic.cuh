// K1 antisym_pack, the CUDA-core (SIMT) convolution kernels, the gradient fold and the
// memory-bound tails (relu/scale/residual, BatchNorm, column sums, Adam).
//
// Reference semantics (paths relative to the reference repository):
//   assembly   layers/tfkeras_layer_Conv2DAntisymmetric3By3.py:113-141, 210-293
//              layers/tfkeras_layer_Conv2DAntisymmetric.py:107-145, 216-270
//   conv+bias  layers/tfkeras_layer_Conv2DAntisymmetric3By3.py:157-171
//   Euler tail models/tfkeras_resnets.py:85-92
//   backward   training/training.py:300 (TF autodiff), closed forms in SURVEY.md App. A.3/A.4
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200ode {

constexpr int MAX_K = 7;

// Where each tap of a diagonal block K[:,:,o,o] comes from.
struct DiagTab {
  int8_t slot[MAX_K * MAX_K];  // free-scalar slot index, or -1 = the constant gamma centre
  int8_t sign[MAX_K * MAX_K];  // +1 / -1
  int nd;                      // free scalars per diagonal block
};

struct LayerGeom {
  int C, k, layout, antisym, use_bias;
  float gamma;
  long long nparams, bias_off;
  DiagTab tab;
};

// flat offset of the first free scalar that belongs to output channel o (W_o block start)
__host__ __device__ inline long long w_block_off(const LayerGeom& g, int o) {
  const long long kk = (long long)g.k * g.k;
  const long long pairs = (long long)o * g.C - (long long)o * (o + 1) / 2;  // sum_{o'<o} (C-o'-1)
  if (g.layout == 0) return 4LL * g.C + kk * pairs;                          // [a,b,c,d | W_0.. | bias]
  return (long long)o * g.tab.nd + kk * pairs + g.tab.nd;                    // per-o: diag scalars, W_o
}
__host__ __device__ inline long long diag_param_off(const LayerGeom& g, int o, int slot) {
  if (g.layout == 0) return (long long)slot * g.C + o;
  return (long long)o * g.tab.nd + (long long)g.k * g.k * ((long long)o * g.C - (long long)o * (o + 1) / 2) + slot;
}

// value of K[a,b,ci,o] from the packed free parameters (closed form)
__device__ __forceinline__ float kernel_entry(const LayerGeom& g, const float* __restrict__ p, int a, int b, int ci,
                                              int o) {
  const int k = g.k;
  if (ci == o) {
    const int t = a * k + b;
    const int s = g.tab.slot[t];
    if (s < 0) return g.gamma;
    const float v = p[diag_param_off(g, o, s)];
    return g.tab.sign[t] > 0 ? v : -v;
  }
  if (ci > o) return p[w_block_off(g, o) + (long long)(a * k + b) * (g.C - o - 1) + (ci - o - 1)];
  // ci < o: negated, 180-degree rotated copy of W_ci[:, :, o-ci-1]   (3By3.py:133-135, 277-293)
  return -p[w_block_off(g, ci) + (long long)((k - 1 - a) * k + (k - 1 - b)) * (g.C - ci - 1) + (o - ci - 1)];
}

__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// One thread per staged element (tap, o, ci), ci fastest: staged tensor-core operand copies
// W[tap][o][ci] (K-major B operand) in up to three formats plus the dense HWIO kernel.
__global__ void pack_kernel(LayerGeom g, const float* __restrict__ params, float* __restrict__ K_dense,
                            float* __restrict__ K_user, float* __restrict__ w_hi, float* __restrict__ w_lo,
                            __nv_bfloat16* __restrict__ w_bf, float* __restrict__ bias_out, int strict) {
  const long long total = (long long)g.k * g.k * g.C * g.C;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.C && bias_out) bias_out[i] = g.use_bias ? params[g.bias_off + i] : 0.0f;
  if (i >= total) return;
  const int ci = (int)(i % g.C);
  const int o = (int)((i / g.C) % g.C);
  const int tap = (int)(i / ((long long)g.C * g.C));
  const float v = kernel_entry(g, params, tap / g.k, tap % g.k, ci, o);
  const long long dense = ((long long)tap * g.C + ci) * g.C + o;  // [a,b,ci,o]
  if (K_dense) K_dense[dense] = v;
  if (K_user) K_user[dense] = v;
  if (w_hi) {
    if (strict) {  // hi + lo == v exactly; the tensor core truncates its tf32 inputs (probe)
      const float hi = tf32_trunc(v);
      w_hi[i] = hi;
      w_lo[i] = tf32_rna(v - hi);
    } else {
      w_hi[i] = tf32_rna(v);
    }
  }
  if (w_bf) w_bf[i] = __float2bfloat16_rn(v);
}

// Chain variant: blockIdx.y = layer; tf32-rounded staged weights [L][taps][o][ci] and biases [L][C].
// strict != 0 (strict chains): [L][taps][2][o][ci] -- per tap the entries truncated to tf32 (what the tensor core reads)
// followed by their remainders, so that one 2C-row B tile per tap carries W_hi and W_lo.
__global__ void pack_chain_kernel(LayerGeom g, const float* __restrict__ params, long long param_layer_stride,
                                  float* __restrict__ w_hi, int strict, float* __restrict__ bias_out) {
  const long long total = (long long)g.k * g.k * g.C * g.C;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int l = blockIdx.y;
  params += (long long)l * param_layer_stride;
  if (i < g.C) bias_out[(long long)l * g.C + i] = g.use_bias ? params[g.bias_off + i] : 0.0f;
  if (i >= total) return;
  const int ci = (int)(i % g.C);
  const int o = (int)((i / g.C) % g.C);
  const int tap = (int)(i / ((long long)g.C * g.C));
  const float v = kernel_entry(g, params, tap / g.k, tap % g.k, ci, o);
  if (strict) {
    const float hi = tf32_trunc(v);
    const long long cc = (long long)g.C * g.C;
    float* dst = w_hi + ((long long)l * g.k * g.k + tap) * 2 * cc + (i - (long long)tap * cc);
    dst[0] = hi;
    dst[cc] = tf32_rna(v - hi);
  } else {
    w_hi[(long long)l * total + i] = tf32_rna(v);
  }
}

// ---------------------------------------------------------------------------------------------
// SIMT convolution family (any C, odd k, any stride; TF SAME padding).  fp32.
// ---------------------------------------------------------------------------------------------
struct ConvGeom {
  int N, H, W, C, Ho, Wo, k, sh, sw, pt, pl;
};

// z[n,oy,ox,o] = bias[o] + sum_{a,b,ci} x[n, oy*sh+a-pt, ox*sw+b-pl, ci] * K[a,b,ci,o]
__global__ void simt_conv_fwd(ConvGeom g, const float* __restrict__ x, const float* __restrict__ Kd,
                              const float* __restrict__ bias, float* __restrict__ z) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.N * g.Ho * g.Wo * g.C;
  if (idx >= total) return;
  const int o = (int)(idx % g.C);
  long long p = idx / g.C;
  const int ox = (int)(p % g.Wo); p /= g.Wo;
  const int oy = (int)(p % g.Ho);
  const int n = (int)(p / g.Ho);
  float acc = bias ? bias[o] : 0.0f;
  for (int a = 0; a < g.k; ++a) {
    const int iy = oy * g.sh + a - g.pt;
    if (iy < 0 || iy >= g.H) continue;
    for (int b = 0; b < g.k; ++b) {
      const int ix = ox * g.sw + b - g.pl;
      if (ix < 0 || ix >= g.W) continue;
      const float* xp = x + (((long long)n * g.H + iy) * g.W + ix) * g.C;
      const float* kp = Kd + ((long long)(a * g.k + b) * g.C) * g.C + o;
      for (int ci = 0; ci < g.C; ++ci) acc = fmaf(xp[ci], kp[(long long)ci * g.C], acc);
    }
  }
  z[idx] = acc;
}

// dx[n,iy,ix,ci] = skip + sum_{a,b,o} dz[n,oy,ox,o] * K[a,b,ci,o],  oy*sh + a - pt == iy
__global__ void simt_conv_dgrad(ConvGeom g, const float* __restrict__ dz, const float* __restrict__ Kd,
                                const float* __restrict__ skip, float* __restrict__ dx) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.N * g.H * g.W * g.C;
  if (idx >= total) return;
  const int ci = (int)(idx % g.C);
  long long p = idx / g.C;
  const int ix = (int)(p % g.W); p /= g.W;
  const int iy = (int)(p % g.H);
  const int n = (int)(p / g.H);
  float acc = skip ? skip[idx] : 0.0f;
  for (int a = 0; a < g.k; ++a) {
    const int ty = iy + g.pt - a;
    if (ty < 0 || ty % g.sh) continue;
    const int oy = ty / g.sh;
    if (oy >= g.Ho) continue;
    for (int b = 0; b < g.k; ++b) {
      const int tx = ix + g.pl - b;
      if (tx < 0 || tx % g.sw) continue;
      const int ox = tx / g.sw;
      if (ox >= g.Wo) continue;
      const float* dp = dz + (((long long)n * g.Ho + oy) * g.Wo + ox) * g.C;
      const float* kp = Kd + ((long long)(a * g.k + b) * g.C + ci) * g.C;
      for (int o = 0; o < g.C; ++o) acc = fmaf(dp[o], kp[o], acc);
    }
  }
  dx[idx] = acc;
}

// partial dense weight gradient: Gpart[part][a,b,ci,o] = sum over this part's output pixels
__global__ void simt_conv_wgrad(ConvGeom g, const float* __restrict__ x, const float* __restrict__ dz,
                                float* __restrict__ Gpart, int nparts) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.k * g.k * g.C * g.C;
  if (e >= total) return;
  const int part = blockIdx.y;
  const int o = (int)(e % g.C);
  const int ci = (int)((e / g.C) % g.C);
  const int tap = (int)(e / ((long long)g.C * g.C));
  const int a = tap / g.k, b = tap % g.k;
  const long long npix = (long long)g.N * g.Ho * g.Wo;
  const long long per = (npix + nparts - 1) / nparts;
  const long long p0 = part * per, p1 = min(npix, p0 + per);
  float acc = 0.0f;
  for (long long p = p0; p < p1; ++p) {
    const int ox = (int)(p % g.Wo);
    const int oy = (int)((p / g.Wo) % g.Ho);
    const int n = (int)(p / ((long long)g.Wo * g.Ho));
    const int iy = oy * g.sh + a - g.pt, ix = ox * g.sw + b - g.pl;
    if (iy < 0 || iy >= g.H || ix < 0 || ix >= g.W) continue;
    acc = fmaf(x[(((long long)n * g.H + iy) * g.W + ix) * g.C + ci], dz[p * g.C + o], acc);
  }
  Gpart[(long long)part * total + e] = acc;
}

// deterministic reduction of split-K partials: G[e] = sum_part Gpart[part][e]
__global__ void reduce_parts(const float* __restrict__ Gpart, int nparts, long long total, float* __restrict__ G,
                             float* __restrict__ G_user) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  float acc = 0.0f;
  for (int p = 0; p < nparts; ++p) acc += Gpart[(long long)p * total + e];
  G[e] = acc;
  if (G_user) G_user[e] = acc;
}

__device__ __forceinline__ float pair_gather(const float* __restrict__ Gp, long long e, int P);
// dense G from the raw pixel-pair partials of wgrad_tc_kernel (tests / diagnostics)
__global__ void reduce_parts_pair(const float* __restrict__ Gpart, int nparts, long long part_stride, long long total,
                                  float* __restrict__ G, float* __restrict__ G_user, int P);

// Which kernel entries a free scalar feeds: every free scalar appears in exactly
// two entries of K (one for a trainable centre); dL/dparam = s1*G[e1] + s2*G[e2].
__device__ __forceinline__ void param_entries(const LayerGeom& g, long long i, long long& e1, float& s1, long long& e2,
                                              float& s2) {
  const int k = g.k, C = g.C;
  const long long kk = (long long)k * k;
  int o, slot = 0;
  long long rem = 0;
  bool is_diag;
  if (g.layout == 0) {
    if (i < 4LL * C) { is_diag = true; slot = (int)(i / C); o = (int)(i % C); }
    else {
      is_diag = false;
      int lo = 0, hi = C - 1;  // largest o with w_block_off(o) <= i
      while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (w_block_off(g, mid) <= i) lo = mid; else hi = mid - 1; }
      o = lo; rem = i - w_block_off(g, o);
    }
  } else {
    int lo = 0, hi = C - 1;  // block of o starts at w_block_off(o) - nd
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (w_block_off(g, mid) - g.tab.nd <= i) lo = mid; else hi = mid - 1; }
    o = lo;
    rem = i - (w_block_off(g, o) - g.tab.nd);
    if (rem < g.tab.nd) { is_diag = true; slot = (int)rem; }
    else { is_diag = false; rem -= g.tab.nd; }
  }
  e1 = e2 = 0; s1 = s2 = 0.0f;
  if (is_diag) {
    int found = 0;
    for (int t = 0; t < k * k; ++t)
      if (g.tab.slot[t] == slot) {
        const long long e = ((long long)t * C + o) * C + o;
        const float sg = g.tab.sign[t] > 0 ? 1.0f : -1.0f;
        if (found == 0) { e1 = e; s1 = sg; } else { e2 = e; s2 = sg; }
        ++found;
      }
  } else {
    const int n = C - o - 1;
    const int tap = (int)(rem / n), j = (int)(rem % n), ci = o + 1 + j;
    const int rt = (int)(kk - 1 - tap);  // (k-1-a)*k + (k-1-b)
    e1 = ((long long)tap * C + ci) * C + o; s1 = 1.0f;
    e2 = ((long long)rt * C + o) * C + ci; s2 = -1.0f;
  }
}

// Power-of-two scale S of the fp16 backward strips of the persistent chains (kernels_chain_f16.cuh):
// 2^10 <= S*|h|*amax < 2^11, i.e. 32x headroom below the fp16 maximum for gradient growth along the chain and
// 2^-35 of the maximum before an element underflows to zero.  Every consumer (chain kernel, gradient fold)
// evaluates the same expression on the same device scalar `amax`, so the scale is undone exactly.
__host__ __device__ inline float chain_grad_scale(float h, float amax) {
  const float t = fabsf(h) * amax;
  if (!(t > 0.0f) || !(t < 3.0e38f)) return 1.0f;
  int e;
  frexpf(t, &e);                       // t = m * 2^e, m in [0.5, 1)
  int k = 11 - e;
  k = k < -100 ? -100 : k > 100 ? 100 : k;
  return ldexpf(1.0f, k);
}
// multiplier that undoes it in the gradient fold (1 when the gradients are unscaled: amax == nullptr)
__device__ __forceinline__ float fold_out_scale(const float* amax, float h) {
  return amax ? 1.0f / chain_grad_scale(h, *amax) : 1.0f;
}

// Fold a (reduced) dense gradient onto the free parameters.
__global__ void fold_grad(LayerGeom g, const float* __restrict__ G, float* __restrict__ grad, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nfree = g.use_bias ? g.bias_off : g.nparams;
  if (i >= nfree) return;
  long long e1, e2; float s1, s2;
  param_entries(g, i, e1, s1, e2, s2);
  const float val = s1 * G[e1] + s2 * G[e2];
  grad[i] = accumulate ? grad[i] + val : val;
}

// Split-K reduction + fold + bias gradient in one launch: `lanes` (1,2,4,8,16) threads per output
// sum the partials (lane l takes parts l, l+lanes, ...) and combine through a fixed-order shuffle
// tree -> deterministic.
// blockIdx.y = layer of a batched (chain) weight gradient; the three layer strides are in floats.
// pair != 0: the partials are the raw pixel-pair accumulators of wgrad_tc_kernel
// ([3 kernel rows][128 lanes (c,px,ci)][32 cols (pd,o)], C = 16): with t = beta - 1 + pd,
// G[alpha,beta,ci,o] = sum over pd of lane (c = floor(t/2) + 1, px = t & 1, ci), col (pd, o).
__device__ __forceinline__ float pair_gather(const float* __restrict__ Gp, long long e, int) {
  const int o = (int)(e & 15), ci = (int)((e >> 4) & 15), tap = (int)(e >> 8);
  const int alpha = tap / 3, beta = tap - 3 * alpha;
  const float* base = Gp + (long long)alpha * 128 * 32;
  float acc = 0.0f;
#pragma unroll
  for (int pd = 0; pd < 2; ++pd) {
    const int t = beta + pd + 1;                 // (beta - 1 + pd) + 2, so that >> and & act on a non-negative value
    const int c = (t >> 1), px = t & 1;          // c = floor((beta-1+pd)/2) + 1
    acc += base[(c * 32 + px * 16 + ci) * 32 + pd * 16 + o];
  }
  return acc;
}
__global__ void reduce_parts_pair(const float* __restrict__ Gpart, int nparts, long long part_stride, long long total,
                                  float* __restrict__ G, float* __restrict__ G_user, int P) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  float acc = 0.0f;
  for (int p = 0; p < nparts; ++p) acc += pair_gather(Gpart + (long long)p * part_stride, e, P);
  G[e] = acc;
  if (G_user) G_user[e] = acc;
}
__global__ void fold_reduce_kernel(LayerGeom g, const float* __restrict__ Gpart, int nparts, long long part_stride,
                                   const float* __restrict__ bias_part, float* __restrict__ grad, int accumulate,
                                   long long part_layer_stride, long long bias_layer_stride, long long grad_layer_stride,
                                   int lanes, int pair_P, int diag_only, const float* __restrict__ amax = nullptr, float h = 1.0f) {
  Gpart += (long long)blockIdx.y * part_layer_stride;
  if (bias_part) bias_part += (long long)blockIdx.y * bias_layer_stride;
  grad += (long long)blockIdx.y * grad_layer_stride;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long i = gid / lanes;
  const int l = (int)(gid % lanes);
  const long long nfree = g.use_bias ? g.bias_off : g.nparams;
  const long long total = nfree + ((g.use_bias && bias_part) ? g.C : 0);
  // diag_only (3By3 layout): the off-diagonal blocks are folded by fold_reduce_tiled_kernel; this launch covers
  // the 4C diagonal scalars [0, 4C) and the bias [nfree, total) only
  if (diag_only && i >= 4LL * g.C) i += nfree - 4LL * g.C;
  if (i >= total) return;
  const unsigned mask = __activemask();
  float val = 0.0f;
  if (i < nfree) {
    long long e1, e2; float s1, s2;
    param_entries(g, i, e1, s1, e2, s2);
    for (int p = l; p < nparts; p += lanes) {
      const float* Gp = Gpart + (long long)p * part_stride;
      if (pair_P) val += s1 * pair_gather(Gp, e1, pair_P) + s2 * pair_gather(Gp, e2, pair_P);
      else val += s1 * Gp[e1] + s2 * Gp[e2];
    }
  } else {
    const int c = (int)(i - nfree);
    for (int p = l; p < nparts; p += lanes) val += bias_part[(long long)p * g.C + c];
  }
  for (int off = lanes >> 1; off > 0; off >>= 1) val += __shfl_xor_sync(mask, val, off, 32);
  val *= fold_out_scale(amax, h);
  if (l == 0) grad[i] = accumulate ? grad[i] + val : val;
}

// Off-diagonal free parameters in 32x32 (ci, o) tiles:
//   dL/dW_o[tap][ci-o-1] = sum_p G_p[tap][ci][o] - sum_p G_p[kk-1-tap][o][ci]        
// The first term is read coalesced along o and transposed through shared memory, the second term and the
// write are coalesced along ci (the per-parameter kernel gathers the first term with a stride of C floats:
// 80 us for C = 256 / 12 partials where the partials are 28 MB).  Fixed summation order -> deterministic.
// grid: (tile pairs to <= tc, taps, layers); PZ lanes over the partials per block.
template <int PZ>
__global__ void __launch_bounds__(256 * PZ) fold_reduce_tiled_kernel(LayerGeom g, const float* __restrict__ Gpart, int nparts,
                                                                     long long part_stride, float* __restrict__ grad, int accumulate,
                                                                     long long part_layer_stride, long long grad_layer_stride,
                                                                     const float* __restrict__ amax = nullptr, float h = 1.0f) {
  // blockDim = (32, 8, PZ): threadIdx.z deals the partials (part p goes to lane p % PZ); the PZ partial sums are
  // combined in a fixed order through shared memory.
  __shared__ float t1[PZ][32][33];
  __shared__ float t2[PZ][32][33];
  Gpart += (long long)blockIdx.z * part_layer_stride;
  grad += (long long)blockIdx.z * grad_layer_stride;
  int tc = 0;
  while ((tc + 1) * (tc + 2) / 2 <= (int)blockIdx.x) ++tc;
  const int to = (int)blockIdx.x - tc * (tc + 1) / 2;
  const int tap = blockIdx.y, rt = g.k * g.k - 1 - tap;
  const int C = g.C, ci0 = tc * 32, o0 = to * 32;
  const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    float a = 0.0f, b = 0.0f;
    {
      const int ci = ci0 + r, o = o0 + tx;      // first term: row ci, contiguous along o
      if (ci < C && o < C) {
        const float* src = Gpart + ((long long)tap * C + ci) * C + o;
        for (int p = tz; p < nparts; p += PZ) a += src[(long long)p * part_stride];
      }
    }
    {
      const int o = o0 + r, ci = ci0 + tx;      // second term: row o, contiguous along ci
      if (o < C && ci < C && ci > o) {
        const float* src = Gpart + ((long long)rt * C + o) * C + ci;
        for (int p = tz; p < nparts; p += PZ) b += src[(long long)p * part_stride];
      }
    }
    t1[tz][r][tx] = a;
    t2[tz][r][tx] = b;
  }
  __syncthreads();
  if (tz == 0) {
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int o = o0 + r, ci = ci0 + tx;
      if (o < C && ci < C && ci > o) {
        float a = 0.0f, b = 0.0f;
#pragma unroll
        for (int z = 0; z < PZ; ++z) { a += t1[z][tx][r]; b += t2[z][r][tx]; }
        const float val = (a - b) * fold_out_scale(amax, h);
        const long long i = w_block_off(g, o) + (long long)tap * (C - o - 1) + (ci - o - 1);
        grad[i] = accumulate ? grad[i] + val : val;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Memory-bound tails
// ---------------------------------------------------------------------------------------------
// y = [h *] relu?(z*scale[c]+shift[c]) [+ x];  mask bit = (u > 0).  One thread per 8 channels.
__global__ void euler_tail_kernel(const float* __restrict__ z, const float* __restrict__ scale,
                                  const float* __restrict__ shift, const float* __restrict__ x, float* __restrict__ y,
                                  uint8_t* __restrict__ mask, long long pixels, int C, float h, int flags) {
  const int groups = (C + 7) / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= pixels * groups) return;
  const long long p = idx / groups;
  const int c0 = (int)(idx % groups) * 8;
  const int nc = min(8, C - c0);
  const long long base = p * C + c0;
  float u[8], xr[8];
  if (nc == 8 && (C % 4) == 0) {
    const float4 a = *reinterpret_cast<const float4*>(z + base), b = *reinterpret_cast<const float4*>(z + base + 4);
    u[0] = a.x; u[1] = a.y; u[2] = a.z; u[3] = a.w; u[4] = b.x; u[5] = b.y; u[6] = b.z; u[7] = b.w;
    if (x && (flags & 8)) {
      const float4 c = *reinterpret_cast<const float4*>(x + base), d = *reinterpret_cast<const float4*>(x + base + 4);
      xr[0] = c.x; xr[1] = c.y; xr[2] = c.z; xr[3] = c.w; xr[4] = d.x; xr[5] = d.y; xr[6] = d.z; xr[7] = d.w;
    }
  } else {
    for (int j = 0; j < nc; ++j) { u[j] = z[base + j]; if (x && (flags & 8)) xr[j] = x[base + j]; }
  }
  uint32_t bits = 0;
  for (int j = 0; j < nc; ++j) {
    float v = u[j];
    if (scale) v = fmaf(v, scale[c0 + j], shift[c0 + j]);
    if (v > 0.0f) bits |= 1u << j;
    if (flags & 2) v = fmaxf(v, 0.0f);
    if (flags & 4) v = h * v;              // Lambda(h*x): own rounding, then add (two roundings)
    if (x && (flags & 8)) v = v + xr[j];
    u[j] = v;
  }
  if (mask) mask[p * groups + c0 / 8] = (uint8_t)bits;
  if (y) {
    if (nc == 8 && (C % 4) == 0) {
      *reinterpret_cast<float4*>(y + base) = make_float4(u[0], u[1], u[2], u[3]);
      *reinterpret_cast<float4*>(y + base + 4) = make_float4(u[4], u[5], u[6], u[7]);
    } else {
      for (int j = 0; j < nc; ++j) y[base + j] = u[j];
    }
  }
}

// dz = h * dy * mask      (fp32 or bf16 I/O)
template <typename T>
__global__ void relu_scale_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ mask, T* __restrict__ dz,
                                      long long pixels, int C, float h) {
  const int groups = (C + 7) / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= pixels * groups) return;
  const long long p = idx / groups;
  const int c0 = (int)(idx % groups) * 8;
  const int nc = min(8, C - c0);
  const uint32_t bits = mask[p * groups + c0 / 8];
  for (int j = 0; j < nc; ++j) {
    const float g = (bits >> j) & 1u ? h * static_cast<float>(dy[p * C + c0 + j]) : 0.0f;
    dz[p * C + c0 + j] = static_cast<T>(g);
  }
}

// Column sums over pixels, two deterministic stages.  mode 0: (a, a*b or a*a);
// mode 1 (BN backward): du = h*dy*[z*scale+shift > 0], zhat = (z-mean)*inv -> (du, du*zhat).
struct ColsumArgs {
  const float* a; const float* b;
  const float* scale; const float* shift; const float* mean; const float* inv;
  float h; int mode;
};
__device__ __forceinline__ void colsum_vals(const ColsumArgs& A, long long p, int C, int c, float& v0, float& v1) {
  const float a = A.a[p * C + c];
  if (A.mode == 0) {
    v0 = a; v1 = A.b ? a * A.b[p * C + c] : a * a;
  } else {
    const float zz = A.b[p * C + c];
    const float u = fmaf(zz, A.scale[c], A.shift[c]);
    const float du = u > 0.0f ? A.h * a : 0.0f;
    v0 = du; v1 = du * ((zz - A.mean[c]) * A.inv[c]);
  }
}
__global__ void colsum_stage1(ColsumArgs A, float* __restrict__ ws, long long pixels, int C, int nparts) {
  // block = one part; threads stride over (row, channel) with channel = tid % C when C <= blockDim
  extern __shared__ float sm[];
  const int part = blockIdx.x;
  const long long per = (pixels + nparts - 1) / nparts;
  const long long p0 = part * per, p1 = min(pixels, p0 + per);
  const int T = blockDim.x;
  for (int cbase = 0; cbase < C; cbase += T) {
    const int lanes = min(C - cbase, T);          // channels handled this pass
    const int rows = T / lanes;                   // pixel rows processed in parallel
    const int c = cbase + threadIdx.x % lanes;
    const int r = threadIdx.x / lanes;
    float s0 = 0.0f, s1 = 0.0f;
    if (r < rows)
      for (long long p = p0 + r; p < p1; p += rows) { float v0, v1; colsum_vals(A, p, C, c, v0, v1); s0 += v0; s1 += v1; }
    sm[threadIdx.x] = s0; sm[T + threadIdx.x] = s1;
    __syncthreads();
    if (threadIdx.x < lanes) {
      float t0 = 0.0f, t1 = 0.0f;
      for (int rr = 0; rr < rows; ++rr) { t0 += sm[rr * lanes + threadIdx.x]; t1 += sm[T + rr * lanes + threadIdx.x]; }
      ws[(long long)part * C + c] = t0;
      ws[(long long)(nparts + part) * C + c] = t1;
    }
    __syncthreads();
  }
}
__global__ void colsum_stage2(const float* __restrict__ ws, float* __restrict__ out0, float* __restrict__ out1, int C,
                              int nparts) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t0 = 0.0f, t1 = 0.0f;
  for (int p = 0; p < nparts; ++p) { t0 += ws[(long long)p * C + c]; t1 += ws[(long long)(nparts + p) * C + c]; }
  if (out0) out0[c] = t0;
  if (out1) out1[c] = t1;
}

__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq,
                                   const float* __restrict__ gam, const float* __restrict__ bet, float* __restrict__ mean,
                                   float* __restrict__ inv, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mmean, float* __restrict__ mvar, long long pixels, int C, float eps,
                                   float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float m = sum[c] / (float)pixels;
  const float var = fmaxf(sumsq[c] / (float)pixels - m * m, 0.0f);  // biased variance
  const float is = rsqrtf(var + eps);
  mean[c] = m; inv[c] = is;
  const float sc = gam[c] * is;
  scale[c] = sc; shift[c] = bet[c] - m * sc;
  if (mmean) {
    const float unb = var * ((float)pixels / fmaxf((float)pixels - 1.0f, 1.0f));
    mmean[c] = mmean[c] * momentum + m * (1.0f - momentum);
    mvar[c] = mvar[c] * momentum + unb * (1.0f - momentum);
  }
}

// dz = gamma*inv * (du - mean(du) - zhat*mean(du*zhat)),  du = h*dy*[u>0]
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ mean, const float* __restrict__ inv,
                                    const float* __restrict__ gam, const float* __restrict__ dgamma,
                                    const float* __restrict__ dbeta, float* __restrict__ dz, long long pixels, int C,
                                    float h) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= pixels * C) return;
  const int c = (int)(idx % C);
  const float zz = z[idx];
  const float u = fmaf(zz, scale[c], shift[c]);
  const float du = u > 0.0f ? h * dy[idx] : 0.0f;
  const float zhat = (zz - mean[c]) * inv[c];
  const float invM = 1.0f / (float)pixels;
  dz[idx] = gam[c] * inv[c] * (du - dbeta[c] * invM - zhat * dgamma[c] * invM);
}

// tf.train.AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, const int* __restrict__ step, float lr, float b1, float b2,
                            float eps, float gscale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float t = (float)(*step);
  const float lr_t = lr * sqrtf(1.0f - powf(b2, t)) / (1.0f - powf(b1, t));
  const float gi = g[i] * gscale;
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  p[i] -= lr_t * mi / (sqrtf(vi) + eps);
}
__global__ void increment_kernel(int* c) { if (threadIdx.x == 0 && blockIdx.x == 0) *c += 1; }

// Per-layer gradient metric of the reference trainer (training/training.py:385-407): ||g||_2 / size over the slice
// [offset, offset + size) of the flat gradient bucket, one block per slice, fixed summation order (deterministic).
// `scale` = 1/world_size turns the all-reduced sum into the mean gradient first.
__global__ void __launch_bounds__(256) segment_mean_norm_kernel(const float* __restrict__ g, const long long* __restrict__ offsets,
                                                                const long long* __restrict__ sizes, float scale,
                                                                float* __restrict__ out) {
  __shared__ float red[256];
  const long long off = offsets[blockIdx.x], n = sizes[blockIdx.x];
  float acc = 0.0f;
  for (long long i = threadIdx.x; i < n; i += 256) { const float v = g[off + i] * scale; acc = fmaf(v, v, acc); }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = n > 0 ? sqrtf(red[0]) / (float)n : 0.0f;
}

}  // namespace b200ode
