// HBM-bound tails of the BatchNorm Euler step (reference: models/tfkeras_resnets.py:85-92, Keras
// BatchNormalization(axis=3): eps 1e-3, momentum 0.99, batch mean / biased variance in training) and the vectorised
// optimiser step.  Every kernel here moves 16 bytes per load / store and keeps several independent loads in flight per
// thread; the per-channel reductions are two deterministic stages (fixed-order partials, fixed-order combine).
//
//   forward : conv epilogue (kernels_conv_tc.cuh, ConvTcParams::bn_part) or colsum_vec_stage1 -> partial rows
//             bn_stats_finalize_kernel: rows -> sum, sumsq [-> mean, inv_std, scale, shift, moving stats]
//             euler_tail_kernel (kernels_basic.cuh): y = x + h*relu(z*scale+shift)
//   backward: colsum_vec_stage1<1> (du, du*zhat partials) -> bn_stats_finalize_kernel (dbeta, dgamma)
//             bn_bwd_apply_vec_kernel: dz = gamma*inv*(du - dbeta/M - zhat*dgamma/M)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_basic.cuh"

namespace b200ode {

// Stage 1 of the per-channel sums over pixels, 4 channels per thread (C % 4 == 0, C / 4 <= blockDim.x):
// MODE 0: (a, a*b or a*a); MODE 1 (BN backward): du = h*dy*[z*scale+shift > 0], zhat = (z-mean)*inv -> (du, du*zhat).
// Partials: ws[part][C] sums, ws[nparts + part][C] second sums (the layout colsum_stage2 / bn_stats_finalize_kernel read).
template <int MODE>
__global__ void __launch_bounds__(256) colsum_vec_stage1(ColsumArgs A, float* __restrict__ ws, long long pixels, int C, int nparts) {
  __shared__ float4 sm0[256], sm1[256];
  const int C4 = C >> 2;
  const int rows = 256 / C4;                       // pixel rows a block reads in parallel
  const int cg = threadIdx.x % C4, r = threadIdx.x / C4;
  const long long per = (pixels + nparts - 1) / nparts;
  const long long p0 = (long long)blockIdx.x * per, p1 = min(pixels, p0 + per);
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  float4 sc = s0, sh = s0, mu = s0, iv = s0;
  if (MODE == 1) {
    sc = reinterpret_cast<const float4*>(A.scale)[cg]; sh = reinterpret_cast<const float4*>(A.shift)[cg];
    mu = reinterpret_cast<const float4*>(A.mean)[cg]; iv = reinterpret_cast<const float4*>(A.inv)[cg];
  }
  auto add = [&](const float4 a, const float4 b) {
    if (MODE == 0) {
      s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
      s1.x = fmaf(a.x, b.x, s1.x); s1.y = fmaf(a.y, b.y, s1.y); s1.z = fmaf(a.z, b.z, s1.z); s1.w = fmaf(a.w, b.w, s1.w);
    } else {
      const float dx = fmaf(b.x, sc.x, sh.x) > 0.f ? A.h * a.x : 0.f, dy = fmaf(b.y, sc.y, sh.y) > 0.f ? A.h * a.y : 0.f;
      const float dz = fmaf(b.z, sc.z, sh.z) > 0.f ? A.h * a.z : 0.f, dw = fmaf(b.w, sc.w, sh.w) > 0.f ? A.h * a.w : 0.f;
      s0.x += dx; s0.y += dy; s0.z += dz; s0.w += dw;
      s1.x = fmaf(dx, (b.x - mu.x) * iv.x, s1.x); s1.y = fmaf(dy, (b.y - mu.y) * iv.y, s1.y);
      s1.z = fmaf(dz, (b.z - mu.z) * iv.z, s1.z); s1.w = fmaf(dw, (b.w - mu.w) * iv.w, s1.w);
    }
  };
  if (r < rows) {
    const float4* a4 = reinterpret_cast<const float4*>(A.a);
    const float4* b4 = reinterpret_cast<const float4*>(A.b ? A.b : A.a);
    long long p = p0 + r;
    for (; p + 3LL * rows < p1; p += 4LL * rows) {       // four independent 16-byte loads per operand in flight
      float4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = a4[(p + (long long)u * rows) * C4 + cg];
      if (MODE == 1 || A.b) {
#pragma unroll
        for (int u = 0; u < 4; ++u) b[u] = b4[(p + (long long)u * rows) * C4 + cg];
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) b[u] = a[u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) add(a[u], b[u]);
    }
    for (; p < p1; p += rows) {
      const float4 a = a4[p * C4 + cg];
      add(a, (MODE == 1 || A.b) ? b4[p * C4 + cg] : a);
    }
  }
  sm0[threadIdx.x] = s0; sm1[threadIdx.x] = s1;
  __syncthreads();
  if (threadIdx.x < C4) {
    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
    for (int rr = 0; rr < rows; ++rr) {
      const float4 u = sm0[rr * C4 + threadIdx.x], v = sm1[rr * C4 + threadIdx.x];
      t0.x += u.x; t0.y += u.y; t0.z += u.z; t0.w += u.w;
      t1.x += v.x; t1.y += v.y; t1.z += v.z; t1.w += v.w;
    }
    reinterpret_cast<float4*>(ws + (long long)blockIdx.x * C)[threadIdx.x] = t0;
    reinterpret_cast<float4*>(ws + (long long)(nparts + blockIdx.x) * C)[threadIdx.x] = t1;
  }
}

// Stage 2: sums the partial rows in a fixed order (one warp per channel: lane l adds rows l, l+32, ...; fixed shuffle
// tree) and, when `gam` != NULL, finishes training-mode BatchNorm in the same launch: mean, inv_std, the affine
// (scale, shift) the tail consumes, moving statistics.  block = 8 warps = 8 channels; grid = ceil(C / 8).
__global__ void __launch_bounds__(256) bn_stats_finalize_kernel(const float* __restrict__ ws, int nrows, int C, float* __restrict__ out_sum,
                                                                float* __restrict__ out_sumsq, const float* __restrict__ gam,
                                                                const float* __restrict__ bet, float* __restrict__ mean,
                                                                float* __restrict__ inv, float* __restrict__ scale,
                                                                float* __restrict__ shift, float* __restrict__ mmean,
                                                                float* __restrict__ mvar, long long pixels, float eps, float momentum) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= C) return;                      // whole warps leave together
  float t0 = 0.f, t1 = 0.f;
  for (int r = lane; r < nrows; r += 32) { t0 += ws[(long long)r * C + c]; t1 += ws[(long long)(nrows + r) * C + c]; }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) { t0 += __shfl_xor_sync(0xFFFFFFFFu, t0, o); t1 += __shfl_xor_sync(0xFFFFFFFFu, t1, o); }
  if (lane != 0) return;
  if (out_sum) out_sum[c] = t0;
  if (out_sumsq) out_sumsq[c] = t1;
  if (!gam) return;
  const float m = t0 / (float)pixels;
  const float var = fmaxf(t1 / (float)pixels - m * m, 0.0f);   // biased variance (training mode)
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

// y = [h *] relu?(z*scale[c]+shift[c]) [+ x] with an optional relu bit mask: 8 channels (two 16-byte vectors) per thread,
// per-channel affine read as vectors.  C % 8 == 0.  Same arithmetic as euler_tail_kernel.
__global__ void __launch_bounds__(256) euler_tail_vec_kernel(const float4* __restrict__ z, const float4* __restrict__ scale,
                                                             const float4* __restrict__ shift, const float4* __restrict__ x,
                                                             float4* __restrict__ y, uint8_t* __restrict__ mask, long long n8, int C8,
                                                             float h, int flags) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n8) return;
  const int g = (int)(idx % C8);
  const float4 a = z[2 * idx], b = z[2 * idx + 1];
  float u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float xr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool res = x != nullptr && (flags & 8);
  if (res) {
    const float4 c = x[2 * idx], d = x[2 * idx + 1];
    xr[0] = c.x; xr[1] = c.y; xr[2] = c.z; xr[3] = c.w; xr[4] = d.x; xr[5] = d.y; xr[6] = d.z; xr[7] = d.w;
  }
  if (scale) {
    const float4 s0 = __ldg(scale + 2 * g), s1 = __ldg(scale + 2 * g + 1), t0 = __ldg(shift + 2 * g), t1 = __ldg(shift + 2 * g + 1);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, sh[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = fmaf(u[j], sc[j], sh[j]);
  }
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = u[j];
    if (v > 0.0f) bits |= 1u << j;
    if (flags & 2) v = fmaxf(v, 0.0f);
    if (flags & 4) v = h * v;              // Lambda(h*x): own rounding, then add (two roundings)
    if (res) v = v + xr[j];
    u[j] = v;
  }
  if (mask) mask[idx] = (uint8_t)bits;
  if (y) {
    y[2 * idx] = make_float4(u[0], u[1], u[2], u[3]);
    y[2 * idx + 1] = make_float4(u[4], u[5], u[6], u[7]);
  }
}

// dz = h * dy * mask, fp32, 8 channels per thread (C % 8 == 0)
__global__ void __launch_bounds__(256) relu_scale_bwd_vec_kernel(const float4* __restrict__ dy, const uint8_t* __restrict__ mask,
                                                                 float4* __restrict__ dz, long long n8, float h) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = dy[2 * i], b = dy[2 * i + 1];
    const uint32_t m = mask[i];
    dz[2 * i] = make_float4(m & 1u ? h * a.x : 0.f, m & 2u ? h * a.y : 0.f, m & 4u ? h * a.z : 0.f, m & 8u ? h * a.w : 0.f);
    dz[2 * i + 1] = make_float4(m & 16u ? h * b.x : 0.f, m & 32u ? h * b.y : 0.f, m & 64u ? h * b.z : 0.f, m & 128u ? h * b.w : 0.f);
  }
}

// dz = gamma*inv * (du - dbeta/M - zhat*dgamma/M),  du = h*dy*[z*scale+shift > 0];  4 channels per thread, 4 loads in flight
__global__ void __launch_bounds__(256) bn_bwd_apply_vec_kernel(const float4* __restrict__ dy, const float4* __restrict__ z,
                                                               const float* __restrict__ scale, const float* __restrict__ shift,
                                                               const float* __restrict__ mean, const float* __restrict__ inv,
                                                               const float* __restrict__ gam, const float* __restrict__ dgamma,
                                                               const float* __restrict__ dbeta, float4* __restrict__ dz,
                                                               long long n4, int C4, long long pixels, float h) {
  const float invM = 1.0f / (float)pixels;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // stride is a multiple of C4 (host guarantees blockDim.x * gridDim.x % C4 == 0): a thread keeps its channel group
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = (int)(i0 % C4);
  const float4 sc = reinterpret_cast<const float4*>(scale)[cg], sh = reinterpret_cast<const float4*>(shift)[cg];
  const float4 mu = reinterpret_cast<const float4*>(mean)[cg], iv = reinterpret_cast<const float4*>(inv)[cg];
  const float4 ga = reinterpret_cast<const float4*>(gam)[cg], dg = reinterpret_cast<const float4*>(dgamma)[cg];
  const float4 db = reinterpret_cast<const float4*>(dbeta)[cg];
  auto one = [&](float d, float zz, float s, float t, float m, float is, float g, float dgm, float dbt) {
    const float du = fmaf(zz, s, t) > 0.0f ? h * d : 0.0f;
    const float zhat = (zz - m) * is;
    return g * is * (du - dbt * invM - zhat * dgm * invM);
  };
  long long i = i0;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { a[u] = dy[i + u * stride]; b[u] = z[i + u * stride]; }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      dz[i + u * stride] = make_float4(one(a[u].x, b[u].x, sc.x, sh.x, mu.x, iv.x, ga.x, dg.x, db.x), one(a[u].y, b[u].y, sc.y, sh.y, mu.y, iv.y, ga.y, dg.y, db.y),
                                       one(a[u].z, b[u].z, sc.z, sh.z, mu.z, iv.z, ga.z, dg.z, db.z), one(a[u].w, b[u].w, sc.w, sh.w, mu.w, iv.w, ga.w, dg.w, db.w));
  }
  for (; i < n4; i += stride) {
    const float4 a = dy[i], b = z[i];
    dz[i] = make_float4(one(a.x, b.x, sc.x, sh.x, mu.x, iv.x, ga.x, dg.x, db.x), one(a.y, b.y, sc.y, sh.y, mu.y, iv.y, ga.y, dg.y, db.y),
                        one(a.z, b.z, sc.z, sh.z, mu.z, iv.z, ga.z, dg.z, db.z), one(a.w, b.w, sc.w, sh.w, mu.w, iv.w, ga.w, dg.w, db.w));
  }
}

// tf.train.AdamOptimizer over the flat bucket, 4 parameters per thread (n4 = n / 4 vectors; the host runs the scalar
// kernel on the n % 4 tail): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps).  Same arithmetic as adam_kernel.
__global__ void __launch_bounds__(256) adam_vec_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                       float4* __restrict__ v, long long n4, const int* __restrict__ step, float lr,
                                                       float b1, float b2, float eps, float gscale) {
  const float t = (float)(*step);
  const float lr_t = lr * sqrtf(1.0f - powf(b2, t)) / (1.0f - powf(b1, t));
  const long long stride = (long long)gridDim.x * blockDim.x;
  auto one = [&](float& pp, float gg, float& mm, float& vv) {
    const float gi = gg * gscale;
    mm = b1 * mm + (1.0f - b1) * gi;
    vv = b2 * vv + (1.0f - b2) * gi * gi;
    pp -= lr_t * mm / (sqrtf(vv) + eps);
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = p[i], mm = m[i], vv = v[i];
    const float4 gg = g[i];
    one(pp.x, gg.x, mm.x, vv.x); one(pp.y, gg.y, mm.y, vv.y); one(pp.z, gg.z, mm.z, vv.z); one(pp.w, gg.w, mm.w, vv.w);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Gradient all-reduce FUSED into the optimiser step over NVLink / NVSwitch peer memory: every rank's
// gradient bucket is mapped into all ranks (CUDA IPC); the Adam kernel of rank r reads the N replicas of each gradient
// vector straight from peer memory, sums them in rank order (the same order on every rank -> parameters stay
// bit-identical across ranks) and updates its own parameter replica.  No all-reduce launch, no NCCL kernel competing
// with the persistent chain kernels for SMs; 3.5 MB x (N-1) of NVLink reads per step for the cfg3 net.
// Synchronisation = two flag barriers in peer memory with a monotonically increasing epoch kept on the device:
//   arrive: rank r stores the epoch into slot r of every rank's arrive[] (its gradients are complete: stream order),
//           every block waits until all slots of the LOCAL arrive[] carry the epoch, then reads the peers' gradients;
//   done:   the last block to finish stores the epoch into slot r of every rank's done[] and waits for all local slots:
//           the kernel does not complete before every peer has finished reading this rank's gradients, so the next
//           step's weight-gradient kernels may overwrite them.
// Every rank must issue the same sequence of calls (same slices, same order).
// ---------------------------------------------------------------------------------------------------------------
#ifndef B200ODE_MAX_RANKS
#define B200ODE_MAX_RANKS 8
#endif
struct P2PAdamArgs {
  const float4* g[B200ODE_MAX_RANKS];
  unsigned* arrive[B200ODE_MAX_RANKS];
  unsigned* done[B200ODE_MAX_RANKS];
  float4* pw[B200ODE_MAX_RANKS];   // two-shot: every rank's parameter replica at the slice offset
  unsigned* ctl;       // local: [0] epoch of the last completed call, [1] finished-block counter
  int nranks, rank;
};
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(256) adam_p2p_kernel(P2PAdamArgs A, float4* __restrict__ p, float4* __restrict__ m, float4* __restrict__ v,
                                                       long long n4, const int* __restrict__ step, float lr, float b1, float b2, float eps,
                                                       float gscale) {
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(A.ctl) + 1u;   // stable until the last block of THIS call bumps it
  if (blockIdx.x == 0 && threadIdx.x < A.nranks) st_release_sys(A.arrive[threadIdx.x] + A.rank, epoch);
  if (threadIdx.x < A.nranks) {
    const unsigned* slot = A.arrive[A.rank] + threadIdx.x;
    while ((int)(ld_acquire_sys(slot) - epoch) < 0) __nanosleep(40);
  }
  __syncthreads();
  const float t = (float)(*step);
  const float lr_t = lr * sqrtf(1.0f - powf(b2, t)) / (1.0f - powf(b1, t));
  auto one = [&](float& pp, float gg, float& mm, float& vv) {
    const float gi = gg * gscale;
    mm = b1 * mm + (1.0f - b1) * gi;
    vv = b2 * vv + (1.0f - b2) * gi * gi;
    pp -= lr_t * mm / (sqrtf(vv) + eps);
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 gr[B200ODE_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < B200ODE_MAX_RANKS; ++r)
      if (r < A.nranks) gr[r] = ld_relaxed_sys_v4(A.g[r] + i);          // all replicas in flight together
    float4 pp = p[i], mm = m[i], vv = v[i];
    float4 gs = gr[0];
#pragma unroll
    for (int r = 1; r < B200ODE_MAX_RANKS; ++r)
      if (r < A.nranks) { gs.x += gr[r].x; gs.y += gr[r].y; gs.z += gr[r].z; gs.w += gr[r].w; }   // rank order: identical on every rank
    one(pp.x, gs.x, mm.x, vv.x); one(pp.y, gs.y, mm.y, vv.y); one(pp.z, gs.z, mm.z, vv.z); one(pp.w, gs.w, mm.w, vv.w);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
  __syncthreads();
  __shared__ unsigned last;
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(A.ctl + 1, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x < A.nranks) {
    st_release_sys(A.done[threadIdx.x] + A.rank, epoch);
    const unsigned* slot = A.done[A.rank] + threadIdx.x;
    while ((int)(ld_acquire_sys(slot) - epoch) < 0) __nanosleep(40);
  }
  __syncthreads();
  if (threadIdx.x == 0) { A.ctl[1] = 0u; __threadfence(); A.ctl[0] = epoch; }
}

// Two-shot form (parameters live in the peer-mapped region too): rank r sums the N gradient replicas of ITS 1/N shard of
// the slice (reduce-scatter over NVLink reads), applies Adam to that shard (its moments m, v are the only ones it ever
// touches: optimiser state is sharded for free) and stores the new parameters into every rank's replica (all-gather
// over NVLink writes).  (N-1)/N of the slice crosses NVLink in each direction instead of (N-1) x the slice.  The `done`
// barrier now also means: every rank's shard has landed in this rank's parameter replica.
__global__ void __launch_bounds__(256) adam_p2p_shard_kernel(P2PAdamArgs A, float4* __restrict__ m, float4* __restrict__ v, long long n4,
                                                             const int* __restrict__ step, float lr, float b1, float b2, float eps,
                                                             float gscale) {
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(A.ctl) + 1u;
  if (blockIdx.x == 0 && threadIdx.x < A.nranks) st_release_sys(A.arrive[threadIdx.x] + A.rank, epoch);
  if (threadIdx.x < A.nranks) {
    const unsigned* slot = A.arrive[A.rank] + threadIdx.x;
    while ((int)(ld_acquire_sys(slot) - epoch) < 0) __nanosleep(40);
  }
  __syncthreads();
  const float t = (float)(*step);
  const float lr_t = lr * sqrtf(1.0f - powf(b2, t)) / (1.0f - powf(b1, t));
  auto one = [&](float& pp, float gg, float& mm, float& vv) {
    const float gi = gg * gscale;
    mm = b1 * mm + (1.0f - b1) * gi;
    vv = b2 * vv + (1.0f - b2) * gi * gi;
    pp -= lr_t * mm / (sqrtf(vv) + eps);
  };
  const long long per = (n4 + A.nranks - 1) / A.nranks;
  const long long lo = per * A.rank, hi = lo + per < n4 ? lo + per : n4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float4* p = A.pw[A.rank];
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    float4 gr[B200ODE_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < B200ODE_MAX_RANKS; ++r)
      if (r < A.nranks) gr[r] = ld_relaxed_sys_v4(A.g[r] + i);
    float4 pp = p[i], mm = m[i], vv = v[i];
    float4 gs = gr[0];
#pragma unroll
    for (int r = 1; r < B200ODE_MAX_RANKS; ++r)
      if (r < A.nranks) { gs.x += gr[r].x; gs.y += gr[r].y; gs.z += gr[r].z; gs.w += gr[r].w; }
    one(pp.x, gs.x, mm.x, vv.x); one(pp.y, gs.y, mm.y, vv.y); one(pp.z, gs.z, mm.z, vv.z); one(pp.w, gs.w, mm.w, vv.w);
    m[i] = mm; v[i] = vv;
#pragma unroll
    for (int r = 0; r < B200ODE_MAX_RANKS; ++r)
      if (r < A.nranks) A.pw[r][i] = pp;                 // own replica and all peers'
  }
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned last;
  if (threadIdx.x == 0) last = atomicAdd(A.ctl + 1, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!last) return;
  if (threadIdx.x < A.nranks) {
    st_release_sys(A.done[threadIdx.x] + A.rank, epoch);
    const unsigned* slot = A.done[A.rank] + threadIdx.x;
    while ((int)(ld_acquire_sys(slot) - epoch) < 0) __nanosleep(40);
  }
  __syncthreads();
  if (threadIdx.x == 0) { A.ctl[1] = 0u; __threadfence(); A.ctl[0] = epoch; }
}

}  // namespace b200ode
