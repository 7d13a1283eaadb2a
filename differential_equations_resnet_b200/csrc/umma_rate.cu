// tcgen05.mma issue/throughput probe: cycles per MMA for the operand layouts the kernels use
// (K-major vs MN-major, tf32 vs bf16, M, N).  One CTA per SM-count argument; shared memory holds
// zeros (timing does not depend on the values).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate umma_rate.cu
//   run  : ./umma_rate [ctas]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sm100_ptx.cuh"

using namespace b200ode;

struct RateCase {
  const char* name;
  int bf16, M, N, a_mn, b_mn;
  uint32_t a_lt, b_lt;      // layout_type fields
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_step, b_step;  // bytes added per MMA (cycled over `span` steps)
  int span;
  int nacc;                 // accumulators cycled through (N columns each)
  uint32_t a_off;           // bytes added to the A start address (row-shifted taps of the conv kernels)
  int commit_every;         // 0: never; k: tcgen05.commit to a scratch mbarrier after every k MMAs
};

template <int VARIANT, int BF16 = -1>
__global__ void __launch_bounds__(128, 1) rate_kernel(RateCase c, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
  __shared__ uint32_t tmem_base_s;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x * 16; i < 192 * 1024; i += blockDim.x * 16) *reinterpret_cast<uint4*>(base + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  const int warp = threadIdx.x / 32;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  if (threadIdx.x == 32) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_mbar_init(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t b32 = smem_u32(base);
    const uint32_t idesc = make_instr_desc(c.bf16 ? FMT_BF16 : FMT_TF32, c.M, c.N, c.a_mn, c.b_mn);
    const uint64_t da0 = make_smem_desc(b32 + c.a_off, c.a_lbo, c.a_sbo, c.a_lt);
    const uint64_t db0 = make_smem_desc(b32 + 100 * 1024, c.b_lbo, c.b_sbo, c.b_lt);
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int pass = 0; pass < 2; ++pass) {   // pass 0 warms up
      t0 = clock64();
      int s = 0, a = 0;
      if (VARIANT == 0) {          // whole warp loops, elected lane issues inside the loop
        for (int r = 0; r < reps; ++r) {
          const uint64_t da = da0 + ((uint64_t)(s * c.a_step) >> 4), db = db0 + ((uint64_t)(s * c.b_step) >> 4);
          if (leader) {
            if (c.bf16) umma_f16(tmem + a * c.N, da, db, idesc, 1);
            else        umma_tf32(tmem + a * c.N, da, db, idesc, 1);
          }
          if (++s == c.span) s = 0;
          if (++a == c.nacc) a = 0;
        }
      } else if (VARIANT == 1) {   // only the elected lane runs the loop
        if (leader) {
          for (int r = 0; r < reps; ++r) {
            const uint64_t da = da0 + ((uint64_t)(s * c.a_step) >> 4), db = db0 + ((uint64_t)(s * c.b_step) >> 4);
            if (c.bf16) umma_f16(tmem + a * c.N, da, db, idesc, 1);
            else        umma_tf32(tmem + a * c.N, da, db, idesc, 1);
            if (++s == c.span) s = 0;
            if (++a == c.nacc) a = 0;
          }
        }
        __syncwarp();
      } else if (VARIANT == 3) {   // like 2 but dtype is a compile-time constant (no branch per MMA)
        if (leader) {
          uint64_t das[8], dbs[8];
          uint32_t dd[8];
          for (int j = 0; j < 8; ++j) {
            das[j] = da0 + ((uint64_t)((j % c.span) * c.a_step) >> 4);
            dbs[j] = db0 + ((uint64_t)((j % c.span) * c.b_step) >> 4);
            dd[j] = tmem + (j % c.nacc) * c.N;
          }
          if (c.span > 8) {   // long walks: descriptors computed on the fly
            int sidx = 0;
            for (int r = 0; r < reps; ++r) {
              const uint64_t da = da0 + ((uint64_t)(sidx * c.a_step) >> 4), db = db0 + ((uint64_t)((sidx % 9) * c.b_step) >> 4);
              umma_tf32(tmem, da, db, idesc, 1);
              if (++sidx == c.span) sidx = 0;
            }
          } else
          for (int r = 0; r < reps; r += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (BF16) umma_f16(dd[j], das[j], dbs[j], idesc, 1);
              else      umma_tf32(dd[j], das[j], dbs[j], idesc, 1);
            }
            if (c.commit_every && ((r + 8) % c.commit_every) == 0) umma_commit(&bar2);
          }
        }
        __syncwarp();
      } else {                     // elected lane, loop unrolled by 8 with precomputed descriptors
        if (leader) {
          uint64_t das[8], dbs[8];
          for (int j = 0; j < 8; ++j) {
            das[j] = da0 + ((uint64_t)((j % c.span) * c.a_step) >> 4);
            dbs[j] = db0 + ((uint64_t)((j % c.span) * c.b_step) >> 4);
          }
          for (int r = 0; r < reps; r += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (c.bf16) umma_f16(tmem + (j % c.nacc) * c.N, das[j], dbs[j], idesc, 1);
              else        umma_tf32(tmem + (j % c.nacc) * c.N, das[j], dbs[j], idesc, 1);
            }
          }
        }
        __syncwarp();
      }
      t1 = clock64();
      if (leader) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, pass & 1);
      t2 = clock64();
    }
    if (threadIdx.x == 32) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  const int ctas = argc > 1 ? atoi(argv[1]) : 1;
  const int reps = 2048;
  std::vector<RateCase> cases;
  // ---- K-major (conv fwd/dgrad): A rows of ROWB bytes, 128B/64B/32B swizzle, SBO = 8*ROWB
  for (int bf = 0; bf < 2; ++bf)
    for (int rowb : {32, 64, 128})
      for (int N : {16, 32, 64, 128, 256}) {
        const uint32_t lt = rowb == 128 ? SWZ_128B : rowb == 64 ? SWZ_64B : SWZ_32B;
        char* nm = new char[96];
        snprintf(nm, 96, "Kmajor %s rowB=%3d M=128 N=%3d", bf ? "bf16" : "tf32", rowb, N);
        cases.push_back({nm, bf, 128, N, 0, 0, lt, lt, 16, (uint32_t)8 * rowb, 16, (uint32_t)8 * rowb, 32u, 32u, rowb / 32,
                         512 / N > 8 ? 8 : 512 / N});
      }
  // ---- conv-kernel conditions: row-shifted A start (a tap shift is a whole number of position rows) and
  //      ONE accumulator (every MMA of a segment accumulates into the same TMEM columns)
  for (int rowb : {64, 128})
    for (int N : {16, 32, 64})
      for (int shift_rows : {0, 1})
        for (int nacc : {1, 2, 4}) {
          if (rowb == 64 && N != 16) continue;
          if (rowb == 128 && N == 16) continue;
          const uint32_t lt = rowb == 128 ? SWZ_128B : SWZ_64B;
          char* nm = new char[96];
          snprintf(nm, 96, "conv tf32 rowB=%3d N=%3d A shifted %d rows, %d accumulators", rowb, N, shift_rows, nacc);
          cases.push_back({nm, 0, 128, N, 0, 0, lt, lt, 16, (uint32_t)8 * rowb, 16, (uint32_t)8 * rowb, 32u, 32u, rowb / 32, nacc,
                           (uint32_t)(shift_rows * rowb), 0});
        }
  for (int N : {16, 32, 64, 128})
    for (int ce : {8, 16, 32}) {
      const int rowb = N == 16 ? 64 : 128;
      const uint32_t lt = rowb == 128 ? SWZ_128B : SWZ_64B;
      char* nm = new char[96];
      snprintf(nm, 96, "commit tf32 rowB=%3d N=%3d commit every %2d MMAs", rowb, N, ce);
      cases.push_back({nm, 0, 128, N, 0, 0, lt, lt, 16, (uint32_t)8 * rowb, 16, (uint32_t)8 * rowb, 32u, 32u, rowb / 32, 1, 0u, ce});
    }
  // ---- many distinct operand tiles (the conv kernels walk 9 taps x k-steps of A windows and weight tiles)
  for (int N : {32, 64})
    for (int mode = 0; mode < 3; ++mode) {
      char* nm = new char[96];
      snprintf(nm, 96, "walk tf32 rowB=128 N=%3d %s", N, mode == 0 ? "A walks 36 windows" : mode == 1 ? "B walks 36 tiles" : "A and B walk");
      // A: +1 row and +1 k-step per MMA (160 B); B: next weight tile (N*128 B) every MMA
      cases.push_back({nm, 0, 128, N, 0, 0, SWZ_128B, SWZ_128B, 16, 1024, 16, 1024, mode == 1 ? 32u : 160u,
                       mode == 0 ? 32u : (uint32_t)(N * 128), mode == 1 ? 4 : 36, 1, 0u, 0});
      if (mode == 1) cases.back().span = 36;
    }
  // ---- MN-major tf32 (wgrad): 128B swizzle / 32B atoms, 32-channel chunks, SBO=512, LBO = chunk stride
  for (int N : {32, 64, 128, 256})
    for (int M : {64, 128}) {
      char* nm = new char[96];
      snprintf(nm, 96, "MNmajor tf32 atom32 M=%3d N=%3d", M, N);
      cases.push_back({nm, 0, M, N, 1, 1, 1u, 1u, 16384, 512, 8192, 512, 1024u, 1024u, 16, 512 / N > 8 ? 8 : 512 / N});
    }
  // trick variant: LBO = one position row (128 B)
  {
    cases.push_back({"MNmajor tf32 atom32 M=128 N= 32 LBO=128 (beta trick)", 0, 128, 32, 1, 1, 1u, 1u, 128, 512, 16384, 512, 1024u,
                     1024u, 16, 8});
  }
  // ---- MN-major bf16 (wgrad): 128B swizzle, 64-channel chunks, 8-position groups (SBO = 1024), K=16 positions
  for (int N : {64, 128, 256})
    for (int M : {64, 128}) {
      char* nm = new char[96];
      snprintf(nm, 96, "MNmajor bf16 swz128 M=%3d N=%3d", M, N);
      cases.push_back({nm, 1, M, N, 1, 1, SWZ_128B, SWZ_128B, 16384, 1024, 8192, 1024, 2048u, 2048u, 8, 512 / N > 8 ? 8 : 512 / N});
    }
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 2 * ctas);
  const int variant = argc > 2 ? atoi(argv[2]) : 0;
  cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(rate_kernel<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(rate_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("variant %d, %d CTAs\n", variant, ctas);
  std::vector<long long> h(2 * ctas);
  printf("%-56s %10s %10s %12s\n", "case", "issue/MMA", "total/MMA", "TFLOP/s@1.9G/SM*148");
  for (auto& c : cases) {
    if (variant == 0) rate_kernel<0><<<ctas, 128, 193 * 1024 + 1024>>>(c, reps, d_out);
    else if (variant == 1) rate_kernel<1><<<ctas, 128, 193 * 1024 + 1024>>>(c, reps, d_out);
    else if (variant == 2) rate_kernel<2><<<ctas, 128, 193 * 1024 + 1024>>>(c, reps, d_out);
    else if (c.bf16) rate_kernel<3, 1><<<ctas, 128, 193 * 1024 + 1024>>>(c, reps, d_out);
    else rate_kernel<3, 0><<<ctas, 128, 193 * 1024 + 1024>>>(c, reps, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-56s CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d_out, sizeof(long long) * 2 * ctas, cudaMemcpyDeviceToHost);
    double iss = 0, tot = 0;
    for (int i = 0; i < ctas; ++i) { iss += h[2 * i]; tot += h[2 * i + 1]; }
    iss /= ctas * (double)reps; tot /= ctas * (double)reps;
    const double kel = c.bf16 ? 16 : 8;
    const double flops = 2.0 * c.M * c.N * kel;
    printf("%-56s %10.1f %10.1f %12.1f\n", c.name, iss, tot, flops / tot * 1.9e9 * 148 * 1e-12);
  }
  return 0;
}
