// Hardware probe for the sm_100a descriptor semantics this project relies on.
//
// The container that builds this repo has no GPU, so every assumption about how
// tcgen05.mma interprets shared-memory descriptors (swizzle phase of row-shifted
// start addresses, MN-major LBO/SBO strides, M=64 TMEM lane mapping, tf32 input
// truncation, operand negation) and how TMA lays out swizzled 4-D boxes with
// out-of-bounds halo coordinates is checked here against a host model.
//
// Method: the host fills a shared-memory image with small random integers,
// predicts which bytes the MMA reads for logical (row, k) through `model_addr`,
// and compares the TMEM dump with the exact integer product.
//
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
//   run  : ./umma_probe            (prints one PASS/FAIL line per case)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_bf16.h>

#include "sm100_ptx.cuh"

using namespace b200ode;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

struct Operand {
  int mn_major;          // 0: K-major, 1: MN-major
  int swz;               // 0/32/64/128
  uint32_t lbo, sbo;     // bytes
  uint32_t region;       // byte offset of the operand region in the image (1024-aligned)
  uint32_t start_off;    // bytes added to region for the descriptor start address
  uint32_t kstep;        // bytes added to the start address per K step
  int base_offset;       // descriptor base_offset field
};

struct Case {
  const char* name;
  int bf16;              // 0: tf32 (4-byte elements), 1: bf16
  int M, N, ksteps;
  Operand a, b;
  int neg_a;
  int special;           // 1: tf32 truncation test values
};

constexpr int IMG_BYTES = 160 * 1024;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint8_t* __restrict__ image, int image_bytes, Case c, float* __restrict__ out /*[128][N]*/) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x * 16; i < image_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(base + i) = *reinterpret_cast<const uint4*>(image + i);
  fence_proxy_async_smem();
  const int warp = threadIdx.x / 32;
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 256);
    tmem_relinquish();
  }
  if (threadIdx.x == 32) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t b32 = smem_u32(base);
    const uint32_t idesc = make_instr_desc(c.bf16 ? FMT_BF16 : FMT_TF32, c.M, c.N, c.a.mn_major, c.b.mn_major,
                                           c.neg_a, 0);
    for (int s = 0; s < c.ksteps; ++s) {
      uint64_t da = make_smem_desc(b32 + c.a.region + c.a.start_off + s * c.a.kstep, c.a.lbo, c.a.sbo,
                                   c.a.swz == 1283 ? 1u : swz_layout_type(c.a.swz), c.a.base_offset);
      uint64_t db = make_smem_desc(b32 + c.b.region + c.b.start_off + s * c.b.kstep, c.b.lbo, c.b.sbo,
                                   c.b.swz == 1283 ? 1u : swz_layout_type(c.b.swz), c.b.base_offset);
      if (c.bf16) umma_f16(tmem, da, db, idesc, s > 0);
      else        umma_tf32(tmem, da, db, idesc, s > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  // dump all 128 lanes x N columns
  const int lane_row = warp * 32 + (threadIdx.x & 31);
  for (int c0 = 0; c0 < c.N; c0 += 8) {
    uint32_t r[8];
    tmem_ld_x8(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 8; ++j) out[lane_row * c.N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---- host model -------------------------------------------------------------
static uint32_t model_addr(const Operand& o, int eb, int mn, int k) {
  // k is the GLOBAL k index; UK elements per step
  const int S = o.swz == 1283 ? 128 : o.swz ? o.swz : 16;
  const int UK = 32 / eb;
  const int s = k / UK, kk = k % UK;
  uint32_t a = o.region + o.start_off + s * o.kstep;
  if (!o.mn_major) {
    if (o.swz) a += (mn / 8) * o.sbo + (mn % 8) * S + kk * eb;
    else       a += (mn / 8) * o.sbo + (mn % 8) * 16 + (kk * eb / 16) * o.lbo + (kk * eb) % 16;
  } else {
    const int E = S / eb;
    if (o.swz == 1283) a += (mn / 32) * o.lbo + (mn % 32) * eb + (kk % 4) * 128 + (kk / 4) * o.sbo;
    else if (o.swz) a += (mn / E) * o.lbo + (mn % E) * eb + (kk % 8) * S + (kk / 8) * o.sbo;
    else       a += (mn / E) * o.sbo + (mn % E) * eb + (kk % 8) * 16 + (kk / 8) * o.lbo;
  }
  if (o.swz == 1283) return a ^ (((a >> 7) & 3u) << 5);   // SWIZZLE_128B_BASE32B = Swizzle<2,5,2>
  return swizzle_addr(a, o.swz);
}

static float bf16_to_f(uint16_t h) { uint32_t u = uint32_t(h) << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t f_to_bf16(float f) { uint32_t u; memcpy(&u, &f, 4); return uint16_t(u >> 16); }

static float read_elem(const std::vector<uint8_t>& img, uint32_t addr, int bf16, int trunc_tf32) {
  if (addr + 4 > img.size()) { printf("model address %u out of image\n", addr); exit(3); }
  if (bf16) { uint16_t h; memcpy(&h, &img[addr], 2); return bf16_to_f(h); }
  uint32_t u; memcpy(&u, &img[addr], 4);
  if (trunc_tf32) u &= 0xFFFFE000u;
  float f; memcpy(&f, &u, 4); return f;
}

static int run_case(const Case& c, uint8_t* d_img, float* d_out, unsigned seed) {
  std::vector<uint8_t> img(IMG_BYTES);
  srand(seed);
  const int eb = c.bf16 ? 2 : 4;
  for (int i = 0; i < IMG_BYTES; i += eb) {
    float v = float((rand() % 9) - 4);
    if (c.special) {  // value with low mantissa bits set: 1 + 2^-11 + 2^-12 (+ sign)
      uint32_t u = 0x3F800000u | (1u << 12) | (1u << 11) | ((rand() & 1) << 10);
      memcpy(&v, &u, 4);
    }
    if (c.bf16) { uint16_t h = f_to_bf16(v); memcpy(&img[i], &h, 2); }
    else memcpy(&img[i], &v, 4);
  }
  CK(cudaMemcpy(d_img, img.data(), IMG_BYTES, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0, 128 * 256 * 4));
  probe_kernel<<<1, 128, IMG_BYTES + 1024>>>(d_img, IMG_BYTES, c, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-58s LAUNCH-ERROR %s\n", c.name, cudaGetErrorString(e)); return -1; }
  std::vector<float> out(128 * c.N);
  CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  const int K = c.ksteps * (32 / eb);
  int bad_trunc = 0, bad_rn = 0;
  double maxerr = 0;
  for (int variant = 0; variant < (c.special ? 2 : 1); ++variant) {
    int bad = 0;
    for (int m = 0; m < c.M; ++m) {
      const int lane = c.M == 128 ? m : (m % 16) + 32 * (m / 16);
      for (int n = 0; n < c.N; ++n) {
        double acc = 0;
        for (int k = 0; k < K; ++k) {
          float av = read_elem(img, model_addr(c.a, eb, m, k), c.bf16, c.special && variant == 0);
          float bv = read_elem(img, model_addr(c.b, eb, n, k), c.bf16, c.special && variant == 0);
          if (c.special && variant == 1) {  // round-to-nearest-even to tf32
            auto rn = [](float f) { uint32_t u; memcpy(&u, &f, 4); u += 0xFFFu + ((u >> 13) & 1u); u &= 0xFFFFE000u; memcpy(&f, &u, 4); return f; };
            av = rn(av); bv = rn(bv);
          }
          acc += double(av) * double(bv);
        }
        if (c.neg_a) acc = -acc;
        double err = fabs(acc - double(out[lane * c.N + n]));
        if (err > (c.special ? 1e-4 : 0.0)) ++bad;
        if (err > maxerr) maxerr = err;
      }
    }
    if (variant == 0) bad_trunc = bad; else bad_rn = bad;
  }
  if (c.special)
    printf("%-58s %s (mismatch: truncation-model %d, round-nearest-model %d)\n", c.name,
           bad_trunc == 0 ? "TRUNCATES" : bad_rn == 0 ? "ROUNDS-NEAREST" : "NEITHER", bad_trunc, bad_rn);
  else
    printf("%-58s %s (%d / %d mismatches, max err %.3g)\n", c.name, bad_trunc == 0 ? "PASS" : "FAIL", bad_trunc,
           c.M * c.N, maxerr);
  return bad_trunc;
}

// ---- TMA probe ----------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// bounded, non-trapping wait so that a TMA that never completes is reported, not fatal
__device__ bool soft_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 22); ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}

__device__ void tma_probe_body(const CUtensorMap* in_map, const CUtensorMap* out_map, int4 lc, int bytes, uint8_t* dump,
                               int do_store, int4 sc, int* dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ int ok_s;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) base[i] = 0xEE;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    dbg[0] = 1;
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, bytes);
    tma_load_4d(base, in_map, &bar, lc.x, lc.y, lc.z, lc.w);
    dbg[0] = 2;
    ok_s = soft_wait(&bar, 0) ? 1 : 0;
    dbg[1] = ok_s;
    dbg[0] = 3;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) dump[i] = base[i];
  __syncthreads();
  if (ok_s && do_store && threadIdx.x == 0) {
    fence_proxy_async_smem();
    tma_store_4d(out_map, base, sc.x, sc.y, sc.z, sc.w);
    tma_store_commit();
    tma_store_wait_all0();
    dbg[0] = 4;
  }
}
__global__ void tma_probe_param(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                                int4 lc, int bytes, uint8_t* dump, int do_store, int4 sc, int* dbg) {
  tma_probe_body(&in_map, &out_map, lc, bytes, dump, do_store, sc, dbg);
}
__global__ void tma_probe_gmem(const CUtensorMap* in_map, const CUtensorMap* out_map, int4 lc, int bytes, uint8_t* dump,
                               int do_store, int4 sc, int* dbg) {
  tma_probe_body(in_map, out_map, lc, bytes, dump, do_store, sc, dbg);
}

// NHWC fp32 tensor [N=2][H=6][W=10][C]; box {CB, BW, BH, 1}
static int tma_probe(EncodeTiledFn enc, int C, int CB, int swz_mode, int halo, int gmem_map, int do_store, const char* name) {
  const int N = 2, H = 6, W = 10, BW = halo ? W + 2 : 8, BH = 4;
  std::vector<float> h(size_t(N) * H * W * C);
  for (size_t i = 0; i < h.size(); ++i) h[i] = float(i % 997) + 1.0f;
  float *d_in, *d_out; uint8_t* d_dump; int* d_dbg; CUtensorMap* d_maps;
  const int bytes = CB * 4 * BW * BH;
  CK(cudaMalloc(&d_in, h.size() * 4)); CK(cudaMalloc(&d_out, h.size() * 4)); CK(cudaMalloc(&d_dump, bytes));
  CK(cudaMalloc(&d_dbg, 16)); CK(cudaMemset(d_dbg, 0, 16)); CK(cudaMalloc(&d_maps, 2 * sizeof(CUtensorMap)));
  CK(cudaMemcpy(d_in, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0, h.size() * 4));
  cuuint64_t dims[4] = {cuuint64_t(C), cuuint64_t(W), cuuint64_t(H), cuuint64_t(N)};
  cuuint64_t strides[3] = {cuuint64_t(C) * 4, cuuint64_t(W) * C * 4, cuuint64_t(H) * W * C * 4};
  cuuint32_t box[4] = {cuuint32_t(CB), cuuint32_t(BW), cuuint32_t(BH), 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = swz_mode == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swz_mode == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                        : swz_mode == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                        : swz_mode == 1283 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  alignas(64) CUtensorMap mi, mo;
  CUresult r1 = enc(&mi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d_in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d_out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("%-58s ENCODE-ERROR %d %d\n", name, int(r1), int(r2)); return -1; }
  CK(cudaMemcpy(d_maps, &mi, sizeof(mi), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_maps + 1, &mo, sizeof(mo), cudaMemcpyHostToDevice));
  const int off = halo ? -1 : 0;
  int4 lc = make_int4(0, off, off, 1), sc = make_int4(0, off, 3, 0);
  if (gmem_map) tma_probe_gmem<<<1, 128, bytes + 1024>>>(d_maps, d_maps + 1, lc, bytes, d_dump, do_store, sc, d_dbg);
  else          tma_probe_param<<<1, 128, bytes + 1024>>>(mi, mo, lc, bytes, d_dump, do_store, sc, d_dbg);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-58s LAUNCH-ERROR %s\n", name, cudaGetErrorString(e)); return -1; }
  int dbg[4]; CK(cudaMemcpy(dbg, d_dbg, 16, cudaMemcpyDeviceToHost));
  std::vector<uint8_t> dump(bytes);
  CK(cudaMemcpy(dump.data(), d_dump, bytes, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int r = 0; r < BH; ++r) for (int q = 0; q < BW; ++q) for (int c = 0; c < CB; ++c) {
    int y = r + off, x = q + off;
    float want = (y >= 0 && y < H && x >= 0 && x < W && c < C) ? h[((size_t(1) * H + y) * W + x) * C + c] : 0.0f;
    uint32_t a = uint32_t(((r * BW + q) * CB + c) * 4);
    uint32_t addr = swz_mode == 1283 ? (a ^ (((a >> 7) & 3u) << 5)) : swizzle_addr(a, swz_mode);
    float got; memcpy(&got, &dump[addr], 4);
    if (got != want) ++bad;
  }
  std::vector<float> o(h.size());
  CK(cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost));
  int bad_store = 0;
  if (do_store) for (int n = 0; n < N; ++n) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < C; ++c) {
    float want = 0.0f;
    int r = y - 3, q = x - off;          // stored box element (r, q)
    if (n == 0 && r >= 0 && r < BH && q >= 0 && q < BW) {
      int sy = r + off, sx = q + off;    // what the load put there
      want = (sy >= 0 && sy < H && sx >= 0 && sx < W) ? h[((size_t(1) * H + sy) * W + sx) * C + c] : 0.0f;
    }
    if (o[((size_t(n) * H + y) * W + x) * C + c] != want) ++bad_store;
  }
  printf("%-58s stage %d complete %d | load %s (%d bad) store %s (%d bad)\n", name, dbg[0], dbg[1], bad ? "FAIL" : "PASS",
         bad, !do_store ? "-" : bad_store ? "FAIL" : "PASS", bad_store);
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_dump); cudaFree(d_dbg); cudaFree(d_maps);
  return bad + bad_store;
}

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;  // run a single case (fresh context per case)
  uint8_t* d_img; float* d_out;
  CK(cudaMalloc(&d_img, IMG_BYTES)); CK(cudaMalloc(&d_out, 128 * 256 * 4));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG_BYTES + 1024));
  const uint32_t RB = 64 * 1024;  // B operand region
  std::vector<Case> cases;
  auto K128 = [&](uint32_t region, uint32_t start, int bo) { return Operand{0, 128, 16, 1024, region, start, 32, bo}; };
  auto K64 = [&](uint32_t region, uint32_t start, int bo) { return Operand{0, 64, 16, 512, region, start, 32, bo}; };
  auto K32 = [&](uint32_t region, uint32_t start, int bo) { return Operand{0, 32, 16, 256, region, start, 32, bo}; };
  // --- K-major, tf32
  cases.push_back({"tf32 K-major SW128 M128 N64 K32", 0, 128, 64, 4, K128(0, 0, 0), K128(RB, 0, 0), 0, 0});
  static char names[64][80];
  int ni = 0;
  for (int r : {1, 2, 3, 5, 7, 34, 35}) {
    snprintf(names[ni], 80, "tf32 K-major SW128 A row-shift %d base_off 0", r);
    cases.push_back({names[ni++], 0, 128, 64, 4, K128(0, r * 128, 0), K128(RB, 0, 0), 0, 0});
    snprintf(names[ni], 80, "tf32 K-major SW128 A row-shift %d base_off r%%8", r);
    cases.push_back({names[ni++], 0, 128, 64, 4, K128(0, r * 128, r % 8), K128(RB, 0, 0), 0, 0});
  }
  cases.push_back({"tf32 K-major SW128 M128 N256 K32", 0, 128, 256, 4, K128(0, 0, 0), K128(RB, 0, 0), 0, 0});
  cases.push_back({"tf32 K-major SW128 M128 N16 K32", 0, 128, 16, 4, K128(0, 0, 0), K128(RB, 0, 0), 0, 0});
  cases.push_back({"tf32 K-major SW64 M128 N16 K16", 0, 128, 16, 2, K64(0, 0, 0), K64(RB, 0, 0), 0, 0});
  for (int r : {1, 3, 18}) {
    snprintf(names[ni], 80, "tf32 K-major SW64 A row-shift %d base_off 0", r);
    cases.push_back({names[ni++], 0, 128, 16, 2, K64(0, r * 64, 0), K64(RB, 0, 0), 0, 0});
  }
  cases.push_back({"tf32 K-major SW128 M64 N64 (TMEM lane map)", 0, 64, 64, 4, K128(0, 0, 0), K128(RB, 0, 0), 0, 0});
  cases.push_back({"tf32 K-major SW128 negate-A", 0, 128, 64, 4, K128(0, 0, 0), K128(RB, 0, 0), 1, 0});
  cases.push_back({"tf32 input conversion (trunc vs round)", 0, 128, 16, 1, K128(0, 0, 0), K128(RB, 0, 0), 0, 1});
  // --- K-major, bf16
  cases.push_back({"bf16 K-major SW128 M128 N64 K64", 1, 128, 64, 4, K128(0, 0, 0), K128(RB, 0, 0), 0, 0});
  cases.push_back({"bf16 K-major SW128 A row-shift 3", 1, 128, 64, 4, K128(0, 3 * 128, 0), K128(RB, 0, 0), 0, 0});
  cases.push_back({"bf16 K-major SW64 M128 N32 K32", 1, 128, 32, 2, K64(0, 0, 0), K64(RB, 0, 0), 0, 0});
  cases.push_back({"bf16 K-major SW32 M128 N16 K16", 1, 128, 16, 1, K32(0, 0, 0), K32(RB, 0, 0), 0, 0});
  cases.push_back({"bf16 K-major SW32 A row-shift 5", 1, 128, 16, 1, K32(0, 5 * 32, 0), K32(RB, 0, 0), 0, 0});
  // --- MN-major: smem = [k rows][swz bytes of MN], chunks of MN at LBO, 8-k-row groups at SBO
  auto MN = [&](int swz, uint32_t lbo, uint32_t region, uint32_t start, uint32_t kstep) {
    return Operand{1, swz, lbo, uint32_t(8 * swz), region, start, kstep, 0};
  };
  // tf32: UK=8 k rows per step -> kstep = 8*swz
  cases.push_back({"tf32 MN-major SW128 A(4 chunks) B(2 chunks) K32", 0, 128, 64, 4, MN(128, 8192, 0, 0, 1024),
                   MN(128, 8192, RB, 0, 1024), 0, 0});
  cases.push_back({"tf32 MN-major SW128 A k-row-shift 3", 0, 128, 64, 4, MN(128, 8192, 0, 3 * 128, 1024),
                   MN(128, 8192, RB, 0, 1024), 0, 0});
  cases.push_back({"tf32 MN-major SW128 A+B k-row-shift 37/2", 0, 128, 64, 4, MN(128, 8192, 0, 37 * 128, 1024),
                   MN(128, 8192, RB, 2 * 128, 1024), 0, 0});
  cases.push_back({"tf32 MN-major SW64 M128(8 chunks) N16 K32", 0, 128, 16, 4, MN(64, 4096, 0, 0, 512),
                   MN(64, 4096, RB, 0, 512), 0, 0});
  cases.push_back({"tf32 MN-major SW64 M64 chunk-stride = 1 k-row (LBO 64)", 0, 64, 16, 4, MN(64, 64, 0, 0, 512),
                   MN(64, 4096, RB, 0, 512), 0, 0});
  cases.push_back({"tf32 MN-major SW128 M128 chunk-stride = 1 k-row (LBO 128)", 0, 128, 32, 4, MN(128, 128, 0, 0, 1024),
                   MN(128, 8192, RB, 0, 1024), 0, 0});
  cases.push_back({"tf32 A MN-major SW128 x B K-major SW128", 0, 128, 64, 4, MN(128, 8192, 0, 0, 1024), K128(RB, 0, 0),
                   0, 0});
  // bf16: UK=16 k rows per step (two 8-row groups at SBO) -> kstep = 16*swz
  cases.push_back({"bf16 MN-major SW128 A(2 chunks) B(1 chunk) K64", 1, 128, 64, 4, MN(128, 16384, 0, 0, 2048),
                   MN(128, 16384, RB, 0, 2048), 0, 0});
  cases.push_back({"bf16 MN-major SW128 A k-row-shift 5", 1, 128, 64, 4, MN(128, 16384, 0, 5 * 128, 2048),
                   MN(128, 16384, RB, 0, 2048), 0, 0});
  cases.push_back({"bf16 MN-major SW32 M128(8 chunks) N16 K64", 1, 128, 16, 4, MN(32, 4096, 0, 0, 512),
                   MN(32, 4096, RB, 0, 512), 0, 0});

  // tf32 MN-major with the 128B swizzle / 32B atom layout (4 k-rows per group, SBO between groups)
  auto MN32 = [&](uint32_t lbo, uint32_t sbo, uint32_t region, uint32_t start, uint32_t kstep) {
    return Operand{1, 1283, lbo, sbo, region, start, kstep, 0};
  };
  cases.push_back({"tf32 MN-major SW128/32B A(4 chunks) B(2 chunks) K32", 0, 128, 64, 4, MN32(8192, 512, 0, 0, 1024),
                   MN32(8192, 512, RB, 0, 1024), 0, 0});
  cases.push_back({"tf32 MN-major SW128/32B A k-row-shift 3", 0, 128, 64, 4, MN32(8192, 512, 0, 3 * 128, 1024),
                   MN32(8192, 512, RB, 0, 1024), 0, 0});
  cases.push_back({"tf32 MN-major SW128/32B A+B k-row-shift 37/2", 0, 128, 64, 4, MN32(8192, 512, 0, 37 * 128, 1024),
                   MN32(8192, 512, RB, 2 * 128, 1024), 0, 0});
  cases.push_back({"tf32 MN-major SW128/32B M128 chunk-stride 1 k-row (LBO 128)", 0, 128, 32, 4, MN32(128, 512, 0, 0, 1024),
                   MN32(8192, 512, RB, 0, 1024), 0, 0});
  cases.push_back({"tf32 MN-major SW128/32B M64 N32", 0, 64, 32, 4, MN32(8192, 512, 0, 0, 1024),
                   MN32(8192, 512, RB, 0, 1024), 0, 0});
  cases.push_back({"tf32 A MN-major SW128/32B x B K-major SW128", 0, 128, 64, 4, MN32(8192, 512, 0, 0, 1024), K128(RB, 0, 0),
                   0, 0});
  cases.push_back({"bf16 MN-major SW128 M128 chunk-stride 1 k-row (LBO 128)", 1, 128, 64, 4, MN(128, 128, 0, 0, 2048),
                   MN(128, 16384, RB, 0, 2048), 0, 0});
  cases.push_back({"bf16 MN-major SW64 M128(4 chunks) N32 K64", 1, 128, 32, 4, MN(64, 8192, 0, 0, 1024),
                   MN(64, 8192, RB, 0, 1024), 0, 0});

  int failures = 0;
  unsigned seed = 1;
  int idx = 0;
  for (const Case& c : cases) {
    ++seed;
    if (only >= 0 && idx++ != only) continue;
    int r = run_case(c, d_img, d_out, seed);
    if (r != 0 && !c.special) ++failures;
    if (r < 0) { printf("aborting after launch error (context is poisoned)\n"); return 1; }
  }
  EncodeTiledFn enc = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&enc), cudaEnableDefault, &qres));
  if (!enc) { printf("cuTensorMapEncodeTiled not found\n"); return 1; }
  const int nc = int(cases.size());
  struct T { int C, CB, swz, halo, gmem, store; const char* name; };
  const T tcases[] = {
    {32, 32, 0,   0, 0, 0, "TMA in-bounds box C=32 no swizzle, param map"},
    {32, 32, 0,   0, 1, 0, "TMA in-bounds box C=32 no swizzle, gmem map"},
    {32, 32, 128, 0, 0, 0, "TMA in-bounds box C=32 SW128, param map"},
    {32, 32, 128, 1, 0, 0, "TMA halo box C=32 SW128 (OOB zero fill), param map"},
    {32, 32, 128, 1, 1, 0, "TMA halo box C=32 SW128 (OOB zero fill), gmem map"},
    {32, 32, 128, 1, 0, 1, "TMA halo box C=32 SW128 + clipped store, param map"},
    {16, 16, 64,  1, 0, 1, "TMA halo box C=16 SW64 + store"},
    {8,  8,  32,  1, 0, 1, "TMA halo box C=8 SW32 + store"},
    {32, 32, 1283,1, 0, 1, "TMA halo box C=32 SW128/ATOM_32B + store"},
    {16, 32, 1283,1, 0, 0, "TMA halo box C=16 padded to 32 (channel OOB) SW128/ATOM_32B"},
    {16, 32, 128, 1, 0, 0, "TMA halo box C=16 padded to 32 (channel OOB) SW128"},
  };
  const int nt = int(sizeof(tcases) / sizeof(tcases[0]));
  if (only >= nc + nt) return 77;
  for (int t = 0; t < nt; ++t)
    if (only < 0 || only == nc + t)
      failures += tma_probe(enc, tcases[t].C, tcases[t].CB, tcases[t].swz, tcases[t].halo, tcases[t].gmem, tcases[t].store, tcases[t].name) != 0;
  if (only < 0) printf("probe finished: %d failing case(s)\n", failures);
  return 0;
}
