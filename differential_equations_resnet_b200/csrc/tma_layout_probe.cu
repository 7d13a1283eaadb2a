// Hardware probe: where does TMA put a box whose inner extent (64 B) is smaller than the swizzle span
// (128 B)?  Loads a [16 ch, P, rows] box of an NHWC fp32 tensor whose value is its own linear index
// and dumps shared memory.  usage: tma_layout_probe <swizzle: 0 none,1 32B,2 64B,3 128B,4 128B_ATOM_32B>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "sm100_ptx.cuh"
using namespace b200ode;

__global__ void probe(const __grid_constant__ CUtensorMap m, float* out, int nfloats, int bytes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  float* f = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) f[i] = -1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, bytes);
    tma_load_4d(smem, &m, &bar, 0, -1, -1, 0);
  }
  mbar_wait(&bar, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = f[i];
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 4;
  const int C = 16, W = 8, H = 4, N = 1, P = W + 1, RB = 3;
  std::vector<float> h(N * H * W * C);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int nfloats = 2048;
  cudaMalloc(&o, nfloats * 4);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap m;
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {C * 4, W * C * 4, H * W * C * 4};
  cuuint32_t box[4] = {C, P, RB, 1}, es[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw[5] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B};
  CUresult r = ((Enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw[mode], CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode swizzle mode %d -> %d\n", mode, (int)r);
  if (r != CUDA_SUCCESS) return 0;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<1, 128, 16384>>>(m, o, nfloats, C * P * RB * 4);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float> out(nfloats);
  cudaMemcpy(out.data(), o, nfloats * 4, cudaMemcpyDeviceToHost);
  // print per 32-byte chunk: first float of the chunk (value -> pixel = v/16, channel = v%16), -1 = untouched, 0 = zero fill
  for (int line = 0; line < 16; ++line) {
    printf("line %2d (byte %4d):", line, line * 128);
    for (int c = 0; c < 4; ++c) {
      const float v = out[line * 32 + c * 8];
      if (v < 0) printf("   ----  ");
      else printf(" p%02d.c%02d ", (int)v / 16, (int)v % 16);
    }
    printf("\n");
  }
  return 0;
}
