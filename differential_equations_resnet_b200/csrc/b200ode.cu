// C-ABI implementation (include/b200ode.h): handles, planning, tensor-map encoding, launches.
#include "../../include/b200ode.h"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "kernels_basic.cuh"
#include "kernels_conv_tc.cuh"
#include "kernels_chain_tc.cuh"
#include "kernels_chain_f16.cuh"
#include "kernels_glue.cuh"
#include "kernels_glue_mma.cuh"
#include "kernels_wgrad_tc.cuh"
#include "kernels_bn.cuh"

using namespace b200ode;

// ------------------------------------------------------------------------------------------------
// errors, launch accounting
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(x)                                                                                     \
  do {                                                                                                  \
    cudaError_t e__ = (x);                                                                              \
    if (e__ != cudaSuccess) return fail(B200ODE_ERR_CUDA, "%s failed: %s", #x, cudaGetErrorString(e__)); \
  } while (0)
#define LAUNCH_CHECK(name)                                                                                 \
  do {                                                                                                     \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                                    \
    cudaError_t e__ = cudaGetLastError();                                                                  \
    if (e__ != cudaSuccess) return fail(B200ODE_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

static uint64_t* g_trace = nullptr;   // debug timeline buffer (device), see b200ode_debug_set_trace
extern "C" int b200ode_debug_set_trace(void* device_buffer) { g_trace = (uint64_t*)device_buffer; return 0; }

static inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

// Programmatic dependent launch (sm100_ptx.cuh): kernels of the train step's main stream that wait for their predecessor
// only after their prologue.  OPT-IN (B200ODE_PDL=1): measured SLOWER on the cfg3 train step (0.977 -> 0.999 ms per step inside
// the CUDA graph, 11 programmatic edges): the early-resident CTAs of the next chain kernel hold an SM's shared memory and TMEM
// while they wait, which keeps the side-stream weight-gradient CTAs off those SMs; the launch gap they hide is smaller than that.
static bool pdl_enabled() {
  static const bool on = getenv("B200ODE_PDL") && atoi(getenv("B200ODE_PDL")) != 0;
  return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
template <typename... Ts>
static inline bool aligned16(Ts... ptrs) {   // every non-NULL pointer is 16-byte aligned (vector kernels)
  uintptr_t acc = 0;
  ((acc |= reinterpret_cast<uintptr_t>(ptrs)), ...);
  return (acc & 15) == 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
  });
  return fn;
}

static int g_num_sms = 0;
static int device_check() {
  static int state = 0;  // 0 unknown, 1 ok, -1 bad
  if (state == 1) return 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(B200ODE_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  int dev = 0, major = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CUDA_TRY(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  if (major != 10) return fail(B200ODE_ERR_UNSUPPORTED, "compute capability %d.x: kernels are built for sm_100a only", major);
  state = 1;
  return 0;
}

// Device scratch of ONE compute call (split-K partials of the weight gradients, SIMT pre-activations, glue partials).
// The caller owns it: a block bound to the handle (b200ode_*_set_workspace) or passed with the call (glue entry points),
// sized by the *_workspace_bytes queries.  When the caller gives none the block comes from the device's stream-ordered
// memory pool (cudaMallocAsync / cudaFreeAsync on the call's stream): no device synchronisation, every call gets its own
// block (calls on different streams do not share scratch), capturable in a CUDA graph.  Nothing is retained past the call.
struct WsLease {
  void* ptr = nullptr;
  bool pooled = false;
  cudaStream_t st = nullptr;
  ~WsLease() { if (pooled && ptr) cudaFreeAsync(ptr, st); }
};
static int lease_ws(void* bound, size_t bound_bytes, size_t need, cudaStream_t st, WsLease* out) {
  if (need == 0) need = 16;
  if (bound) {
    if (bound_bytes < need)
      return fail(B200ODE_ERR_INVALID, "workspace too small: %zu bytes bound, this call needs %zu (see b200ode_*_workspace_bytes)", bound_bytes, need);
    out->ptr = bound;
    return 0;
  }
  static std::once_flag once;
  std::call_once(once, [] {   // keep freed blocks in the pool instead of returning them to the OS at every synchronisation
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
  });
  void* p = nullptr;
  cudaError_t e = cudaMallocAsync(&p, need, st);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(B200ODE_ERR_CUDA, "stream-ordered workspace allocation of %zu bytes failed: %s", need, cudaGetErrorString(e));
  }
  out->ptr = p; out->pooled = true; out->st = st;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// layer handle
// ------------------------------------------------------------------------------------------------
struct b200ode_layer {
  LayerGeom g;
  int sh, sw, mode_req, mode_eff;
  bool packed;
  float* Kdense;   // [k,k,C,C]
  float* Gdense;   // [k,k,C,C]
  float* w_hi;     // [taps][C][C] tf32 (hi)
  float* w_lo;     // strict only
  __nv_bfloat16* w_bf;
  float* bias;     // [C]
  void* ws;        // caller-owned workspace (b200ode_layer_set_workspace), may be NULL
  size_t ws_bytes;
  CUtensorMap map_w_hi, map_w_lo, map_w_bf;
  CUtensorMap map_w_bf_half, map_w_hi_half;   // CTA-pair kernels (bf16 / tf32, C >= 128): boxes of C/2 output channels
};

static void build_diag_tab(LayerGeom& g) {
  const int k = g.k;
  DiagTab& t = g.tab;
  memset(&t, 0, sizeof(t));
  for (int i = 0; i < MAX_K * MAX_K; ++i) { t.slot[i] = -1; t.sign[i] = 1; }
  if (g.layout == B200ODE_LAYOUT_3BY3) {
    // [[a,b,c],[d,gamma,-d],[-c,-b,-a]]  (tfkeras_layer_Conv2DAntisymmetric3By3.py:262-275)
    const int pos[4][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 0}};
    for (int s = 0; s < 4; ++s) {
      const int a = pos[s][0], b = pos[s][1];
      t.slot[a * 3 + b] = (int8_t)s; t.sign[a * 3 + b] = 1;
      t.slot[(2 - a) * 3 + (2 - b)] = (int8_t)s; t.sign[(2 - a) * 3 + (2 - b)] = -1;
    }
    t.nd = 4;
    return;
  }
  // general layer: free scalars in creation order (tfkeras_layer_Conv2DAntisymmetric.py:231-264)
  int nd = 0;
  for (int i = 0; i < k; ++i)
    for (int j = i; j < k; ++j) {
      if (j > i || (j == i && i <= k / 2 - 1)) {
        t.slot[i * k + j] = (int8_t)nd; t.sign[i * k + j] = 1;
        t.slot[(k - 1 - i) * k + (k - 1 - j)] = (int8_t)nd;
        t.sign[(k - 1 - i) * k + (k - 1 - j)] = g.antisym ? -1 : 1;
        ++nd;
      } else if (j == i && i == k / 2 && (k % 2) == 1 && !g.antisym) {
        t.slot[i * k + j] = (int8_t)nd; t.sign[i * k + j] = 1;
        ++nd;
      }
    }
  t.nd = nd;
}

static bool tc_channels_ok(int C) { return C == 16 || C == 32 || C == 64 || C == 128 || C == 256; }

static int make_w_map(CUtensorMap* m, void* ptr, int C, int eb, int kb, int tw, int ntaps = 9, int box_rows = 0) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(B200ODE_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)C, (cuuint64_t)ntaps};
  cuuint64_t strides[2] = {(cuuint64_t)C * eb, (cuuint64_t)C * C * eb};
  cuuint32_t box[3] = {(cuuint32_t)kb, (cuuint32_t)(box_rows > 0 ? box_rows : C), (cuuint32_t)tw};
  cuuint32_t es[3] = {1, 1, 1};
  const int rowb = kb * eb;
  CUtensorMapSwizzle sw = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(m, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ptr, dims, strides, box,
                   es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200ODE_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
  return 0;
}

static int taps_per_w_stage(int mode, int C) {
  const int eb = mode == MODE_BF16 ? 2 : 4;
  const int rowb = C * eb >= 128 ? 128 : C * eb;
  const int per_tap = C * rowb * (mode == MODE_STRICT ? 2 : 1);
  if (per_tap * 9 <= 40 * 1024) return 9;
  if (per_tap * 3 <= 56 * 1024) return 3;
  return 1;
}

extern "C" int b200ode_version(void) { return B200ODE_VERSION; }
extern "C" const char* b200ode_last_error(void) { return g_err.c_str(); }
extern "C" int b200ode_device_ok(void) { return device_check() == 0 ? 1 : 0; }
// ------------------------------------------------------------------------------------------------
// gradient exchange: NCCL bound at run time
// ------------------------------------------------------------------------------------------------
namespace {
struct NcclId { char internal[128]; };          // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed by value
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
int load_nccl() {
  if (g_nccl.lib) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(B200ODE_ERR_UNSUPPORTED, "cannot load libnccl.so.2: %s", dlerror());
  NcclApi a;
  a.lib = h;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.AllGather || !a.CommDestroy || !a.GetErrorString)
    return fail(B200ODE_ERR_UNSUPPORTED, "libnccl.so.2 lacks an expected symbol");
  g_nccl = a;
  return 0;
}
}  // namespace
struct b200ode_comm {
  void* nccl;
  int nranks, rank;
  // peer-memory gradient bucket (b200ode_comm_shared_alloc): [flags 4 KB | n floats], mapped into every rank (CUDA IPC)
  void* region;                       // this rank's allocation
  size_t region_floats;
  void* peer[B200ODE_MAX_RANKS];      // peer[r] = rank r's region in this process' address space (peer[rank] = region)
  unsigned* ctl;                      // local control words: [0] epoch, [1] block counter
};
#define NCCL_TRY(x)                                                                                          \
  do {                                                                                                       \
    int e__ = (x);                                                                                           \
    if (e__ != 0) return fail(B200ODE_ERR_CUDA, "%s failed: %s", #x, g_nccl.GetErrorString(e__));            \
  } while (0)

extern "C" int b200ode_comm_unique_id(void* id_out) {
  if (!id_out) return fail(B200ODE_ERR_INVALID, "id_out is NULL");
  if (int rc = load_nccl()) return rc;
  NCCL_TRY(g_nccl.GetUniqueId(reinterpret_cast<NcclId*>(id_out)));
  return 0;
}
extern "C" int b200ode_comm_init(int nranks, int rank, const void* nccl_unique_id, b200ode_comm_t** out) {
  if (!out || !nccl_unique_id) return fail(B200ODE_ERR_INVALID, "unique id / out is NULL");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(B200ODE_ERR_INVALID, "rank %d of %d", rank, nranks);
  if (int rc = device_check()) return rc;
  if (int rc = load_nccl()) return rc;
  NcclId id;
  memcpy(&id, nccl_unique_id, sizeof(id));
  void* c = nullptr;
  NCCL_TRY(g_nccl.CommInitRank(&c, nranks, id, rank));
  b200ode_comm* cm = new b200ode_comm();
  memset(cm, 0, sizeof(*cm));
  cm->nccl = c; cm->nranks = nranks; cm->rank = rank;
  *out = cm;
  return 0;
}
extern "C" int b200ode_comm_allreduce_bucket(b200ode_comm_t* comm, float* buf, size_t n, void* stream) {
  if (!comm || !buf) return fail(B200ODE_ERR_INVALID, "comm/buf is NULL");
  if (n == 0) return 0;
  NCCL_TRY(g_nccl.AllReduce(buf, buf, n, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm->nccl, (cudaStream_t)stream));
  return 0;
}
// ---- peer-memory gradient bucket + all-reduce fused into the optimiser step (kernels_bn.cuh: adam_p2p_kernel) ----
enum { P2P_FLAG_BYTES = 4096 };   // arrive[MAX_RANKS] at +0, done[MAX_RANKS] at +1024 (uint32 each), then the floats

extern "C" int b200ode_comm_shared_alloc(b200ode_comm_t* comm, size_t n_floats, float** local_out, float** params_out) {
  if (!comm || !local_out || n_floats == 0) return fail(B200ODE_ERR_INVALID, "comm/local_out is NULL or n_floats == 0");
  if (comm->region) return fail(B200ODE_ERR_INVALID, "this communicator already owns a shared bucket");
  if (comm->nranks > B200ODE_MAX_RANKS) return fail(B200ODE_ERR_UNSUPPORTED, "peer-memory exchange supports up to %d ranks", B200ODE_MAX_RANKS);
  const size_t n_pad = (n_floats + 3) / 4 * 4;
  const size_t bytes = P2P_FLAG_BYTES + 2 * n_pad * sizeof(float);      // [flags | gradient bucket | parameter replica]
  void* reg = nullptr;
  CUDA_TRY(cudaMalloc(&reg, bytes));
  CUDA_TRY(cudaMemset(reg, 0, bytes));
  cudaIpcMemHandle_t mine;
  CUDA_TRY(cudaIpcGetMemHandle(&mine, reg));
  // exchange the 64-byte handles with the one out-of-band channel the library has: an all-gather of bytes
  char* dev = nullptr;
  CUDA_TRY(cudaMalloc(&dev, (size_t)(comm->nranks + 1) * sizeof(mine)));
  CUDA_TRY(cudaMemcpy(dev, &mine, sizeof(mine), cudaMemcpyHostToDevice));
  NCCL_TRY(g_nccl.AllGather(dev, dev + sizeof(mine), sizeof(mine), /*ncclInt8*/ 0, comm->nccl, nullptr));
  CUDA_TRY(cudaStreamSynchronize(nullptr));
  cudaIpcMemHandle_t all[B200ODE_MAX_RANKS];
  CUDA_TRY(cudaMemcpy(all, dev + sizeof(mine), (size_t)comm->nranks * sizeof(mine), cudaMemcpyDeviceToHost));
  cudaFree(dev);
  for (int r = 0; r < comm->nranks; ++r) {
    if (r == comm->rank) { comm->peer[r] = reg; continue; }
    cudaError_t e = cudaIpcOpenMemHandle(&comm->peer[r], all[r], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(B200ODE_ERR_UNSUPPORTED, "cannot map rank %d's gradient bucket (cudaIpcOpenMemHandle: %s): no peer access between these GPUs",
                  r, cudaGetErrorString(e));
    }
  }
  CUDA_TRY(cudaMalloc((void**)&comm->ctl, 64));
  CUDA_TRY(cudaMemset(comm->ctl, 0, 64));
  comm->region = reg; comm->region_floats = n_floats;
  *local_out = reinterpret_cast<float*>(static_cast<char*>(reg) + P2P_FLAG_BYTES);
  if (params_out) *params_out = *local_out + n_pad;
  // nobody touches a peer's flags before every rank has zeroed its region
  float* tmp = *local_out;
  NCCL_TRY(g_nccl.AllReduce(tmp, tmp, 4, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm->nccl, nullptr));
  CUDA_TRY(cudaStreamSynchronize(nullptr));
  CUDA_TRY(cudaMemset(tmp, 0, 16));
  return 0;
}

extern "C" int b200ode_comm_adam_step(b200ode_comm_t* comm, float* params, const float* grads_local, float* m, float* v, int64_t n,
                                      float lr, float beta1, float beta2, float eps, const int32_t* step_counter, void* stream) {
  if (!comm || !params || !grads_local || !m || !v || !step_counter) return fail(B200ODE_ERR_INVALID, "NULL argument");
  if (!comm->region) return fail(B200ODE_ERR_INVALID, "b200ode_comm_shared_alloc must run first");
  if (n == 0) return 0;
  const float* base = reinterpret_cast<const float*>(static_cast<char*>(comm->region) + P2P_FLAG_BYTES);
  const long long off = grads_local - base;
  if (off < 0 || (size_t)(off + n) > comm->region_floats) return fail(B200ODE_ERR_INVALID, "grads_local is not inside the shared bucket");
  if ((off & 3) || (n & 3) || !aligned16(params, m, v)) return fail(B200ODE_ERR_INVALID, "slice offset / length must be multiples of 4 floats, pointers 16-byte aligned");
  const size_t n_pad = (comm->region_floats + 3) / 4 * 4;
  // parameters inside the shared region at the gradient's offset -> two-shot exchange (each rank reduces and updates
  // 1/nranks of the slice and writes the new parameters into every replica); private parameters -> one-shot
  const bool two_shot = params == base + n_pad + off;
  P2PAdamArgs A;
  memset(&A, 0, sizeof(A));
  A.nranks = comm->nranks; A.rank = comm->rank;
  for (int r = 0; r < comm->nranks; ++r) {
    char* reg = static_cast<char*>(comm->peer[r]);
    A.arrive[r] = reinterpret_cast<unsigned*>(reg);
    A.done[r] = reinterpret_cast<unsigned*>(reg + 1024);
    A.g[r] = reinterpret_cast<const float4*>(reg + P2P_FLAG_BYTES) + off / 4;
    A.pw[r] = reinterpret_cast<float4*>(reg + P2P_FLAG_BYTES) + (n_pad + off) / 4;
  }
  A.ctl = comm->ctl;
  const int sms = g_num_sms > 0 ? g_num_sms : 148;
  const long long n4 = n / 4;
  if (two_shot) {
    const long long per = (n4 + comm->nranks - 1) / comm->nranks, want = (per + 255) / 256;
    adam_p2p_shard_kernel<<<(unsigned)(want < sms ? want : sms), 256, 0, (cudaStream_t)stream>>>(A, (float4*)m, (float4*)v, n4, step_counter, lr,
                                                                                               beta1, beta2, eps, 1.0f / (float)comm->nranks);
    LAUNCH_CHECK("adam_p2p_shard_kernel");
    return 0;
  }
  const long long want = (n4 + 255) / 256;
  adam_p2p_kernel<<<(unsigned)(want < sms ? want : sms), 256, 0, (cudaStream_t)stream>>>(A, (float4*)params, (float4*)m, (float4*)v, n4,
                                                                                       step_counter, lr, beta1, beta2, eps,
                                                                                       1.0f / (float)comm->nranks);
  LAUNCH_CHECK("adam_p2p_kernel");
  return 0;
}

extern "C" int b200ode_comm_destroy(b200ode_comm_t* comm) {
  if (!comm) return 0;
  if (comm->region) {
    cudaDeviceSynchronize();
    for (int r = 0; r < comm->nranks; ++r)
      if (r != comm->rank && comm->peer[r]) cudaIpcCloseMemHandle(comm->peer[r]);
    cudaFree(comm->region);
    cudaFree(comm->ctl);
    comm->region = nullptr;
  }
  int e = g_nccl.CommDestroy ? g_nccl.CommDestroy(comm->nccl) : 0;
  delete comm;
  return e ? fail(B200ODE_ERR_CUDA, "ncclCommDestroy failed: %s", g_nccl.GetErrorString(e)) : 0;
}

extern "C" int64_t b200ode_launch_count(void) { return g_launches.load(); }

extern "C" int b200ode_layer_create(int C, int ksize, float gamma, int stride_h, int stride_w, int use_bias,
                                    int antisymmetric, int precision_mode, int param_layout, b200ode_layer_t** out) {
  if (!out) return fail(B200ODE_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (C < 1) return fail(B200ODE_ERR_INVALID, "channels must be >= 1 (got %d)", C);
  if (ksize < 1 || ksize > MAX_K || (ksize % 2) == 0) return fail(B200ODE_ERR_UNSUPPORTED, "kernel_size %d: odd sizes 1..%d are supported", ksize, MAX_K);
  if (stride_h < 1 || stride_w < 1) return fail(B200ODE_ERR_INVALID, "strides must be >= 1");
  if (precision_mode < 0 || precision_mode > 3) return fail(B200ODE_ERR_INVALID, "unknown precision_mode %d", precision_mode);
  if (param_layout != B200ODE_LAYOUT_3BY3 && param_layout != B200ODE_LAYOUT_GENERAL) return fail(B200ODE_ERR_INVALID, "unknown param_layout %d", param_layout);
  if (param_layout == B200ODE_LAYOUT_3BY3 && (ksize != 3 || !antisymmetric)) return fail(B200ODE_ERR_INVALID, "LAYOUT_3BY3 implies kernel_size 3 and antisymmetric");
  if (int rc = device_check()) return rc;
  b200ode_layer* L = new b200ode_layer();
  memset(L, 0, sizeof(*L));
  LayerGeom& g = L->g;
  g.C = C; g.k = ksize; g.layout = param_layout; g.antisym = antisymmetric ? 1 : 0; g.use_bias = use_bias ? 1 : 0; g.gamma = gamma;
  build_diag_tab(g);
  const long long kk = (long long)ksize * ksize;
  g.bias_off = (long long)g.tab.nd * C + kk * C * (C - 1) / 2;
  g.nparams = g.bias_off + (use_bias ? C : 0);
  L->sh = stride_h; L->sw = stride_w; L->mode_req = precision_mode;
  // tensor path: k = 3 everywhere; k = 5 / 7 forward and data gradient in the fp32-I/O modes (the weight gradient of
  // those layers stays on the CUDA-core kernel, which needs fp32 operands)
  const bool tc_ok = tc_channels_ok(C) && stride_h == 1 && stride_w == 1 && antisymmetric &&
                     (ksize == 3 || ((ksize == 5 || ksize == 7) && precision_mode != B200ODE_PREC_FAST_BF16));
  L->mode_eff = (precision_mode != B200ODE_PREC_SIMT_FP32 && tc_ok) ? precision_mode : B200ODE_PREC_SIMT_FP32;
  if (precision_mode == B200ODE_PREC_FAST_BF16 && !tc_ok) {
    delete L;
    return fail(B200ODE_ERR_UNSUPPORTED, "FAST_BF16 needs C in {16,32,64,128,256}, k=3, strides (1,1), antisymmetric");
  }
  const size_t kbytes = (size_t)kk * C * C * sizeof(float);
  cudaError_t e = cudaMalloc(&L->Kdense, kbytes);
  if (e == cudaSuccess) e = cudaMalloc(&L->Gdense, kbytes);
  if (e == cudaSuccess) e = cudaMalloc(&L->bias, (size_t)C * sizeof(float));
  if (e == cudaSuccess && L->mode_eff != B200ODE_PREC_SIMT_FP32) {
    if (L->mode_eff == B200ODE_PREC_FAST_BF16) e = cudaMalloc(&L->w_bf, (size_t)kk * C * C * 2);
    else {
      e = cudaMalloc(&L->w_hi, kbytes);
      if (e == cudaSuccess && L->mode_eff == B200ODE_PREC_STRICT) e = cudaMalloc(&L->w_lo, kbytes);
    }
  }
  if (e != cudaSuccess) {
    b200ode_layer_destroy(L);
    return fail(B200ODE_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
  }
  if (L->mode_eff != B200ODE_PREC_SIMT_FP32) {
    const int eb = L->mode_eff == B200ODE_PREC_FAST_BF16 ? 2 : 4;
    const int kb = (C * eb >= 128 ? 128 : C * eb) / eb;
    const int tw = ksize == 3 ? taps_per_w_stage(L->mode_eff, C) : 1;
    int rc = 0;
    if (L->w_bf) rc = make_w_map(&L->map_w_bf, L->w_bf, C, 2, kb, tw, (int)kk);
    if (!rc && L->w_bf && C >= 128 && ksize == 3) rc = make_w_map(&L->map_w_bf_half, L->w_bf, C, 2, kb, tw, 9, C / 2);
    if (!rc && L->w_hi) rc = make_w_map(&L->map_w_hi, L->w_hi, C, 4, kb, tw, (int)kk);
    if (!rc && L->w_hi && C >= 128 && ksize == 3) rc = make_w_map(&L->map_w_hi_half, L->w_hi, C, 4, kb, tw, 9, C / 2);
    if (!rc && L->w_lo) rc = make_w_map(&L->map_w_lo, L->w_lo, C, 4, kb, tw, (int)kk);
    if (rc) { b200ode_layer_destroy(L); return rc; }
  }
  *out = L;
  return 0;
}

extern "C" int b200ode_layer_destroy(b200ode_layer_t* L) {
  if (!L) return 0;
  cudaFree(L->Kdense); cudaFree(L->Gdense); cudaFree(L->w_hi); cudaFree(L->w_lo); cudaFree(L->w_bf);
  cudaFree(L->bias);
  delete L;
  return 0;
}
extern "C" int64_t b200ode_layer_num_params(const b200ode_layer_t* L) { return L ? L->g.nparams : -1; }
extern "C" int b200ode_layer_effective_mode(const b200ode_layer_t* L) { return L ? L->mode_eff : -1; }

extern "C" int b200ode_pack_kernel(b200ode_layer_t* L, const float* params, float* K_dense_hwio, void* stream) {
  if (!L || !params) return fail(B200ODE_ERR_INVALID, "layer/params is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)L->g.k * L->g.k * L->g.C * L->g.C;
  const long long n = total > L->g.C ? total : L->g.C;
  pack_kernel<<<blocks_for(n, 256), 256, 0, st>>>(L->g, params, L->Kdense, K_dense_hwio, L->w_hi, L->w_lo, L->w_bf, L->bias,
                                                 L->mode_eff == B200ODE_PREC_STRICT);
  LAUNCH_CHECK("pack_kernel");
  L->packed = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// tensor-core convolution planning
// ------------------------------------------------------------------------------------------------
struct TcPlan {
  ConvTcParams p;
  size_t smem;
  int grid;
  int box_rows, box_imgs;
};

static inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// two != 0: plan for CTA pairs (kernels_conv_tc.cuh, TWO): one image per tile, each CTA stages half of every weight tile,
// tiles are counted in pairs (the same tile-in-image of two consecutive images), the grid is even
static int plan_conv_tc(int mode, int C, int N, int H, int W, TcPlan* plan, int ksize = 3, int two = 0) {
  const int eb = mode == MODE_BF16 ? 2 : 4;
  const int rowb = C * eb >= 128 ? 128 : C * eb;
  const int nkb = C * eb / rowb;
  const bool strict = mode == MODE_STRICT;
  const int pad = ksize / 2, ntaps = ksize * ksize;
  const int P = W + pad;
  if (P > 256 || H + 2 * pad + 1 > 256)
    return fail(B200ODE_ERR_UNSUPPORTED, "tensor path supports H <= %d and W <= %d (got %dx%d)", 255 - 2 * pad, 256 - pad, H, W);
  const long long Q = (long long)H * P;
  const int tw = ksize == 3 ? taps_per_w_stage(mode, C) : 1;
  const uint32_t w_bytes = (uint32_t)tw * (two ? C / 2 : C) * rowb;
  const uint32_t w_stride = w_bytes * (strict ? 2 : 1);
  const int max_smem = 227 * 1024 - 2048;
  double best_cost = 1e30;
  TcPlan best;
  bool found = false;
  for (int whole = 0; whole < 2; ++whole) {
    for (int a = 1; a <= 32; ++a) {
      int spi, nimg, tpi, RB;
      static const int spi_env = getenv("B200ODE_CONV_SPI") ? atoi(getenv("B200ODE_CONV_SPI")) : 0;   // debug: force segments per tile
      if (spi_env > 0 && (whole || a != spi_env)) continue;
      if (whole) {
        // H+2 halo rows plus one: tap (2,2) of the last pixel reads the first pixel of halo row H+2
        // (the shared zero column), which must come from TMA zero fill, not stale shared memory
        spi = (int)((Q + 127) / 128); nimg = a; tpi = 1; RB = H + 2 * pad + 1;
        if (nimg > N && nimg > 1) break;
      } else {
        spi = a; nimg = 1; tpi = (int)((Q + 128LL * spi - 1) / (128LL * spi)); RB = (128 * spi) / P + 2 + 2 * pad;
        if (128LL * (spi - 1) >= Q) break;
        if (tpi == 1) continue;  // covered by whole-image mode
      }
      const int mt = nimg * spi;
      const int accw = strict ? 2 * C : C;
      if (mt * accw > 512 || RB > 256 || nimg > 256) continue;
      if (two && nimg != 1) continue;
      const int acc_stages = 2 * mt * accw <= 512 ? 2 : 1;
      const uint32_t a_bytes = (uint32_t)nimg * RB * P * rowb;
      const uint32_t a_lo_off = align_up(a_bytes, 1024);
      const uint32_t a_stride = strict ? 2 * a_lo_off : a_lo_off;
      // stages: at least 2 of each when possible
      int sw = ntaps / tw >= 3 ? 3 : 2;
      if (tw == 9) sw = 2;
      int sa = 2;
      long long need = (long long)sa * a_stride + (long long)sw * w_stride + 1024;
      if (need > max_smem) { sa = 1; need = (long long)sa * a_stride + (long long)sw * w_stride + 1024; }
      if (need > max_smem && sw > 2) { sw = 2; need = (long long)sa * a_stride + (long long)sw * w_stride + 1024; }   // strict C = 256
      if (need > max_smem) continue;
      while (sw < 6 && tw == 1 && need + w_stride <= max_smem) { ++sw; need += w_stride; }
      if (nkb > 1 || true) while (sa < 3 && need + a_stride <= max_smem) { ++sa; need += a_stride; }
      const long long tiles = two ? (long long)((N + 1) / 2) * tpi : whole ? (N + nimg - 1) / nimg : (long long)N * tpi;
      // Per-tile cycle model, calibrated on per-CTA timelines (tools/gpu_trace.py): MMA issue (M = 128: 40 cycles up
      // to N = 32, 48 at N = 64, N/2 above), L2 -> smem fill at ~40 B/clk, and the drain of the accumulators at ~48
      // cycles per (segment, channel) (C = 16: 7.8k cycles for 9 segments, C = 256 bf16: 22k for 2).  With double-
      // buffered accumulators the drain overlaps the next tile's MMAs; one tile's worth of the shorter phase is
      // exposed as pipeline fill/drain, which is what favours more, smaller tiles at small C.
      static const int model_env = getenv("B200ODE_CONV_COSTMODEL") ? atoi(getenv("B200ODE_CONV_COSTMODEL")) : 1;   // 0: previous model (A/B runs)
      const double per_mma = model_env ? (C >= 128 ? C / 2 : C == 64 ? 48 : 40) * (strict ? 3.0 : 1.0)
                                       : (C / 2 > 32 ? C / 2 : 32) * (strict ? 3.0 : 1.0);
      const double mma_clk = (double)mt * nkb * ntaps * (rowb / 32) * per_mma;
      const double load_clk = ((double)nimg * RB * P * C * eb + (double)ntaps * C * C * eb * (strict ? 2 : 1)) / 40.0;
      const double epi_clk = model_env ? (double)mt * C * 48.0 : (double)mt * 128 * C * 4 * 2 / 40.0;
      double tile_clk = mma_clk > load_clk ? mma_clk : load_clk;
      if (acc_stages == 1) tile_clk += epi_clk; else if (epi_clk > tile_clk) tile_clk = epi_clk;
      // single-buffered accumulators serialise MMAs and drain; measured tiles (C = 256: 125k cycles tf32, 61k bf16) run
      // ~25 % over the sum of the two phases
      if (model_env && acc_stages == 1) tile_clk *= 1.25;
      if (sa == 1) tile_clk = mma_clk + load_clk + (acc_stages == 1 ? epi_clk : 0);
      tile_clk += 600;
      const int sms = g_num_sms > 0 ? g_num_sms : 148;
      const double waves = std::ceil((double)tiles / (two ? sms / 2 : sms));
      double cost = waves * tile_clk;
      if (model_env && acc_stages == 2) cost += mma_clk < epi_clk ? mma_clk : epi_clk;
      if (cost < best_cost) {
        best_cost = cost; found = true;
        memset(&best, 0, sizeof(best));
        ConvTcParams& p = best.p;
        p.N = N; p.H = H; p.W = W; p.P = P; p.RB = RB; p.nimg = nimg; p.spi = spi; p.tpi = tpi; p.total_tiles = (int)tiles;
        p.sa = sa; p.sw = sw; p.tw = tw;
        p.ksize = ksize; p.ntaps = ntaps; p.pad = pad;
        p.a_bytes = a_bytes; p.a_lo_off = a_lo_off; p.a_stride = a_stride; p.w_bytes = w_bytes; p.w_stride = w_stride;
        p.a_off = 0; p.w_off = (uint32_t)sa * a_stride; p.bar_off = p.w_off + (uint32_t)sw * w_stride;
        uint32_t cols = (uint32_t)acc_stages * mt * accw, pc = 32;
        while (pc < cols) pc <<= 1;
        p.tmem_cols = pc;
        best.smem = (size_t)p.bar_off + 512 + 1024;
        // Cluster multicast of the weight stream (B200ODE_CONV_CLUSTER=2|4) is implemented but OFF by default:
        // measured at C = 128/256 it does not help (the limit is the shared-memory port: TMA stage writes +
        // MMA operand reads, which multicast does not reduce) and clusters of 4 CTAs of ~220 KB do not all fit
        // one wave on 148 SMs (2x slower).  Ghost tiles pad the last round when a cluster is used.
        static const int cs_env = getenv("B200ODE_CONV_CLUSTER") ? atoi(getenv("B200ODE_CONV_CLUSTER")) : 0;
        int cs = cs_env > 0 ? cs_env : 1;
        while (cs > 1 && (cs > tiles || sms % cs)) cs >>= 1;
        int grid = (int)(tiles < sms ? tiles : sms);
        grid = (grid + cs - 1) / cs * cs;
        if (grid > sms) grid = sms / cs * cs;
        p.cs = cs;
        p.iters = (int)((tiles + grid - 1) / grid);
        if (two) {   // one cluster of 2 per pair-tile
          const long long pairs = tiles < sms / 2 ? tiles : sms / 2;
          grid = (int)(2 * pairs);
          p.cs = 2;
          p.iters = (int)((tiles + pairs - 1) / pairs);
        }
        best.grid = grid;
        best.box_rows = RB; best.box_imgs = nimg;
      }
    }
  }
  if (!found) return fail(B200ODE_ERR_UNSUPPORTED, "no tensor-core tiling fits shared memory for C=%d H=%d W=%d", C, H, W);
  static const bool plan_dbg = getenv("B200ODE_PLAN_DEBUG") != nullptr;
  if (plan_dbg)
    fprintf(stderr, "[b200ode] conv plan mode=%d C=%d N=%d %dx%d: nimg=%d spi=%d tpi=%d tiles=%d grid=%d iters=%d sa=%d sw=%d tw=%d tmem=%u smem=%zu cost=%.0f\n",
            mode, C, N, H, W, best.p.nimg, best.p.spi, best.p.tpi, best.p.total_tiles, best.grid, best.p.iters, best.p.sa, best.p.sw,
            best.p.tw, best.p.tmem_cols, best.smem, best_cost);
  *plan = best;
  return 0;
}

static int make_act_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int eb, int kb, int box_w, int box_rows,
                        int box_imgs, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(B200ODE_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(B200ODE_ERR_INVALID, "activation pointer must be 16-byte aligned");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * eb, (cuuint64_t)W * C * eb, (cuuint64_t)H * W * C * eb};
  cuuint32_t box[4] = {(cuuint32_t)kb, (cuuint32_t)box_w, (cuuint32_t)box_rows, (cuuint32_t)box_imgs};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr),
                   dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200ODE_ERR_CUDA, "cuTensorMapEncodeTiled(activations) failed with %d", (int)r);
  return 0;
}

template <int MODE, int C, bool BN = false, bool GENK = false, bool TWO = false>
static int launch_conv_tc_t(const b200ode_layer* L, const TcPlan& plan, const CUtensorMap& map_a, cudaStream_t st, bool two = false) {
  if constexpr (!BN && !GENK && MODE != MODE_BF16) {
    if (plan.p.ksize != 3) return launch_conv_tc_t<MODE, C, false, true>(L, plan, map_a, st);
    if (plan.p.bn_part) return launch_conv_tc_t<MODE, C, true>(L, plan, map_a, st);
  }
  if constexpr (!TWO && !BN && !GENK && MODE != MODE_STRICT && C >= 128) {
    if (two) return launch_conv_tc_t<MODE, C, false, false, true>(L, plan, map_a, st);
  }
  auto kern = conv_tc_kernel<MODE, C, BN, GENK, TWO>;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const CUtensorMap& mw = MODE == MODE_BF16 ? (TWO ? L->map_w_bf_half : L->map_w_bf) : (TWO ? L->map_w_hi_half : L->map_w_hi);
  const CUtensorMap& mwl = MODE == MODE_STRICT ? L->map_w_lo : mw;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(plan.grid); cfg.blockDim = dim3(ConvTcCfg<MODE, C>::NWARPS * 32); cfg.dynamicSmemBytes = plan.smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = plan.p.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = plan.p.cs > 1 ? 1 : 0;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, map_a, mw, mwl, plan.p));
  LAUNCH_CHECK("conv_tc_kernel");
  return 0;
}

template <int MODE>
static int launch_conv_tc_m(const b200ode_layer* L, const TcPlan& plan, const CUtensorMap& map_a, cudaStream_t st, bool two = false) {
  switch (L->g.C) {
    case 16: return launch_conv_tc_t<MODE, 16>(L, plan, map_a, st);
    case 32: return launch_conv_tc_t<MODE, 32>(L, plan, map_a, st);
    case 64: return launch_conv_tc_t<MODE, 64>(L, plan, map_a, st);
    case 128: return launch_conv_tc_t<MODE, 128>(L, plan, map_a, st, two);
    case 256: return launch_conv_tc_t<MODE, 256>(L, plan, map_a, st, two);
  }
  return fail(B200ODE_ERR_UNSUPPORTED, "tensor path: unsupported channel count %d", L->g.C);
}

// conv_K(input) with a fused epilogue; shared by forward and data gradient
static int run_conv_tc(const b200ode_layer* L, const void* input, int N, int H, int W, ConvTcParams epi, cudaStream_t st) {
  const int mode = L->mode_eff, C = L->g.C;
  TcPlan plan;
  // CTA pairs (cta_group::2) for the wide bf16 layers; B200ODE_CONV_2CTA=0 keeps the single-CTA kernel (A/B runs)
  static const int two_env = getenv("B200ODE_CONV_2CTA") ? atoi(getenv("B200ODE_CONV_2CTA")) : 1;
  bool two = two_env && mode != MODE_STRICT && C >= 128 && L->g.k == 3 && !epi.bn_part && N >= 2;
  if (two && plan_conv_tc(mode, C, N, H, W, &plan, 3, 1) != 0) two = false;     // no pair plan fits: single-CTA kernel
  if (!two)
  if (int rc = plan_conv_tc(mode, C, N, H, W, &plan, L->g.k)) return rc;
  if (L->g.k != 3 && epi.bn_part) return fail(B200ODE_ERR_UNSUPPORTED, "internal: BatchNorm statistics from the epilogue exist for k = 3");
  ConvTcParams& p = plan.p;
  p.in = epi.in; p.skip = epi.skip; p.out = epi.out; p.z_out = epi.z_out; p.mask = epi.mask; p.bias = epi.bias;
  p.acc_scale = epi.acc_scale; p.c_in = epi.c_in; p.h = epi.h; p.relu = epi.relu; p.scale_h = epi.scale_h;
  p.mask2 = epi.mask2; p.out2 = epi.out2; p.h2 = epi.h2;
  p.bn_part = epi.bn_part;
  if (epi.bn_rows_out) *epi.bn_rows_out = plan.grid * 8;   // one row of partial sums per epilogue warp
  p.trace = g_trace;
  const int eb = mode == MODE_BF16 ? 2 : 4;
  const int rowb = C * eb >= 128 ? 128 : C * eb;
  CUtensorMap map_a;
  CUtensorMapSwizzle sw = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  if (int rc = make_act_map(&map_a, input, N, H, W, C, eb, rowb / eb, p.P, plan.box_rows, plan.box_imgs, sw)) return rc;
  switch (mode) {
    case MODE_STRICT: return launch_conv_tc_m<MODE_STRICT>(L, plan, map_a, st);
    case MODE_TF32: return launch_conv_tc_m<MODE_TF32>(L, plan, map_a, st, two);
    case MODE_BF16: return launch_conv_tc_m<MODE_BF16>(L, plan, map_a, st, two);
  }
  return fail(B200ODE_ERR_INVALID, "bad mode");
}

static size_t colsum_ws_bytes(int C) { return (size_t)2 * B200ODE_COLSUM_PARTS * C * sizeof(float); }

static ConvGeom conv_geom(const b200ode_layer* L, int N, int H, int W) {
  ConvGeom g;
  g.N = N; g.H = H; g.W = W; g.C = L->g.C; g.k = L->g.k; g.sh = L->sh; g.sw = L->sw;
  g.Ho = (H + L->sh - 1) / L->sh; g.Wo = (W + L->sw - 1) / L->sw;
  const int ph = (g.Ho - 1) * L->sh + g.k - H, pw = (g.Wo - 1) * L->sw + g.k - W;
  g.pt = (ph > 0 ? ph : 0) / 2; g.pl = (pw > 0 ? pw : 0) / 2;  // TF SAME: pad_before = total // 2
  return g;
}

extern "C" int b200ode_euler_tail(const float* z, const float* scale, const float* shift, const float* x, float* y,
                                  uint8_t* relu_mask, int64_t pixels, int channels, float h, int fuse_flags, void* stream) {
  if (!z) return fail(B200ODE_ERR_INVALID, "z is NULL");
  if ((fuse_flags & B200ODE_F_RESIDUAL) && !x) return fail(B200ODE_ERR_INVALID, "RESIDUAL needs x");
  const long long n = pixels * ((channels + 7) / 8);
  if (n == 0) return 0;
  if ((channels % 8) == 0 && aligned16(z, scale, shift, x, y)) {
    euler_tail_vec_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)z, (const float4*)scale, (const float4*)shift,
                                                                                (const float4*)x, (float4*)y, relu_mask, n, channels / 8, h,
                                                                                fuse_flags);
    LAUNCH_CHECK("euler_tail_vec_kernel");
    return 0;
  }
  euler_tail_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(z, scale, shift, x, y, relu_mask, pixels, channels, h,
                                                                          fuse_flags);
  LAUNCH_CHECK("euler_tail_kernel");
  return 0;
}

extern "C" int b200ode_euler_fwd(b200ode_layer_t* L, const void* x, void* y, uint8_t* relu_mask, float* z_out, int N, int H,
                                 int W, float h, int fuse_flags, void* stream) {
  if (!L || !x) return fail(B200ODE_ERR_INVALID, "layer/x is NULL");
  if (!L->packed) return fail(B200ODE_ERR_NOT_PACKED, "b200ode_pack_kernel must run before compute calls");
  if (N < 0 || H < 1 || W < 1) return fail(B200ODE_ERR_INVALID, "bad shape N=%d H=%d W=%d", N, H, W);
  if (!y && !z_out) return fail(B200ODE_ERR_INVALID, "need y or z_out");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool residual = fuse_flags & B200ODE_F_RESIDUAL;
  if (residual && (L->sh != 1 || L->sw != 1)) return fail(B200ODE_ERR_INVALID, "RESIDUAL needs strides (1,1)");
  // Lambda(h*x) exists only when h != 1.0 (models/tfkeras_resnets.py:90)
  const int scale_h = (fuse_flags & B200ODE_F_SCALE) && h != 1.0f;
  if (L->mode_eff != B200ODE_PREC_SIMT_FP32) {
    ConvTcParams e;
    memset(&e, 0, sizeof(e));
    e.in = residual ? x : nullptr; e.skip = nullptr; e.out = y; e.z_out = z_out; e.mask = relu_mask;
    e.bias = (fuse_flags & B200ODE_F_BIAS) && L->g.use_bias ? L->bias : nullptr;
    e.acc_scale = 1.0f; e.c_in = 1.0f; e.h = h; e.relu = (fuse_flags & B200ODE_F_RELU) ? 1 : 0; e.scale_h = scale_h;
    return run_conv_tc(L, x, N, H, W, e, st);
  }
  const ConvGeom g = conv_geom(L, N, H, W);
  const long long total = (long long)N * g.Ho * g.Wo * g.C;
  const bool plain = !(fuse_flags & (B200ODE_F_RELU | B200ODE_F_RESIDUAL)) && !scale_h && !relu_mask;
  float* zbuf = z_out;
  WsLease lease;
  if (!zbuf) {
    if (plain) zbuf = (float*)y;
    else {
      if (int rc = lease_ws(L->ws, L->ws_bytes, (size_t)total * sizeof(float), st, &lease)) return rc;
      zbuf = (float*)lease.ptr;
    }
  }
  simt_conv_fwd<<<blocks_for(total, 256), 256, 0, st>>>(g, (const float*)x, L->Kdense,
                                                       (fuse_flags & B200ODE_F_BIAS) && L->g.use_bias ? L->bias : nullptr, zbuf);
  LAUNCH_CHECK("simt_conv_fwd");
  if (plain) {
    if (y && zbuf != y) CUDA_TRY(cudaMemcpyAsync(y, zbuf, (size_t)total * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  if (!y && !relu_mask) return 0;
  return b200ode_euler_tail(zbuf, nullptr, nullptr, residual ? (const float*)x : nullptr, (float*)y, relu_mask,
                            (long long)N * g.Ho * g.Wo, g.C, h, (fuse_flags & ~B200ODE_F_SCALE) | (scale_h ? B200ODE_F_SCALE : 0), st);
}

extern "C" int b200ode_euler_dgrad(b200ode_layer_t* L, const void* dz, const void* dy_skip, void* dx, int N, int H, int W,
                                   void* stream) {
  if (!L || !dz || !dx) return fail(B200ODE_ERR_INVALID, "layer/dz/dx is NULL");
  if (!L->packed) return fail(B200ODE_ERR_NOT_PACKED, "b200ode_pack_kernel must run before compute calls");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (L->mode_eff != B200ODE_PREC_SIMT_FP32) {
    // rot180(K)^T = -K + 2*gamma*delta  =>  dx = dy_skip - conv_K(dz) + 2*gamma*dz
    ConvTcParams e;
    memset(&e, 0, sizeof(e));
    e.in = L->g.gamma != 0.0f ? dz : nullptr; e.skip = dy_skip; e.out = dx;
    e.acc_scale = -1.0f; e.c_in = 2.0f * L->g.gamma; e.h = 1.0f;
    return run_conv_tc(L, dz, N, H, W, e, st);
  }
  const ConvGeom g = conv_geom(L, N, H, W);
  const long long total = (long long)N * H * W * g.C;
  simt_conv_dgrad<<<blocks_for(total, 256), 256, 0, st>>>(g, (const float*)dz, L->Kdense, (const float*)dy_skip, (float*)dx);
  LAUNCH_CHECK("simt_conv_dgrad");
  return 0;
}

// Data gradient of one Euler step fused with the relu/h backward of the step below it:
//   dx = dy_skip - conv_K(dz) + 2*gamma*dz          (= dY of the previous step)
//   dz_prev = h * dx * prev_relu_mask                (= the dz input of the previous step's dgrad / wgrad)
extern "C" int b200ode_euler_dgrad_fused(b200ode_layer_t* L, const void* dz, const void* dy_skip, void* dx,
                                         const uint8_t* prev_relu_mask, void* dz_prev, float h, int N, int H, int W, void* stream) {
  if (!L || !dz || !dx || !prev_relu_mask || !dz_prev) return fail(B200ODE_ERR_INVALID, "layer/dz/dx/prev_relu_mask/dz_prev is NULL");
  if (!L->packed) return fail(B200ODE_ERR_NOT_PACKED, "b200ode_pack_kernel must run before compute calls");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (L->mode_eff != B200ODE_PREC_SIMT_FP32) {
    ConvTcParams e;
    memset(&e, 0, sizeof(e));
    e.in = L->g.gamma != 0.0f ? dz : nullptr; e.skip = dy_skip; e.out = dx;
    e.acc_scale = -1.0f; e.c_in = 2.0f * L->g.gamma; e.h = 1.0f;
    e.mask2 = prev_relu_mask; e.out2 = dz_prev; e.h2 = h;
    return run_conv_tc(L, dz, N, H, W, e, st);
  }
  if (int rc = b200ode_euler_dgrad(L, dz, dy_skip, dx, N, H, W, stream)) return rc;
  return b200ode_relu_scale_bwd(dx, prev_relu_mask, dz_prev, (int64_t)N * H * W, L->g.C, h, 0, stream);
}

// stage 1 of the per-channel sums: partial rows into ws ([2][parts][C]); returns the number of rows
static int colsum_stage1_launch(ColsumArgs A, float* ws, long long pixels, int C, cudaStream_t st, int* rows_out) {
  const int sms = g_num_sms > 0 ? g_num_sms : 148;
  const int C4 = C / 4;
  if ((C % 4) == 0 && C4 <= 256 && aligned16(A.a, A.b, A.scale, A.shift, A.mean, A.inv)) {
    // 16-byte loads, four per operand in flight; at least 16 pixel rows per thread (few partial rows for small
    // tensors: the second stage is pure latency), up to 4 blocks per SM for large ones
    const int rows = 256 / C4;
    long long parts = (pixels + 16LL * rows - 1) / (16LL * rows);
    parts = parts < 1 ? 1 : parts > 4LL * sms ? 4LL * sms : parts;
    if (parts > B200ODE_COLSUM_PARTS) parts = B200ODE_COLSUM_PARTS;
    if (A.mode == 0) colsum_vec_stage1<0><<<(unsigned)parts, 256, 0, st>>>(A, ws, pixels, C, (int)parts);
    else colsum_vec_stage1<1><<<(unsigned)parts, 256, 0, st>>>(A, ws, pixels, C, (int)parts);
    LAUNCH_CHECK("colsum_vec_stage1");
    *rows_out = (int)parts;
    return 0;
  }
  const int parts = (int)(pixels < 256 ? (pixels > 0 ? pixels : 1) : 256);
  colsum_stage1<<<parts, 256, 2 * 256 * sizeof(float), st>>>(A, ws, pixels, C, parts);
  LAUNCH_CHECK("colsum_stage1");
  *rows_out = parts;
  return 0;
}
static int colsum_impl(ColsumArgs A, float* out0, float* out1, float* ws, long long pixels, int C, cudaStream_t st) {
  int rows = 0;
  if (int rc = colsum_stage1_launch(A, ws, pixels, C, st, &rows)) return rc;
  bn_stats_finalize_kernel<<<blocks_for(C, 8), 256, 0, st>>>(ws, rows, C, out0, out1, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                             nullptr, nullptr, nullptr, pixels, 0.0f, 0.0f);
  LAUNCH_CHECK("bn_stats_finalize_kernel");
  return 0;
}

extern "C" int b200ode_colsum(const float* a, const float* b, float* out_sum, float* out_sumprod, float* workspace,
                              int64_t pixels, int channels, void* stream) {
  if (!a || !workspace) return fail(B200ODE_ERR_INVALID, "a/workspace is NULL");
  ColsumArgs A;
  memset(&A, 0, sizeof(A));
  A.a = a; A.b = b; A.mode = 0;
  return colsum_impl(A, out_sum, out_sumprod, workspace, pixels, channels, (cudaStream_t)stream);
}

// BatchNorm Euler step, forward part 1: z = conv_K(x) + b (fp32) and the partial sums of z, z*z per channel.
// Tensor path: the sums come from the epilogue registers of the convolution kernel (one row per epilogue warp);
// CUDA-core path: a vectorised pass over z.
extern "C" int b200ode_euler_fwd_bn_stats(b200ode_layer_t* L, const void* x, float* z_out, float* stats_ws, int* rows_out, int N,
                                          int H, int W, void* stream) {
  if (!L || !x || !z_out || !stats_ws || !rows_out) return fail(B200ODE_ERR_INVALID, "NULL argument");
  if (!L->packed) return fail(B200ODE_ERR_NOT_PACKED, "b200ode_pack_kernel must run before compute calls");
  if (N < 1 || H < 1 || W < 1) return fail(B200ODE_ERR_INVALID, "bad shape N=%d H=%d W=%d", N, H, W);
  if (L->mode_eff == B200ODE_PREC_FAST_BF16) return fail(B200ODE_ERR_UNSUPPORTED, "BatchNorm needs fp32 activations (STRICT / FAST_TF32 / SIMT)");
  cudaStream_t st = (cudaStream_t)stream;
  if (L->mode_eff != B200ODE_PREC_SIMT_FP32 && L->g.k == 3) {
    ConvTcParams e;
    memset(&e, 0, sizeof(e));
    e.z_out = z_out; e.bias = L->g.use_bias ? L->bias : nullptr; e.acc_scale = 1.0f; e.c_in = 1.0f; e.h = 1.0f;
    e.bn_part = stats_ws; e.bn_rows_out = rows_out;
    if (int rc = run_conv_tc(L, x, N, H, W, e, st)) return rc;
    if (*rows_out > B200ODE_COLSUM_PARTS) return fail(B200ODE_ERR_INVALID, "internal: %d statistic rows", *rows_out);
    return 0;
  }
  if (int rc = b200ode_euler_fwd(L, x, nullptr, nullptr, z_out, N, H, W, 1.0f, B200ODE_F_BIAS, stream)) return rc;
  const ConvGeom g = conv_geom(L, N, H, W);
  ColsumArgs A;
  memset(&A, 0, sizeof(A));
  A.a = z_out; A.mode = 0;
  return colsum_stage1_launch(A, stats_ws, (long long)N * g.Ho * g.Wo, g.C, st, rows_out);
}

extern "C" int b200ode_bn_stats_finalize(const float* stats_ws, int rows, float* out_sum, float* out_sumsq, const float* bn_gamma,
                                         const float* bn_beta, float* mean, float* inv_std, float* scale, float* shift,
                                         float* moving_mean, float* moving_var, int64_t pixels, int channels, float eps,
                                         float momentum, void* stream) {
  if (!stats_ws || rows < 1 || channels < 1) return fail(B200ODE_ERR_INVALID, "bad statistics workspace");
  if (bn_gamma && (!bn_beta || !mean || !inv_std || !scale || !shift)) return fail(B200ODE_ERR_INVALID, "NULL argument");
  if (!bn_gamma && !out_sum && !out_sumsq) return fail(B200ODE_ERR_INVALID, "nothing to compute");
  bn_stats_finalize_kernel<<<blocks_for(channels, 8), 256, 0, (cudaStream_t)stream>>>(stats_ws, rows, channels, out_sum, out_sumsq, bn_gamma,
                                                                                      bn_beta, mean, inv_std, scale, shift, moving_mean,
                                                                                      moving_var, pixels, eps, momentum);
  LAUNCH_CHECK("bn_stats_finalize_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// tensor-core weight gradient planning + launch
// ------------------------------------------------------------------------------------------------
// One launch computes the weight gradients of L layers of equal shape (L = 1: a single layer).
// Layer 0 reads its input from x0, layer l >= 1 from xrest + (l-1)*N*H*W*C (the saved outputs of
// a chain); dz of layer l is dz + l*N*H*W*C; gradients go to grad_params + l*grad_layer_stride.
// mode MODE_F16: 16-bit operands are fp16 (fp16 chains); `amax` != nullptr: the gradients carry the chain's backward
// scale chain_grad_scale(h, *amax), undone by the fold kernels.
enum { MODE_F16 = 3 };
struct WsArg {          // the caller's workspace (ptr may be NULL: stream-ordered pool); query != NULL: plan only, report the bytes
  void* ptr;
  size_t bytes;
  size_t* query;
};
static int run_wgrad_tc(int mode, const LayerGeom& lg, const void* x0, const void* xrest, const void* dz, int L, int N, int H,
                        int W, float* G, float* G_user, float* grad_params, long long grad_layer_stride, int accumulate,
                        cudaStream_t st, const WsArg& wsa, int force_mgroups = 0, const float* amax = nullptr, float amax_h = 1.0f) {
  const int C = lg.C;
  const bool f16 = mode == MODE_F16;
  if (f16) mode = MODE_BF16;
  const bool bf16 = mode == MODE_BF16, strict = mode == MODE_STRICT;
  const int eb = bf16 ? 2 : 4;
  int UKP = bf16 ? 16 : 8;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W; p.C = C; p.P = W + 1; p.L = L;
  if (p.P > 256) return fail(B200ODE_ERR_UNSUPPORTED, "tensor path supports W <= 255");
  static const int pair_env = getenv("B200ODE_WGRAD_PAIR") ? atoi(getenv("B200ODE_WGRAD_PAIR")) : 1;   // debug switch
  // strict mode too (B200ODE_WGRAD_STRICT_PAIR=0: off): half the shared memory per position of the 32-channel chunks that
  // C = 16 is otherwise padded to -> tiles twice as long beside the lo strips, half the MMAs
  static const int spair_env = getenv("B200ODE_WGRAD_STRICT_PAIR") ? atoi(getenv("B200ODE_WGRAD_STRICT_PAIR")) : 1;
  // 16-bit operands too (B200ODE_WGRAD_PAIR16=0: off): an operand row of two positions is 64 bytes instead of 32 (half the
  // shared-memory wavefronts per position) and ONE M = 128 x N = 32 MMA per kernel row covers 32 positions, where the beta
  // trick issues two M = 64 (half-rate) x N = 16 ones
  static const int pair16_env = getenv("B200ODE_WGRAD_PAIR16") ? atoi(getenv("B200ODE_WGRAD_PAIR16")) : 1;
  p.pair = ((bf16 ? pair16_env != 0 : (!strict || spair_env)) && C == 16 && (W % 2) == 0 && pair_env) ? 1 : 0;
  if (p.pair) p.P = W + 2;
  p.CH = bf16 ? (C < 64 ? C : 64) : 32;
  p.RWB = p.CH * eb;
  p.PB = p.RWB;
  if (p.pair) { p.CH = 16; p.RWB = bf16 ? 64 : 128; p.PB = p.RWB / 2; UKP = bf16 ? 32 : 16; }
  const int Cpad = C > p.CH ? C : p.CH;
  p.xchunks = Cpad / p.CH;
  p.trick = (p.pair || (p.xchunks == 1 && 4 * p.CH <= 128)) ? 1 : 0;
  p.mgroups = 1;
  // double-shift mode (kernels_wgrad_tc.cuh): 16-bit operands, one chunk holds all channels
  static const int s2_env = getenv("B200ODE_WGRAD_SHIFT2") ? atoi(getenv("B200ODE_WGRAD_SHIFT2")) : 1;   // debug switch
  // On for C = 64 (measured, 36 layers of 128x8x8x64: 100 -> 67 us; nine M = 64 MMAs per 16 positions become two M = 128 ones).
  // C = 16 / 32 (B200ODE_WGRAD_SHIFT2=2 forces it): SLOWER than the beta trick (213 vs 180 us, 85 vs 78 us) -- with 32 / 64-byte
  // operand rows an MMA costs one shared-memory wavefront per (position, chunk) unless the chunks share a 128-byte line, and
  // chunks one kernel row apart never do; the beta trick's chunks (one position apart) do.
  p.shift2 = (bf16 && !strict && !p.pair && C == p.CH && (C == 64 || (s2_env == 2 && (C == 16 || C == 32))) && s2_env) ? 1 : 0;
  if (p.shift2) {
    p.P = W + 2;
    if (p.P > 256) return fail(B200ODE_ERR_UNSUPPORTED, "tensor path supports W <= 254");
    p.trick = 0;
    p.s2_M = C == 64 ? 128 : 4 * C; p.s2_N = C == 64 ? 192 : 4 * C; p.s2_nmma = C == 64 ? 2 : 1;
    p.TG = 9; p.NT = C; p.ntapgroups = 1; p.nngroups = 1; p.MB = 1; p.Mblk = p.s2_M; p.dchunks = 1;
  } else if (p.pair) {
    p.TG = 9; p.NT = 32; p.ntapgroups = 1; p.nngroups = 1; p.MB = 1; p.Mblk = 128; p.dchunks = 1;
  } else if (p.trick) {
    p.TG = 9; p.NT = p.CH; p.ntapgroups = 1; p.nngroups = 1; p.MB = 1; p.Mblk = 4 * p.CH; p.dchunks = 1;
  } else {
    p.Mblk = Cpad < 128 ? Cpad : 128;
    // strict, C >= 128: 64-channel M blocks dealt to different CTAs -- hi + lo strips of 128 input channels leave no room
    // for a second stage of even ONE image row; M = 64 MMAs run at half rate, but this kernel is bound by its staging
    static const int sm64_env = getenv("B200ODE_WGRAD_STRICT_M64") ? atoi(getenv("B200ODE_WGRAD_STRICT_M64")) : 1;
    const bool strict_m64 = strict && sm64_env && Cpad >= 128;
    if (strict_m64) p.Mblk = 64;
    p.MB = Cpad / p.Mblk;
    // More than one M block (C = 256): deal the blocks to different CTAs.  Staging all C input channels leaves
    // room for only ~64 positions per stage (3x halo over-read, L2-bound); one 128-channel block per CTA fits
    // four times as many and lets a CTA keep 3 taps x 128 output channels in TMEM.
    // Measured (N=256, C=256) together with the row-aligned tiles below: bf16 32x32 841 -> 283 us, bf16 64x64
    // 702 -> 285 us, tf32 32x32 1264 -> 527 us.  tf32 at W = 64 leaves room for one stage only (position-granular
    // tiles stay), strict doubles every strip (hi + lo).
    static const int mg_env = getenv("B200ODE_WGRAD_MGROUPS") ? atoi(getenv("B200ODE_WGRAD_MGROUPS")) : -1;   // debug override
    const bool use_mg = force_mgroups || strict_m64 || (mg_env >= 0 ? mg_env != 0 : ((bf16 && W <= 64) || (!strict && W <= 32)));
    if (p.MB > 1 && use_mg) { p.mgroups = p.MB; p.MB = 1; p.xchunks = p.Mblk / p.CH; }
    double best = 1e30;
    for (int NT = p.CH; NT <= (C < 256 ? C : 256); NT *= 2)
      for (int TG : {9, 3, 1}) {
        if (TG * p.MB * NT * (strict ? 2 : 1) > 512) continue;
        const double load = (double)(p.Mblk * p.MB + NT) * UKP * eb / 40.0;     // channels staged per position
        const double per = (double)(NT / 2 > p.Mblk / 4 ? NT / 2 : p.Mblk / 4);
        const double mma = (double)TG * p.MB * per * (strict ? 3 : 1);
        const double t = (load > mma ? load : mma) * (9 / TG) * (C / NT) * p.mgroups;
        if (t < best - 1e-9) { best = t; p.TG = TG; p.NT = NT; }
      }
    if (getenv("B200ODE_WGRAD_TG") && getenv("B200ODE_WGRAD_NT")) {   // debug override of the tiling search
      const int tg = atoi(getenv("B200ODE_WGRAD_TG")), nt = atoi(getenv("B200ODE_WGRAD_NT"));
      if (tg * p.MB * nt * (strict ? 2 : 1) <= 512 && nt >= p.CH && nt <= C) { p.TG = tg; p.NT = nt; best = 0; }
    }
    if (best > 1e29) return fail(B200ODE_ERR_UNSUPPORTED, "no wgrad tiling for C=%d", C);
    // five taps per CTA when TMEM allows: two tap groups (5 + 4) re-read the strips twice instead of three times
    if (p.TG == 3 && 5 * p.MB * p.NT * (strict ? 2 : 1) <= 512) p.TG = 5;
    p.ntapgroups = (9 + p.TG - 1) / p.TG; p.nngroups = C / p.NT; p.dchunks = p.NT / p.CH;
  }
  // CTA pairs (kernels_wgrad_tc.cuh, TWO): the two 128-channel M blocks of C = 256 as a cluster of 2 issuing M = 256 MMAs, the
  // dz strip split between them.  B200ODE_WGRAD_2CTA=0 keeps two independent CTAs (A/B runs).
  static const int w2_env = getenv("B200ODE_WGRAD_2CTA") ? atoi(getenv("B200ODE_WGRAD_2CTA")) : 1;
  // Measured (N=256, C=256 bf16): 32x32 images 286 -> 257 us (1201 TFLOP/s = 73 % of the measured bf16 peak); 64x64 images 578 -> 664 us
  // (the row-aligned tiles of wide images are already load-balanced around whole dz strips; =2 forces pairs there for A/B runs)
  const bool two = w2_env && bf16 && !f16 && !strict && !p.trick && !p.pair && !p.shift2 && p.mgroups == 2 && p.Mblk == 128 && p.MB == 1 &&
                   (p.dchunks % 2) == 0 && (W <= 32 || w2_env == 2);
  // ... and, for ONE 128-channel M block (C = 128), pairs by TAPS: kernel rows (0 | 1) and (2 | -) per tap group
  // OFF by default (B200ODE_WGRAD_TAPPAIR=1|2 to try): measured SLOWER than three independent tap groups (N=256, C=128: 32x32 104.6 ->
  // 111.3 us, 64x64 167 -> 241 us) -- a quarter of the pair's MMA rows is junk (kernel row 3) and two tap groups balance worse than three
  static const int w2t_env = getenv("B200ODE_WGRAD_TAPPAIR") ? atoi(getenv("B200ODE_WGRAD_TAPPAIR")) : 0;
  const bool two_tap = w2t_env && w2_env && bf16 && !f16 && !strict && !p.trick && !p.pair && !p.shift2 && C == 128 && p.mgroups == 1 &&
                       p.Mblk == 128 && p.MB == 1 && (W <= 32 || w2t_env == 2);
  if (two_tap) { p.TG = 3; p.NT = 128; p.ntapgroups = 2; p.nngroups = 1; p.dchunks = p.NT / p.CH; p.tappair = 1; }
  if (two || two_tap) p.dchunks /= 2;      // chunks of the dz strip THIS CTA stages
  const int ngroups = p.ntapgroups * p.nngroups * p.mgroups;
  const int nent = p.trick ? 3 : p.TG * p.MB;
  if (p.trick && strict && 3 * p.NT * 2 > 512) return fail(B200ODE_ERR_UNSUPPORTED, "wgrad TMEM");
  uint32_t cols = p.shift2 ? (uint32_t)(p.s2_nmma * p.s2_N) : (uint32_t)nent * p.NT * (strict ? 2 : 1), pc = 32;
  while (pc < cols) pc <<= 1;
  p.tmem_cols = pc;
  const long long Q = (long long)H * p.P;
  // Layer-batched launches (the chains' weight gradients) run on a side stream under the next stage's backward sweep
  // (training.py): capped at ~half an SM's shared memory their CTAs co-reside with the chain CTAs (88-90 KB for the
  // 16- and 32-channel stages) instead of waiting for them.
  static const int wg_cap_env = getenv("B200ODE_WGRAD_SMEM_KB") ? atoi(getenv("B200ODE_WGRAD_SMEM_KB")) : 0;
  const int cap_kb = wg_cap_env > 0 ? wg_cap_env : 227;
  const int max_smem = (L > 1 && cap_kb < 227 ? cap_kb : 227) * 1024 - 2048;
  // Row-aligned tiles (bf16 / tf32, 128-channel operand blocks): a tile is R whole image rows, so the strips hold exactly
  // R+2 / R rows (x over-read (R+2)/R instead of the ~3x of position-granular tiles on row-granular strips) and
  // the k-steps run into a zeroed pad up to the next multiple of 16 positions.
  static const int rows_env = getenv("B200ODE_WGRAD_ROWS") ? atoi(getenv("B200ODE_WGRAD_ROWS")) : -1;   // debug override (0 = off)
  int rows = 0;
  // Strict mode too (round 2): its position-granular tiles had degenerated to 24 positions per 6 staged rows (hi + lo strips
  // double the shared memory, the planner insisted on two stages): 8x over-read, a TMA round trip per 24 positions, 12 % of the
  // 3xTF32 roof.  Row-aligned tiles stage R + 2 / R whole rows.
  static const int srow_env = getenv("B200ODE_WGRAD_STRICT_ROWS") ? atoi(getenv("B200ODE_WGRAD_STRICT_ROWS")) : 1;
  // ... and for the beta-trick tiling of strict mode (C <= 32; the chains' layer-batched weight gradients): position-granular
  // tiles leave room for two stages of ~120 positions beside the lo strips (7 + 5 staged rows per 3.6 useful ones, one TMA round
  // trip per tile: 650 of the stage-1 launch's 1171 us with the MMAs switched off); R whole rows per tile stage R + 2 / R rows
  // and leave room for three or more stages in flight.
  static const int strow_env = getenv("B200ODE_WGRAD_STRICT_TRICK_ROWS") ? atoi(getenv("B200ODE_WGRAD_STRICT_TRICK_ROWS")) : 0;
  const bool strict_trick_rows = strict && p.trick && !p.pair && !p.shift2 && strow_env > 0;
  if ((((!strict || (srow_env && C >= 64)) && !p.trick && !p.pair && !p.shift2 && (p.Mblk == 128 || strict)) || strict_trick_rows) && rows_env != 0) {
    auto stage_bytes = [&](int R) -> long long {
      const int kt = (R * p.P + UKP - 1) / UKP * UKP;
      const uint32_t xs = align_up((uint32_t)(kt + 2 * p.P + 3) * p.PB, 1024), ds = align_up((uint32_t)kt * p.PB, 1024);
      return ((long long)p.xchunks * xs + (long long)p.dchunks * ds) * (strict ? 2 : 1);
    };
    if (rows_env > 0) rows = rows_env;
    else if (strict_trick_rows) rows = strow_env < H ? strow_env : H;
    else {   // most rows that still leave two stages; prefer a row count with little k-step padding
      double best_cost = 1e30;
      for (int R = 1; R <= H && R <= 64; ++R) {
        if (stage_bytes(R) * 2 + 1024 + 4608 > max_smem) break;
        const int kt = (R * p.P + UKP - 1) / UKP * UKP;
        const int tpi = (H + R - 1) / R;
        // bytes staged per useful position (x rows R+2, dz padded) and MMA k-steps per useful position
        const double bytes = ((double)p.xchunks * (R + 2) * p.P + (double)p.dchunks * R * p.P) * tpi / (double)(H * p.P);
        const double mma = (double)kt * tpi / (double)(H * p.P);
        const double cost = bytes * 0.5 + mma * (p.xchunks + p.dchunks);
        if (cost < best_cost - 1e-9) { best_cost = cost; rows = R; }
      }
    }
    if (rows > 0 && (stage_bytes(rows) + 1024 + 4608 > max_smem || rows + 2 > 256)) rows = 0;
  }
  if (rows > 0) {
    p.rowtiles = 1;
    p.KT = (rows * p.P + UKP - 1) / UKP * UKP;
    p.tstride = rows * p.P;
    p.tpi = (H + rows - 1) / rows;
    p.RBx = rows + 2; p.RBd = rows;
    p.x_chunk_bytes = (uint32_t)p.RBx * p.P * p.PB; p.d_chunk_bytes = (uint32_t)p.RBd * p.P * p.PB;
    p.x_chunk_stride = align_up((uint32_t)(p.KT + 2 * p.P + 3) * p.PB, 1024); p.d_chunk_stride = align_up((uint32_t)p.KT * p.PB, 1024);
    const long long stage = ((long long)p.xchunks * p.x_chunk_stride + (long long)p.dchunks * p.d_chunk_stride) * (strict ? 2 : 1);
    p.stages = 1;
    while (p.stages < 6 && stage * (p.stages + 1) + 1024 + 4608 <= max_smem) ++p.stages;
  } else {
  // positions per tile: as large as fits two stages (one as a fallback)
  const int off_max = p.P - 1;   // largest offset of a tile start inside its strip
  // positions a tile's MMAs reach beyond its last position: x two kernel rows (+ the junk row of the double-shift
  // mode's fourth chunk), dz nothing (double shift: the three shifted chunks)
  const int xreach = (p.shift2 ? 3 : 2) * p.P + 8, dreach = p.shift2 ? 3 : 0;
  int KT = 0, stages = 0;
  for (int st_try = 2; st_try >= 1 && !KT; --st_try) {
    for (int kt = p.pair ? 1024 : 512; kt >= UKP; kt -= UKP) {
      const int RBx = (off_max + kt + xreach + p.P - 1) / p.P, RBd = (off_max + kt + dreach + p.P - 1) / p.P;
      if (RBx > 256) continue;
      const uint32_t xs = align_up((uint32_t)RBx * p.P * p.PB, 1024), ds = align_up((uint32_t)RBd * p.P * p.PB, 1024);
      const long long stage = ((long long)p.xchunks * xs + (long long)p.dchunks * ds) * (strict ? 2 : 1) + (p.pair ? (strict ? 2048 : 1024) : 0);
      if (stage * st_try + 1024 + 4608 <= max_smem) { KT = kt; stages = st_try; break; }
    }
  }
  if (!KT && p.MB > 1 && !force_mgroups)   // all-channel strips do not fit (strict C = 256, tf32 C = 256 at W = 64): one M block per CTA
    return run_wgrad_tc(f16 ? MODE_F16 : mode, lg, x0, xrest, dz, L, N, H, W, G, G_user, grad_params, grad_layer_stride, accumulate, st, wsa, 1,
                        amax, amax_h);
  if (!KT) return fail(B200ODE_ERR_UNSUPPORTED, "wgrad strips do not fit shared memory (C=%d W=%d)", C, W);
  if (KT > Q) KT = (int)((Q + UKP - 1) / UKP * UKP);
  p.tpi = (int)((Q + KT - 1) / KT);
  KT = (int)(((Q + p.tpi - 1) / p.tpi + UKP - 1) / UKP * UKP);
  {  // small tiles (small images): deepen the TMA pipeline with the shared memory that is left
    const int RBx = (off_max + KT + xreach + p.P - 1) / p.P, RBd = (off_max + KT + dreach + p.P - 1) / p.P;
    const uint32_t xs = align_up((uint32_t)RBx * p.P * p.PB, 1024), ds = align_up((uint32_t)RBd * p.P * p.PB, 1024);
    const long long stage = ((long long)p.xchunks * xs + (long long)p.dchunks * ds) * (strict ? 2 : 1) + (p.pair ? (strict ? 2048 : 1024) : 0);
    while (stages < 6 && stage * (stages + 1) + 1024 + 4608 <= max_smem) ++stages;
  }
  p.KT = KT; p.tstride = KT; p.stages = stages;
  p.RBx = (off_max + KT + xreach + p.P - 1) / p.P;
  p.RBd = (off_max + KT + dreach + p.P - 1) / p.P;
  p.x_chunk_bytes = (uint32_t)p.RBx * p.P * p.PB; p.d_chunk_bytes = (uint32_t)p.RBd * p.P * p.PB;
  p.x_chunk_stride = align_up(p.x_chunk_bytes, 1024); p.d_chunk_stride = align_up(p.d_chunk_bytes, 1024);
  }
  p.x_off = p.pair ? 1024 : 0; p.d_off = p.x_off + p.xchunks * p.x_chunk_stride;
  const uint32_t hi_bytes = p.d_off + p.dchunks * p.d_chunk_stride;
  p.x_lo_off = hi_bytes + p.x_off; p.d_lo_off = hi_bytes + p.d_off;      // the lo region mirrors the hi region (incl. the pair pad)
  p.stage_stride = strict ? 2 * hi_bytes : hi_bytes;
  p.ent_off = p.stage_stride * p.stages;
  p.bsum_off = p.ent_off + 128;
  p.bar_off = p.bsum_off + 4096;
  const size_t smem = (size_t)p.bar_off + 256 + 1024;
  p.total_tiles = N * p.tpi;
  const int sms = g_num_sms > 0 ? g_num_sms : 148;
  int nparts = sms / (ngroups * L * (two_tap ? 2 : 1));
  if (nparts < 1) nparts = 1;
  // debug / tuning: more position slices per (layer, group) than one wave of CTAs (B200ODE_WGRAD_PARTS_MULT)
  static const int pm_env = getenv("B200ODE_WGRAD_PARTS_MULT") ? atoi(getenv("B200ODE_WGRAD_PARTS_MULT")) : 1;
  if (pm_env > 1 && L > 1) nparts *= pm_env;
  if (strict) {
    // The tensor core accumulates with truncation (csrc/umma_probe.cu): the error of one TMEM accumulator grows
    // linearly with the number of k-steps it absorbs (measured at N=256, 32x32: 1.9e-9 per position, 9e-5 at the
    // 65k positions a C=256 part used to cover).  Strict mode therefore caps a part at ~2048 positions (<= 4e-6)
    // and lets the fp32 round-to-nearest reduction of the partials (fold_reduce*) carry the rest: more CTAs than
    // SMs, several waves.
    const long long want = ((long long)p.total_tiles * p.tstride + 2047) / 2048;
    if (want > nparts) nparts = (int)(want < 65535 ? want : 65535);
  }
  if (nparts > p.total_tiles) nparts = p.total_tiles;
  p.nparts = nparts;
  const long long total = 9LL * C * C;
  const long long pstride = p.pair ? 3LL * 128 * 32 : total;   // floats per partial
  p.part_stride = pstride;
  const size_t ws_need = (size_t)L * nparts * (pstride + C) * sizeof(float);
  if (wsa.query) { *wsa.query = ws_need; return 0; }
  WsLease lease;
  if (int rc = lease_ws(wsa.ptr, wsa.bytes, ws_need, st, &lease)) return rc;
  float* ws = (float*)lease.ptr;
  p.partials = ws;
  p.trace = g_trace;
  static const int wdbg_env = getenv("B200ODE_WGRAD_DBG") ? atoi(getenv("B200ODE_WGRAD_DBG")) : 0;
  p.dbg = wdbg_env;
  static const int nostack_env = getenv("B200ODE_WGRAD_NOSTACK") ? atoi(getenv("B200ODE_WGRAD_NOSTACK")) : 0;
  p.dbg_nostack = nostack_env;
  p.bias_partials = ws + (size_t)L * nparts * pstride;
  p.part_layer_stride = (long long)nparts * pstride;
  p.bias_layer_stride = (long long)nparts * C;
  CUtensorMap mx0, mx, md;
  const CUtensorMapSwizzle sw = bf16 ? (p.RWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : p.RWB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B)
                                     : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  if (p.pair) {
    // pixel-pair view: [images, H, W/2, 32 floats]; the box is P/2 = W/2+1 pairs wide (last pair out of bounds -> zeros)
    if (int rc = make_act_map(&mx0, x0, N, H, W / 2, 32, eb, 32, p.P / 2, p.RBx, 1, sw)) return rc;
    if (L > 1) { if (int rc = make_act_map(&mx, xrest, (L - 1) * N, H, W / 2, 32, eb, 32, p.P / 2, p.RBx, 1, sw)) return rc; }
    else mx = mx0;
    if (int rc = make_act_map(&md, dz, L * N, H, W / 2, 32, eb, 32, p.P / 2, p.RBd, 1, sw)) return rc;
  } else {
    if (int rc = make_act_map(&mx0, x0, N, H, W, C, eb, p.CH, p.P, p.RBx, 1, sw)) return rc;
    if (L > 1) { if (int rc = make_act_map(&mx, xrest, (L - 1) * N, H, W, C, eb, p.CH, p.P, p.RBx, 1, sw)) return rc; }
    else mx = mx0;
    if (int rc = make_act_map(&md, dz, L * N, H, W, C, eb, p.CH, p.P, p.RBd, 1, sw)) return rc;
  }
  dim3 grid(nparts, ngroups, L);
  if (two || two_tap) {
    static bool attr_set2 = false;
    if (!attr_set2) {
      CUDA_TRY(cudaFuncSetAttribute(wgrad_tc_kernel<MODE_BF16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_set2 = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * nparts, two_tap ? ngroups : ngroups / 2, L); cfg.blockDim = dim3(6 * 32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;   // x = 2 * part + M group
    cfg.attrs = attr; cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<MODE_BF16, false, true>, mx0, mx, md, p));
  } else {
#define WG_LAUNCH(M_, F_)                                                                                      \
  do {                                                                                                         \
    static bool attr_set = false;                                                                              \
    if (!attr_set) {                                                                                           \
      CUDA_TRY(cudaFuncSetAttribute(wgrad_tc_kernel<M_, F_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr_set = true;                                                                                         \
    }                                                                                                          \
    wgrad_tc_kernel<M_, F_><<<grid, (M_ == MODE_STRICT ? 10 : 6) * 32, smem, st>>>(mx0, mx, md, p);            \
  } while (0)
  if (mode == MODE_STRICT) WG_LAUNCH(MODE_STRICT, false);
  else if (mode == MODE_TF32) WG_LAUNCH(MODE_TF32, false);
  else if (f16) WG_LAUNCH(MODE_BF16, true);
  else WG_LAUNCH(MODE_BF16, false);
#undef WG_LAUNCH
  }
  LAUNCH_CHECK("wgrad_tc_kernel");
  if (G_user && L == 1) {   // dense gradient requested (tests / diagnostics)
    if (p.pair) reduce_parts_pair<<<blocks_for(total, 256), 256, 0, st>>>(ws, nparts, pstride, total, G, G_user, p.P);
    else reduce_parts<<<blocks_for(total, 256), 256, 0, st>>>(ws, nparts, total, G, G_user);
    LAUNCH_CHECK("reduce_parts");
  }
  const long long nout = (lg.use_bias ? lg.nparams : lg.bias_off);
  int fl = 1;                                   // lanes per output: power of two >= nparts, at most 16
  while (fl < 16 && fl < nparts) fl <<= 1;
  // Off-diagonal blocks through the coalesced tile kernel when it has enough threads to fill the GPU; PZ lanes per
  // block share the partials (1 / 2 / 4 for <= 4 / <= 16 / more partials).  Otherwise the per-parameter kernel with up
  // to 16 lanes per output.  Measured: C=256 / 12 partials 88 -> ~10 us; C=32 / 148 partials (9 blocks) stays on the
  // per-parameter kernel (10 us there, 110 us tiled).
  const int T = (C + 31) / 32;
  const int PZ = nparts <= 4 ? 1 : nparts <= 16 ? 2 : 4;
  const long long tblocks = (long long)T * (T + 1) / 2 * lg.k * lg.k * L;
  const bool tiled = !p.pair && lg.layout == 0 && lg.antisym && C >= 32 && (nparts <= 16 * PZ || (strict && C >= 64)) && tblocks * PZ >= 296;
  if (tiled) {
    const dim3 tgrid(T * (T + 1) / 2, lg.k * lg.k, L), tblock(32, 8, PZ);
    if (PZ == 1) fold_reduce_tiled_kernel<1><<<tgrid, tblock, 0, st>>>(lg, ws, nparts, pstride, grad_params, accumulate, p.part_layer_stride, grad_layer_stride, amax, amax_h);
    else if (PZ == 2) fold_reduce_tiled_kernel<2><<<tgrid, tblock, 0, st>>>(lg, ws, nparts, pstride, grad_params, accumulate, p.part_layer_stride, grad_layer_stride, amax, amax_h);
    else fold_reduce_tiled_kernel<4><<<tgrid, tblock, 0, st>>>(lg, ws, nparts, pstride, grad_params, accumulate, p.part_layer_stride, grad_layer_stride, amax, amax_h);
    LAUNCH_CHECK("fold_reduce_tiled_kernel");
  }
  const long long nfold = tiled ? 4LL * C + (lg.use_bias ? C : 0) : nout;
  dim3 fgrid(blocks_for(nfold * fl, 256), L);
  fold_reduce_kernel<<<fgrid, 256, 0, st>>>(lg, ws, nparts, pstride, p.bias_partials, grad_params, accumulate,
                                            p.part_layer_stride, p.bias_layer_stride, grad_layer_stride, fl, p.pair ? p.P : 0,
                                            tiled ? 1 : 0, amax, amax_h);
  LAUNCH_CHECK("fold_reduce_kernel");
  return 0;
}

extern "C" int b200ode_euler_wgrad(b200ode_layer_t* L, const void* x, const void* dz, float* grad_params, float* G_dense,
                                   int N, int H, int W, int accumulate, void* stream) {
  if (!L || !x || !dz || !grad_params) return fail(B200ODE_ERR_INVALID, "layer/x/dz/grad_params is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const ConvGeom g = conv_geom(L, N, H, W);
  const long long total = (long long)g.k * g.k * g.C * g.C;
  const long long npix = (long long)N * g.Ho * g.Wo;
  WsLease lease;
  float* colsum_ws = nullptr;
  if (L->mode_eff != B200ODE_PREC_SIMT_FP32 && L->g.k == 3) {
    return run_wgrad_tc(L->mode_eff, L->g, x, x, dz, 1, N, H, W, L->Gdense, G_dense, grad_params, 0, accumulate, st,
                        WsArg{L->ws, L->ws_bytes, nullptr});
  } else {
    int parts = (int)(npix / 64);
    if (parts < 1) parts = 1;
    if (parts > 128) parts = 128;
    if (int rc = lease_ws(L->ws, L->ws_bytes, (size_t)parts * total * sizeof(float) + colsum_ws_bytes(g.C), st, &lease)) return rc;
    float* ws = (float*)lease.ptr;
    colsum_ws = ws + (size_t)parts * total;
    dim3 grid(blocks_for(total, 128), parts);
    simt_conv_wgrad<<<grid, 128, 0, st>>>(g, (const float*)x, (const float*)dz, ws, parts);
    LAUNCH_CHECK("simt_conv_wgrad");
    reduce_parts<<<blocks_for(total, 256), 256, 0, st>>>(ws, parts, total, L->Gdense, G_dense);
    LAUNCH_CHECK("reduce_parts");
  }
  const long long nfree = L->g.use_bias ? L->g.bias_off : L->g.nparams;
  fold_grad<<<blocks_for(nfree, 256), 256, 0, st>>>(L->g, L->Gdense, grad_params, accumulate);
  LAUNCH_CHECK("fold_grad");
  if (L->g.use_bias) {
    if (accumulate) return fail(B200ODE_ERR_UNSUPPORTED, "accumulate with bias is not implemented");
    if (L->mode_eff == B200ODE_PREC_FAST_BF16) {
      return fail(B200ODE_ERR_UNSUPPORTED, "bias gradient in FAST_BF16 mode: use b200ode_colsum on an fp32 dz");
    }
    ColsumArgs A;
    memset(&A, 0, sizeof(A));
    A.a = (const float*)dz; A.mode = 0;
    return colsum_impl(A, grad_params + L->g.bias_off, nullptr, colsum_ws, npix, g.C, st);
  }
  return 0;
}

extern "C" int b200ode_relu_scale_bwd(const void* dy, const uint8_t* relu_mask, void* dz, int64_t pixels, int channels, float h,
                                      int is_bf16, void* stream) {
  if (!dy || !relu_mask || !dz) return fail(B200ODE_ERR_INVALID, "NULL argument");
  const long long n = pixels * ((channels + 7) / 8);
  if (n == 0) return 0;
  if (!is_bf16 && (channels % 8) == 0 && aligned16(dy, dz)) {
    const int sms = g_num_sms > 0 ? g_num_sms : 148;
    const long long want = (n + 255) / 256;
    relu_scale_bwd_vec_kernel<<<(unsigned)(want < 16LL * sms ? want : 16LL * sms), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)dy, relu_mask, (float4*)dz, n, h);
    LAUNCH_CHECK("relu_scale_bwd_vec_kernel");
    return 0;
  }
  if (is_bf16)
    relu_scale_bwd_kernel<__nv_bfloat16><<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dy, relu_mask, (__nv_bfloat16*)dz, pixels, channels, h);
  else
    relu_scale_bwd_kernel<float><<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const float*)dy, relu_mask, (float*)dz,
                                                                                       pixels, channels, h);
  LAUNCH_CHECK("relu_scale_bwd_kernel");
  return 0;
}

extern "C" int b200ode_bn_finalize(const float* sum, const float* sumsq, const float* bn_gamma, const float* bn_beta, float* mean,
                                   float* inv_std, float* scale, float* shift, float* moving_mean, float* moving_var,
                                   int64_t pixels, int channels, float eps, float momentum, void* stream) {
  if (!sum || !sumsq || !bn_gamma || !bn_beta || !mean || !inv_std || !scale || !shift) return fail(B200ODE_ERR_INVALID, "NULL argument");
  bn_finalize_kernel<<<blocks_for(channels, 128), 128, 0, (cudaStream_t)stream>>>(sum, sumsq, bn_gamma, bn_beta, mean, inv_std, scale,
                                                                                  shift, moving_mean, moving_var, pixels, channels,
                                                                                  eps, momentum);
  LAUNCH_CHECK("bn_finalize_kernel");
  return 0;
}

extern "C" int b200ode_bn_bwd_reduce(const float* dy, const float* z, const float* scale, const float* shift, const float* mean,
                                     const float* inv_std, float* dgamma, float* dbeta, float* workspace, int64_t pixels,
                                     int channels, float h, void* stream) {
  if (!dy || !z || !workspace) return fail(B200ODE_ERR_INVALID, "NULL argument");
  ColsumArgs A;
  A.a = dy; A.b = z; A.scale = scale; A.shift = shift; A.mean = mean; A.inv = inv_std; A.h = h; A.mode = 1;
  return colsum_impl(A, dbeta, dgamma, workspace, pixels, channels, (cudaStream_t)stream);
}

extern "C" int b200ode_bn_bwd_apply(const float* dy, const float* z, const float* scale, const float* shift, const float* mean,
                                    const float* inv_std, const float* bn_gamma, const float* dgamma, const float* dbeta,
                                    float* dz, int64_t pixels, int channels, float h, void* stream) {
  if (!dy || !z || !dz) return fail(B200ODE_ERR_INVALID, "NULL argument");
  const long long n = pixels * channels;
  if (n == 0) return 0;
  const int C4 = channels / 4;
  if ((channels % 4) == 0 && C4 <= 256 && (256 % C4) == 0 &&
      aligned16(dy, z, scale, shift, mean, inv_std, bn_gamma, dgamma, dbeta, dz)) {   // a thread keeps its 4-channel group: per-channel terms in registers
    const int sms = g_num_sms > 0 ? g_num_sms : 148;
    const long long n4 = n / 4, want = (n4 + 1023) / 1024;
    bn_bwd_apply_vec_kernel<<<(unsigned)(want < 8LL * sms ? want : 8LL * sms), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)dy, (const float4*)z, scale, shift, mean, inv_std, bn_gamma, dgamma, dbeta, (float4*)dz, n4, C4, pixels, h);
    LAUNCH_CHECK("bn_bwd_apply_vec_kernel");
    return 0;
  }
  bn_bwd_apply_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, z, scale, shift, mean, inv_std, bn_gamma, dgamma,
                                                                            dbeta, dz, pixels, channels, h);
  LAUNCH_CHECK("bn_bwd_apply_kernel");
  return 0;
}

extern "C" int b200ode_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, const int32_t* step_counter,
                                 float lr, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (!params || !grads || !m || !v || !step_counter) return fail(B200ODE_ERR_INVALID, "NULL argument");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  long long done = 0;
  if (n >= 1024 && !(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)m | (uintptr_t)v) & 15)) {   // 16-byte vectors + scalar tail
    const int sms = g_num_sms > 0 ? g_num_sms : 148;
    const long long n4 = n / 4, want = (n4 + 255) / 256;
    adam_vec_kernel<<<(unsigned)(want < 8LL * sms ? want : 8LL * sms), 256, 0, st>>>((float4*)params, (const float4*)grads, (float4*)m,
                                                                                   (float4*)v, n4, step_counter, lr, beta1, beta2, eps, grad_scale);
    LAUNCH_CHECK("adam_vec_kernel");
    done = n4 * 4;
  }
  if (done < n) {
    adam_kernel<<<blocks_for(n - done, 256), 256, 0, st>>>(params + done, grads + done, m + done, v + done, n - done, step_counter, lr,
                                                          beta1, beta2, eps, grad_scale);
    LAUNCH_CHECK("adam_kernel");
  }
  return 0;
}

// Host-only view of the tile planner (no device needed): lets the CPU test-suite pin the tiling decisions the
// measurements in DESIGN.md rest on.  out = {nimg, spi, tpi, total_tiles, grid, sa, sw, accumulator stages}.
extern "C" int b200ode_debug_conv_plan(int precision_mode, int channels, int N, int H, int W, int* out8) {
  if (!out8) return fail(B200ODE_ERR_INVALID, "out8 is NULL");
  if (precision_mode < 0 || precision_mode > 2 || !tc_channels_ok(channels))
    return fail(B200ODE_ERR_UNSUPPORTED, "tensor-core plans exist for STRICT / FAST_TF32 / FAST_BF16 and C in {16,32,64,128,256}");
  TcPlan plan;
  if (int rc = plan_conv_tc(precision_mode, channels, N, H, W, &plan)) return rc;
  const int accw = precision_mode == MODE_STRICT ? 2 * channels : channels, mt = plan.p.nimg * plan.p.spi;
  out8[0] = plan.p.nimg; out8[1] = plan.p.spi; out8[2] = plan.p.tpi; out8[3] = plan.p.total_tiles; out8[4] = plan.grid;
  out8[5] = plan.p.sa; out8[6] = plan.p.sw; out8[7] = 2 * mt * accw <= 512 ? 2 : 1;
  return 0;
}

extern "C" int b200ode_gradient_mean_norms(const float* grads, const int64_t* offsets, const int64_t* sizes, int n_slices,
                                           float grad_scale, float* out, void* stream) {
  if (!grads || !offsets || !sizes || !out) return fail(B200ODE_ERR_INVALID, "NULL argument");
  if (n_slices <= 0) return 0;
  static_assert(sizeof(long long) == sizeof(int64_t), "int64_t layout");
  segment_mean_norm_kernel<<<n_slices, 256, 0, (cudaStream_t)stream>>>(grads, reinterpret_cast<const long long*>(offsets),
                                                                       reinterpret_cast<const long long*>(sizes), grad_scale, out);
  LAUNCH_CHECK("segment_mean_norm_kernel");
  return 0;
}

extern "C" int b200ode_increment(int32_t* counter, void* stream) {
  if (!counter) return fail(B200ODE_ERR_INVALID, "NULL argument");
  increment_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counter);
  LAUNCH_CHECK("increment_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// persistent Euler-step chains (kernels_chain_tc.cuh): one launch per direction per residual stage
// ------------------------------------------------------------------------------------------------
struct b200ode_chain {
  LayerGeom g;
  int L;           // distinct weight layers
  int mode;        // B200ODE_PREC_FAST_TF32 / B200ODE_PREC_STRICT (kernels_chain_tc.cuh) or B200ODE_PREC_FAST_F16 (kernels_chain_f16.cuh)
  bool packed;
  float* w_hi;     // FAST_TF32: [L][9][C][C] tf32-rounded, K-major B operand; STRICT: [L][9][2][C][C] = per tap the entries
                   // truncated to tf32 followed by their remainders (3xTF32; one 2C-row B tile per tap)
  float* w_lo;     // unused (kept NULL)
  __half* w16;     // FAST_F16:  [L][9][C][C] fp16, K-major B operand
  float* bias;     // [L][C]
  const float* amax_cur;   // FAST_F16: the scalar the last backward sweep used (amax, or the caller's of b200ode_chain_dgrad_amax)
  float* amax;     // FAST_F16: device scalar max|dy| of the last backward sweep (scale of dz_all)
  float amax_h;    //           and the step size it was taken with
  void* ws;        // caller-owned workspace (b200ode_chain_set_workspace), may be NULL
  size_t ws_bytes;
};

struct ChainPlan {
  ChainParams p;
  size_t smem;
};

static bool chain_channels_ok(int C) { return C == 16 || C == 32 || C == 64; }

// Shared-memory plan of chain_tc_kernel; returns false when a whole image does not fit.
static bool plan_chain(int C, int H, int W, int dir, ChainPlan* plan, bool strict = false) {
  if (!chain_channels_ok(C)) return false;
  const int rowb = C * 4 >= 128 ? 128 : C * 4;
  const int nkb = C * 4 / rowb;
  const int P = W + 1;
  if (P > 256 || H + 2 > 256) return false;
  const int maxseg = C == 16 ? 9 : C == 32 ? 5 : 2;
  const int nseg = (H * P + 127) / 128;
  if (nseg > maxseg || nseg * C > 512) return false;
  ChainParams p;
  memset(&p, 0, sizeof(p));
  p.H = H; p.W = W; p.P = P; p.nseg = nseg;
  p.plane_bytes = align_up((uint32_t)((H + 2) * P + 1) * rowb, 1024);
  p.strip_stride = (uint32_t)nkb * p.plane_bytes;
  p.x_bytes = (uint32_t)(H + 2) * P * rowb;
  p.tw = taps_per_w_stage(MODE_TF32, C);
  p.w_stage_bytes = (uint32_t)p.tw * C * rowb * (strict ? 2 : 1);     // strict: 2C rows per tap (W_hi rows, then W_lo rows)
  p.seg_outer = (nkb == 1 && p.tw == 9) ? 1 : 0;
  p.e_off = 2 * p.strip_stride;
  const uint32_t e_bytes = dir ? align_up((uint32_t)H * W * C * 4, 1024) : 0;
  p.w_off = p.e_off + e_bytes;
  const long long max_smem = 227 * 1024 - 1024 - 512;
  const int per_layer = (9 / p.tw) * nkb;           // ring entries one layer consumes
  int sw = 2 * per_layer;                            // ideally: the next layer fully prefetched
  while (sw > 1 && (long long)p.w_off + (long long)sw * p.w_stage_bytes > max_smem) --sw;
  if ((long long)p.w_off + (long long)sw * p.w_stage_bytes > max_smem) return false;
  if (sw < per_layer && !strict) return false;       // the fast kernel's MMA warp waits for a whole layer's entries up front
  p.sw = sw;
  p.bar_off = p.w_off + (uint32_t)sw * p.w_stage_bytes;
  // the MMAs of the last segment read (junk rows) up to 128*nseg + 2P + 2 positions of strip 1: keep that inside the allocation
  const long long reach = (long long)p.strip_stride + (long long)(nkb - 1) * p.plane_bytes + (long long)(128 * nseg + 2 * P + 3) * rowb;
  if (reach > (long long)p.bar_off) return false;
  uint32_t cols = (uint32_t)nseg * C * (strict ? 2 : 1), pc = 32;      // strict: [main | correction] column ranges per segment
  if (cols > 512) return false;
  while (pc < cols) pc <<= 1;
  p.tmem_cols = pc;
  plan->p = p;
  plan->smem = (size_t)p.bar_off + 512 + 1024;
  return true;
}

struct ChainF16Plan {
  ChainF16Params p;
  size_t smem;
};

// Shared-memory plan of chain_f16_kernel (both directions use the same layout); false when a whole image does not fit.
static bool plan_chain_f16(int C, int H, int W, ChainF16Plan* plan) {
  if (!chain_channels_ok(C)) return false;
  const int rowb = C * 2;
  const int P = W + 1;
  if (P > 256 || H + 2 > 256) return false;
  const int maxseg = C == 16 ? 9 : C == 32 ? 5 : 2;
  const int nseg = (H * P + 127) / 128;
  if (nseg > maxseg || nseg * C > 512) return false;
  ChainF16Params p;
  memset(&p, 0, sizeof(p));
  p.H = H; p.W = W; p.P = P; p.nseg = nseg;
  // rows a strip must hold: the halo image, and what the MMAs of the last segment reach (junk rows included)
  const long long rows_img = (long long)(H + 2) * P + 1, rows_mma = 128LL * nseg + 2 * P + 3;
  p.strip_stride = align_up((uint32_t)((rows_img > rows_mma ? rows_img : rows_mma) * rowb), 1024);
  if (2 * P + 2 > 128 + P + 1) return false;          // wavefront hand-over: a segment's A rows reach into the NEXT segment only
  p.w_off = 2 * p.strip_stride;                       // (the fp32 residual lives in registers)
  p.w_box_bytes = (uint32_t)9 * C * rowb;
  p.w_layer_bytes = align_up(p.w_box_bytes, 1024);
  p.sw = 2;                                           // the next layer's weights are in flight while this one runs
  p.bar_off = p.w_off + (uint32_t)p.sw * p.w_layer_bytes;
  uint32_t cols = (uint32_t)nseg * C, pc = 32;
  while (pc < cols) pc <<= 1;
  p.tmem_cols = pc;
  plan->p = p;
  plan->smem = (size_t)p.bar_off + 512 + 1024;
  return plan->smem <= (size_t)227 * 1024;
}

extern "C" int b200ode_chain_supported(int channels, int H, int W, int precision_mode) {
  if (precision_mode == B200ODE_PREC_FAST_F16) {
    ChainF16Plan pl;
    return plan_chain_f16(channels, H, W, &pl) ? 1 : 0;
  }
  if (precision_mode != B200ODE_PREC_FAST_TF32 && precision_mode != B200ODE_PREC_STRICT) return 0;
  ChainPlan pl;
  const bool strict = precision_mode == B200ODE_PREC_STRICT;
  return plan_chain(channels, H, W, 0, &pl, strict) && plan_chain(channels, H, W, 1, &pl, strict) ? 1 : 0;
}

extern "C" int b200ode_chain_create(int channels, int n_layers, float gamma, int use_bias, int precision_mode,
                                    b200ode_chain_t** out) {
  if (!out) return fail(B200ODE_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (n_layers < 1) return fail(B200ODE_ERR_INVALID, "n_layers must be >= 1");
  if (precision_mode != B200ODE_PREC_FAST_TF32 && precision_mode != B200ODE_PREC_FAST_F16 && precision_mode != B200ODE_PREC_STRICT)
    return fail(B200ODE_ERR_UNSUPPORTED, "chains run in STRICT, FAST_TF32 or FAST_F16 mode");
  if (!chain_channels_ok(channels)) return fail(B200ODE_ERR_UNSUPPORTED, "chains need C in {16,32,64} (got %d)", channels);
  if (int rc = device_check()) return rc;
  b200ode_chain* ch = new b200ode_chain();
  memset(ch, 0, sizeof(*ch));
  LayerGeom& g = ch->g;
  g.C = channels; g.k = 3; g.layout = B200ODE_LAYOUT_3BY3; g.antisym = 1; g.use_bias = use_bias ? 1 : 0; g.gamma = gamma;
  build_diag_tab(g);
  g.bias_off = (long long)g.tab.nd * channels + 9LL * channels * (channels - 1) / 2;
  g.nparams = g.bias_off + (use_bias ? channels : 0);
  ch->L = n_layers; ch->mode = precision_mode; ch->amax_h = 1.0f;
  const size_t wn = (size_t)n_layers * 9 * channels * channels;
  cudaError_t e = precision_mode == B200ODE_PREC_FAST_F16 ? cudaMalloc(&ch->w16, wn * sizeof(__half))
                                                          : cudaMalloc(&ch->w_hi, wn * sizeof(float) * (precision_mode == B200ODE_PREC_STRICT ? 2 : 1));
  if (e == cudaSuccess) e = cudaMalloc(&ch->bias, (size_t)n_layers * channels * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&ch->amax, sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(ch->amax, 0, sizeof(float));
  if (e != cudaSuccess) {
    b200ode_chain_destroy(ch);
    return fail(B200ODE_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
  }
  *out = ch;
  return 0;
}

extern "C" int b200ode_chain_destroy(b200ode_chain_t* ch) {
  if (!ch) return 0;
  cudaFree(ch->w_hi); cudaFree(ch->w_lo); cudaFree(ch->w16); cudaFree(ch->bias); cudaFree(ch->amax);
  delete ch;
  return 0;
}
extern "C" int64_t b200ode_chain_layer_params(const b200ode_chain_t* ch) { return ch ? ch->g.nparams : -1; }

extern "C" int b200ode_chain_pack(b200ode_chain_t* ch, const float* params, int64_t param_layer_stride, void* stream) {
  if (!ch || !params) return fail(B200ODE_ERR_INVALID, "chain/params is NULL");
  const long long total = 9LL * ch->g.C * ch->g.C;
  dim3 grid(blocks_for(total, 256), ch->L);
  if (ch->mode == B200ODE_PREC_FAST_F16)
    pack_chain_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ch->g, params, param_layer_stride, ch->w16, ch->bias);
  else
    pack_chain_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ch->g, params, param_layer_stride, ch->w_hi, ch->mode == B200ODE_PREC_STRICT ? 1 : 0, ch->bias);
  LAUNCH_CHECK("pack_chain_kernel");
  ch->packed = true;
  return 0;
}

static int make_chain_w_map(CUtensorMap* m, const b200ode_chain* ch, int tw) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(B200ODE_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  const int C = ch->g.C;
  const int rowb = C * 4 >= 128 ? 128 : C * 4;
  const int rows = ch->mode == B200ODE_PREC_STRICT ? 2 * C : C;       // strict: W_hi rows then W_lo rows per tap
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)9 * ch->L};
  cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)rows * C * 4};
  cuuint32_t box[3] = {(cuuint32_t)(rowb / 4), (cuuint32_t)rows, (cuuint32_t)tw};
  cuuint32_t es[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ch->w_hi, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200ODE_ERR_CUDA, "cuTensorMapEncodeTiled(chain weights) failed with %d", (int)r);
  return 0;
}

// Grid and cluster size: one CTA per image (up to the SM count); clusters share the weight stream.
static int chain_grid(ChainParams& p, int C, int N) {
  static const int cs_env = getenv("B200ODE_CHAIN_CLUSTER") ? atoi(getenv("B200ODE_CHAIN_CLUSTER")) : 0;   // debug override
  const int sms = g_num_sms > 0 ? g_num_sms : 148;
  int cs = cs_env > 0 ? cs_env : (C >= 32 ? 4 : 1);   // 8 halves the resident CTAs (measured); 2/4 cost nothing and cut L2 reads
  while (cs > 1 && cs > N) cs >>= 1;
  int grid = N < sms ? N : sms;
  grid = (grid + cs - 1) / cs * cs;
  if (grid > sms) grid = sms / cs * cs;
  p.cs = cs;
  p.dbg_skip_w = getenv("B200ODE_CHAIN_SKIPW") ? 1 : 0;
  p.iters = (N + grid - 1) / grid;
  return grid;
}

template <int DIR, bool ST = false>
static int launch_chain(const b200ode_chain* ch, const ChainPlan& plan, const CUtensorMap& mx, const CUtensorMap& mw, int grid,
                        cudaStream_t st) {
#define CH_LAUNCH(C_)                                                                                              \
  do {                                                                                                             \
    static bool attr_set = false;                                                                                  \
    if (!attr_set) {                                                                                               \
      CUDA_TRY(cudaFuncSetAttribute(chain_tc_kernel<C_, DIR, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr_set = true;                                                                                             \
    }                                                                                                              \
    cudaLaunchConfig_t cfg;                                                                                        \
    memset(&cfg, 0, sizeof(cfg));                                                                                  \
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = plan.smem; cfg.stream = st;          \
    cudaLaunchAttribute attr[1];                                                                                   \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                                              \
    attr[0].val.clusterDim.x = plan.p.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;              \
    cfg.attrs = attr; cfg.numAttrs = plan.p.cs > 1 ? 1 : 0;                                                         \
    CUDA_TRY(cudaLaunchKernelEx(&cfg, chain_tc_kernel<C_, DIR, ST>, mx, mw, plan.p));                               \
  } while (0)
  switch (ch->g.C) {
    case 16: CH_LAUNCH(16); break;
    case 32: CH_LAUNCH(32); break;
    case 64: CH_LAUNCH(64); break;
    default: return fail(B200ODE_ERR_UNSUPPORTED, "chain: unsupported channel count %d", ch->g.C);
  }
#undef CH_LAUNCH
  LAUNCH_CHECK("chain_tc_kernel");
  return 0;
}

static int make_chain_w16_map(CUtensorMap* m, const b200ode_chain* ch) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(B200ODE_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  const int C = ch->g.C, rowb = C * 2;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)C, (cuuint64_t)9 * ch->L};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)C * C * 2};
  cuuint32_t box[3] = {(cuuint32_t)C, (cuuint32_t)C, 9};          // all nine taps of a layer in one box
  cuuint32_t es[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, ch->w16, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200ODE_ERR_CUDA, "cuTensorMapEncodeTiled(fp16 chain weights) failed with %d", (int)r);
  return 0;
}

static int chain_f16_grid(ChainF16Params& p, int C, int N) {
  static const int cs_env = getenv("B200ODE_CHAIN_CLUSTER") ? atoi(getenv("B200ODE_CHAIN_CLUSTER")) : 0;   // debug override
  const int sms = g_num_sms > 0 ? g_num_sms : 148;
  int cs = cs_env > 0 ? cs_env : (C >= 32 ? 4 : 1);
  while (cs > 1 && cs > N) cs >>= 1;
  int grid = N < sms ? N : sms;
  grid = (grid + cs - 1) / cs * cs;
  if (grid > sms) grid = sms / cs * cs;
  p.cs = cs;
  p.iters = (N + grid - 1) / grid;
  return grid;
}

template <int DIR>
static int launch_chain_f16(const b200ode_chain* ch, const ChainF16Plan& plan, const CUtensorMap& mw, int grid, cudaStream_t st) {
  static const int ew_env = getenv("B200ODE_CHAIN_EW") ? atoi(getenv("B200ODE_CHAIN_EW")) : 0;   // debug: epilogue warps (8 | 16)
  const int ew = ew_env == 8 || ew_env == 16 || (ew_env == 12 && ch->g.C == 16) ? ew_env : 8;
#define CHF_LAUNCH(C_, EW_)                                                                                       \
  do {                                                                                                            \
    static bool attr_set = false;                                                                                 \
    if (!attr_set) {                                                                                              \
      CUDA_TRY(cudaFuncSetAttribute(chain_f16_kernel<C_, DIR, EW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr_set = true;                                                                                            \
    }                                                                                                             \
    cudaLaunchConfig_t cfg;                                                                                       \
    memset(&cfg, 0, sizeof(cfg));                                                                                 \
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + EW_ * 32); cfg.dynamicSmemBytes = plan.smem; cfg.stream = st; \
    cudaLaunchAttribute attr[2];                                                                                  \
    int na = 0;                                                                                                   \
    if (plan.p.cs > 1) {                                                                                          \
      attr[na].id = cudaLaunchAttributeClusterDimension;                                                          \
      attr[na].val.clusterDim.x = plan.p.cs; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;        \
      ++na;                                                                                                       \
    }                                                                                                             \
    if (pdl_enabled()) {                                                                                          \
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                           \
      attr[na].val.programmaticStreamSerializationAllowed = 1;                                                    \
      ++na;                                                                                                       \
    }                                                                                                             \
    cfg.attrs = attr; cfg.numAttrs = na;                                                                          \
    CUDA_TRY(cudaLaunchKernelEx(&cfg, chain_f16_kernel<C_, DIR, EW_>, mw, plan.p));                                \
  } while (0)
#define CHF_EW(C_) do { if (ew == 8) CHF_LAUNCH(C_, 8); else CHF_LAUNCH(C_, 16); } while (0)
  switch (ch->g.C) {
    case 16: if (ew == 12) CHF_LAUNCH(16, 12); else CHF_EW(16); break;
    case 32: CHF_EW(32); break;
    case 64: CHF_EW(64); break;
    default: return fail(B200ODE_ERR_UNSUPPORTED, "chain: unsupported channel count %d", ch->g.C);
  }
#undef CHF_EW
#undef CHF_LAUNCH
  LAUNCH_CHECK("chain_f16_kernel");
  return 0;
}

extern "C" int b200ode_chain_fwd(b200ode_chain_t* ch, const float* x0, void* acts, uint8_t* relu_masks, float* y_final, int N,
                                 int H, int W, float h, int n_steps, void* stream) {
  if (!ch || !x0) return fail(B200ODE_ERR_INVALID, "chain/x0 is NULL");
  if (!ch->packed) return fail(B200ODE_ERR_NOT_PACKED, "b200ode_chain_pack must run before compute calls");
  if (!acts && !y_final) return fail(B200ODE_ERR_INVALID, "need acts or y_final");
  if (n_steps < 1 || N < 0 || H < 1 || W < 1) return fail(B200ODE_ERR_INVALID, "bad shape");
  if (N == 0) return 0;
  if (ch->mode == B200ODE_PREC_FAST_F16) {
    if (!y_final) return fail(B200ODE_ERR_INVALID, "FAST_F16 chains: y_final (fp32 output of the last step) is required; acts holds the fp16 INPUTS of the steps");
    ChainF16Plan plan;
    if (!plan_chain_f16(ch->g.C, H, W, &plan))
      return fail(B200ODE_ERR_UNSUPPORTED, "chain: a %dx%dx%d image does not fit shared memory (use the per-layer entry points)", H, W, ch->g.C);
    ChainF16Params& p = plan.p;
    p.N = N; p.L = n_steps; p.Lw = ch->L; p.h = h; p.gamma = ch->g.gamma;
    p.x0 = x0; p.acts = (__half*)acts; p.masks = relu_masks; p.y_final = y_final; p.bias = ch->bias; p.trace = g_trace;
    CUtensorMap mw;
    if (int rc = make_chain_w16_map(&mw, ch)) return rc;
    return launch_chain_f16<0>(ch, plan, mw, chain_f16_grid(p, ch->g.C, N), (cudaStream_t)stream);
  }
  ChainPlan plan;
  const bool strict = ch->mode == B200ODE_PREC_STRICT;
  if (!plan_chain(ch->g.C, H, W, 0, &plan, strict))
    return fail(B200ODE_ERR_UNSUPPORTED, "chain: a %dx%dx%d image does not fit shared memory (use the per-layer entry points)", H, W, ch->g.C);
  ChainParams& p = plan.p;
  p.N = N; p.L = n_steps; p.Lw = ch->L; p.h = h; p.gamma = ch->g.gamma;
  p.acts = (float*)acts; p.masks = relu_masks; p.y_final = y_final; p.bias = ch->bias; p.trace = g_trace;
  const int C = ch->g.C, rowb = C * 4 >= 128 ? 128 : C * 4;
  CUtensorMap mx, mw;
  CUtensorMapSwizzle sw = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  if (int rc = make_act_map(&mx, x0, N, H, W, C, 4, rowb / 4, p.P, H + 2, 1, sw)) return rc;
  if (int rc = make_chain_w_map(&mw, ch, p.tw)) return rc;
  if (strict) return launch_chain<0, true>(ch, plan, mx, mw, chain_grid(p, C, N), (cudaStream_t)stream);
  return launch_chain<0>(ch, plan, mx, mw, chain_grid(p, C, N), (cudaStream_t)stream);
}

static int chain_dgrad_impl(b200ode_chain_t* ch, const float* dy, const uint8_t* relu_masks, void* dz_all, float* dx,
                           int N, int H, int W, float h, const float* dy_amax, void* stream);
extern "C" int b200ode_chain_dgrad(b200ode_chain_t* ch, const float* dy, const uint8_t* relu_masks, void* dz_all, float* dx,
                                   int N, int H, int W, float h, void* stream) {
  return chain_dgrad_impl(ch, dy, relu_masks, dz_all, dx, N, H, W, h, nullptr, stream);
}
extern "C" int b200ode_chain_dgrad_amax(b200ode_chain_t* ch, const float* dy, const uint8_t* relu_masks, void* dz_all, float* dx,
                                        int N, int H, int W, float h, const float* dy_amax, void* stream) {
  if (!dy_amax) return fail(B200ODE_ERR_INVALID, "dy_amax is NULL");
  return chain_dgrad_impl(ch, dy, relu_masks, dz_all, dx, N, H, W, h, dy_amax, stream);
}
static int chain_dgrad_impl(b200ode_chain_t* ch, const float* dy, const uint8_t* relu_masks, void* dz_all, float* dx,
                            int N, int H, int W, float h, const float* dy_amax, void* stream) {
  if (!ch || !dy || !relu_masks || !dz_all || !dx) return fail(B200ODE_ERR_INVALID, "NULL argument");
  if (!ch->packed) return fail(B200ODE_ERR_NOT_PACKED, "b200ode_chain_pack must run before compute calls");
  if (N == 0) return 0;
  if (ch->mode == B200ODE_PREC_FAST_F16) {
    ChainF16Plan plan;
    if (!plan_chain_f16(ch->g.C, H, W, &plan))
      return fail(B200ODE_ERR_UNSUPPORTED, "chain: a %dx%dx%d image does not fit shared memory (use the per-layer entry points)", H, W, ch->g.C);
    cudaStream_t st = (cudaStream_t)stream;
    // scale of the fp16 backward strips: max|dy| -> device scalar (read by the chain kernel and by the gradient fold)
    if (dy_amax) {      // the kernel that produced dy already left max|dy| there (b200ode_*_amax entries): two launches less
      ch->amax_cur = dy_amax;
    } else {
      const long long n4 = (long long)N * H * W * ch->g.C / 4;
      CUDA_TRY(cudaMemsetAsync(ch->amax, 0, sizeof(float), st));
      const int sms = g_num_sms > 0 ? g_num_sms : 148;
      const long long want = (n4 + 255) / 256;
      amax_abs_kernel<<<(unsigned)(want < 2LL * sms ? want : 2LL * sms), 256, 0, st>>>((const float4*)dy, n4, (unsigned int*)ch->amax);
      LAUNCH_CHECK("amax_abs_kernel");
      ch->amax_cur = ch->amax;
    }
    ch->amax_h = h;
    ChainF16Params& p = plan.p;
    p.N = N; p.L = ch->L; p.Lw = ch->L; p.h = h; p.gamma = ch->g.gamma;
    p.bias = ch->bias; p.dy = dy; p.masks_r = relu_masks; p.dz_all = (__half*)dz_all; p.dx = dx; p.amax = ch->amax_cur; p.trace = g_trace;
    CUtensorMap mw;
    if (int rc = make_chain_w16_map(&mw, ch)) return rc;
    return launch_chain_f16<1>(ch, plan, mw, chain_f16_grid(p, ch->g.C, N), st);
  }
  ChainPlan plan;
  const bool strict = ch->mode == B200ODE_PREC_STRICT;
  if (!plan_chain(ch->g.C, H, W, 1, &plan, strict))
    return fail(B200ODE_ERR_UNSUPPORTED, "chain: a %dx%dx%d image does not fit shared memory (use the per-layer entry points)", H, W, ch->g.C);
  ChainParams& p = plan.p;
  p.N = N; p.L = ch->L; p.Lw = ch->L; p.h = h; p.gamma = ch->g.gamma;
  p.bias = ch->bias; p.dy = dy; p.masks_r = relu_masks; p.dz_all = (float*)dz_all; p.dx = dx; p.trace = g_trace;
  CUtensorMap mw;
  if (int rc = make_chain_w_map(&mw, ch, p.tw)) return rc;
  if (strict) return launch_chain<1, true>(ch, plan, mw, mw, chain_grid(p, ch->g.C, N), (cudaStream_t)stream);
  return launch_chain<1>(ch, plan, mw, mw, chain_grid(p, ch->g.C, N), (cudaStream_t)stream);
}

extern "C" int b200ode_chain_wgrad(b200ode_chain_t* ch, const float* x0, const void* acts, const void* dz_all,
                                   float* grad_params, int64_t grad_layer_stride, int N, int H, int W, void* stream) {
  if (!ch || !dz_all || !grad_params) return fail(B200ODE_ERR_INVALID, "NULL argument");
  if (N == 0) return 0;
  if (ch->mode == B200ODE_PREC_FAST_F16) {
    // acts[l] = fp16 input of layer l (written by b200ode_chain_fwd), dz_all[l] = fp16(S*dZ_l) (b200ode_chain_dgrad)
    if (!acts) return fail(B200ODE_ERR_INVALID, "acts is NULL");
    const __half* a16 = (const __half*)acts;
    return run_wgrad_tc(MODE_F16, ch->g, a16, a16 + (size_t)N * H * W * ch->g.C, dz_all, ch->L, N, H, W, nullptr, nullptr, grad_params,
                        grad_layer_stride, 0, (cudaStream_t)stream, WsArg{ch->ws, ch->ws_bytes, nullptr}, 0, ch->amax_cur ? ch->amax_cur : ch->amax, ch->amax_h);
  }
  if (!x0) return fail(B200ODE_ERR_INVALID, "x0 is NULL");
  if (ch->L > 1 && !acts) return fail(B200ODE_ERR_INVALID, "acts is NULL");
  return run_wgrad_tc(ch->mode == B200ODE_PREC_STRICT ? MODE_STRICT : MODE_TF32, ch->g, x0, acts, dz_all, ch->L, N, H, W, nullptr, nullptr,
                      grad_params, grad_layer_stride, 0, (cudaStream_t)stream, WsArg{ch->ws, ch->ws_bytes, nullptr});
}


// ------------------------------------------------------------------------------------------------
// stem / transition / head layers (kernels_glue.cuh)
// ------------------------------------------------------------------------------------------------
static int reduce_rows(const float* ws, int R, long long stride, long long n, float* out, cudaStream_t st);

enum { STEM_WGRAD_ROWS = 8 };
static size_t stem_wgrad_ws_bytes(int N, int H, int Cin, int Cout) {
  const int bands = (H + STEM_WGRAD_ROWS - 1) / STEM_WGRAD_ROWS;
  return (size_t)N * bands * (9 * (size_t)Cin * Cout + Cout) * sizeof(float);
}
static size_t head_ws_bytes(int N, int C, int K) { return (size_t)N * ((size_t)C * K + K + 1) * sizeof(float); }

static GlueConv glue_geom(int N, int H, int W, int Cin, int Cout, int sh, int sw) {
  GlueConv g;
  g.N = N; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.sh = sh; g.sw = sw;
  g.Ho = (H + sh - 1) / sh; g.Wo = (W + sw - 1) / sw;
  const int ph = (g.Ho - 1) * sh + 3 - H, pw = (g.Wo - 1) * sw + 3 - W;
  g.pt = (ph > 0 ? ph : 0) / 2; g.pl = (pw > 0 ? pw : 0) / 2;   // TF SAME: pad_before = total // 2
  return g;
}

extern "C" int b200ode_stem_fwd(const void* images, int images_are_u8, float subtract_mean, float divide_by_stddev, int normalize,
                                const float* kernel_hwio, const float* bias, float* out, int N, int H, int W, int Cin, int Cout,
                                void* stream) {
  if (!images || !kernel_hwio || !bias || !out) return fail(B200ODE_ERR_INVALID, "NULL argument");
  if (Cout % 4) return fail(B200ODE_ERR_UNSUPPORTED, "stem: filters must be a multiple of 4 (got %d)", Cout);
  if (int rc = device_check()) return rc;
  if (N == 0) return 0;
  const GlueConv g = glue_geom(N, H, W, Cin, Cout, 1, 1);
  const size_t psm = ((size_t)9 * Cin * Cout + Cout + 256) * sizeof(float);
  if ((Cout == 8 || Cout == 16 || Cout == 32) && psm <= 48 * 1024 && aligned16(kernel_hwio, out)) {
    // one thread per pixel, all filters: weights + the uint8 normalisation table in shared memory
    const long long npix = (long long)N * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cout == 8) stem_fwd_pixel_kernel<8><<<blocks_for(npix, 256), 256, psm, st>>>(g, images, images_are_u8, subtract_mean, divide_by_stddev, normalize, kernel_hwio, bias, out);
    else if (Cout == 16) stem_fwd_pixel_kernel<16><<<blocks_for(npix, 256), 256, psm, st>>>(g, images, images_are_u8, subtract_mean, divide_by_stddev, normalize, kernel_hwio, bias, out);
    else stem_fwd_pixel_kernel<32><<<blocks_for(npix, 256), 256, psm, st>>>(g, images, images_are_u8, subtract_mean, divide_by_stddev, normalize, kernel_hwio, bias, out);
    LAUNCH_CHECK("stem_fwd_pixel_kernel");
    return 0;
  }
  const long long total = (long long)N * H * W * (Cout / 4);
  stem_fwd_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(g, images, images_are_u8, subtract_mean, divide_by_stddev,
                                                                           normalize, kernel_hwio, bias, out);
  LAUNCH_CHECK("stem_fwd_kernel");
  return 0;
}

extern "C" int b200ode_stem_wgrad(const void* images, int images_are_u8, float subtract_mean, float divide_by_stddev, int normalize,
                                  const float* out, const float* dout, float* dparams, int N, int H, int W, int Cin, int Cout,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  if (!images || !out || !dout || !dparams) return fail(B200ODE_ERR_INVALID, "NULL argument");
  if (N == 0) return 0;
  const GlueConv g = glue_geom(N, H, W, Cin, Cout, 1, 1);
  const int rows = STEM_WGRAD_ROWS, bands = (H + rows - 1) / rows;
  const long long nout = 9LL * Cin * Cout + Cout;
  WsLease lease;
  if (int rc = lease_ws(workspace, workspace_bytes, stem_wgrad_ws_bytes(N, H, Cin, Cout), (cudaStream_t)stream, &lease)) return rc;
  float* ws = (float*)lease.ptr;
  const size_t smem = ((size_t)rows * W * Cout + (size_t)(rows + 2) * (W + 2) * Cin) * sizeof(float);
  if (smem > 48 * 1024) return fail(B200ODE_ERR_UNSUPPORTED, "stem_wgrad: image too wide (W=%d, filters=%d)", W, Cout);
  stem_wgrad_partial<<<dim3(N, bands), 256, smem, (cudaStream_t)stream>>>(g, images, images_are_u8, subtract_mean, divide_by_stddev,
                                                                         normalize, out, dout, ws, rows);
  LAUNCH_CHECK("stem_wgrad_partial");
  return reduce_rows(ws, N * bands, nout, nout, dparams, (cudaStream_t)stream);
}

static int transition_check(int Cin, int Cout) {
  if (Cin % 4 || Cout % 8) return fail(B200ODE_ERR_UNSUPPORTED, "transition: Cin %% 4 and Cout %% 8 must be 0 (got %d -> %d)", Cin, Cout);
  if (Cout > 256 || 256 % Cout || Cin % (256 / Cout) || Cin / (256 / Cout) > 16)
    return fail(B200ODE_ERR_UNSUPPORTED, "transition: unsupported channel pair %d -> %d", Cin, Cout);
  return 0;
}

// reduction of per-block partials: 32 / 8 row lanes per output when there are many rows (whole warps only: the
// combine is a warp shuffle, so n * LANES must fill the last warp)
static int reduce_rows(const float* ws, int R, long long stride, long long n, float* out, cudaStream_t st) {
  if (R >= 64 && n <= 2048) {   // few outputs: 32 row lanes (a warp's lanes read 32 different rows: 4 useful bytes per sector, only worth it when the launch is latency-bound)
    reduce_rows_wide_kernel<32><<<blocks_for(n * 32, 256), 256, 0, st>>>(ws, R, stride, n, out);
    LAUNCH_CHECK("reduce_rows_wide_kernel");
  } else if (R >= 32 && (n * 8) % 32 == 0) {
    reduce_rows_wide_kernel<8><<<blocks_for(n * 8, 256), 256, 0, st>>>(ws, R, stride, n, out);
    LAUNCH_CHECK("reduce_rows_wide_kernel");
  } else {
    reduce_rows_kernel<<<blocks_for(n, 128), 128, 0, st>>>(ws, R, stride, n, out);
    LAUNCH_CHECK("reduce_rows_kernel");
  }
  return 0;
}

#define GLUE_SMEM_LAUNCH(kern, grid, block, smem, st, ...)                                                       \
  do {                                                                                                           \
    static bool attr_set = false;                                                                                \
    if (!attr_set) {                                                                                             \
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));             \
      attr_set = true;                                                                                           \
    }                                                                                                            \
    kern<<<grid, block, smem, st>>>(__VA_ARGS__);                                                                \
  } while (0)


// Tensor-core transitions (kernels_glue_mma.cuh): the reference's stride-2 transitions 16 -> 32 and 32 -> 64 channels.
// B200ODE_GLUE_SIMT=1 keeps the CUDA-core kernels (A/B measurements); other shapes always use them.
// fast = one tf32 MMA on round-to-nearest operands instead of the 3xTF32 split (the *_fast entry points / B200ODE_TR_FAST=1)
static bool transition_fast_env() {
  static const bool on = getenv("B200ODE_TR_FAST") && atoi(getenv("B200ODE_TR_FAST")) != 0;
  return on;
}
static bool transition_mma_ok(int Cin, int Cout, int sh, int sw) {
  static const bool simt = getenv("B200ODE_GLUE_SIMT") && atoi(getenv("B200ODE_GLUE_SIMT")) != 0;
  return !simt && sh == 2 && sw == 2 && ((Cin == 16 && Cout == 32) || (Cin == 32 && Cout == 64));
}
#define GLUE_MMA_LAUNCH(kern, grid, block, smem, st, ...)                                                        \
  do {                                                                                                           \
    static bool attr_set = false;                                                                                \
    if (!attr_set) {                                                                                             \
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));             \
      attr_set = true;                                                                                           \
    }                                                                                                            \
    kern<<<grid, block, smem, st>>>(__VA_ARGS__);                                                                \
  } while (0)

// returns 0 = launched, > 0 = error, -1 = shape does not fit (caller falls back to the CUDA-core kernel)
template <int CIN, int COUT, int NSPLIT, bool FAST = false>
static int transition_fwd_mma(const GlueConv& g, const float* x, const float* Wm, const float* bm, const float* Ws, const float* bs,
                              float* out, uint8_t* mask, cudaStream_t st) {
  int orows = (128 / NSPLIT) / g.Wo;            // 8 warps x 16 positions / NSPLIT channel slices per block
  if (orows < 1) orows = 1;
  if (orows > g.Ho) orows = g.Ho;
  while (orows > 1 && TrFwdMma<CIN, COUT>::smem_bytes(orows, g.W) > 160 * 1024) orows >>= 1;
  const size_t smem = TrFwdMma<CIN, COUT>::smem_bytes(orows, g.W);
  if (smem > 160 * 1024) return -1;
  const dim3 grid(g.N, (g.Ho + orows - 1) / orows);
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(transition_fwd_mma_kernel<CIN, COUT, NSPLIT, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  CUDA_TRY(launch_pdl(transition_fwd_mma_kernel<CIN, COUT, NSPLIT, FAST>, grid, dim3(256), smem, st, g, x, Wm, bm, Ws, bs, out, mask, orows));
  LAUNCH_CHECK("transition_fwd_mma_kernel");
  return 0;
}
template <int CIN, int COUT, int NSPLIT, bool FAST = false>
static int transition_dgrad_mma(const GlueConv& g, const float* dout, const uint8_t* mask, const float* Wm, const float* Ws, float* dx,
                                unsigned int* dx_amax, cudaStream_t st) {
  const int CH = (g.H - 1 + g.pt) / 2 + 1, CW = (g.W - 1 + g.pl) / 2 + 1;
  int crows = (128 / NSPLIT) / CW;
  if (crows < 1) crows = 1;
  if (crows > CH) crows = CH;
  while (crows > 1 && TrDgradMma<CIN, COUT>::smem_bytes(crows, g.Wo) > 160 * 1024) crows >>= 1;
  const size_t smem = TrDgradMma<CIN, COUT>::smem_bytes(crows, g.Wo);
  if (smem > 160 * 1024) return -1;
  const dim3 grid(g.N, (CH + crows - 1) / crows);
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(transition_dgrad_mma_kernel<CIN, COUT, NSPLIT, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  CUDA_TRY(launch_pdl(transition_dgrad_mma_kernel<CIN, COUT, NSPLIT, FAST>, grid, dim3(256), smem, st, g, dout, mask, Wm, Ws, dx, crows, dx_amax));
  LAUNCH_CHECK("transition_dgrad_mma_kernel");
  return 0;
}
// partial rows the tensor-core weight gradient writes (one per block)
static int transition_wgrad_mma_rows(int N) { const int ipb = (N + 147) / 148; return (N + ipb - 1) / ipb; }
template <int CIN, int COUT, bool FAST = false>
static int transition_wgrad_mma(const GlueConv& g, const float* x, const float* dout, const uint8_t* mask, float* part, cudaStream_t st) {
  int orows = g.Ho;
  while (orows > 1 && TrWgradMma<CIN, COUT>::smem_bytes(orows, g.W, g.Wo) > 200 * 1024) orows = (orows + 1) >> 1;
  const size_t smem = TrWgradMma<CIN, COUT>::smem_bytes(orows, g.W, g.Wo);
  if (smem > 200 * 1024) return -1;
  const int ipb = (g.N + 147) / 148;
  constexpr int threads = 32 * TrWgradMma<CIN, COUT>::NWARP;
  GLUE_MMA_LAUNCH((transition_wgrad_mma_kernel<CIN, COUT, FAST>), dim3((g.N + ipb - 1) / ipb), threads, smem, st, g, x, dout, mask, part, orows, ipb);
  LAUNCH_CHECK("transition_wgrad_mma_kernel");
  return 0;
}

static int transition_fwd_impl(const float* x, const float* main_kernel, const float* main_bias, const float* short_kernel,
                               const float* short_bias, float* out, uint8_t* relu_mask, int N, int H, int W, int Cin,
                               int Cout, int stride_h, int stride_w, bool fast, void* stream);
extern "C" int b200ode_transition_fwd(const float* x, const float* main_kernel, const float* main_bias, const float* short_kernel,
                                      const float* short_bias, float* out, uint8_t* relu_mask, int N, int H, int W, int Cin,
                                      int Cout, int stride_h, int stride_w, void* stream) {
  return transition_fwd_impl(x, main_kernel, main_bias, short_kernel, short_bias, out, relu_mask, N, H, W, Cin, Cout, stride_h, stride_w,
                             transition_fast_env(), stream);
}
extern "C" int b200ode_transition_fwd_fast(const float* x, const float* main_kernel, const float* main_bias, const float* short_kernel,
                                           const float* short_bias, float* out, uint8_t* relu_mask, int N, int H, int W, int Cin,
                                           int Cout, int stride_h, int stride_w, void* stream) {
  return transition_fwd_impl(x, main_kernel, main_bias, short_kernel, short_bias, out, relu_mask, N, H, W, Cin, Cout, stride_h, stride_w,
                             true, stream);
}
static int transition_fwd_impl(const float* x, const float* main_kernel, const float* main_bias, const float* short_kernel,
                               const float* short_bias, float* out, uint8_t* relu_mask, int N, int H, int W, int Cin,
                               int Cout, int stride_h, int stride_w, bool fast, void* stream) {
  if (!x || !main_kernel || !main_bias || !short_kernel || !short_bias || !out || !relu_mask) return fail(B200ODE_ERR_INVALID, "NULL argument");
  if (int rc = transition_check(Cin, Cout)) return rc;
  if (int rc = device_check()) return rc;
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const GlueConv g = glue_geom(N, H, W, Cin, Cout, stride_h, stride_w);
  if (transition_mma_ok(Cin, Cout, stride_h, stride_w)) {
    const int rc = fast ? (Cin == 16 ? transition_fwd_mma<16, 32, 1, true>(g, x, main_kernel, main_bias, short_kernel, short_bias, out, relu_mask, st)
                                     : transition_fwd_mma<32, 64, 2, true>(g, x, main_kernel, main_bias, short_kernel, short_bias, out, relu_mask, st))
                        : (Cin == 16 ? transition_fwd_mma<16, 32, 1>(g, x, main_kernel, main_bias, short_kernel, short_bias, out, relu_mask, st)
                                     : transition_fwd_mma<32, 64, 2>(g, x, main_kernel, main_bias, short_kernel, short_bias, out, relu_mask, st));
    if (rc >= 0) return rc;
  }
  // bands of 32*PT output pixels (PT per lane); each warp of a block owns one 8-channel slice.  Small bands = many
  // blocks: the kernel is latency-bound (ncu: 11 % warps active, IPC 0.4 with 128-pixel bands on 256 blocks).
  constexpr int COT = 8;
  static const int pt_env = getenv("B200ODE_TR_FWD_PT") ? atoi(getenv("B200ODE_TR_FWD_PT")) : 1;
  const int PT = pt_env == 4 ? 4 : pt_env == 2 ? 2 : 1;
  int orows = (32 * PT + g.Wo - 1) / g.Wo;
  if (orows > g.Ho) orows = g.Ho;
  int G = Cout / COT < 8 ? Cout / COT : 8;
  while ((Cout / COT) % G) --G;
  size_t smem = 0;
  for (; orows >= 1; orows >>= 1) {
    smem = ((size_t)((orows - 1) * stride_h + 3) * W * (Cin + 4) + (size_t)G * 10 * Cin * COT) * sizeof(float);
    if (smem <= 160 * 1024) break;
  }
  if (orows < 1) return fail(B200ODE_ERR_UNSUPPORTED, "transition_fwd: rows too wide (W=%d, Cin=%d)", W, Cin);
  const int bands = (g.Ho + orows - 1) / orows;
  const dim3 grid(N, bands, Cout / COT / G);
  const bool s2 = stride_h == 2 && stride_w == 2;     // the reference's transitions (compile-time strides)
  if (PT == 4) GLUE_SMEM_LAUNCH((transition_fwd_kernel<COT, 4>), grid, 32 * G, smem, st, g, x, main_kernel, main_bias, short_kernel, short_bias, out, relu_mask, orows);
  else if (PT == 2) GLUE_SMEM_LAUNCH((transition_fwd_kernel<COT, 2>), grid, 32 * G, smem, st, g, x, main_kernel, main_bias, short_kernel, short_bias, out, relu_mask, orows);
  else if (s2) GLUE_SMEM_LAUNCH((transition_fwd_kernel<COT, 1, 2>), grid, 32 * G, smem, st, g, x, main_kernel, main_bias, short_kernel, short_bias, out, relu_mask, orows);
  else GLUE_SMEM_LAUNCH((transition_fwd_kernel<COT, 1>), grid, 32 * G, smem, st, g, x, main_kernel, main_bias, short_kernel, short_bias, out, relu_mask, orows);
  LAUNCH_CHECK("transition_fwd_kernel");
  return 0;
}

static int transition_dgrad_impl(const float* dout, const uint8_t* relu_mask, const float* main_kernel, const float* short_kernel, float* dx,
                                 int N, int H, int W, int Cin, int Cout, int stride_h, int stride_w, float* dx_amax, bool fast, void* stream);
extern "C" int b200ode_transition_dgrad_fast(const float* dout, const uint8_t* relu_mask, const float* main_kernel,
                                             const float* short_kernel, float* dx, int N, int H, int W, int Cin, int Cout, int stride_h,
                                             int stride_w, float* dx_amax, void* stream) {      /* dx_amax may be NULL */
  return transition_dgrad_impl(dout, relu_mask, main_kernel, short_kernel, dx, N, H, W, Cin, Cout, stride_h, stride_w, dx_amax, true, stream);
}
extern "C" int b200ode_transition_dgrad(const float* dout, const uint8_t* relu_mask, const float* main_kernel,
                                        const float* short_kernel, float* dx, int N, int H, int W, int Cin, int Cout, int stride_h,
                                        int stride_w, void* stream) {
  return transition_dgrad_impl(dout, relu_mask, main_kernel, short_kernel, dx, N, H, W, Cin, Cout, stride_h, stride_w, nullptr, transition_fast_env(), stream);
}
extern "C" int b200ode_transition_dgrad_amax(const float* dout, const uint8_t* relu_mask, const float* main_kernel,
                                             const float* short_kernel, float* dx, int N, int H, int W, int Cin, int Cout, int stride_h,
                                             int stride_w, float* dx_amax, void* stream) {
  if (!dx_amax) return fail(B200ODE_ERR_INVALID, "dx_amax is NULL");
  return transition_dgrad_impl(dout, relu_mask, main_kernel, short_kernel, dx, N, H, W, Cin, Cout, stride_h, stride_w, dx_amax, transition_fast_env(), stream);
}
// dx_amax != NULL: *dx_amax = max(*dx_amax, max|dx|) (the tensor-core kernel does it in its epilogue; other shapes by one more launch)
static int amax_after(const float* v, long long n, float* amax, cudaStream_t st) {
  if (n % 4) return fail(B200ODE_ERR_UNSUPPORTED, "amax: element count must be a multiple of 4");
  const int sms = g_num_sms > 0 ? g_num_sms : 148;
  const long long n4 = n / 4, want = (n4 + 255) / 256;
  amax_abs_kernel<<<(unsigned)(want < 2LL * sms ? want : 2LL * sms), 256, 0, st>>>((const float4*)v, n4, (unsigned int*)amax);
  LAUNCH_CHECK("amax_abs_kernel");
  return 0;
}
static int transition_dgrad_impl(const float* dout, const uint8_t* relu_mask, const float* main_kernel, const float* short_kernel, float* dx,
                                 int N, int H, int W, int Cin, int Cout, int stride_h, int stride_w, float* dx_amax, bool fast, void* stream) {
  if (!dout || !relu_mask || !main_kernel || !short_kernel || !dx) return fail(B200ODE_ERR_INVALID, "NULL argument");
  if (int rc = transition_check(Cin, Cout)) return rc;
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const GlueConv g = glue_geom(N, H, W, Cin, Cout, stride_h, stride_w);
  if (transition_mma_ok(Cin, Cout, stride_h, stride_w)) {
    const int rc = fast ? (Cin == 16 ? transition_dgrad_mma<16, 32, 1, true>(g, dout, relu_mask, main_kernel, short_kernel, dx, (unsigned int*)dx_amax, st)
                                     : transition_dgrad_mma<32, 64, 2, true>(g, dout, relu_mask, main_kernel, short_kernel, dx, (unsigned int*)dx_amax, st))
                        : (Cin == 16 ? transition_dgrad_mma<16, 32, 1>(g, dout, relu_mask, main_kernel, short_kernel, dx, (unsigned int*)dx_amax, st)
                                     : transition_dgrad_mma<32, 64, 2>(g, dout, relu_mask, main_kernel, short_kernel, dx, (unsigned int*)dx_amax, st));
    if (rc >= 0) return rc;
  }
  // band height: the staged output rows (dout + masked copy) and the weight slice must fit shared memory
  constexpr int CIT = 8;
  if (Cin % CIT) return fail(B200ODE_ERR_UNSUPPORTED, "transition_dgrad: Cin must be a multiple of %d (got %d)", CIT, Cin);
  static const int rows_env = getenv("B200ODE_TR_DGRAD_ROWS") ? atoi(getenv("B200ODE_TR_DGRAD_ROWS")) : 4;
  static const int pt_env = getenv("B200ODE_TR_DGRAD_PT") ? atoi(getenv("B200ODE_TR_DGRAD_PT")) : 1;
  const int PT = pt_env == 4 ? 4 : pt_env == 2 ? 2 : 1;
  int rows = H < rows_env ? H : rows_env;
  size_t smem = 0;
  for (; rows >= 1; rows >>= 1) {
    const int nor = (rows - 1 + 2) / stride_h + 2;
    smem = ((size_t)2 * nor * g.Wo * (Cout + 4) + (size_t)10 * CIT * Cout) * sizeof(float);
    if (smem <= 160 * 1024) break;
  }
  if (rows < 1) return fail(B200ODE_ERR_UNSUPPORTED, "transition_dgrad: rows too wide (Wo=%d, Cout=%d)", g.Wo, Cout);
  const int bands = (H + rows - 1) / rows;
  const dim3 grid(N, bands, Cin / CIT);
  if (PT == 4) GLUE_SMEM_LAUNCH((transition_dgrad_kernel<CIT, 4>), grid, 128, smem, st, g, dout, relu_mask, main_kernel, short_kernel, dx, rows);
  else if (PT == 2) GLUE_SMEM_LAUNCH((transition_dgrad_kernel<CIT, 2>), grid, 128, smem, st, g, dout, relu_mask, main_kernel, short_kernel, dx, rows);
  // compile-time strides: measured 39 -> 34 us for 16 -> 32 channels at 32x32, but 53 -> 59 us for 32 -> 64 at 16x16 (the fully
  // unrolled tap loop is larger than what the second shape amortises): on for narrow inputs only
  else if (stride_h == 2 && stride_w == 2 && Cin <= 16) GLUE_SMEM_LAUNCH((transition_dgrad_kernel<CIT, 1, 2>), grid, 128, smem, st, g, dout, relu_mask, main_kernel, short_kernel, dx, rows);
  else GLUE_SMEM_LAUNCH((transition_dgrad_kernel<CIT, 1>), grid, 128, smem, st, g, dout, relu_mask, main_kernel, short_kernel, dx, rows);
  LAUNCH_CHECK("transition_dgrad_kernel");
  if (dx_amax) return amax_after(dx, (long long)N * H * W * Cin, dx_amax, st);
  return 0;
}

struct TransWgradPlan { int tpg, orows, bands; size_t smem, ws_bytes; long long nout; };
static int plan_transition_wgrad(const GlueConv& g, TransWgradPlan* pl) {
  const int Cin = g.Cin, Cout = g.Cout;
  pl->nout = 9LL * Cin * Cout + Cout + (long long)Cin * Cout + Cout;
  const int ntile = (Cin / 4) * (Cout / 4);
  if (ntile > 256 || 256 % ntile) return fail(B200ODE_ERR_UNSUPPORTED, "transition_wgrad: unsupported channel pair %d -> %d", Cin, Cout);
  const int ngroups = 256 / ntile > 10 ? 10 : 256 / ntile;
  pl->tpg = (10 + ngroups - 1) / ngroups;
  // band of output rows per block: dout + masked copy + the input rows they touch
  static const int wrows_env = getenv("B200ODE_TR_WGRAD_ROWS") ? atoi(getenv("B200ODE_TR_WGRAD_ROWS")) : 4;
  int orows = g.Ho < wrows_env ? g.Ho : wrows_env;
  size_t smem = 0;
  for (; orows >= 1; orows >>= 1) {
    const int nir = (orows - 1) * g.sh + 3;
    smem = ((size_t)2 * orows * g.Wo * Cout + (size_t)nir * g.W * Cin) * sizeof(float);
    if (smem <= 100 * 1024) break;
  }
  if (orows < 1) return fail(B200ODE_ERR_UNSUPPORTED, "transition_wgrad: rows too wide (W=%d)", g.W);
  pl->orows = orows; pl->smem = smem;
  pl->bands = (g.Ho + orows - 1) / orows;
  pl->ws_bytes = (size_t)g.N * pl->bands * pl->nout * sizeof(float);
  return 0;
}

static int transition_wgrad_impl(const float* x, const float* dout, const uint8_t* relu_mask, float* dparams, int N, int H,
                                 int W, int Cin, int Cout, int stride_h, int stride_w, void* workspace,
                                 size_t workspace_bytes, bool fast, void* stream);
extern "C" int b200ode_transition_wgrad(const float* x, const float* dout, const uint8_t* relu_mask, float* dparams, int N, int H,
                                        int W, int Cin, int Cout, int stride_h, int stride_w, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  return transition_wgrad_impl(x, dout, relu_mask, dparams, N, H, W, Cin, Cout, stride_h, stride_w, workspace, workspace_bytes,
                               transition_fast_env(), stream);
}
extern "C" int b200ode_transition_wgrad_fast(const float* x, const float* dout, const uint8_t* relu_mask, float* dparams, int N, int H,
                                             int W, int Cin, int Cout, int stride_h, int stride_w, void* workspace,
                                             size_t workspace_bytes, void* stream) {
  return transition_wgrad_impl(x, dout, relu_mask, dparams, N, H, W, Cin, Cout, stride_h, stride_w, workspace, workspace_bytes, true, stream);
}
static int transition_wgrad_impl(const float* x, const float* dout, const uint8_t* relu_mask, float* dparams, int N, int H,
                                 int W, int Cin, int Cout, int stride_h, int stride_w, void* workspace,
                                 size_t workspace_bytes, bool fast, void* stream) {
  if (!x || !dout || !relu_mask || !dparams) return fail(B200ODE_ERR_INVALID, "NULL argument");
  if (int rc = transition_check(Cin, Cout)) return rc;
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const GlueConv g = glue_geom(N, H, W, Cin, Cout, stride_h, stride_w);
  TransWgradPlan pl;
  if (int rc = plan_transition_wgrad(g, &pl)) return rc;
  const int tpg = pl.tpg, orows = pl.orows, bands = pl.bands;
  const size_t smem = pl.smem;
  const long long nout = pl.nout;
  WsLease lease;
  if (int rc = lease_ws(workspace, workspace_bytes, pl.ws_bytes, st, &lease)) return rc;
  float* ws = (float*)lease.ptr;
  if (transition_mma_ok(Cin, Cout, stride_h, stride_w)) {      // rows <= N <= N * bands: the SIMT plan's workspace covers it
    const int rc = fast ? (Cin == 16 ? transition_wgrad_mma<16, 32, true>(g, x, dout, relu_mask, ws, st) : transition_wgrad_mma<32, 64, true>(g, x, dout, relu_mask, ws, st))
                        : (Cin == 16 ? transition_wgrad_mma<16, 32>(g, x, dout, relu_mask, ws, st) : transition_wgrad_mma<32, 64>(g, x, dout, relu_mask, ws, st));
    if (rc > 0) return rc;
    if (rc == 0) return reduce_rows(ws, transition_wgrad_mma_rows(N), nout, nout, dparams, st);
  }
  switch (tpg) {
    case 1: GLUE_SMEM_LAUNCH(transition_wgrad_partial<1>, dim3(N, bands), 256, smem, st, g, x, dout, relu_mask, ws, orows); break;
    case 2: GLUE_SMEM_LAUNCH(transition_wgrad_partial<2>, dim3(N, bands), 256, smem, st, g, x, dout, relu_mask, ws, orows); break;
    case 3: case 4: case 5: GLUE_SMEM_LAUNCH(transition_wgrad_partial<5>, dim3(N, bands), 256, smem, st, g, x, dout, relu_mask, ws, orows); break;
    default: GLUE_SMEM_LAUNCH(transition_wgrad_partial<10>, dim3(N, bands), 256, smem, st, g, x, dout, relu_mask, ws, orows); break;
  }
  LAUNCH_CHECK("transition_wgrad_partial");
  return reduce_rows(ws, N * bands, nout, nout, dparams, st);
}

static int head_impl(const float* x, const float* fc_kernel, const float* fc_bias, const float* onehot, float eps, float* probs, float* loss,
                     float* dx, float* dparams, int N, int HW, int C, int K, void* workspace, size_t workspace_bytes, float* dx_amax,
                     void* stream);
extern "C" int b200ode_head_fwd_bwd(const float* x, const float* fc_kernel, const float* fc_bias, const float* onehot, float eps,
                                    float* probs, float* loss, float* dx, float* dparams, int N, int HW, int C, int K,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  return head_impl(x, fc_kernel, fc_bias, onehot, eps, probs, loss, dx, dparams, N, HW, C, K, workspace, workspace_bytes, nullptr, stream);
}
extern "C" int b200ode_head_fwd_bwd_amax(const float* x, const float* fc_kernel, const float* fc_bias, const float* onehot, float eps,
                                         float* probs, float* loss, float* dx, float* dparams, int N, int HW, int C, int K,
                                         void* workspace, size_t workspace_bytes, float* dx_amax, void* stream) {
  if (!dx || !dx_amax) return fail(B200ODE_ERR_INVALID, "dx / dx_amax is NULL");
  return head_impl(x, fc_kernel, fc_bias, onehot, eps, probs, loss, dx, dparams, N, HW, C, K, workspace, workspace_bytes, dx_amax, stream);
}
static int head_impl(const float* x, const float* fc_kernel, const float* fc_bias, const float* onehot, float eps, float* probs, float* loss,
                     float* dx, float* dparams, int N, int HW, int C, int K, void* workspace, size_t workspace_bytes, float* dx_amax,
                     void* stream) {
  if (!x || !fc_kernel || !fc_bias || !onehot || !loss) return fail(B200ODE_ERR_INVALID, "NULL argument");
  if (K > 32 || K > C || C > 1024 || (C % 32)) return fail(B200ODE_ERR_UNSUPPORTED, "head: need classes <= 32 <= channels (multiple of 32, <= 1024)");
  if (int rc = device_check()) return rc;
  if (N == 0) return 0;
  const long long nout = (long long)C * K + K + 1;
  WsLease lease;
  if (int rc = lease_ws(workspace, workspace_bytes, head_ws_bytes(N, C, K), (cudaStream_t)stream, &lease)) return rc;
  float* ws = (float*)lease.ptr;
  CUDA_TRY(launch_pdl(head_kernel, dim3(N), dim3(C), (C + 33) * sizeof(float), (cudaStream_t)stream, x, HW, C, K, fc_kernel, fc_bias, onehot, eps,
                      N, probs, dx, ws, (unsigned int*)dx_amax));
  LAUNCH_CHECK("head_kernel");
  if (dparams && N >= 64 && nout <= 2048) {     // parameter gradients and the loss in ONE reduction launch (32 row lanes per output)
    reduce_rows_wide_kernel<32><<<blocks_for(nout * 32, 256), 256, 0, (cudaStream_t)stream>>>(ws, N, nout, nout, dparams, loss);
    LAUNCH_CHECK("reduce_rows_wide_kernel");
    return 0;
  }
  if (dparams)
    if (int rc = reduce_rows(ws, N, nout, nout - 1, dparams, (cudaStream_t)stream)) return rc;
  return reduce_rows(ws + (nout - 1), N, nout, 1, loss, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// workspace queries / binding (the caller owns the workspace)
// ------------------------------------------------------------------------------------------------
extern "C" int b200ode_layer_workspace_bytes(const b200ode_layer_t* L, int N, int H, int W, size_t* bytes_out) {
  if (!L || !bytes_out) return fail(B200ODE_ERR_INVALID, "layer/bytes_out is NULL");
  if (N < 0 || H < 1 || W < 1) return fail(B200ODE_ERR_INVALID, "bad shape N=%d H=%d W=%d", N, H, W);
  *bytes_out = 0;
  if (N == 0) return 0;
  if (L->mode_eff != B200ODE_PREC_SIMT_FP32 && L->g.k == 3) {   // tensor path: only the weight gradient needs scratch (split-K partials)
    return run_wgrad_tc(L->mode_eff, L->g, nullptr, nullptr, nullptr, 1, N, H, W, nullptr, nullptr, nullptr, 0, 0, nullptr,
                        WsArg{nullptr, 0, bytes_out});
  }
  const ConvGeom g = conv_geom(L, N, H, W);
  const long long npix = (long long)N * g.Ho * g.Wo;
  long long parts = npix / 64;
  parts = parts < 1 ? 1 : parts > 128 ? 128 : parts;
  const size_t fwd = (size_t)npix * g.C * sizeof(float);                              // pre-activations of a fused forward
  const size_t wg = (size_t)parts * g.k * g.k * g.C * g.C * sizeof(float) + colsum_ws_bytes(g.C);   // weight-gradient partials + bias sums
  *bytes_out = fwd > wg ? fwd : wg;
  return 0;
}
extern "C" int b200ode_layer_set_workspace(b200ode_layer_t* L, void* workspace, size_t bytes) {
  if (!L) return fail(B200ODE_ERR_INVALID, "layer is NULL");
  if (workspace && ((uintptr_t)workspace & 255)) return fail(B200ODE_ERR_INVALID, "workspace must be 256-byte aligned");
  L->ws = workspace; L->ws_bytes = workspace ? bytes : 0;
  return 0;
}
extern "C" int b200ode_chain_workspace_bytes(const b200ode_chain_t* ch, int N, int H, int W, size_t* bytes_out) {
  if (!ch || !bytes_out) return fail(B200ODE_ERR_INVALID, "chain/bytes_out is NULL");
  if (N < 0 || H < 1 || W < 1) return fail(B200ODE_ERR_INVALID, "bad shape N=%d H=%d W=%d", N, H, W);
  *bytes_out = 0;
  if (N == 0) return 0;
  return run_wgrad_tc(ch->mode == B200ODE_PREC_FAST_F16 ? MODE_F16 : ch->mode == B200ODE_PREC_STRICT ? MODE_STRICT : MODE_TF32, ch->g,
                      nullptr, nullptr, nullptr, ch->L, N, H, W,
                      nullptr, nullptr, nullptr, 0, 0, nullptr, WsArg{nullptr, 0, bytes_out});
}
extern "C" int b200ode_chain_set_workspace(b200ode_chain_t* ch, void* workspace, size_t bytes) {
  if (!ch) return fail(B200ODE_ERR_INVALID, "chain is NULL");
  if (workspace && ((uintptr_t)workspace & 255)) return fail(B200ODE_ERR_INVALID, "workspace must be 256-byte aligned");
  ch->ws = workspace; ch->ws_bytes = workspace ? bytes : 0;
  return 0;
}
extern "C" int b200ode_glue_workspace_bytes(int op, int N, int H, int W, int Cin, int Cout, int stride_h, int stride_w,
                                            size_t* bytes_out) {
  if (!bytes_out) return fail(B200ODE_ERR_INVALID, "bytes_out is NULL");
  if (N < 0 || H < 1 || W < 1 || Cin < 1 || Cout < 1) return fail(B200ODE_ERR_INVALID, "bad shape");
  *bytes_out = 0;
  if (N == 0) return 0;
  switch (op) {
    case B200ODE_GLUE_STEM_WGRAD: *bytes_out = stem_wgrad_ws_bytes(N, H, Cin, Cout); return 0;
    case B200ODE_GLUE_TRANSITION_WGRAD: {
      if (int rc = transition_check(Cin, Cout)) return rc;
      TransWgradPlan pl;
      if (int rc = plan_transition_wgrad(glue_geom(N, H, W, Cin, Cout, stride_h, stride_w), &pl)) return rc;
      *bytes_out = pl.ws_bytes;
      return 0;
    }
    case B200ODE_GLUE_HEAD: *bytes_out = head_ws_bytes(N, Cin, Cout); return 0;
    default: return fail(B200ODE_ERR_INVALID, "unknown glue op %d", op);
  }
}
