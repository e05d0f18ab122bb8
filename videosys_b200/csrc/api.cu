// vsb200 -- C-ABI plumbing: error text, device check, tensor-map encoding, PAB integer gate.
#include <string.h>

#include "vsb_host.h"

namespace vsb {
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
static EncodeTiledFn g_encode = nullptr;
static int g_sms = 0;

EncodeTiledFn encode_tiled() { return g_encode; }

int num_sms() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

static int load_encode() {
  if (g_encode) return VSB_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr)
    return fail(VSB_ERR_CUDA, "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return VSB_OK;
}

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const unsigned long long* dims,
                   const unsigned long long* strides_bytes, const unsigned* box, CUtensorMapSwizzle swz) {
  int rc = load_encode();
  if (rc) return rc;
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      if (gstr[i - 1] % 16) return fail(VSB_ERR_UNSUPPORTED, "tensor map stride %llu not a multiple of 16 bytes", (unsigned long long)gstr[i - 1]);
    }
  }
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VSB_ERR_CUDA, "cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
  return VSB_OK;
}
}  // namespace vsb

using namespace vsb;

extern "C" int vsb_version(void) { return 100; }
extern "C" const char* vsb_last_error(void) { return g_err; }
extern "C" unsigned long long vsb_launch_count(void) { return g_launches.load(); }

extern "C" int vsb_init(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(VSB_ERR_NO_DEVICE, "no CUDA device: vsb200 has no CPU fallback");
  }
  if (device < 0 || device >= n) return fail(VSB_ERR_INVALID, "device %d out of range (%d devices)", device, n);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(VSB_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(VSB_ERR_NO_DEVICE, "device %d is sm_%d%d; vsb200 kernels are sm_100a only", device, prop.major, prop.minor);
  if (cudaSetDevice(device) != cudaSuccess) return fail(VSB_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  g_sms = prop.multiProcessorCount;
  return load_encode();
}

namespace vsb {
extern int g_opt_gemm_2sm;
extern int g_opt_attn_variant;
extern int g_opt_attn_pingpong;
extern int g_opt_attn_poly;
}
extern "C" int vsb_set_option(const char* name, int value) {
  if (!name) return fail(VSB_ERR_INVALID, "set_option: null name");
  if (!strcmp(name, "gemm_2sm")) {
    g_opt_gemm_2sm = value;
    return VSB_OK;
  }
  if (!strcmp(name, "attn_poly_exp")) {
    g_opt_attn_poly = value;
    return VSB_OK;
  }
  if (!strcmp(name, "attn_pingpong")) {
    g_opt_attn_pingpong = value;
    return VSB_OK;
  }
  if (!strcmp(name, "attn_variant")) {
    g_opt_attn_variant = value;
    return VSB_OK;
  }
  return fail(VSB_ERR_INVALID, "set_option: unknown option '%s'", name);
}

// core/pab/pab_mgr.py:54-91: flag = on && t is not None && count % range != 0 && lo < t < hi; count = (count+1) % steps
extern "C" int vsb_pab_gate(int broadcast_on, int has_timestep, int timestep, int* count, int range, int lo, int hi,
                            int steps) {
  if (!count || steps <= 0) return fail(VSB_ERR_INVALID, "pab_gate: bad args");
  int flag = 0;
  if (broadcast_on && has_timestep) {
    if (range <= 0) return fail(VSB_ERR_INVALID, "pab_gate: range must be positive");
    flag = ((*count % range) != 0) && (lo < timestep) && (timestep < hi);
  }
  *count = (*count + 1) % steps;
  return flag;
}
