#include "common.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace bp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct DevInfo {
  int sms = 0, smem = 0, major = 0, minor = 0;
  bool ok = false;
};
static DevInfo& devinfo() {
  static DevInfo d[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) {
    static DevInfo bad;
    return bad;
  }
  if (!d[dev].ok) {
    cudaDeviceGetAttribute(&d[dev].sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d[dev].smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&d[dev].major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&d[dev].minor, cudaDevAttrComputeCapabilityMinor, dev);
    d[dev].ok = d[dev].sms > 0;
  }
  return d[dev];
}
int sm_count() { return devinfo().sms; }
int max_smem_optin() {
#ifdef BPB_ROLE_PROFILE
  // the role-profiling build keeps a small static shared array
  return devinfo().smem > 1024 ? devinfo().smem - 1024 : devinfo().smem;
#else
  return devinfo().smem;
#endif
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("BPB_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

static unsigned* g_dbg_host = nullptr;
static unsigned* g_dbg_dev = nullptr;
unsigned* debug_buffer() {
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    if (cudaHostAlloc(&h, 2048 * sizeof(unsigned), cudaHostAllocMapped) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    memset(h, 0, 2048 * sizeof(unsigned));
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    g_dbg_host = static_cast<unsigned*>(h);
    g_dbg_dev = static_cast<unsigned*>(d);
  });
  return g_dbg_dev;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

bool make_tmap_3d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                  uint32_t box0, uint32_t box1) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return false;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides,
                  box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d %llu x %llu x %llu) failed with %d",
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (int)r);
    return false;
  }
  return true;
}

// [32 channels x 32 rows] boxes, 64-byte swizzle: the per-warp output tiles of the conv epilogue
bool make_tmap_3d_sw64(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return false;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides,
                  box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d sw64 %llu x %llu x %llu) failed with %d",
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (int)r);
    return false;
  }
  return true;
}

bool make_tmap_3d_strided(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                          uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0,
                          uint32_t box1) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return false;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides,
                  box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d strided %llu x %llu x %llu, strides %llu / %llu) failed "
              "with %d", (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
              (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, (int)r);
    return false;
  }
  return true;
}

bool make_tmap_2d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint32_t box0,
                  uint32_t box1) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return false;
  }
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {d0 * 2};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                  box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d %llu x %llu) failed with %d", (unsigned long long)d0,
              (unsigned long long)d1, (int)r);
    return false;
  }
  return true;
}

}  // namespace bp

extern "C" {

int bpb_version(void) { return 1; }

const char* bpb_last_error(void) { return bp::g_err; }

// sizeof of the argument structs, so that a binding can verify its own layout (tests/test_abi.py)
int bpb_abi_struct_size(int which) {
  switch (which) {
    case 0: return static_cast<int>(sizeof(bpb_conv3_args));
    case 1: return static_cast<int>(sizeof(bpb_wgrad3_args));
    case 2: return static_cast<int>(sizeof(bpb_input_conv_args));
    case 3: return static_cast<int>(sizeof(bpb_heads_dgrad_args));
    case 4: return static_cast<int>(sizeof(bpb_prep_table));
    case 5: return static_cast<int>(sizeof(bpb_peer_args));
    default: return -1;
  }
}

int bpb_debug_dump(unsigned* out, int n) {
  if (bp::g_dbg_host == nullptr) return 0;
  if (n > 2048) n = 2048;
  for (int i = 0; i < n; ++i) out[i] = bp::g_dbg_host[i];
  return n;
}

int bpb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  bp::DevInfo& d = bp::devinfo();
  if (!d.ok) {
    bp::set_error("no CUDA device available");
    return 1;
  }
  if (sm_count) *sm_count = d.sms;
  if (cc_major) *cc_major = d.major;
  if (cc_minor) *cc_minor = d.minor;
  return 0;
}

}  // extern "C"
