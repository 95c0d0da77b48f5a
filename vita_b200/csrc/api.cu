// Library plumbing: error reporting, launch accounting, device query, TMA descriptor encoding.
#include <atomic>
#include <cstdlib>
#include <mutex>

#include "common.h"

namespace vita {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

void set_last_error(const std::string& msg) { g_last_error = msg; }

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return VITA_OK;
    set_last_error(std::string(what) + ": " + cudaGetErrorString(e));
    return VITA_ERR_CUDA;
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return check_cuda(cudaGetLastError(), what);
}

int num_sms() {
    static int cached = -1;
    if (cached < 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
            (void)cudaGetLastError();
            return 0;
        }
        cached = n;
    }
    return cached;
}

bool use_pdl() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("VITA_B200_PDL");
        cached = (e != nullptr && e[0] == '1') ? 1 : 0;   // measured: no gain (5.27 ms/token off vs 5.31 on) -> opt-in
    }
    return cached == 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    });
    return fn;
}

int make_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                         const uint32_t* box, bool swizzle128) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
        return VITA_ERR_CUDA;
    }
    cuuint64_t gdims[5];
    cuuint64_t gstrides[4];
    cuuint32_t gbox[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
        estr[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstrides[i] = strides[i];
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                          gdims, gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
        return VITA_ERR_CUDA;
    }
    return VITA_OK;
}

}  // namespace vita

extern "C" int vita_version(void) { return 100; }
extern "C" const char* vita_last_error(void) { return vita::g_last_error.c_str(); }
extern "C" int vita_num_sms(void) { return vita::num_sms(); }
extern "C" int64_t vita_launch_count(int reset) {
    const long long v = vita::g_launches.load();
    if (reset) vita::g_launches.store(0);
    return v;
}
