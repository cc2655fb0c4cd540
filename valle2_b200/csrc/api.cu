// Housekeeping entry points of the C ABI: version, error string, device query, TMA descriptor factory.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_set>

#include "common.cuh"

static thread_local char g_err[512] = "";

void vb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int vb_version(void) { return 100; }

extern "C" const char* vb_last_error_string(void) { return g_err; }

int vb_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

bool vb_pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VALLE_B200_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

void vb_prefer_max_carveout(const void* kern) {
    static std::mutex mu;
    static std::unordered_set<const void*> done;
    static int enabled = -1;
    std::lock_guard<std::mutex> lock(mu);
    if (enabled < 0) {
        const char* e = getenv("VALLE_B200_CARVEOUT");
        enabled = (e && e[0] == '1') ? 1 : 0;
    }
    if (!enabled || !done.insert(kern).second) return;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();     // a kernel that cannot take the hint keeps its default
}

extern "C" int vb_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* smem_optin_bytes) {
    int dev = 0, n = 0, maj = 0, min = 0, smem = 0;
    VB_CUDA(cudaGetDevice(&dev));
    VB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    VB_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    VB_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    VB_CUDA(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (sm_count) *sm_count = n;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    if (smem_optin_bytes) *smem_optin_bytes = smem;
    VB_REQUIRE(maj == 10, VB_ERR_UNSUPPORTED, "libvalle_b200 is built for sm_100a only; device is sm_%d%d", maj, min);
    return VB_OK;
}

// ---- TMA descriptors --------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// Tensor maps are pure functions of (base, rows, cols, ld, box): the decode / NAR / training paths ask for the same few
// hundred over and over (weights, fixed workspaces), so the encoded 128-byte descriptors are kept in a small direct-mapped
// cache per thread instead of calling into the driver on every GEMM launch (~600 encodes per NAR request otherwise).
namespace {
struct TmapKey {
    const void* base; int64_t rows, cols, ld; int box_rows, box_cols;
    bool operator==(const TmapKey& o) const {
        return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && box_cols == o.box_cols;
    }
};
struct TmapSlot { TmapKey key; CUtensorMap map; bool valid; };
constexpr int kTmapSlots = 1024;
thread_local TmapSlot g_tmap_cache[kTmapSlots];
inline size_t tmap_hash(const TmapKey& k) {
    uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= (static_cast<uint64_t>(k.rows) * 0xBF58476D1CE4E5B9ull) ^ (static_cast<uint64_t>(k.cols) << 21) ^ (static_cast<uint64_t>(k.ld) << 42);
    h ^= (static_cast<uint64_t>(k.box_rows) << 7) ^ static_cast<uint64_t>(k.box_cols);
    h ^= h >> 29;
    return static_cast<size_t>(h % kTmapSlots);
}
}  // namespace

int vb_make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                         int box_cols) {
    const TmapKey key{base, rows, cols, ld, box_rows, box_cols};
    TmapSlot& slot = g_tmap_cache[tmap_hash(key)];
    if (slot.valid && slot.key == key) {
        *map = slot.map;
        return VB_OK;
    }
    PFN_encodeTiled enc = get_encode();
    VB_REQUIRE(enc != nullptr, VB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    VB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0, VB_ERR_BAD_ARG,
               "TMA operand must be 16-byte aligned with a 16-byte multiple pitch (base=%p ld=%lld)", base, (long long)ld);
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VB_REQUIRE(r == CUDA_SUCCESS, VB_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed: CUresult %d (rows=%lld cols=%lld ld=%lld box=%dx%d)",
               (int)r, (long long)rows, (long long)cols, (long long)ld, box_rows, box_cols);
    slot.key = key;
    slot.map = *map;
    slot.valid = true;
    return VB_OK;
}

int vb_make_tmap_bf16_3d(CUtensorMap* map, const void* base, int64_t d0, int64_t d1, int64_t d2, int64_t s1, int64_t s2,
                         int box0, int box1) {
    PFN_encodeTiled enc = get_encode();
    VB_REQUIRE(enc != nullptr, VB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    VB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (s1 * 2) % 16 == 0 && (s2 * 2) % 16 == 0, VB_ERR_BAD_ARG,
               "TMA operand must be 16-byte aligned with 16-byte multiple strides");
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(s1) * 2, static_cast<cuuint64_t>(s2) * 2};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(box0), static_cast<cuuint32_t>(box1), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VB_REQUIRE(r == CUDA_SUCCESS, VB_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed: CUresult %d", (int)r);
    return VB_OK;
}

// 3D fp32 [d2][d1][d0] (d0 contiguous), strides in elements; box = [1][box1][box0], no swizzle: the destination of TMA stores
int vb_make_tmap_f32_3d(CUtensorMap* map, const void* base, int64_t d0, int64_t d1, int64_t d2, int64_t s1, int64_t s2, int box0,
                        int box1) {
    PFN_encodeTiled enc = get_encode();
    VB_REQUIRE(enc != nullptr, VB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    VB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (s1 * 4) % 16 == 0 && (s2 * 4) % 16 == 0, VB_ERR_BAD_ARG,
               "TMA operand must be 16-byte aligned with 16-byte multiple strides");
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(s1) * 4, static_cast<cuuint64_t>(s2) * 4};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(box0), static_cast<cuuint32_t>(box1), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VB_REQUIRE(r == CUDA_SUCCESS, VB_ERR_CUDA, "cuTensorMapEncodeTiled(f32 3d) failed: CUresult %d", (int)r);
    return VB_OK;
}
