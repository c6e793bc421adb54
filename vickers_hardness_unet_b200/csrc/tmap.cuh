// Host-side TMA tensor-map construction (driver entry point fetched at run time: no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace ub {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

inline CUtensorMapSwizzle swizzle_for_bytes(int inner_bytes) {
    if (inner_bytes >= 128) return CU_TENSOR_MAP_SWIZZLE_128B;
    if (inner_bytes == 64) return CU_TENSOR_MAP_SWIZZLE_64B;
    if (inner_bytes == 32) return CU_TENSOR_MAP_SWIZZLE_32B;
    return CU_TENSOR_MAP_SWIZZLE_NONE;
}

// Generic bf16 tensor map of rank 2..5.  dims[0] is the contiguous dimension; strides_bytes[i] is the byte stride of
// dims[i+1] (rank-1 entries); box/estride per dimension.  Returns an error string ("" on success).
inline std::string make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                                  const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estride,
                                  CUtensorMapSwizzle swz, CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return "cuTensorMapEncodeTiled entry point unavailable";
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) {
        d[i] = dims[i];
        b[i] = box[i];
        e[i] = estride[i];
    }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = enc(out, dt, rank, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[512];
        snprintf(buf, sizeof buf,
                 "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] strides [%llu %llu %llu] box "
                 "[%u %u %u %u] es [%u %u %u %u] base %p",
                 int(r), rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
                 (unsigned long long)(rank > 2 ? d[2] : 0), (unsigned long long)(rank > 3 ? d[3] : 0),
                 (unsigned long long)s[0], (unsigned long long)(rank > 2 ? s[1] : 0),
                 (unsigned long long)(rank > 3 ? s[2] : 0), b[0], rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0,
                 rank > 3 ? b[3] : 0, e[0], rank > 1 ? e[1] : 0, rank > 2 ? e[2] : 0, rank > 3 ? e[3] : 0, base);
        return buf;
    }
    return "";
}

}  // namespace ub
