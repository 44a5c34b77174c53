// Thread-local cache of encoded CUtensorMaps keyed by every argument of the encode call.  A tensor map holds nothing but the
// base address and the geometry, so a cached map stays valid for as long as the caller passes the same pointer and shape -
// which is every frame for a resident grid / feature slab (ADVICE r1 / VERDICT r1: cuTensorMapEncodeTiled used to run on the
// host at every call of the single-episode path).
#pragma once
#include <cuda.h>
#include <string.h>

typedef CUresult (*PFN_eodEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                       const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct EodTmapKey {
    const void *ptr;
    cuuint64_t gdim[3], gstr[2];
    cuuint32_t box[3];
    int rank, dtype, promo, swizzle;
};

static inline CUresult eod_encode_tmap_cached(PFN_eodEncodeTiled enc, CUtensorMap *tm, CUtensorMapDataType dtype, cuuint32_t rank, const void *ptr,
                                             const cuuint64_t *gdim, const cuuint64_t *gstr, const cuuint32_t *box, CUtensorMapL2promotion promo,
                                             CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B)
{
    constexpr int N = 16;
    static thread_local EodTmapKey keys[N];
    static thread_local CUtensorMap maps[N];
    static thread_local int used = 0, next = 0;
    EodTmapKey k;
    memset(&k, 0, sizeof(k));
    k.ptr = ptr; k.rank = (int)rank; k.dtype = (int)dtype; k.promo = (int)promo; k.swizzle = (int)swizzle;
    for (cuuint32_t i = 0; i < rank; ++i) { k.gdim[i] = gdim[i]; k.box[i] = box[i]; }
    for (cuuint32_t i = 0; i + 1 < rank; ++i) k.gstr[i] = gstr[i];
    for (int i = 0; i < used; ++i)
        if (memcmp(&keys[i], &k, sizeof(k)) == 0) {
            *tm = maps[i];
            return CUDA_SUCCESS;
        }
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(tm, dtype, rank, const_cast<void *>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, promo,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return r;
    const int slot = used < N ? used++ : (next = (next + 1) % N);
    keys[slot] = k;
    maps[slot] = *tm;
    return r;
}
