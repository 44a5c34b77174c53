// Shared device/host helpers for libeod_memory.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/eod_memory.h"

#define EOD_NUM_SMS_FALLBACK 148

void eod_set_error(const char *fmt, ...);
int eod_check_launch(const char *what);
int eod_num_sms();
int *eod_work_tickets();      // zeroed {next, done} pair for one dynamically scheduled launch (api.cu)

#define EOD_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            eod_set_error(__VA_ARGS__);   \
            return (code);                \
        }                                 \
    } while (0)

static inline bool eod_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: `done` is a per-kernel bitmask over device
// ordinals (one bit per device of the process), so a second GPU used from the same process gets the attribute too.
// Returns EOD_OK or EOD_ERR_LAUNCH (message set).  Racing threads at worst set the attribute twice.
template <typename F>
static inline int eod_ensure_dyn_smem(F func, int bytes, unsigned long long *done, const char *what)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(done, __ATOMIC_ACQUIRE) & bit) return EOD_OK;
    const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        eod_set_error("%s: cannot raise dynamic shared memory to %d bytes: %s", what, bytes, cudaGetErrorString(e));
        return EOD_ERR_LAUNCH;
    }
    __atomic_fetch_or(done, bit, __ATOMIC_RELEASE);
    return EOD_OK;
}

// ---- device helpers -------------------------------------------------------------------------------

// 128-bit fp32 reduction to global memory (sm_90+): one L2 atomic transaction for four floats.
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

__device__ __forceinline__ void red_add_f32(float *addr, float a)
{
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

__device__ __forceinline__ float4 ldg_stream_f4(const float *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- mbarrier / TMA (inline PTX) --------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 3-D tiled TMA load global -> shared, completion signalled on an mbarrier (SASS: UTMALDG).
__device__ __forceinline__ void tma_load_3d(void *dst, const void *tmap, int x, int y, int z, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}
// 1-D bulk copy global -> shared (SASS: UBLKCP); bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
