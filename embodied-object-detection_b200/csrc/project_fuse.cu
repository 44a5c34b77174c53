// 1x1 projection + weighted fusion in ONE kernel on the 5th-generation tensor cores:
//
//     mem = Conv2d(C_mem -> C_out, 1x1, bias)(level.float())      timm.py:174   (fp32, eval, autocast off)
//     mem = mem * MAP_FEATURE_WEIGHT                               timm.py:177
//     out = res + mem | mem                                        timm.py:181-184
//
// `level` is what eod_read_pool wrote: (E, h, w, K) fp16 channels-last, i.e. an (M = E*h*w, K) K-major matrix whose
// elements are EXACT fp16 numbers (the reference rounds the pooled level to half before the conv, timm.py:168).  The
// fp32 weight is split once on the device into two fp16 terms,  W = W_hi + 2^-11 * W_lo  (W_hi = half(W),
// W_lo = half((W - W_hi) * 2^11); 22 mantissa bits), so that  x . W = x . W_hi + 2^-11 * (x . W_lo)  with every product
// exact in the fp32 accumulator: two kind::f16 UMMAs against the SAME staged A tile give fp32-GEMM accuracy
// (measured against an fp64 reference in the tests) at fp16 tensor-core rate, instead of a CUDA-core SGEMM + a
// separate elementwise pass.
//
// Tile: 128 rows (pixels) x 128 output channels; the UMMA is M=128, N=256 (columns 0..127 = W_hi block, 128..255 =
// W_lo block of the same output channels), K=16 per instruction, fp32 accumulators in 256 TMEM columns.
// One CTA (128 threads) per tile, two CTAs resident per SM (2 x 256 TMEM columns, 2 x 97 KB shared memory): while one
// CTA streams its epilogue (res in, out out: the HBM-bound part), the other runs its main loop.
//   warp 0 / lane 0 : TMA producer  - A box {64 K, 128 rows}, B box {64 K, 256 rows}, SWIZZLE_128B, 2-stage ring
//   warp 1 / lane 0 : MMA issuer    - 4 x tcgen05.mma per stage, tcgen05.commit releases the stage / publishes the tile
//   warps 0-3       : epilogue      - tcgen05.ld 32x32b (thread = one pixel row), + bias, * weight, + res, NCHW store
//                                     (for a fixed channel the 32 lanes of a warp touch 32 consecutive pixels = 128 B)
#include <cuda.h>

#include "eod_common.cuh"

namespace {

constexpr int PF_BM = 128;                    // pixels per tile
constexpr int PF_BN = 128;                    // output channels per tile
constexpr int PF_BK = 64;                     // fp16 elements per 128-byte swizzle row
constexpr int PF_STAGES = 2;
constexpr int PF_A_BYTES = PF_BM * PF_BK * 2;             // 16 KB
constexpr int PF_B_BYTES = 2 * PF_BN * PF_BK * 2;         // 32 KB (hi + lo rows)
constexpr int PF_STAGE_BYTES = PF_A_BYTES + PF_B_BYTES;
constexpr int PF_TMEM_COLS = 2 * PF_BN;                   // 256 fp32 columns
constexpr int PF_SMEM_BYTES = PF_STAGES * PF_STAGE_BYTES + 1024 /* alignment slack */ + 64 /* barriers + tmem ptr */;
constexpr float PF_LO_SCALE = 2048.f, PF_LO_INV = 1.f / 2048.f;

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc], fp16 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all tcgen05.mma issued so far by this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address >> 4 in
// bits [0,14), leading byte offset (unused for swizzled K-major; 1) in [16,30), stride byte offset = 1024 B (one 8-row
// swizzle atom) >> 4 in [32,46), version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (InstrDescriptor): D = F32 (1 << 4), A = B = F16 (0), both K-major, N >> 3 in [17,23), M >> 4 in [24,29)
constexpr uint32_t PF_IDESC = (1u << 4) | ((uint32_t)(2 * PF_BN >> 3) << 17) | ((uint32_t)(PF_BM >> 4) << 24);

// kSum: out = res + w * mem, else out = w * mem
template <bool kSum>
__global__ void __launch_bounds__(128) project_fuse_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                                                           const float *__restrict__ bias, const float *__restrict__ res, float *__restrict__ out,
                                                           float weight, int M, int hw, int N, int n_kblocks)
{
    extern __shared__ uint8_t pf_smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(pf_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + PF_STAGES * PF_STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + PF_STAGES, *tmem_full = bars + 2 * PF_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * PF_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_nblocks = N / PF_BN;
    const int mb = blockIdx.x / n_nblocks, nb = blockIdx.x % n_nblocks;      // the n-blocks of one m-block are neighbours: A comes from L2 the 2nd time

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < PF_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(tmem_full, 1);
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b) : "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, PF_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        for (int kb = 0; kb < n_kblocks; ++kb) {
            const int s = kb % PF_STAGES;
            mbar_wait(empty + s, ((kb / PF_STAGES) & 1) ^ 1);
            uint8_t *a = smem + s * PF_STAGE_BYTES, *b = a + PF_A_BYTES;
            mbar_expect_tx(full + s, PF_STAGE_BYTES);
            tma_load_2d(a, &tm_a, kb * PF_BK, mb * PF_BM, full + s);
            tma_load_2d(b, &tm_b, kb * PF_BK, nb * 2 * PF_BN, full + s);
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        for (int kb = 0; kb < n_kblocks; ++kb) {
            const int s = kb % PF_STAGES;
            mbar_wait(full + s, (kb / PF_STAGES) & 1);
            tc_fence_after();
            const uint32_t a = smem_u32(smem + s * PF_STAGE_BYTES), b = a + PF_A_BYTES;
#pragma unroll
            for (int k = 0; k < PF_BK / 16; ++k)       // 16 fp16 = 32 bytes along K inside the swizzle row
                umma_f16(tmem_base, umma_desc_sw128(a + k * 32), umma_desc_sw128(b + k * 32), PF_IDESC, (kb | k) != 0);
            umma_commit(empty + s);                     // the stage may be refilled once these MMAs have read it
        }
        umma_commit(tmem_full);                         // accumulators complete
    }
    __syncwarp();

    // ===== epilogue: all four warps; warp w owns TMEM lanes [32w, 32w + 32) =====
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int m = mb * PF_BM + warp * 32 + lane;
    const bool live = m < M;
    const int e = live ? m / hw : 0, pix = live ? m - e * hw : 0;
    const size_t base = ((size_t)e * N + (size_t)nb * PF_BN) * hw + pix;      // NCHW: + channel * hw
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float *bias_n = bias ? bias + nb * PF_BN : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < PF_BN; c0 += 16) {
        uint32_t hi[16], lo[16];
        float r[16];
        tmem_ld16(taddr + c0, hi);
        tmem_ld16(taddr + PF_BN + c0, lo);
        if (kSum) {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = live ? __ldg(res + base + (size_t)(c0 + j) * hw) : 0.f;
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float v = __fmaf_rn(__uint_as_float(lo[j]), PF_LO_INV, __uint_as_float(hi[j]));      // x.W_hi + 2^-11 x.W_lo: the GEMM result
            if (bias_n) v = __fadd_rn(v, __ldg(bias_n + c0 + j));                                 // conv bias
            v = __fmul_rn(v, weight);                                                             // timm.py:177
            if (kSum) v = __fadd_rn(r[j], v);                                                     // timm.py:182
            if (live) out[base + (size_t)(c0 + j) * hw] = v;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, PF_TMEM_COLS);
}

// W (N,K) f32 -> (2N,K) f16: per block of 128 output channels, 128 rows of W_hi then 128 rows of W_lo
__global__ void __launch_bounds__(256) split_weights_kernel(const float *__restrict__ w, int N, int K, __half *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * K) return;
    const int n = (int)(i / K), k = (int)(i - (int64_t)n * K);
    const float x = __ldg(w + i);
    const __half hi = __float2half_rn(x);
    const __half lo = __float2half_rn(__fmul_rn(__fsub_rn(x, __half2float(hi)), PF_LO_SCALE));
    const int64_t row_hi = (int64_t)(n / PF_BN) * (2 * PF_BN) + (n % PF_BN);
    out[row_hi * K + k] = hi;
    out[(row_hi + PF_BN) * K + k] = lo;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled pf_encode_fn()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// (rows, K) fp16 row-major -> 2-D map with a {64, box_rows} SWIZZLE_128B box; rows beyond `rows` read as zero
int pf_make_tmap(const void *ptr, int64_t rows, int K, int box_rows, CUtensorMap *tm)
{
    PFN_encodeTiled enc = pf_encode_fn();
    EOD_REQUIRE(enc, EOD_ERR_LAUNCH, "eod_project_fuse: cuTensorMapEncodeTiled entry point unavailable");
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {PF_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    EOD_REQUIRE(r == CUDA_SUCCESS, EOD_ERR_LAUNCH, "eod_project_fuse: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return EOD_OK;
}

}  // namespace

extern "C" int eod_project_split_weights(const float *weight, int N, int K, void *w_split, eod_stream_t stream)
{
    EOD_REQUIRE(weight && w_split, EOD_ERR_BADARG, "eod_project_split_weights: null pointer");
    EOD_REQUIRE(N > 0 && K > 0, EOD_ERR_BADARG, "eod_project_split_weights: bad sizes");
    EOD_REQUIRE(N % PF_BN == 0 && K % PF_BK == 0, EOD_ERR_UNSUPPORTED, "eod_project_split_weights: N %% 128 == 0 and K %% 64 == 0 required");
    const int64_t n = (int64_t)N * K;
    split_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(weight, N, K, reinterpret_cast<__half *>(w_split));
    return eod_check_launch("eod_project_split_weights");
}

extern "C" int eod_project_fuse(const void *level, const void *w_split, const float *bias, const float *res, float weight, int mode,
                                int n_episodes, int hw, int K, int N, float *out, eod_stream_t stream)
{
    EOD_REQUIRE(level && w_split && out, EOD_ERR_BADARG, "eod_project_fuse: null pointer");
    EOD_REQUIRE(mode == EOD_FUSE_SUM || mode == EOD_FUSE_MEM_ONLY, EOD_ERR_BADARG, "eod_project_fuse: mode must be sum or mem_only");
    EOD_REQUIRE(mode != EOD_FUSE_SUM || res, EOD_ERR_BADARG, "eod_project_fuse: res is required for sum");
    EOD_REQUIRE(n_episodes > 0 && hw > 0 && (int64_t)n_episodes * hw < (1ll << 31) - PF_BM, EOD_ERR_BADARG, "eod_project_fuse: bad sizes");
    EOD_REQUIRE(N > 0 && K > 0 && N % PF_BN == 0 && K % PF_BK == 0, EOD_ERR_UNSUPPORTED, "eod_project_fuse: N %% 128 == 0 and K %% 64 == 0 required");
    EOD_REQUIRE(eod_aligned16(level) && eod_aligned16(w_split), EOD_ERR_ALIGN, "eod_project_fuse: level and w_split must be 16-byte aligned");
    const int M = n_episodes * hw;
    CUtensorMap tm_a, tm_b;
    int rc = pf_make_tmap(level, M, K, PF_BM, &tm_a);
    if (rc) return rc;
    rc = pf_make_tmap(w_split, 2 * (int64_t)N, K, 2 * PF_BN, &tm_b);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(project_fuse_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM_BYTES);
        cudaFuncSetAttribute(project_fuse_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM_BYTES);
        attr_set = true;
    }
    const unsigned grid = (unsigned)((M + PF_BM - 1) / PF_BM) * (unsigned)(N / PF_BN);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == EOD_FUSE_SUM)
        project_fuse_kernel<true><<<grid, 128, PF_SMEM_BYTES, st>>>(tm_a, tm_b, bias, res, out, weight, M, hw, N, K / PF_BK);
    else
        project_fuse_kernel<false><<<grid, 128, PF_SMEM_BYTES, st>>>(tm_a, tm_b, bias, res, out, weight, M, hw, N, K / PF_BK);
    return eod_check_launch("eod_project_fuse");
}
