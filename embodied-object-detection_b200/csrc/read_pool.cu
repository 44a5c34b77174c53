// Memory read: normalise (sum/count where count>1) -> fp16 -> gather to the
// image plane -> avg-pool 4 -> three times (avg-pool 2 -> fp16), fused into one kernel.
//
// Work decomposition: one CTA per 32x32-pixel block of one episode's frame (= one L2 output pixel, four L1,
// sixteen L0).  Thread (q, g): q = 16x16 quadrant of the block, g = group of 4 consecutive channels, so a
// warp is 32 channel groups of the SAME quadrant and walks the SAME pixels: the cell id is warp-uniform,
// every table access is one coalesced 512 B (fp32) / 256 B (fp16) segment of a cell row, and consecutive
// pixels that hit the same cell reuse the value from registers.  The 1.2 MB int index plane is read once;
// table rows come from L1/L2 after the first touch (a frame sees ~5k distinct cells).
//
// Summation order == ATen CPU avg_pool2d: fp32, start from 0, row-major over the window, then / k^2;
// the fp16 roundings between levels (timm.py:168) are reproduced, so outputs are bit-identical.
#include "eod_common.cuh"

namespace {

template <typename IdxT>
__device__ __forceinline__ int load_cell(const IdxT *p) { return (int)__ldg(p); }

__device__ __forceinline__ float4 round_to_half(float4 v)
{
    v.x = __half2float(__float2half_rn(v.x));
    v.y = __half2float(__float2half_rn(v.y));
    v.z = __half2float(__float2half_rn(v.z));
    v.w = __half2float(__float2half_rn(v.w));
    return v;
}

// Split row fetch: the global loads of a batch of pixels are issued first (load_raw, load_n), their
// consumers (finish: normalise + fp16 rounding) run afterwards, so one memory latency is exposed per batch
// instead of one per cell change.
struct RawF32 { float4 v; float n; };
struct RawF16 { uint2 v; };

__device__ __forceinline__ RawF32 load_raw(const float *table, const float *counts, size_t cell, int C, int g)
{
    RawF32 r;
    r.v = __ldg(reinterpret_cast<const float4 *>(table + cell * C) + g);
    r.n = counts ? __ldg(counts + cell) : 0.f;
    return r;
}
__device__ __forceinline__ RawF16 load_raw(const __half *table, const float *, size_t cell, int C, int g)
{
    RawF16 r;
    r.v = __ldg(reinterpret_cast<const uint2 *>(table + cell * C) + g);
    return r;
}
__device__ __forceinline__ float4 finish(const RawF32 &r)
{
    float4 v = r.v;
    if (r.n > 1.0f) {            // custom_rcnn.py:774 (cells seen once or never are left as-is)
        v.x = __fdiv_rn(v.x, r.n); v.y = __fdiv_rn(v.y, r.n); v.z = __fdiv_rn(v.z, r.n); v.w = __fdiv_rn(v.w, r.n);
    }
    return round_to_half(v);     // custom_rcnn.py:1036
}
__device__ __forceinline__ float4 finish(const RawF16 &r)
{
    const __half2 a = *reinterpret_cast<const __half2 *>(&r.v.x), b = *reinterpret_cast<const __half2 *>(&r.v.y);
    const float2 fa = __half22float2(a), fb = __half22float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> struct RawOf;
template <> struct RawOf<float> { using type = RawF32; };
template <> struct RawOf<__half> { using type = RawF16; };

__device__ __forceinline__ void store_half4(__half *dst, float4 v)
{
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<uint32_t *>(&a);
    raw.y = *reinterpret_cast<uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(dst) = raw;
}

// packed fp32x2 arithmetic (sm_100 FADD2 / FMUL2): two IEEE round-to-nearest operations per instruction, same bits as the scalar ones
__device__ __forceinline__ float4 add4(float4 a, float4 b)
{
    const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)), hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 scale4(float4 a, float s)
{
    const float2 lo = __fmul2_rn(make_float2(a.x, a.y), make_float2(s, s)), hi = __fmul2_rn(make_float2(a.z, a.w), make_float2(s, s));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

template <int C, typename TableT, typename IdxT>
__global__ void __launch_bounds__(C, 1024 / C) read_pool_kernel(const TableT *__restrict__ table, const float *__restrict__ counts,
                                                      const IdxT *__restrict__ idx, int H, int W, int64_t n_cells,
                                                      __half *__restrict__ L0, __half *__restrict__ L1, __half *__restrict__ L2)
{
    constexpr int G = C / 4;                 // channel groups == threads per quadrant
    __shared__ int s_idx[32 * 32];
    __shared__ int s_wcell[64];              // per 4x4 window (8x8 of them): the cell id if all 16 pixels agree, else -1
    __shared__ float4 s_l1[4][G];

    const int e = blockIdx.z, by = blockIdx.y, bx = blockIdx.x;
    const int q = threadIdx.x / G, g = threadIdx.x % G;
    const int qy = q >> 1, qx = q & 1;

    const IdxT *idx_e = idx + (size_t)e * H * W;
    for (int i = threadIdx.x; i < 1024; i += C) {
        const int r = i >> 5, c = i & 31;
        s_idx[i] = load_cell(idx_e + (size_t)(by * 32 + r) * W + bx * 32 + c);
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        const int wy = threadIdx.x >> 3, wx = threadIdx.x & 7;
        const int4 r0 = *reinterpret_cast<const int4 *>(&s_idx[(wy * 4 + 0) * 32 + wx * 4]);
        const int4 r1 = *reinterpret_cast<const int4 *>(&s_idx[(wy * 4 + 1) * 32 + wx * 4]);
        const int4 r2 = *reinterpret_cast<const int4 *>(&s_idx[(wy * 4 + 2) * 32 + wx * 4]);
        const int4 r3 = *reinterpret_cast<const int4 *>(&s_idx[(wy * 4 + 3) * 32 + wx * 4]);
        const int c0 = r0.x;
        const bool u = (r0.y == c0) & (r0.z == c0) & (r0.w == c0) & (r1.x == c0) & (r1.y == c0) & (r1.z == c0) & (r1.w == c0) &
                       (r2.x == c0) & (r2.y == c0) & (r2.z == c0) & (r2.w == c0) & (r3.x == c0) & (r3.y == c0) & (r3.z == c0) & (r3.w == c0);
        s_wcell[threadIdx.x] = u ? c0 : -1;
    }
    __syncthreads();

    const TableT *table_e = table + (size_t)e * n_cells * C;
    const float *counts_e = counts ? counts + (size_t)e * n_cells : nullptr;
    const int h0 = H / 8, w0 = W / 8, h1 = H / 16, w1 = W / 16, h2 = H / 32, w2 = W / 32;

    int cur_cell = -1;
    float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 l1acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int l0 = 0; l0 < 4; ++l0) {                       // L0 pixels of the quadrant, row-major
        const int l0y = l0 >> 1, l0x = l0 & 1;
        // The four 4x4 windows of this L0 pixel.  A window whose 16 pixels hit ONE cell needs no additions: the
        // gathered value x is an fp16 number, so the sequential fp32 sum x+x+...+x is exact at every step
        // (k*x, k <= 16, has at most 15 significant bits) and avg_pool2d(4) returns x itself.
        const int wbase = (qy * 4 + l0y * 2) * 8 + qx * 4 + l0x * 2;
        const int w00 = s_wcell[wbase];
        float4 v0;
        if (w00 >= 0 && w00 == s_wcell[wbase + 1] && w00 == s_wcell[wbase + 8] && w00 == s_wcell[wbase + 9]) {
            // whole 8x8 block in one cell: pool(4), pool(2) and the fp16 rounding all return the gathered value
            if (w00 != cur_cell) cur = finish(load_raw(table_e, counts_e, (size_t)w00, C, g));
            cur_cell = w00;
            v0 = cur;
        } else {
            float4 s2 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
            for (int win = 0; win < 4; ++win) {            // 4x4 windows of the avg_pool2d(4) (timm.py:152)
                const int wc = s_wcell[wbase + (win >> 1) * 8 + (win & 1)];
                if (wc >= 0) {
                    if (wc != cur_cell) cur = finish(load_raw(table_e, counts_e, (size_t)wc, C, g));
                    cur_cell = wc;
                    s2 = add4(s2, cur);
                    continue;
                }
                const int row0 = qy * 16 + l0y * 8 + (win >> 1) * 4, col0 = qx * 16 + l0x * 8 + (win & 1) * 4;
                float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int half = 0; half < 2; ++half) {     // two rows (8 pixels) per batch of loads
                    const int4 ca = *reinterpret_cast<const int4 *>(&s_idx[(row0 + 2 * half) * 32 + col0]);
                    const int4 cb = *reinterpret_cast<const int4 *>(&s_idx[(row0 + 2 * half + 1) * 32 + col0]);
                    const int cc[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
                    typename RawOf<TableT>::type raw[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k)             // warp-uniform predicates; all loads in flight together
                        if (cc[k] != (k == 0 ? cur_cell : cc[k - 1])) raw[k] = load_raw(table_e, counts_e, (size_t)cc[k], C, g);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (cc[k] != (k == 0 ? cur_cell : cc[k - 1])) cur = finish(raw[k]);
                        s4 = add4(s4, cur);
                    }
                    cur_cell = cc[7];
                }
                s2 = add4(s2, scale4(s4, 0.0625f));        // / 16 (exact)
            }
            v0 = round_to_half(scale4(s2, 0.25f));         // avg_pool2d(2) -> half (timm.py:168, level 0)
        }
        const int y0 = by * 4 + qy * 2 + l0y, x0 = bx * 4 + qx * 2 + l0x;
        store_half4(L0 + (((size_t)e * h0 + y0) * w0 + x0) * C + 4 * g, v0);
        l1acc = add4(l1acc, v0);
    }
    const float4 v1 = round_to_half(scale4(l1acc, 0.25f));  // level 1
    store_half4(L1 + (((size_t)e * h1 + by * 2 + qy) * w1 + bx * 2 + qx) * C + 4 * g, v1);
    s_l1[q][g] = v1;
    __syncthreads();
    if (q == 0) {                                           // level 2: quadrants in row-major order
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) s = add4(s, s_l1[k][g]);
        store_half4(L2 + (((size_t)e * h2 + by) * w2 + bx) * C + 4 * g, scale4(s, 0.25f));
    }
}


// Full-grid normalise (create_implicit_memory, custom_rcnn.py:764-774) for callers that want the table
// itself (the reference API returns it); the read kernel above never needs it.  Streaming, float4.
template <bool HALF_OUT>
__global__ void __launch_bounds__(256) normalize_kernel(const float4 *__restrict__ sums, const float *__restrict__ counts, int64_t n_rows, int C4,
                                                        float4 *__restrict__ out32, uint2 *__restrict__ out16)
{
    const int64_t total = n_rows * C4, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t row = i / C4;
        float4 v = __ldg(sums + i);
        const float n = __ldg(counts + row);
        if (n > 1.0f) { v.x = __fdiv_rn(v.x, n); v.y = __fdiv_rn(v.y, n); v.z = __fdiv_rn(v.z, n); v.w = __fdiv_rn(v.w, n); }
        if (HALF_OUT) {
            __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
            uint2 raw; raw.x = *reinterpret_cast<uint32_t *>(&a); raw.y = *reinterpret_cast<uint32_t *>(&b);
            out16[i] = raw;
        } else out32[i] = v;
    }
}

template <int C>
int launch(const void *table, int mem_is_f16, const float *counts, const void *idx, int idx_is_i64, int E, int H, int W,
           int64_t n_cells, void *L0, void *L1, void *L2, cudaStream_t st)
{
    dim3 grid(W / 32, H / 32, E), block(C);
    __half *l0 = (__half *)L0, *l1 = (__half *)L1, *l2 = (__half *)L2;
    if (mem_is_f16) {
        if (idx_is_i64) read_pool_kernel<C, __half, int64_t><<<grid, block, 0, st>>>((const __half *)table, nullptr, (const int64_t *)idx, H, W, n_cells, l0, l1, l2);
        else read_pool_kernel<C, __half, int32_t><<<grid, block, 0, st>>>((const __half *)table, nullptr, (const int32_t *)idx, H, W, n_cells, l0, l1, l2);
    } else {
        if (idx_is_i64) read_pool_kernel<C, float, int64_t><<<grid, block, 0, st>>>((const float *)table, counts, (const int64_t *)idx, H, W, n_cells, l0, l1, l2);
        else read_pool_kernel<C, float, int32_t><<<grid, block, 0, st>>>((const float *)table, counts, (const int32_t *)idx, H, W, n_cells, l0, l1, l2);
    }
    return eod_check_launch("eod_read_pool");
}

}  // namespace

extern "C" int eod_read_pool(const void *table, int mem_is_f16, const float *counts, const void *idx, int idx_is_i64,
                             int n_episodes, int C, int H, int W, int64_t n_cells, void *L0, void *L1, void *L2,
                             eod_stream_t stream)
{
    EOD_REQUIRE(table && idx && L0 && L1 && L2, EOD_ERR_BADARG, "eod_read_pool: null pointer");
    EOD_REQUIRE(n_episodes > 0 && n_episodes <= 65535 && n_cells > 0, EOD_ERR_BADARG, "eod_read_pool: bad sizes");
    EOD_REQUIRE(H > 0 && W > 0 && H % 32 == 0 && W % 32 == 0, EOD_ERR_BADARG, "eod_read_pool: H and W must be multiples of 32 (got %dx%d)", H, W);
    EOD_REQUIRE(H / 32 <= 65535, EOD_ERR_BADARG, "eod_read_pool: H too large");
    EOD_REQUIRE(eod_aligned16(table) && eod_aligned16(idx) && eod_aligned16(L0) && eod_aligned16(L1) && eod_aligned16(L2),
                EOD_ERR_ALIGN, "eod_read_pool: pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    switch (C) {
    case 128: return launch<128>(table, mem_is_f16, counts, idx, idx_is_i64, n_episodes, H, W, n_cells, L0, L1, L2, st);
    case 256: return launch<256>(table, mem_is_f16, counts, idx, idx_is_i64, n_episodes, H, W, n_cells, L0, L1, L2, st);
    case 512: return launch<512>(table, mem_is_f16, counts, idx, idx_is_i64, n_episodes, H, W, n_cells, L0, L1, L2, st);
    default:
        eod_set_error("eod_read_pool: C=%d not compiled in (128, 256, 512)", C);
        return EOD_ERR_UNSUPPORTED;
    }
}

extern "C" int eod_normalize_memory(const float *sums, const float *counts, int64_t n_rows, int C, void *out, int out_is_f16,
                                    eod_stream_t stream)
{
    EOD_REQUIRE(sums && counts && out, EOD_ERR_BADARG, "eod_normalize_memory: null pointer");
    EOD_REQUIRE(n_rows > 0 && C > 0 && C % 4 == 0, EOD_ERR_BADARG, "eod_normalize_memory: bad sizes (C %% 4 == 0)");
    EOD_REQUIRE(eod_aligned16(sums) && eod_aligned16(out), EOD_ERR_ALIGN, "eod_normalize_memory: pointers must be 16-byte aligned");
    const int64_t total = n_rows * (C / 4);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)eod_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (out_is_f16) normalize_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const float4 *)sums, counts, n_rows, C / 4, nullptr, (uint2 *)out);
    else normalize_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const float4 *)sums, counts, n_rows, C / 4, (float4 *)out, nullptr);
    return eod_check_launch("eod_normalize_memory");
}
