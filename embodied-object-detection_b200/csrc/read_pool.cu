// Memory read: normalise (sum/count where count>1) -> fp16 -> gather to the
// image plane -> avg-pool 4 -> three times (avg-pool 2 -> fp16): one kernel for levels 0 and 1 (the (480,640,C)
// image-plane tensors are never materialised) and a small second one that pools level 2 from level 1.
//
// Work decomposition: a warp owns one 16x16-pixel quadrant of one episode's frame (= four L0 pixels, one L1 pixel)
// and ALL channels (lane = C/32 consecutive channels), so the cell id is warp-uniform, every table access is one
// coalesced C*2-byte (fp16) / C*4-byte (fp32) cell row, and consecutive pixels that hit the same cell reuse the value
// from registers.  The 1.2 MB index plane is read once; table rows come from L1/L2 after the first touch (a frame sees a
// few hundred to a few thousand distinct cells).  The kernel is instruction-issue bound, not bandwidth bound: what it
// optimises is scalar bookkeeping per gathered row (see read_pool_kernel).
//
// Summation order == ATen CPU avg_pool2d: fp32, start from 0, row-major over the window, then / k^2;
// the fp16 roundings between levels (timm.py:168) are reproduced, so outputs are bit-identical.
#include "eod_common.cuh"

namespace {

// ---- V channels per lane (V = 4 or 8) as packed fp32x2 pairs -------------------------------------------------------
// FADD2 / FMUL2 (sm_100): two IEEE round-to-nearest operations per instruction, the same bits as the scalar ones.
template <int V>
struct Vec { float2 p[V / 2]; };

template <int V>
__device__ __forceinline__ Vec<V> vzero()
{
    Vec<V> r;
#pragma unroll
    for (int i = 0; i < V / 2; ++i) r.p[i] = make_float2(0.f, 0.f);
    return r;
}
template <int V>
__device__ __forceinline__ Vec<V> vadd(const Vec<V> &a, const Vec<V> &b)
{
    Vec<V> r;
#pragma unroll
    for (int i = 0; i < V / 2; ++i) r.p[i] = __fadd2_rn(a.p[i], b.p[i]);
    return r;
}
template <int V>
__device__ __forceinline__ Vec<V> vscale(const Vec<V> &a, float s)
{
    Vec<V> r;
#pragma unroll
    for (int i = 0; i < V / 2; ++i) r.p[i] = __fmul2_rn(a.p[i], make_float2(s, s));
    return r;
}
template <int V>
__device__ __forceinline__ Vec<V> vround_half(const Vec<V> &a)
{
    Vec<V> r;
#pragma unroll
    for (int i = 0; i < V / 2; ++i) r.p[i] = __half22float2(__float22half2_rn(a.p[i]));
    return r;
}
// V halves, 2*V bytes, one store
template <int V>
__device__ __forceinline__ void vstore_half(__half *dst, const Vec<V> &a)
{
    uint32_t w[V / 2];
#pragma unroll
    for (int i = 0; i < V / 2; ++i) {
        const __half2 h = __float22half2_rn(a.p[i]);
        w[i] = *reinterpret_cast<const uint32_t *>(&h);
    }
    if (V >= 8) {
#pragma unroll
        for (int i = 0; i < V / 8; ++i) reinterpret_cast<uint4 *>(dst)[i] = make_uint4(w[(4 * i) % (V / 2)], w[(4 * i + 1) % (V / 2)], w[(4 * i + 2) % (V / 2)], w[(4 * i + 3) % (V / 2)]);
    } else {
        *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[1]);
    }
}

// Split row fetch: the global load of a row is issued (load_raw) before its consumer (finish: normalise + fp16
// rounding) runs, so the next run's row is in flight under the current run's adds.
template <int V> struct RawF32 { float4 v[V / 4]; float n; };
template <int V> struct RawF16 { uint32_t w[V / 2]; };

template <int V>
__device__ __forceinline__ RawF32<V> load_raw(const float *table, const float *counts, size_t cell, int C, int g)
{
    RawF32<V> r;
#pragma unroll
    for (int i = 0; i < V / 4; ++i) r.v[i] = __ldg(reinterpret_cast<const float4 *>(table + cell * C + (size_t)g * V) + i);
    r.n = counts ? __ldg(counts + cell) : 0.f;
    return r;
}
template <int V>
__device__ __forceinline__ RawF16<V> load_raw(const __half *table, const float *, size_t cell, int C, int g)
{
    RawF16<V> r;
    if (V >= 8) {
#pragma unroll
        for (int i = 0; i < V / 8; ++i) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(table + cell * C + (size_t)g * V) + i);
            r.w[(4 * i) % (V / 2)] = q.x; r.w[(4 * i + 1) % (V / 2)] = q.y; r.w[(4 * i + 2) % (V / 2)] = q.z; r.w[(4 * i + 3) % (V / 2)] = q.w;
        }
    } else {
        const uint2 q = __ldg(reinterpret_cast<const uint2 *>(table + cell * C) + g);
        r.w[0] = q.x; r.w[1] = q.y;
    }
    return r;
}
template <int V>
__device__ __forceinline__ Vec<V> finish(const RawF32<V> &r)
{
    Vec<V> v;
#pragma unroll
    for (int i = 0; i < V / 4; ++i) {
        float4 x = r.v[i];
        if (r.n > 1.0f) {        // custom_rcnn.py:774 (cells seen once or never are left as-is)
            x.x = __fdiv_rn(x.x, r.n); x.y = __fdiv_rn(x.y, r.n); x.z = __fdiv_rn(x.z, r.n); x.w = __fdiv_rn(x.w, r.n);
        }
        v.p[2 * i] = make_float2(x.x, x.y);
        v.p[2 * i + 1] = make_float2(x.z, x.w);
    }
    return vround_half<V>(v);    // custom_rcnn.py:1036
}
template <int V>
__device__ __forceinline__ Vec<V> finish(const RawF16<V> &r)
{
    Vec<V> v;
#pragma unroll
    for (int i = 0; i < V / 2; ++i) v.p[i] = __half22float2(*reinterpret_cast<const __half2 *>(&r.w[i]));
    return v;
}
template <typename T, int V> struct RawOf;
template <int V> struct RawOf<float, V> { using type = RawF32<V>; };
template <int V> struct RawOf<__half, V> { using type = RawF16<V>; };

// Warp-autonomous read: a work item = one 16x16-pixel quadrant (4 L0 pixels, 1 L1 pixel) x all C channels, owned by ONE
// warp (lane = group of V = C/32 consecutive channels: the per-run bookkeeping below is scalar work that every lane
// repeats, so wider lanes amortise it over more channels - V = 4 -> 8 took C=256 from 0.40 to 0.27 ms, V = 8 -> 16 took
// C=512 from 0.51 to 0.40 ms at E=64).  The warp stages the quadrant's 256 cell ids in its private
// slice of shared memory, classifies its 16 windows itself (lanes 0-15) and walks them - no CTA barrier anywhere, and the
// ids of the NEXT item are already in flight (registers) while the current one is processed.  Level 2 needs four L1
// pixels of different quadrants; it is pooled from the stored L1 by pool_level2_kernel instead of through a CTA-wide
// exchange.  (The CTA-per-32x32-block version spent 35 % of its stall samples on the index staging latency and on two
// barriers per block.)
//
// Summation order == ATen CPU avg_pool2d: fp32, start from 0, row-major over the window, then / k^2; the fp16 roundings
// between levels (timm.py:168) are reproduced, so outputs are bit-identical.
constexpr int kReadWarps = 8;

struct QuadIdx { int v[8]; };

template <typename IdxT>
__device__ __forceinline__ QuadIdx load_quad_idx(const IdxT *idx_e, int W, int qy, int qx, unsigned lane)
{
    // lane -> row lane>>1 of the quadrant, columns (lane&1)*8 .. +7
    const IdxT *src = idx_e + (size_t)(qy * 16 + (lane >> 1)) * W + qx * 16 + (lane & 1) * 8;
    QuadIdx q;
    if (sizeof(IdxT) == 4) {
        const int4 a = __ldg(reinterpret_cast<const int4 *>(src)), b = __ldg(reinterpret_cast<const int4 *>(src) + 1);
        q.v[0] = a.x; q.v[1] = a.y; q.v[2] = a.z; q.v[3] = a.w; q.v[4] = b.x; q.v[5] = b.y; q.v[6] = b.z; q.v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(src) + k);
            q.v[2 * k] = (int)a.x; q.v[2 * k + 1] = (int)a.y;
        }
    }
    return q;
}

template <int C, typename TableT, typename IdxT>
__global__ void __launch_bounds__(kReadWarps * 32, (C >= 512 ? 2 : (C >= 256 ? 3 : 4))) read_pool_kernel(const TableT *__restrict__ table, const float *__restrict__ counts,
                                                                                        const IdxT *__restrict__ idx, int H, int W, int64_t n_cells, int E,
                                                                                        __half *__restrict__ L0, __half *__restrict__ L1)
{
    constexpr int V = C >= 512 ? 16 : (C >= 256 ? 8 : 4);      // channels per lane
    constexpr int NC = C / (32 * V);         // channel chunks per quadrant
    using Raw = typename RawOf<TableT, V>::type;
    __shared__ __align__(16) int s_idx[kReadWarps][16 * 16];
    __shared__ int s_wcell[kReadWarps][16];                    // per 4x4 window: the cell id if all 16 pixels agree, else -1
    __shared__ int s_run_cell[kReadWarps][16][16];             // mixed windows: runs of equal cell id in row-major (= summation) order ...
    __shared__ __align__(16) unsigned char s_run_len[kReadWarps][16][16];   // ... their lengths ...
    __shared__ int s_nruns[kReadWarps][16];                    // ... and how many there are

    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nqy = H / 16, nqx = W / 16;
    const int64_t n_items = (int64_t)E * nqy * nqx * NC;
    const int64_t stride = (int64_t)gridDim.x * kReadWarps;
    const int h0 = H / 8, w0 = W / 8, h1 = H / 16, w1 = W / 16;

    auto decode = [&](int64_t item, int &e, int &qy, int &qx, int &chunk) {
        chunk = (int)(item % NC);
        int64_t q = item / NC;
        qx = (int)(q % nqx); q /= nqx;
        qy = (int)(q % nqy);
        e = (int)(q / nqy);
    };

    int64_t item = (int64_t)blockIdx.x * kReadWarps + warp;
    if (item >= n_items) return;
    int e, qy, qx, chunk;
    decode(item, e, qy, qx, chunk);
    QuadIdx nxt = load_quad_idx<IdxT>(idx + (size_t)e * H * W, W, qy, qx, lane);

    for (; item < n_items; item += stride) {
        decode(item, e, qy, qx, chunk);
        // stage this item's ids, start the next item's loads
        {
            int4 *dst = reinterpret_cast<int4 *>(&s_idx[warp][(lane >> 1) * 16 + (lane & 1) * 8]);
            dst[0] = make_int4(nxt.v[0], nxt.v[1], nxt.v[2], nxt.v[3]);
            dst[1] = make_int4(nxt.v[4], nxt.v[5], nxt.v[6], nxt.v[7]);
        }
        if (item + stride < n_items) {
            int e2, qy2, qx2, c2;
            decode(item + stride, e2, qy2, qx2, c2);
            nxt = load_quad_idx<IdxT>(idx + (size_t)e2 * H * W, W, qy2, qx2, lane);
        }
        __syncwarp();
        if (lane < 16) {                                   // classify window (lane>>2, lane&3) of the quadrant
            const int wy = lane >> 2, wx = lane & 3;
            const int4 r0 = *reinterpret_cast<const int4 *>(&s_idx[warp][(wy * 4 + 0) * 16 + wx * 4]);
            const int4 r1 = *reinterpret_cast<const int4 *>(&s_idx[warp][(wy * 4 + 1) * 16 + wx * 4]);
            const int4 r2 = *reinterpret_cast<const int4 *>(&s_idx[warp][(wy * 4 + 2) * 16 + wx * 4]);
            const int4 r3 = *reinterpret_cast<const int4 *>(&s_idx[warp][(wy * 4 + 3) * 16 + wx * 4]);
            const int c0 = r0.x;
            const bool u = (r0.y == c0) & (r0.z == c0) & (r0.w == c0) & (r1.x == c0) & (r1.y == c0) & (r1.z == c0) & (r1.w == c0) &
                           (r2.x == c0) & (r2.y == c0) & (r2.z == c0) & (r2.w == c0) & (r3.x == c0) & (r3.y == c0) & (r3.z == c0) & (r3.w == c0);
            s_wcell[warp][lane] = u ? c0 : -1;
            if (!u) {
                // a mixed window is typically an edge between two or three cells: 2-8 runs instead of 16 pixels; the walk below
                // pays the fetch / fp16->fp32 conversion / bookkeeping per RUN and V/2 packed adds per pixel
                const int cells[16] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
                int prev = c0, len = 1, nr = 0;
#pragma unroll
                for (int k = 1; k < 16; ++k) {
                    if (cells[k] != prev) {
                        s_run_cell[warp][lane][nr] = prev;
                        s_run_len[warp][lane][nr] = (unsigned char)len;
                        ++nr;
                        prev = cells[k];
                        len = 1;
                    } else {
                        ++len;
                    }
                }
                s_run_cell[warp][lane][nr] = prev;
                s_run_len[warp][lane][nr] = (unsigned char)len;
                s_nruns[warp][lane] = nr + 1;
            }
        }
        __syncwarp();

        const int g = chunk * 32 + (int)lane;               // group of V channels
        const TableT *table_e = table + (size_t)e * n_cells * C;
        const float *counts_e = counts ? counts + (size_t)e * n_cells : nullptr;
        int cur_cell = -1;
        Vec<V> cur = vzero<V>();
        Vec<V> l1acc = vzero<V>();
#pragma unroll 1
        for (int l0 = 0; l0 < 4; ++l0) {                   // L0 pixels of the quadrant, row-major
            const int l0y = l0 >> 1, l0x = l0 & 1;
            // The four 4x4 windows of this L0 pixel.  A window whose 16 pixels hit ONE cell needs no additions: the
            // gathered value x is an fp16 number, so the sequential fp32 sum 0+x+x+...+x is exact at every step
            // (k*x, k <= 16, has at most 15 significant bits) and avg_pool2d(4) returns x itself - except that -0.0
            // becomes +0.0 (0 + -0 = +0), which the shortcuts reproduce by adding +0 (found by profiles/stress_read.py).
            const int wbase = (l0y * 2) * 4 + l0x * 2;
            const int w00 = s_wcell[warp][wbase];
            Vec<V> v0;
            if (w00 >= 0 && w00 == s_wcell[warp][wbase + 1] && w00 == s_wcell[warp][wbase + 4] && w00 == s_wcell[warp][wbase + 5]) {
                // whole 8x8 block in one cell: pool(4), pool(2) and the fp16 rounding all return the gathered value
                if (w00 != cur_cell) cur = finish<V>(load_raw<V>(table_e, counts_e, (size_t)w00, C, g));
                cur_cell = w00;
                v0 = vadd<V>(cur, vzero<V>());               // the reference's sums start from +0: a gathered -0.0 comes out as +0.0

            } else {
                Vec<V> s2 = vzero<V>();
#pragma unroll 1
                for (int win = 0; win < 4; ++win) {        // 4x4 windows of the avg_pool2d(4) (timm.py:152)
                    const int wi = wbase + (win >> 1) * 4 + (win & 1);
                    const int wc = s_wcell[warp][wi];
                    if (wc >= 0) {
                        if (wc != cur_cell) cur = finish<V>(load_raw<V>(table_e, counts_e, (size_t)wc, C, g));
                        cur_cell = wc;
                        s2 = vadd<V>(s2, cur);
                        continue;
                    }
                    const int nr = s_nruns[warp][wi];
                    Vec<V> s4 = vzero<V>();
                    Raw raw = load_raw<V>(table_e, counts_e, (size_t)s_run_cell[warp][wi][0], C, g);
#pragma unroll 1
                    for (int r = 0; r < nr; ++r) {
                        cur = finish<V>(raw);
                        if (r + 1 < nr) raw = load_raw<V>(table_e, counts_e, (size_t)s_run_cell[warp][wi][r + 1], C, g);   // next run's row in flight under the adds
                        int len = s_run_len[warp][wi][r];   // 1..15, warp-uniform
                        // sequential fp32 sum: `len` times + cur
#pragma unroll 1
                        for (; len >= 4; len -= 4) { s4 = vadd<V>(s4, cur); s4 = vadd<V>(s4, cur); s4 = vadd<V>(s4, cur); s4 = vadd<V>(s4, cur); }
                        if (len & 2) { s4 = vadd<V>(s4, cur); s4 = vadd<V>(s4, cur); }
                        if (len & 1) s4 = vadd<V>(s4, cur);
                    }
                    cur_cell = s_run_cell[warp][wi][nr - 1];
                    s2 = vadd<V>(s2, vscale<V>(s4, 0.0625f));    // / 16 (exact)
                }
                v0 = vround_half<V>(vscale<V>(s2, 0.25f));      // avg_pool2d(2) -> half (timm.py:168, level 0)
            }
            const int y0 = qy * 2 + l0y, x0 = qx * 2 + l0x;
            vstore_half<V>(L0 + (((size_t)e * h0 + y0) * w0 + x0) * C + V * g, v0);
            l1acc = vadd<V>(l1acc, v0);
        }
        const Vec<V> v1 = vround_half<V>(vscale<V>(l1acc, 0.25f));  // level 1
        vstore_half<V>(L1 + (((size_t)e * h1 + qy) * w1 + qx) * C + V * g, v1);
        __syncwarp();                                       // the slice is rewritten by the next item
    }
}

// level 2 = avg_pool2d(level 1, 2) -> half (timm.py:168): the four L1 pixels in row-major order, fp32, then / 4
__global__ void __launch_bounds__(256) pool_level2_kernel(const __half *__restrict__ L1, int E, int h1, int w1, int C4, __half *__restrict__ L2)
{
    const int h2 = h1 / 2, w2 = w1 / 2;
    const int64_t total = (int64_t)E * h2 * w2 * C4, step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        const int g = (int)(i % C4);
        int64_t p = i / C4;
        const int x = (int)(p % w2); p /= w2;
        const int y = (int)(p % h2);
        const int e = (int)(p / h2);
        Vec<4> s = vzero<4>();
#pragma unroll
        for (int k = 0; k < 4; ++k)
            s = vadd<4>(s, finish<4>(load_raw<4>(L1, nullptr, ((size_t)e * h1 + 2 * y + (k >> 1)) * w1 + 2 * x + (k & 1), C4 * 4, g)));
        vstore_half<4>(L2 + (((size_t)e * h2 + y) * w2 + x) * C4 * 4 + 4 * g, vscale<4>(s, 0.25f));
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
    const int64_t n_items = (int64_t)E * (H / 16) * (W / 16);      // C in {128, 256, 512}: one warp covers all channels of a quadrant (V = C / 32 per lane)
    int64_t blocks = (n_items + kReadWarps - 1) / kReadWarps;
    const int64_t cap = (int64_t)eod_num_sms() * 4 * 4;          // 3-4 resident CTAs per SM, a few waves: each warp walks several items with prefetch
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks), block(kReadWarps * 32);
    __half *l0 = (__half *)L0, *l1 = (__half *)L1, *l2 = (__half *)L2;
    if (mem_is_f16) {
        if (idx_is_i64) read_pool_kernel<C, __half, int64_t><<<grid, block, 0, st>>>((const __half *)table, nullptr, (const int64_t *)idx, H, W, n_cells, E, l0, l1);
        else read_pool_kernel<C, __half, int32_t><<<grid, block, 0, st>>>((const __half *)table, nullptr, (const int32_t *)idx, H, W, n_cells, E, l0, l1);
    } else {
        if (idx_is_i64) read_pool_kernel<C, float, int64_t><<<grid, block, 0, st>>>((const float *)table, counts, (const int64_t *)idx, H, W, n_cells, E, l0, l1);
        else read_pool_kernel<C, float, int32_t><<<grid, block, 0, st>>>((const float *)table, counts, (const int32_t *)idx, H, W, n_cells, E, l0, l1);
    }
    int rc = eod_check_launch("eod_read_pool");
    if (rc) return rc;
    const int64_t total = (int64_t)E * (H / 32) * (W / 32) * (C / 4);
    int64_t b2 = (total + 255) / 256;
    if (b2 > (int64_t)eod_num_sms() * 8) b2 = (int64_t)eod_num_sms() * 8;
    pool_level2_kernel<<<(unsigned)b2, 256, 0, st>>>(l1, E, H / 16, W / 16, C / 4, l2);
    return eod_check_launch("eod_read_pool[level 2]");
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
