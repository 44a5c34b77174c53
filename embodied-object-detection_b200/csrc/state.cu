// Grid state maintenance: memory_reset (custom_rcnn.py:470-477) without rewriting the whole grid.
//
// Invariant of the state the library maintains: a cell's sums row, its normalised fp16 row and its count are
// non-zero only if the cell was visible in some frame since the last reset (writes touch a subset of the visible
// cells, finalize refreshes norm16 for exactly the visible cells and raises their counts).  A reset therefore only
// has to clear the rows whose count is non-zero: the count plane (4 B/cell) is streamed once and ~10^3 rows per
// episode are cleared, instead of (C*6 + 8) B/cell of memset traffic.
#include "eod_common.cuh"

namespace {

// kRefresh == false: memory_reset - clear count, sums row and norm16 row of every cell with a non-zero count.
// kRefresh == true : re-derive the normalised fp16 row (custom_rcnn.py:764-774 + :1036) of every cell with a non-zero
//                    count from the current sums / counts (TEST_TYPE longterm: the read table is refreshed only at the
//                    first frame of a sequence, custom_rcnn.py:482-486).
// mask (E) i32 nullable: only episodes with mask[e] != 0 are processed (per-slot reset / refresh of a lock-step batch).
template <bool kRefresh>
__global__ void __launch_bounds__(256) touched_rows_kernel(float *__restrict__ counts, float *__restrict__ sums, __half *__restrict__ norm16,
                                                           const int32_t *__restrict__ mask, int64_t n_rows, int64_t n_cells, int C)
{
    const unsigned lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint2 z2 = make_uint2(0u, 0u);
    for (int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < n_rows; base += warps * 32) {
        if (mask) {                                      // warp-uniform early-out: both ends of the 32-row chunk in unselected episodes
            const int64_t last = (base + 31 < n_rows ? base + 31 : n_rows - 1);
            const int e0 = (int)(base / n_cells), e1 = (int)(last / n_cells);
            bool any = false;
            for (int e = e0; e <= e1; ++e) any |= __ldg(mask + e) != 0;
            if (!any) continue;
        }
        const int64_t row = base + lane;
        float n = 0.f;
        bool hit = false;
        if (row < n_rows && (!mask || __ldg(mask + (int)(row / n_cells)) != 0)) {
            n = counts[row];
            hit = n != 0.f;
        }
        if (hit && !kRefresh) counts[row] = 0.f;
        unsigned todo = __ballot_sync(0xffffffffu, hit);
        while (todo) {                                   // the warp serves each flagged row cooperatively
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t r = base + src;
            if (kRefresh) {
                const float wn = __shfl_sync(0xffffffffu, n, src);
                const float4 *s = reinterpret_cast<const float4 *>(sums + r * C);
                uint2 *h = reinterpret_cast<uint2 *>(norm16 + r * C);
                for (int k = lane; k < C / 4; k += 32) {
                    float4 v = s[k];
                    if (wn > 1.0f) { v.x = __fdiv_rn(v.x, wn); v.y = __fdiv_rn(v.y, wn); v.z = __fdiv_rn(v.z, wn); v.w = __fdiv_rn(v.w, wn); }
                    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
                    uint2 raw;
                    raw.x = *reinterpret_cast<const uint32_t *>(&a);
                    raw.y = *reinterpret_cast<const uint32_t *>(&b);
                    h[k] = raw;
                }
            } else {
                float4 *s = reinterpret_cast<float4 *>(sums + r * C);
                for (int k = lane; k < C / 4; k += 32) s[k] = z4;
                if (norm16) {
                    uint2 *h = reinterpret_cast<uint2 *>(norm16 + r * C);
                    for (int k = lane; k < C / 4; k += 32) h[k] = z2;
                }
            }
        }
    }
}

// Range check of an externally supplied index plane (proj_indices from memory_data/*.h5, set_indices ...): every
// kernel of the library uses a cell id as a row offset, so an id outside [0, n_cells) must never reach them.
// err[0] += number of out-of-range ids; out32 (nullable) receives the ids as int32, out-of-range ones CLAMPED into
// the grid so that downstream launches stay memory-safe while the host turns err[0] into an IndexError.
template <typename IdxT>
__global__ void __launch_bounds__(256) check_indices_kernel(const IdxT *__restrict__ idx, int64_t n, int64_t n_cells, int32_t *__restrict__ out32,
                                                            int32_t *__restrict__ err)
{
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const long long v = (long long)idx[i];
        const bool ok = v >= 0 && v < (long long)n_cells;
        bad += ok ? 0 : 1;
        if (out32) out32[i] = (int32_t)(ok ? v : (v < 0 ? 0 : n_cells - 1));
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(err, bad);
}

// explicit_map read mode (SMNet/loader.py:233-246, create_explicit_memory in the 3.9 bytecode of custom_rcnn): the projection index of a
// pixel is replaced by the class of the cell it falls into, shifted by `add` (-1 "empty" -> row 0 of the [zeros; class table] memory).
// A bad cell id or a class outside [0, n_rows) is counted in err[0] and clamped, as in check_indices_kernel.
template <typename IdxT, typename LutT>
__global__ void __launch_bounds__(256) remap_indices_kernel(const IdxT *__restrict__ idx, int64_t n, const LutT *__restrict__ lut, int64_t n_cells,
                                                            int add, int64_t n_rows, int32_t *__restrict__ out32, int32_t *__restrict__ err)
{
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const long long c = (long long)idx[i];
        const bool ok_c = c >= 0 && c < (long long)n_cells;
        const long long v = (long long)__ldg(lut + (ok_c ? c : 0)) + add;
        const bool ok = ok_c && v >= 0 && v < (long long)n_rows;
        bad += ok ? 0 : 1;
        out32[i] = (int32_t)(ok ? v : 0);
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(err, bad);
}

int launch_touched(bool refresh, float *counts, float *sums, void *norm16, const int32_t *mask, int64_t n_rows, int64_t n_cells, int C,
                   cudaStream_t st, const char *what)
{
    int64_t blocks = (n_rows + 255) / 256;
    const int64_t cap = (int64_t)eod_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (refresh) touched_rows_kernel<true><<<(int)blocks, 256, 0, st>>>(counts, sums, (__half *)norm16, mask, n_rows, n_cells, C);
    else touched_rows_kernel<false><<<(int)blocks, 256, 0, st>>>(counts, sums, (__half *)norm16, mask, n_rows, n_cells, C);
    return eod_check_launch(what);
}

}  // namespace

extern "C" int eod_reset_touched(float *counts, float *sums, void *norm16, int64_t n_rows, int C, eod_stream_t stream)
{
    EOD_REQUIRE(counts && sums, EOD_ERR_BADARG, "eod_reset_touched: null pointer");
    EOD_REQUIRE(n_rows > 0 && C > 0 && C % 4 == 0, EOD_ERR_BADARG, "eod_reset_touched: bad sizes (C %% 4 == 0)");
    EOD_REQUIRE(eod_aligned16(sums) && (!norm16 || eod_aligned16(norm16)), EOD_ERR_ALIGN, "eod_reset_touched: rows must be 16-byte aligned");
    return launch_touched(false, counts, sums, norm16, nullptr, n_rows, n_rows, C, (cudaStream_t)stream, "eod_reset_touched");
}

extern "C" int eod_reset_episodes(float *counts, float *sums, void *norm16, const int32_t *mask, int n_episodes, int64_t n_cells, int C,
                                  eod_stream_t stream)
{
    EOD_REQUIRE(counts && sums && mask, EOD_ERR_BADARG, "eod_reset_episodes: null pointer");
    EOD_REQUIRE(n_episodes > 0 && n_cells > 0 && C > 0 && C % 4 == 0, EOD_ERR_BADARG, "eod_reset_episodes: bad sizes (C %% 4 == 0)");
    EOD_REQUIRE(eod_aligned16(sums) && (!norm16 || eod_aligned16(norm16)), EOD_ERR_ALIGN, "eod_reset_episodes: rows must be 16-byte aligned");
    return launch_touched(false, counts, sums, norm16, mask, (int64_t)n_episodes * n_cells, n_cells, C, (cudaStream_t)stream, "eod_reset_episodes");
}

extern "C" int eod_refresh_norm16(const float *counts, const float *sums, void *norm16, const int32_t *mask, int n_episodes, int64_t n_cells,
                                  int C, eod_stream_t stream)
{
    EOD_REQUIRE(counts && sums && norm16, EOD_ERR_BADARG, "eod_refresh_norm16: null pointer");
    EOD_REQUIRE(n_episodes > 0 && n_cells > 0 && C > 0 && C % 4 == 0, EOD_ERR_BADARG, "eod_refresh_norm16: bad sizes (C %% 4 == 0)");
    EOD_REQUIRE(eod_aligned16(sums) && eod_aligned16(norm16), EOD_ERR_ALIGN, "eod_refresh_norm16: rows must be 16-byte aligned");
    return launch_touched(true, const_cast<float *>(counts), const_cast<float *>(sums), norm16, mask, (int64_t)n_episodes * n_cells, n_cells, C,
                          (cudaStream_t)stream, "eod_refresh_norm16");
}

extern "C" int eod_check_indices(const void *idx, int idx_is_i64, int64_t n, int64_t n_cells, int32_t *idx32_out, int32_t *err, eod_stream_t stream)
{
    EOD_REQUIRE(idx && err, EOD_ERR_BADARG, "eod_check_indices: null pointer");
    EOD_REQUIRE(n > 0 && n_cells > 0 && n_cells <= 0x7fffffffll, EOD_ERR_BADARG, "eod_check_indices: bad sizes (cells < 2^31)");
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)eod_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (idx_is_i64) check_indices_kernel<int64_t><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const int64_t *)idx, n, n_cells, idx32_out, err);
    else check_indices_kernel<int32_t><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const int32_t *)idx, n, n_cells, idx32_out, err);
    return eod_check_launch("eod_check_indices");
}

extern "C" int eod_remap_indices(const void *idx, int idx_is_i64, int64_t n, const void *lut, int lut_is_i64, int64_t n_cells, int add, int64_t n_rows,
                                 int32_t *out32, int32_t *err, eod_stream_t stream)
{
    EOD_REQUIRE(idx && lut && out32 && err, EOD_ERR_BADARG, "eod_remap_indices: null pointer");
    EOD_REQUIRE(n > 0 && n_cells > 0 && n_rows > 0 && n_rows <= 0x7fffffffll, EOD_ERR_BADARG, "eod_remap_indices: bad sizes");
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)eod_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (idx_is_i64 && lut_is_i64) remap_indices_kernel<int64_t, int64_t><<<(int)blocks, 256, 0, st>>>((const int64_t *)idx, n, (const int64_t *)lut, n_cells, add, n_rows, out32, err);
    else if (idx_is_i64) remap_indices_kernel<int64_t, int32_t><<<(int)blocks, 256, 0, st>>>((const int64_t *)idx, n, (const int32_t *)lut, n_cells, add, n_rows, out32, err);
    else if (lut_is_i64) remap_indices_kernel<int32_t, int64_t><<<(int)blocks, 256, 0, st>>>((const int32_t *)idx, n, (const int64_t *)lut, n_cells, add, n_rows, out32, err);
    else remap_indices_kernel<int32_t, int32_t><<<(int)blocks, 256, 0, st>>>((const int32_t *)idx, n, (const int32_t *)lut, n_cells, add, n_rows, out32, err);
    return eod_check_launch("eod_remap_indices");
}
