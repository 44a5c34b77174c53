// Grid state maintenance: memory_reset (custom_rcnn.py:470-477) without rewriting the whole grid.
//
// Invariant of the state the library maintains: a cell's sums row, its normalised fp16 row and its count are
// non-zero only if the cell was visible in some frame since the last reset (writes touch a subset of the visible
// cells, finalize refreshes norm16 for exactly the visible cells and raises their counts).  A reset therefore only
// has to clear the rows whose count is non-zero: the count plane (4 B/cell) is streamed once and ~10^3 rows per
// episode are cleared, instead of (C*6 + 8) B/cell of memset traffic.
#include "eod_common.cuh"

namespace {

__global__ void __launch_bounds__(256) reset_touched_kernel(float *__restrict__ counts, float *__restrict__ sums, __half *__restrict__ norm16,
                                                            int64_t n_rows, int C)
{
    const unsigned lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint2 z2 = make_uint2(0u, 0u);
    for (int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < n_rows; base += warps * 32) {
        const int64_t row = base + lane;
        const bool hit = row < n_rows && counts[row] != 0.f;
        if (hit) counts[row] = 0.f;
        unsigned todo = __ballot_sync(0xffffffffu, hit);
        while (todo) {                                   // the warp clears each flagged row cooperatively
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t r = base + src;
            float4 *s = reinterpret_cast<float4 *>(sums + r * C);
            for (int k = lane; k < C / 4; k += 32) s[k] = z4;
            if (norm16) {
                uint2 *h = reinterpret_cast<uint2 *>(norm16 + r * C);
                for (int k = lane; k < C / 4; k += 32) h[k] = z2;
            }
        }
    }
}

}  // namespace

extern "C" int eod_reset_touched(float *counts, float *sums, void *norm16, int64_t n_rows, int C, eod_stream_t stream)
{
    EOD_REQUIRE(counts && sums, EOD_ERR_BADARG, "eod_reset_touched: null pointer");
    EOD_REQUIRE(n_rows > 0 && C > 0 && C % 4 == 0, EOD_ERR_BADARG, "eod_reset_touched: bad sizes (C %% 4 == 0)");
    EOD_REQUIRE(eod_aligned16(sums) && (!norm16 || eod_aligned16(norm16)), EOD_ERR_ALIGN, "eod_reset_touched: rows must be 16-byte aligned");
    int64_t blocks = (n_rows + 255) / 256;
    const int64_t cap = (int64_t)eod_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    reset_touched_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(counts, sums, (__half *)norm16, n_rows, C);
    return eod_check_launch("eod_reset_touched");
}
