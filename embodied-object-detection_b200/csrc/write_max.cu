// Memory write, SMNet height-max mode (SURVEY 8a row A7': SMNet.encode, bytecode only).
//
//   height_map, arg = scatter_max(heights[inliers] + 1000, flat_idx[inliers], out=height_map)
//   m = arg >= 0 ; observed |= m ; state[m] = feature[inliers][arg[m]]            ('replace' update)
//
// Deterministic argmax by packing: key = (orderable(height+1000) << 32) | (raster pixel index + 1);
// atomicMax on a u64 scratch word per cell makes the highest pixel index win among equal heights; the
// winner thread then compares against the persistent fp32 height map with >= (an equal later value
// replaces), which is the canonical torch_scatter-1.4 CPU rule fixed in SURVEY 8(c).
//   pass 1  one thread per lattice pixel: atomicMax(key64[cell], key)                   [index plane only]
//   pass 2  one WARP per lattice pixel: if it is its cell's winner and raises the map, update
//           height_map / arg / observed, copy the pixel's C-vector into state[cell] (lanes over
//           channels) and clear the scratch word.
#include "eod_common.cuh"

namespace {

__device__ __forceinline__ uint32_t orderable(float f)
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(256) max_pass1_kernel(const float *__restrict__ height, const int32_t *__restrict__ idx,
                                                        const uint8_t *__restrict__ outlier, int H, int W, int stride, int Hs,
                                                        int Ws, int64_t n_cells, unsigned long long *__restrict__ key64)
{
    const int e = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Hs * Ws) return;
    const int v = (i / Ws) * stride, u = (i % Ws) * stride;
    const int p = v * W + u;
    const size_t g = (size_t)e * H * W + p;
    if (outlier && __ldg(outlier + g)) return;
    const float h = __fadd_rn(__ldg(height + g), 1000.0f);
    if (h != h) return;          // NaN never satisfies `src >= out`: it may neither raise the cell nor shadow the finite candidates
    const unsigned long long key = ((unsigned long long)orderable(h) << 32) | (unsigned)(p + 1);
    atomicMax(key64 + (size_t)e * n_cells + __ldg(idx + g), key);
}

__global__ void __launch_bounds__(256) max_pass2_kernel(const float *__restrict__ height, const int32_t *__restrict__ idx,
                                                        const uint8_t *__restrict__ outlier, const float *__restrict__ feat,
                                                        int layout, int C, int H, int W, int stride, int Hs, int Ws, int64_t n_cells,
                                                        float *__restrict__ height_map, unsigned long long *__restrict__ key64,
                                                        int32_t *__restrict__ arg_pix, uint8_t *__restrict__ observed,
                                                        float *__restrict__ state)
{
    const int e = blockIdx.y;
    const unsigned lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // each warp scans 32 lattice pixels with one lane each, then serves the winners among them cooperatively
    const int i = warp_global * 32 + lane;
    const int HW = H * W;
    bool win = false;
    int p = 0, cell = 0;
    float h = 0.f;
    if (i < Hs * Ws) {
        const int v = (i / Ws) * stride, u = (i % Ws) * stride;
        p = v * W + u;
        const size_t g = (size_t)e * HW + p;
        if (!(outlier && __ldg(outlier + g))) {
            cell = __ldg(idx + g);
            const unsigned long long k = key64[(size_t)e * n_cells + cell];
            if ((unsigned)(k & 0xffffffffull) == (unsigned)(p + 1)) {      // this pixel won the in-frame contest
                h = __fadd_rn(__ldg(height + g), 1000.0f);
                key64[(size_t)e * n_cells + cell] = 0ull;                   // only the winner resets
                win = h >= height_map[(size_t)e * n_cells + cell];          // '>=': equal later value replaces
                if (win) {
                    height_map[(size_t)e * n_cells + cell] = h;
                    arg_pix[(size_t)e * n_cells + cell] = p;
                    if (observed) observed[(size_t)e * n_cells + cell] = 1;
                }
            }
        }
    }
    if (!feat || !state) return;
    unsigned todo = __ballot_sync(0xffffffffu, win);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int wp = __shfl_sync(0xffffffffu, p, src);
        const int wc = __shfl_sync(0xffffffffu, cell, src);
        float *dst = state + ((size_t)e * n_cells + wc) * C;
        if (layout == EOD_LAYOUT_HWC) {
            const float4 *row = reinterpret_cast<const float4 *>(feat + ((size_t)e * HW + wp) * C);
            for (int k = lane; k < C / 4; k += 32) reinterpret_cast<float4 *>(dst)[k] = __ldg(row + k);
        } else {
            const float *col = feat + (size_t)e * C * HW + wp;
            for (int c = lane; c < C; c += 32) dst[c] = __ldg(col + (size_t)c * HW);
        }
    }
}

}  // namespace

extern "C" int eod_write_max(const float *height, const int32_t *idx, const uint8_t *outlier, const float *feat, int layout,
                             int n_episodes, int C, int H, int W, int pix_stride, int64_t n_cells, float *height_map,
                             uint64_t *key64, int32_t *arg_pix, uint8_t *observed, float *state, eod_stream_t stream)
{
    EOD_REQUIRE(height && idx && height_map && key64 && arg_pix, EOD_ERR_BADARG, "eod_write_max: null pointer");
    EOD_REQUIRE(n_episodes > 0 && n_episodes <= 65535 && H > 0 && W > 0 && n_cells > 0 && pix_stride > 0, EOD_ERR_BADARG, "eod_write_max: bad sizes");
    EOD_REQUIRE(layout == EOD_LAYOUT_CHW || layout == EOD_LAYOUT_HWC, EOD_ERR_BADARG, "eod_write_max: bad layout");
    EOD_REQUIRE(!feat || (C > 0 && C % 4 == 0 && eod_aligned16(feat) && eod_aligned16(state)), EOD_ERR_ALIGN, "eod_write_max: C %% 4 and 16-byte alignment required");
    cudaStream_t st = (cudaStream_t)stream;
    const int Hs = (H + pix_stride - 1) / pix_stride, Ws = (W + pix_stride - 1) / pix_stride;
    const int n = Hs * Ws;
    cudaError_t err = cudaMemsetAsync(arg_pix, 0xff, (size_t)n_episodes * n_cells * sizeof(int32_t), st);   // arg := -1
    EOD_REQUIRE(err == cudaSuccess, EOD_ERR_LAUNCH, "eod_write_max: memset failed: %s", cudaGetErrorString(err));
    dim3 grid((n + 255) / 256, n_episodes);
    max_pass1_kernel<<<grid, 256, 0, st>>>(height, idx, outlier, H, W, pix_stride, Hs, Ws, n_cells, (unsigned long long *)key64);
    int rc = eod_check_launch("eod_write_max[pass1]");
    if (rc) return rc;
    max_pass2_kernel<<<grid, 256, 0, st>>>(height, idx, outlier, feat, layout, C, H, W, pix_stride, Hs, Ws, n_cells, height_map,
                                           (unsigned long long *)key64, arg_pix, observed, state);
    return eod_check_launch("eod_write_max[pass2]");
}
