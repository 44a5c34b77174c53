// Explicit semantic-map decode: custom_rcnn.py:747-756 + visualise_clip_image_features
// (:938-978).  The reference recomputes, every frame and over the WHOLE grid, mean(|S|)/count, its global min/max
// normalisation, a (cells,512)x(512,K) GEMM, softmax and argmax, then copies the result to the host.  Only the cells
// visible in the frame can change, so two small per-cell planes are maintained instead:
//   intensity[c] = mean_ch |S_c| (/ count_c if count_c > 1)            (custom_rcnn.py:747-748)
//   cls[c]       = argmax_{k < n_cls} <S_c, zs_weight[:, k]>           (:946-958; 50*normalize and softmax are monotone)
// eod_semmap_update refreshes them for the frame's visible cells (run it after the write, before eod_finalize_counts,
// while frame_cnt still marks them); eod_semmap_decode does the global min/max (:751) over the 4 B/cell intensity plane
// and emits semmap[c] = cls[c], or -1 where the normalised intensity is below MEMORY_OBS_SCORE_THRESH (:968).
#include <float.h>

#include "eod_common.cuh"

namespace {

constexpr int MAX_CLS = 32;

template <int C>
__global__ void __launch_bounds__(256) semmap_update_kernel(int64_t n_rows, const uint32_t *__restrict__ frame_cnt, const float *__restrict__ counts,
                                                            const float *__restrict__ sums, const float *__restrict__ zs, int ldz, int n_cls,
                                                            float *__restrict__ intensity, int32_t *__restrict__ cls)
{
    constexpr int J = C / 32;
    const unsigned lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < n_rows; base += warps * 32) {
        const int64_t row = base + lane;
        const bool vis = row < n_rows && frame_cnt[row] != 0u;
        unsigned todo = __ballot_sync(0xffffffffu, vis);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t r = base + src;
            const float *x = sums + r * C + lane;
            float abs_sum = 0.f;
            float dot[MAX_CLS];
#pragma unroll
            for (int k = 0; k < MAX_CLS; ++k) dot[k] = 0.f;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const float v = x[32 * j];
                abs_sum += fabsf(v);
                const float *z = zs + (size_t)(lane + 32 * j) * ldz;
#pragma unroll
                for (int k = 0; k < MAX_CLS; ++k)
                    if (k < n_cls) dot[k] = fmaf(v, __ldg(z + k), dot[k]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) abs_sum += __shfl_xor_sync(0xffffffffu, abs_sum, o);
            int best = 0;
            float best_v = -FLT_MAX;
#pragma unroll
            for (int k = 0; k < MAX_CLS; ++k) {
                if (k < n_cls) {
                    float d = dot[k];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                    if (d > best_v) { best_v = d; best = k; }          // first maximum wins, like torch.max
                }
            }
            if (lane == 0) {
                const float n = counts[r] + 1.0f;                       // visibility count once this frame is finalised
                float inten = abs_sum / (float)C;
                if (n > 1.0f) inten = inten / n;
                intensity[r] = inten;
                cls[r] = best;
            }
        }
    }
}

// one CTA per episode: global min / max of the intensity plane (custom_rcnn.py:751)
__global__ void __launch_bounds__(1024) semmap_minmax_kernel(const float *__restrict__ intensity, int64_t n_cells, float *__restrict__ minmax)
{
    __shared__ float s_min[32], s_max[32];
    const int e = blockIdx.x;
    const float *p = intensity + (size_t)e * n_cells;
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (int64_t i = threadIdx.x; i < n_cells; i += 1024) {
        const float v = p[i];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = lo; s_max[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        lo = s_min[threadIdx.x];
        hi = s_max[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (threadIdx.x == 0) { minmax[2 * e] = lo; minmax[2 * e + 1] = hi; }
    }
}

__global__ void __launch_bounds__(256) semmap_decode_kernel(const float *__restrict__ intensity, const int32_t *__restrict__ cls,
                                                            const float *__restrict__ minmax, int64_t n_cells, float thresh,
                                                            int32_t *__restrict__ semmap)
{
    const int e = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    const float lo = minmax[2 * e], hi = minmax[2 * e + 1];
    const size_t g = (size_t)e * n_cells + i;
    const float v = __fdiv_rn(__fsub_rn(intensity[g], lo), __fsub_rn(hi, lo));     // NaN when hi == lo: compares false, keeps the class
    semmap[g] = (v < thresh) ? -1 : cls[g];
}

}  // namespace

extern "C" int eod_semmap_update(const uint32_t *frame_cnt, const float *counts, const float *sums, const float *zs_weight, int ldz, int n_cls,
                                 int n_episodes, int C, int64_t n_cells, float *intensity, int32_t *cls, eod_stream_t stream)
{
    EOD_REQUIRE(frame_cnt && counts && sums && zs_weight && intensity && cls, EOD_ERR_BADARG, "eod_semmap_update: null pointer");
    EOD_REQUIRE(n_episodes > 0 && n_cells > 0 && n_cls > 0 && n_cls <= MAX_CLS && ldz >= n_cls, EOD_ERR_BADARG,
                "eod_semmap_update: bad sizes (n_cls <= 32, ldz >= n_cls)");
    const int64_t n_rows = (int64_t)n_episodes * n_cells;
    int64_t blocks = (n_rows + 255) / 256;
    const int64_t cap = (int64_t)eod_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    switch (C) {
    case 128: semmap_update_kernel<128><<<(int)blocks, 256, 0, st>>>(n_rows, frame_cnt, counts, sums, zs_weight, ldz, n_cls, intensity, cls); break;
    case 256: semmap_update_kernel<256><<<(int)blocks, 256, 0, st>>>(n_rows, frame_cnt, counts, sums, zs_weight, ldz, n_cls, intensity, cls); break;
    case 512: semmap_update_kernel<512><<<(int)blocks, 256, 0, st>>>(n_rows, frame_cnt, counts, sums, zs_weight, ldz, n_cls, intensity, cls); break;
    default:
        eod_set_error("eod_semmap_update: C=%d not compiled in (128, 256, 512)", C);
        return EOD_ERR_UNSUPPORTED;
    }
    return eod_check_launch("eod_semmap_update");
}

extern "C" int eod_semmap_decode(const float *intensity, const int32_t *cls, int n_episodes, int64_t n_cells, float thresh, float *minmax_ws,
                                 int32_t *semmap, eod_stream_t stream)
{
    EOD_REQUIRE(intensity && cls && minmax_ws && semmap, EOD_ERR_BADARG, "eod_semmap_decode: null pointer");
    EOD_REQUIRE(n_episodes > 0 && n_episodes <= 65535 && n_cells > 0, EOD_ERR_BADARG, "eod_semmap_decode: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    semmap_minmax_kernel<<<n_episodes, 1024, 0, st>>>(intensity, n_cells, minmax_ws);
    dim3 grid((unsigned)((n_cells + 255) / 256), n_episodes);
    semmap_decode_kernel<<<grid, 256, 0, st>>>(intensity, cls, minmax_ws, n_cells, thresh, semmap);
    return eod_check_launch("eod_semmap_decode");
}
