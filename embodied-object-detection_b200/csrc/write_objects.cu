// Memory write, object-feature regime, fused (SURVEY 8a rows A6 + A7; 8f rank 1).
//
// The reference builds a (1, C, 480, 640) fp32 image of per-pixel object-feature means (629 MB at C=512,
// custom_rcnn.py:884-901, one host sync per object), compacts the observed pixels, keeps every 8th and reduces
// them per map cell with a dense one-hot matmul (:903-936).  Only the SAMPLED pixels ever reach the grid, so the
// image is never needed:
//
//   eod_masks_observed   observed[p] = any_k masks[k][p]                                  (K*HW bytes read once)
//   eod_sample_mask      every stride-th observed pixel in raster order (write_mean.cu)
//   eod_frame_count      per-cell sample counts + visibility bits; the first sample of a cell claims a compact
//                        SLOT for it (write_mean.cu)
//   eod_write_objects    for each sampled pixel: g = (sum over the covering objects, in index order, of f_k) / n_obj
//                        - the value box_to_image_features would have stored, bit for bit - then
//                        scratch[slot(cell)] += g   (fp32 reductions into a zeroed per-frame row)
//   eod_flush_slots      sums[cell] += scratch[slot] / n_cell ; scratch row, slot map and slot counter := 0
//   eod_finalize_counts  counts += 1 for every visible cell                (write_mean.cu)
//
// The per-frame scratch keeps the arithmetic of the reference: the frame's samples of a cell are summed on their own,
// divided by their count, and only that mean meets the (much larger) running sum - one rounding at grid magnitude per
// frame instead of one per sample.
//
// Work is ~P'(=observed/8) * K * C adds per frame (a few 10^8): the kernel is latency-, not bandwidth-bound;
// what matters is that the 629 MB image and the K host round trips are gone.
#include "eod_common.cuh"

namespace {

// one thread per 4 pixels (uchar4 mask loads): observed = OR over the episode's objects
__global__ void __launch_bounds__(256) masks_observed_kernel(const uint8_t *__restrict__ masks, const int32_t *__restrict__ n_obj, int Kmax,
                                                             int HW, uint8_t *__restrict__ observed)
{
    const int e = blockIdx.y;
    const int p4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (p4 * 4 >= HW) return;
    const int K = n_obj ? min(__ldg(n_obj + e), Kmax) : Kmax;
    const uint8_t *m = masks + (size_t)e * Kmax * HW;
    uint32_t acc = 0;
    for (int k = 0; k < K; ++k) acc |= __ldg(reinterpret_cast<const uint32_t *>(m + (size_t)k * HW) + p4);
    uint32_t out = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) out |= ((acc >> (8 * b)) & 0xffu) ? (1u << (8 * b)) : 0u;
    reinterpret_cast<uint32_t *>(observed + (size_t)e * HW)[p4] = out;
}

// CTA = 256 consecutive pixels of one episode; a warp takes the CTA's sampled pixels round-robin, lanes own
// channels c = lane + 32*j.  The per-object adds happen in object-index order (custom_rcnn.py:890-895).
template <int C>
__global__ void __launch_bounds__(256) write_objects_kernel(const float *__restrict__ box_features, const uint8_t *__restrict__ masks,
                                                            const int32_t *__restrict__ n_obj, int Kmax, const int32_t *__restrict__ idx,
                                                            const uint8_t *__restrict__ samp, const int32_t *__restrict__ slot_of_cell, int HW,
                                                            int64_t n_cells, int S, float *__restrict__ scratch)
{
    constexpr int J = C / 32;
    __shared__ int s_list[256];
    __shared__ int s_n;
    const int e = blockIdx.y;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    if (p < HW && __ldg(samp + (size_t)e * HW + p)) s_list[atomicAdd(&s_n, 1)] = p;      // order is irrelevant: each pixel is independent
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;
    const int K = n_obj ? min(__ldg(n_obj + e), Kmax) : Kmax;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint8_t *m = masks + (size_t)e * Kmax * HW;
    const float *f = box_features + (size_t)e * Kmax * C;
    for (int i = warp; i < n; i += 8) {
        const int px = s_list[i];
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = 0.f;
        int cnt = 0;
        for (int k0 = 0; k0 < K; k0 += 32) {                       // lanes fetch 32 objects' mask bytes at once
            const int k = k0 + (int)lane;
            const unsigned cover = __ballot_sync(0xffffffffu, k < K && __ldg(m + (size_t)k * HW + px) != 0);
            unsigned todo = cover;
            while (todo) {                                         // ascending object index
                const int kk = k0 + __ffs(todo) - 1;
                todo &= todo - 1;
                const float *row = f + (size_t)kk * C + lane;
#pragma unroll
                for (int j = 0; j < J; ++j) acc[j] = __fadd_rn(acc[j], __ldg(row + 32 * j));
            }
            cnt += __popc(cover);
        }
        if (cnt == 0) continue;                                    // cannot happen for a sampled pixel; defensive
        const int cell = __ldg(idx + (size_t)e * HW + px);
        const int slot = __ldg(slot_of_cell + (size_t)e * n_cells + cell) - 1;
        if (slot < 0 || slot >= S) continue;                       // slot table overflow (S < #sampled pixels): caller error
        const float n_px = (float)cnt;
        float *dst = scratch + ((size_t)e * S + slot) * C + lane;
#pragma unroll
        for (int j = 0; j < J; ++j) red_add_f32(dst + 32 * j, __fdiv_rn(acc[j], n_px));
    }
}

// warp per slot: sums[cell] += scratch[slot] / n_cell (custom_rcnn.py:931-934, :696-697,742); scratch and slot map back to zero
__global__ void __launch_bounds__(256) flush_slots_kernel(const uint32_t *__restrict__ frame_cnt, int32_t *__restrict__ slot_of_cell,
                                                          const int32_t *__restrict__ slot_cell, const int32_t *__restrict__ n_slots, int S, int C,
                                                          int64_t n_cells, float *__restrict__ scratch, float *__restrict__ sums)
{
    const int e = blockIdx.y;
    const int slot = blockIdx.x * 8 + (threadIdx.x >> 5);
    const unsigned lane = threadIdx.x & 31;
    const int n = min(__ldg(n_slots + e), S);
    if (slot >= n) return;
    const int cell = __ldg(slot_cell + (size_t)e * S + slot);
    const float n_cell = (float)(__ldg(frame_cnt + (size_t)e * n_cells + cell) & 0x7fffffffu);
    float4 *src = reinterpret_cast<float4 *>(scratch + ((size_t)e * S + slot) * C);
    float4 *dst = reinterpret_cast<float4 *>(sums + ((size_t)e * n_cells + cell) * C);
    for (int k = lane; k < C / 4; k += 32) {
        const float4 a = src[k];
        float4 d = dst[k];
        d.x = __fadd_rn(d.x, __fdiv_rn(a.x, n_cell)); d.y = __fadd_rn(d.y, __fdiv_rn(a.y, n_cell));
        d.z = __fadd_rn(d.z, __fdiv_rn(a.z, n_cell)); d.w = __fadd_rn(d.w, __fdiv_rn(a.w, n_cell));
        dst[k] = d;
        src[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (lane == 0) slot_of_cell[(size_t)e * n_cells + cell] = 0;
}

__global__ void zero_i32_kernel(int32_t *p, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

}  // namespace

extern "C" int eod_masks_observed(const uint8_t *masks, const int32_t *n_obj, int n_episodes, int Kmax, int HW, uint8_t *observed,
                                  eod_stream_t stream)
{
    EOD_REQUIRE(masks && observed, EOD_ERR_BADARG, "eod_masks_observed: null pointer");
    EOD_REQUIRE(n_episodes > 0 && n_episodes <= 65535 && Kmax >= 0 && HW > 0, EOD_ERR_BADARG, "eod_masks_observed: bad sizes");
    EOD_REQUIRE(HW % 4 == 0 && (reinterpret_cast<uintptr_t>(masks) & 3u) == 0 && (reinterpret_cast<uintptr_t>(observed) & 3u) == 0, EOD_ERR_ALIGN,
                "eod_masks_observed: HW %% 4 == 0 and 4-byte aligned planes required");
    dim3 grid((HW / 4 + 255) / 256, n_episodes);
    masks_observed_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(masks, n_obj, Kmax, HW, observed);
    return eod_check_launch("eod_masks_observed");
}

extern "C" int eod_write_objects(const float *box_features, const uint8_t *masks, const int32_t *n_obj, int Kmax, const int32_t *idx,
                                 const uint8_t *samp, const int32_t *slot_of_cell, int n_episodes, int C, int HW, int64_t n_cells,
                                 int n_slots_max, float *scratch, eod_stream_t stream)
{
    EOD_REQUIRE(box_features && masks && idx && samp && slot_of_cell && scratch, EOD_ERR_BADARG, "eod_write_objects: null pointer");
    EOD_REQUIRE(n_episodes > 0 && n_episodes <= 65535 && Kmax > 0 && HW > 0 && n_cells > 0 && n_slots_max > 0, EOD_ERR_BADARG,
                "eod_write_objects: bad sizes");
    dim3 grid((HW + 255) / 256, n_episodes);
    cudaStream_t st = (cudaStream_t)stream;
    switch (C) {
    case 128: write_objects_kernel<128><<<grid, 256, 0, st>>>(box_features, masks, n_obj, Kmax, idx, samp, slot_of_cell, HW, n_cells, n_slots_max, scratch); break;
    case 256: write_objects_kernel<256><<<grid, 256, 0, st>>>(box_features, masks, n_obj, Kmax, idx, samp, slot_of_cell, HW, n_cells, n_slots_max, scratch); break;
    case 512: write_objects_kernel<512><<<grid, 256, 0, st>>>(box_features, masks, n_obj, Kmax, idx, samp, slot_of_cell, HW, n_cells, n_slots_max, scratch); break;
    default:
        eod_set_error("eod_write_objects: C=%d not compiled in (128, 256, 512)", C);
        return EOD_ERR_UNSUPPORTED;
    }
    return eod_check_launch("eod_write_objects");
}

extern "C" int eod_flush_slots(const uint32_t *frame_cnt, int32_t *slot_of_cell, const int32_t *slot_cell, int32_t *n_slots, int n_episodes,
                               int C, int64_t n_cells, int n_slots_max, float *scratch, float *sums, eod_stream_t stream)
{
    EOD_REQUIRE(frame_cnt && slot_of_cell && slot_cell && n_slots && scratch && sums, EOD_ERR_BADARG, "eod_flush_slots: null pointer");
    EOD_REQUIRE(n_episodes > 0 && n_episodes <= 65535 && C > 0 && C % 4 == 0 && n_cells > 0 && n_slots_max > 0, EOD_ERR_BADARG, "eod_flush_slots: bad sizes");
    EOD_REQUIRE(eod_aligned16(scratch) && eod_aligned16(sums), EOD_ERR_ALIGN, "eod_flush_slots: rows must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((n_slots_max + 7) / 8, n_episodes);
    flush_slots_kernel<<<grid, 256, 0, st>>>(frame_cnt, slot_of_cell, slot_cell, n_slots, n_slots_max, C, n_cells, scratch, sums);
    int rc = eod_check_launch("eod_flush_slots");
    if (rc) return rc;
    zero_i32_kernel<<<(n_episodes + 255) / 256, 256, 0, st>>>(n_slots, n_episodes);
    return eod_check_launch("eod_flush_slots[reset]");
}
