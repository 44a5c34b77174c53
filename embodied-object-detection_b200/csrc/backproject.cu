// Depth back-projection + pose transform + quantisation to map-cell indices.
//
// One thread per pixel; every access is coalesced (depth read, idx/mask/height writes).  ~9 B/pixel of
// HBM traffic, so the kernel is a small streaming pass; it exists to be BIT-EXACT with torch-CPU:
//   x_scale = ((u + 0.5) - cx) / fx         IEEE divide          (core.py:107)
//   world_r = fma(T3,1, fma(T2,z, fma(T1,y, T0*x)))              (core.py:175, torch.bmm K=4 on CPU)
//   q       = rint((world - shift0 - shift1) / cell)             (core.py:220, build_memory_data.py:135-136)
// All arithmetic uses explicit _rn intrinsics so nvcc can neither contract nor reassociate.
#include "eod_common.cuh"

namespace {

struct BackprojectParams {
    const float *depth;
    const float *pose;
    const float *shifts;
    int32_t *idx;
    int32_t *q2;
    uint8_t *outlier;
    float *height;
    float *world;
    int H, W;
    float fx, fy, cx, cy, cell, z_clip;
    int map_w, map_h, order;
};

__global__ void __launch_bounds__(256) backproject_quantize_kernel(const BackprojectParams P)
{
    const int e = blockIdx.y;
    const int HW = P.H * P.W;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int v = p / P.W, u = p - v * P.W;

    const float *T = P.pose + 12 * e;       // uniform across the block: served by the constant/L1 path
    const float *S = P.shifts + 6 * e;

    const float xs = __fdiv_rn(__fsub_rn(__fadd_rn((float)u, 0.5f), P.cx), P.fx);
    const float ys = __fdiv_rn(__fsub_rn(__fadd_rn((float)v, 0.5f), P.cy), P.fy);
    const size_t g = (size_t)e * HW + p;
    const float z = __ldg(P.depth + g);
    const float x = __fmul_rn(z, xs);
    const float y = __fmul_rn(z, ys);

    float w[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float t0 = __ldg(T + 4 * r), t1 = __ldg(T + 4 * r + 1), t2 = __ldg(T + 4 * r + 2), t3 = __ldg(T + 4 * r + 3);
        w[r] = __fmaf_rn(t3, 1.0f, __fmaf_rn(t2, z, __fmaf_rn(t1, y, __fmul_rn(t0, x))));
    }
    const float p0x = __fsub_rn(w[0], __ldg(S + 0)), p0y = __fsub_rn(w[1], __ldg(S + 1)), p0z = __fsub_rn(w[2], __ldg(S + 2));
    if (P.world) {
        float *o = P.world + 3 * g;
        o[0] = p0x; o[1] = p0y; o[2] = p0z;
    }
    const float p1x = __fsub_rn(p0x, __ldg(S + 3)), p1y = __fsub_rn(p0y, __ldg(S + 4)), p1z = __fsub_rn(p0z, __ldg(S + 5));
    const float qx = rintf(__fdiv_rn(p1x, P.cell));
    const float qz = rintf(__fdiv_rn(p1z, P.cell));
    if (P.q2) {
        P.q2[2 * g] = (int32_t)qx;
        P.q2[2 * g + 1] = (int32_t)qz;
    }
    if (P.outlier) {
        const float thr = __fadd_rn(__ldg(T + 7), P.z_clip);
        const bool out = (qx >= (float)P.map_w) || (qz >= (float)P.map_h) || (qx < 0.0f) || (qz < 0.0f) ||
                         (p1y > thr) || (z == 0.0f);
        P.outlier[g] = out ? 1 : 0;
    }
    if (P.height) P.height[g] = p1y;
    if (P.idx) {
        const int ix = (int)fminf(fmaxf(qx, 0.0f), (float)(P.map_w - 1));
        const int iz = (int)fminf(fmaxf(qz, 0.0f), (float)(P.map_h - 1));
        P.idx[g] = P.order == EOD_ORDER_XZ ? ix * P.map_h + iz : iz * P.map_w + ix;
    }
}

// World xyz (as stored in sensor_data/*.h5 'projection_indices', SMNet/build_data.py:209-213,280) -> flat clipped cell
// index: exactly SMNet/build_memory_data.py:135-143 (shift, IEEE divide, round-half-even, clip, z*map_w + x).
__global__ void __launch_bounds__(256) quantize_world_kernel(const float *__restrict__ world, int64_t n, float sx, float sz, float cell,
                                                             int map_w, int map_h, int order, int32_t *__restrict__ idx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = __ldg(world + 3 * i), z = __ldg(world + 3 * i + 2);
    const float qx = rintf(__fdiv_rn(__fsub_rn(x, sx), cell));
    const float qz = rintf(__fdiv_rn(__fsub_rn(z, sz), cell));
    const int ix = (int)fminf(fmaxf(qx, 0.0f), (float)(map_w - 1));
    const int iz = (int)fminf(fmaxf(qz, 0.0f), (float)(map_h - 1));
    idx[i] = order == EOD_ORDER_XZ ? ix * map_h + iz : iz * map_w + ix;
}

}  // namespace

extern "C" int eod_quantize_world(const float *world, int64_t n_points, float shift_x, float shift_z, float cell, int map_w, int map_h,
                                  int order, int32_t *idx, eod_stream_t stream)
{
    EOD_REQUIRE(world && idx, EOD_ERR_BADARG, "eod_quantize_world: null pointer");
    EOD_REQUIRE(n_points > 0 && map_w > 0 && map_h > 0 && cell > 0.0f, EOD_ERR_BADARG, "eod_quantize_world: bad sizes");
    EOD_REQUIRE(order == EOD_ORDER_ZX || order == EOD_ORDER_XZ, EOD_ERR_BADARG, "eod_quantize_world: bad order");
    EOD_REQUIRE((int64_t)map_w * map_h < (int64_t)INT32_MAX, EOD_ERR_BADARG, "eod_quantize_world: map too large");
    const int64_t blocks = (n_points + 255) / 256;
    EOD_REQUIRE(blocks <= 0x7fffffff, EOD_ERR_BADARG, "eod_quantize_world: too many points for one launch");
    quantize_world_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(world, n_points, shift_x, shift_z, cell, map_w, map_h, order, idx);
    return eod_check_launch("eod_quantize_world");
}

extern "C" int eod_backproject_quantize(const float *depth, const float *pose, const float *shifts, int n_episodes,
                                        int H, int W, float fx, float fy, float cx, float cy, float cell, int map_w,
                                        int map_h, int order, float z_clip, int32_t *idx, int32_t *q2,
                                        uint8_t *outlier, float *height, float *world, eod_stream_t stream)
{
    EOD_REQUIRE(depth && pose && shifts, EOD_ERR_BADARG, "eod_backproject_quantize: null input");
    EOD_REQUIRE(n_episodes > 0 && H > 0 && W > 0 && map_w > 0 && map_h > 0, EOD_ERR_BADARG,
                "eod_backproject_quantize: non-positive size");
    EOD_REQUIRE(order == EOD_ORDER_ZX || order == EOD_ORDER_XZ, EOD_ERR_BADARG, "eod_backproject_quantize: bad order");
    EOD_REQUIRE(cell > 0.0f && fx != 0.0f && fy != 0.0f, EOD_ERR_BADARG, "eod_backproject_quantize: bad cell/intrinsics");
    EOD_REQUIRE((int64_t)map_w * map_h < (int64_t)INT32_MAX, EOD_ERR_BADARG, "eod_backproject_quantize: map too large");
    EOD_REQUIRE(n_episodes <= 65535, EOD_ERR_BADARG, "eod_backproject_quantize: n_episodes > 65535");
    BackprojectParams P{depth, pose, shifts, idx, q2, outlier, height, world, H, W, fx, fy, cx, cy, cell, z_clip, map_w, map_h, order};
    dim3 grid((H * W + 255) / 256, n_episodes);
    backproject_quantize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
    return eod_check_launch("eod_backproject_quantize");
}
