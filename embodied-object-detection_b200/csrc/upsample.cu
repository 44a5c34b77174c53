// Bilinear upsampling sampled on a regular lattice - the front end of the dense backbone-feature write (SURVEY 8a row
// A7'': `F.interpolate(p3, (480, 640), mode='bilinear', align_corners=True)[:, :, ::8, ::8]`, older CustomMapFPN.forward,
// bytecode only).  Only the lattice points are ever used, so the (C,480,640) upsampled tensor is never built: one
// thread per (channel, lattice point).
//
// Arithmetic == ATen CPU upsample_bilinear2d (align_corners=True), pinned empirically (tests/golden/make_golden.py):
//   scale = float(in - 1) / float(out - 1);  real = scale * o;  i0 = floor(real);  l1 = real - i0;  l0 = 1 - l1
//   row_k = fma(wx0, v[k][x0], wx1 * v[k][x1])   k = 0, 1;    out = fma(wy0, row_0, wy1 * row_1)
#include "eod_common.cuh"

namespace {

__device__ __forceinline__ void lin_weights(int o, int n_in, int n_out, int &i0, int &i1, float &l0, float &l1)
{
    const float scale = n_out > 1 ? __fdiv_rn((float)(n_in - 1), (float)(n_out - 1)) : 0.f;
    const float real = __fmul_rn(scale, (float)o);
    i0 = min((int)floorf(real), n_in - 1);
    i1 = min(i0 + 1, n_in - 1);
    l1 = fminf(fmaxf(__fsub_rn(real, (float)i0), 0.f), 1.f);
    l0 = __fsub_rn(1.f, l1);
}

__global__ void __launch_bounds__(256) bilinear_lattice_kernel(const float *__restrict__ src, int C, int h, int w, int H_out, int W_out,
                                                               int step, int Hl, int Wl, float *__restrict__ out)
{
    const int e = blockIdx.z, c = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Hl * Wl) return;
    const int ly = p / Wl, lx = p - ly * Wl;
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    lin_weights(ly * step, h, H_out, y0, y1, wy0, wy1);
    lin_weights(lx * step, w, W_out, x0, x1, wx0, wx1);
    const float *s = src + ((size_t)e * C + c) * h * w;
    const float v00 = __ldg(s + y0 * w + x0), v01 = __ldg(s + y0 * w + x1), v10 = __ldg(s + y1 * w + x0), v11 = __ldg(s + y1 * w + x1);
    const float r0 = __fmaf_rn(wx0, v00, __fmul_rn(wx1, v01));
    const float r1 = __fmaf_rn(wx0, v10, __fmul_rn(wx1, v11));
    out[((size_t)e * C + c) * Hl * Wl + p] = __fmaf_rn(wy0, r0, __fmul_rn(wy1, r1));
}

}  // namespace

extern "C" int eod_bilinear_lattice(const float *src, int n_episodes, int C, int h, int w, int H_out, int W_out, int step, float *out,
                                    eod_stream_t stream)
{
    EOD_REQUIRE(src && out, EOD_ERR_BADARG, "eod_bilinear_lattice: null pointer");
    EOD_REQUIRE(n_episodes > 0 && n_episodes <= 65535 && C > 0 && C <= 65535 && h > 0 && w > 0 && H_out > 0 && W_out > 0 && step > 0,
                EOD_ERR_BADARG, "eod_bilinear_lattice: bad sizes");
    const int Hl = (H_out + step - 1) / step, Wl = (W_out + step - 1) / step;
    dim3 grid((Hl * Wl + 255) / 256, C, n_episodes);
    bilinear_lattice_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, C, h, w, H_out, W_out, step, Hl, Wl, out);
    return eod_check_launch("eod_bilinear_lattice");
}
