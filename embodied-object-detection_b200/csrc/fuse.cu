// Fusion epilogue (timm.py:177-189): out = res + w*mem | w*mem | res.
// torch computes `mem * w` and `mem + res` as two separately rounded fp32 ops; the explicit _rn intrinsics
// keep nvcc from contracting them into one FMA, so the result is bit-identical.  128-bit vectorised,
// grid-stride, 2 reads + 1 write per element (pure streaming).
#include "eod_common.cuh"

namespace {

template <int MODE>
__global__ void __launch_bounds__(256) fuse_kernel(const float4 *__restrict__ res, const float4 *__restrict__ mem, float w, int64_t n4,
                                                   const float *__restrict__ res_tail, const float *__restrict__ mem_tail, int tail,
                                                   float4 *__restrict__ out, float *__restrict__ out_tail)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 o;
        if (MODE == EOD_FUSE_IMAGE_ONLY) {
            o = __ldg(res + i);
        } else {
            const float4 m = __ldg(mem + i);
            o = make_float4(__fmul_rn(m.x, w), __fmul_rn(m.y, w), __fmul_rn(m.z, w), __fmul_rn(m.w, w));
            if (MODE == EOD_FUSE_SUM) {
                const float4 r = __ldg(res + i);
                o = make_float4(__fadd_rn(o.x, r.x), __fadd_rn(o.y, r.y), __fadd_rn(o.z, r.z), __fadd_rn(o.w, r.w));
            }
        }
        out[i] = o;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) {
        const int i = threadIdx.x;
        float o;
        if (MODE == EOD_FUSE_IMAGE_ONLY) o = res_tail[i];
        else {
            o = __fmul_rn(mem_tail[i], w);
            if (MODE == EOD_FUSE_SUM) o = __fadd_rn(o, res_tail[i]);
        }
        out_tail[i] = o;
    }
}

}  // namespace

extern "C" int eod_fuse(const float *res, const float *mem, float weight, int mode, int64_t n, float *out, eod_stream_t stream)
{
    EOD_REQUIRE(out && n > 0, EOD_ERR_BADARG, "eod_fuse: null output or n <= 0");
    EOD_REQUIRE(mode >= EOD_FUSE_SUM && mode <= EOD_FUSE_IMAGE_ONLY, EOD_ERR_BADARG, "eod_fuse: bad mode %d", mode);
    EOD_REQUIRE(mode == EOD_FUSE_IMAGE_ONLY || mem, EOD_ERR_BADARG, "eod_fuse: mem is null");
    EOD_REQUIRE(mode == EOD_FUSE_MEM_ONLY || res, EOD_ERR_BADARG, "eod_fuse: res is null");
    EOD_REQUIRE(eod_aligned16(out) && (!res || eod_aligned16(res)) && (!mem || eod_aligned16(mem)), EOD_ERR_ALIGN, "eod_fuse: pointers must be 16-byte aligned");
    const int64_t n4 = n / 4;
    const int tail = (int)(n - 4 * n4);
    int64_t blocks = (n4 + 255) / 256;
    const int64_t cap = (int64_t)eod_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const float4 *r4 = (const float4 *)res, *m4 = (const float4 *)mem;
    const float *rt = res ? res + 4 * n4 : nullptr, *mt = mem ? mem + 4 * n4 : nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == EOD_FUSE_SUM) fuse_kernel<EOD_FUSE_SUM><<<(int)blocks, 256, 0, st>>>(r4, m4, weight, n4, rt, mt, tail, (float4 *)out, out + 4 * n4);
    else if (mode == EOD_FUSE_MEM_ONLY) fuse_kernel<EOD_FUSE_MEM_ONLY><<<(int)blocks, 256, 0, st>>>(r4, m4, weight, n4, rt, mt, tail, (float4 *)out, out + 4 * n4);
    else fuse_kernel<EOD_FUSE_IMAGE_ONLY><<<(int)blocks, 256, 0, st>>>(r4, m4, weight, n4, rt, mt, tail, (float4 *)out, out + 4 * n4);
    return eod_check_launch("eod_fuse");
}
