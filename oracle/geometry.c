/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's depth back-projection + pose transform +
 * quantisation to map-cell indices.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call this.
 *
 * Follows (reference paths relative to /root/reference/Detic):
 *   SMNet/projector/core.py:107-108   x_scale = (u + 0.5 - cx) / fx   (fp32, true divide)
 *   SMNet/projector/core.py:142-146   z = d / 1.0 ; x = z * x_scale ; y = z * y_scale
 *   SMNet/projector/core.py:175       world = bmm(T, xyz1)  -- on torch-CPU this is bit-identical to the
 *                                     FMA chain fma(T3,1, fma(T2,z, fma(T1,y, T0*x))) (verified against
 *                                     the imported reference, see tests/golden/make_golden.py)
 *   SMNet/projector/core.py:220       world -= world_shift_origin
 *   SMNet/build_memory_data.py:135-143  p -= map_world_shift ; q = round(p[[0,2]] / (0.02*10)).long() ;
 *                                     clip ; flat = q_z * map_w + q_x
 *   robot_demo.py:526-533             same with flat = q_x * map_h + q_z
 *   SMNet/projector/core.py:258-269   out-of-map + above-camera outlier mask
 *   SMNet/projector/projector.py:90,101  no-depth mask OR-ed in
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no implicit contraction; the only fused
 * operations are the explicit fmaf calls).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define ORACLE_ORDER_ROWMAJOR 0 /* flat = q_z * map_w + q_x   (build_memory_data.py:143) */
#define ORACLE_ORDER_COLMAJOR 1 /* flat = q_x * map_h + q_z   (robot_demo.py:533)        */

static float clipf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
/* build_memory_data.py:136-142: q.round().long() and THEN the clip.  On x86 (the hosts the reference runs on) float -> int64 of NaN,
 * +-inf or anything beyond +-2^63 is INT64_MIN, which clips to cell 0.  (core.py:253-256 compares the rounded floats: no such rule.) */
static int q_overflows(float q) { return !(fabsf(q) < 9.223372036854775808e18f); }
static int32_t clip_cell(float q, int n) { return q_overflows(q) ? 0 : (int32_t)clipf(q, 0.0f, (float)(n - 1)); }

/*
 * depth   (H*W) f32 metres, 0 == no depth
 * T       12 floats: rows 0..2 of the 4x4 camera-to-world matrix (row major)
 * intr    fx, fy, cx, cy
 * shift0  world_shift_origin (core.py:220); shift1 map_world_shift (build_memory_data.py:135)
 * Outputs (any may be NULL): idx (clipped flat index), q2 (H*W*2 unclipped x,z as int32),
 * outlier (u8), height (f32, world y after both shifts), world (H*W*3 f32 after shift0 only).
 */
void oracle_backproject_quantize(const float *depth, int H, int W, const float *T, const float *intr,
                                 const float *shift0, const float *shift1, float cell, int map_w, int map_h,
                                 int order, float z_clip, int32_t *idx, int32_t *q2, uint8_t *outlier,
                                 float *height, float *world)
{
    const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
    const float thr = T[7] + z_clip; /* camera_y + z_clip (core.py:263-266) */
    for (int v = 0; v < H; ++v) {
        const float ys = (((float)v + 0.5f) - cy) / fy;
        for (int u = 0; u < W; ++u) {
            const size_t p = (size_t)v * W + u;
            const float xs = (((float)u + 0.5f) - cx) / fx;
            const float z = depth[p];
            const float x = z * xs;
            const float y = z * ys;
            float w[3];
            for (int r = 0; r < 3; ++r) {
                const float *t = T + 4 * r;
                w[r] = fmaf(t[3], 1.0f, fmaf(t[2], z, fmaf(t[1], y, t[0] * x)));
            }
            float p0x = w[0] - shift0[0], p0y = w[1] - shift0[1], p0z = w[2] - shift0[2];
            if (world) { world[3 * p] = p0x; world[3 * p + 1] = p0y; world[3 * p + 2] = p0z; }
            const float p1x = p0x - shift1[0], p1y = p0y - shift1[1], p1z = p0z - shift1[2];
            const float qx = rintf(p1x / cell); /* torch.round == half-to-even */
            const float qz = rintf(p1z / cell);
            if (q2) { q2[2 * p] = (int32_t)qx; q2[2 * p + 1] = (int32_t)qz; }
            if (outlier) {
                int out = (qx >= (float)map_w) | (qz >= (float)map_h) | (qx < 0.0f) | (qz < 0.0f);
                out |= (p1y > thr);
                out |= (z == 0.0f);
                outlier[p] = (uint8_t)(out != 0);
            }
            if (height) height[p] = p1y;
            if (idx) {
                const int32_t ix = clip_cell(qx, map_w);
                const int32_t iz = clip_cell(qz, map_h);
                idx[p] = order == ORACLE_ORDER_COLMAJOR ? ix * map_h + iz : iz * map_w + ix;
            }
        }
    }
}

/*
 * Read-side pooling chain (timm.py:147-168) for ONE channel-contiguous table, restated with the exact
 * sequential row-major summation order of ATen's CPU avg_pool2d:
 *   E[v,u,:]  = table16[idx[v,u], :]                      (fp16 values, timm.py:147)
 *   P4        = avg_pool2d(float(E), 4, 4)                 (timm.py:152)
 *   L0 = half(avg_pool2d(P4, 2, 2)); L1 = half(avg_pool2d(float(L0), 2, 2)); L2 likewise   (timm.py:168)
 * table16 holds the fp16 bit patterns; outputs are fp16 bit patterns laid out (C, h_l, w_l).
 * Used to cross-check the torch restatement at full size (the torch oracle stays the primary one).
 */
static float half_to_float(uint16_t h)
{
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else { int e = -1; do { man <<= 1; ++e; } while (!(man & 0x400u)); bits = sign | ((uint32_t)(112 - e) << 23) | ((man & 0x3ffu) << 13); }
    } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
    else bits = sign | ((exp + 112) << 23) | (man << 13);
    union { uint32_t u; float f; } c; c.u = bits; return c.f;
}

static uint16_t float_to_half(float f)
{
    union { uint32_t u; float f; } c; c.f = f;
    uint32_t x = c.u, sign = (x >> 16) & 0x8000u; x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0));
    if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);             /* rounds to inf */
    if (x < 0x33000001u) return (uint16_t)sign;                           /* rounds to zero */
    int e = (int)(x >> 23) - 127; uint32_t man = (x & 0x7fffffu) | 0x800000u;
    int shift = e < -14 ? (13 + (-14 - e)) : 13;
    uint32_t half_man = man >> shift, rem = man & ((1u << shift) - 1), halfway = 1u << (shift - 1);
    uint32_t out = e < -14 ? half_man : (((uint32_t)(e + 15) << 10) + (half_man - 0x400u));
    if (rem > halfway || (rem == halfway && (out & 1u))) ++out;            /* RN-even; carries propagate */
    return (uint16_t)(sign | out);
}

void oracle_read_pool_f16(const uint16_t *table16, int C, const int32_t *idx, int H, int W,
                          uint16_t *L0, uint16_t *L1, uint16_t *L2)
{
    const int h4 = H / 4, w4 = W / 4, h0 = h4 / 2, w0 = w4 / 2, h1 = h0 / 2, w1 = w0 / 2, h2 = h1 / 2, w2 = w1 / 2;
    for (int c = 0; c < C; ++c) {
        for (int y = 0; y < h0; ++y)
            for (int x = 0; x < w0; ++x) {
                float s2 = 0.0f;
                for (int dy = 0; dy < 2; ++dy)
                    for (int dx = 0; dx < 2; ++dx) {
                        float s4 = 0.0f;
                        const int v0 = (2 * y + dy) * 4, u0 = (2 * x + dx) * 4;
                        for (int r = 0; r < 4; ++r)
                            for (int k = 0; k < 4; ++k)
                                s4 += half_to_float(table16[(size_t)idx[(size_t)(v0 + r) * W + u0 + k] * C + c]);
                        s2 += s4 / 16.0f;
                    }
                L0[((size_t)c * h0 + y) * w0 + x] = float_to_half(s2 / 4.0f);
            }
        for (int y = 0; y < h1; ++y)
            for (int x = 0; x < w1; ++x) {
                float s = 0.0f;
                for (int dy = 0; dy < 2; ++dy)
                    for (int dx = 0; dx < 2; ++dx) s += half_to_float(L0[((size_t)c * h0 + 2 * y + dy) * w0 + 2 * x + dx]);
                L1[((size_t)c * h1 + y) * w1 + x] = float_to_half(s / 4.0f);
            }
        for (int y = 0; y < h2; ++y)
            for (int x = 0; x < w2; ++x) {
                float s = 0.0f;
                for (int dy = 0; dy < 2; ++dy)
                    for (int dx = 0; dx < 2; ++dx) s += half_to_float(L1[((size_t)c * h1 + 2 * y + dy) * w1 + 2 * x + dx]);
                L2[((size_t)c * h2 + y) * w2 + x] = float_to_half(s / 4.0f);
            }
    }
}

/*
 * Sequential fp32 per-cell mean of sampled pixels in raster order (the fixed summation order the
 * deterministic CUDA variant reproduces bit-for-bit).  feat is (C, H*W) channel-major as in the
 * reference's image_features (custom_rcnn.py:886,905-906); samp[p] != 0 selects the pixel.
 * sum_out (cells*C) and n_out (cells) must be zero-initialised by the caller.
 * Restates custom_rcnn.py:917-934 (one-hot matmul == per-cell sum; mean = sum / count).
 */
void oracle_cell_sums_seq(const float *feat, int C, int HW, const int32_t *idx, const uint8_t *samp,
                          float *sum_out, int32_t *n_out)
{
    for (int p = 0; p < HW; ++p) {
        if (samp && !samp[p]) continue;
        const size_t cell = (size_t)idx[p];
        n_out[cell] += 1;
        for (int c = 0; c < C; ++c) sum_out[cell * C + c] += feat[(size_t)c * HW + p];
    }
}
