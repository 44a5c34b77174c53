/*
 * ORACLE (test infrastructure, not product code): CPU restatement of mask pasting.
 *
 * Reference call site: Detic/detic/modeling/meta_arch/custom_rcnn.py:880
 *     masks = paste_masks_in_image(masks, boxes, (480, 640), threshold=mask_thresh)
 * The callee lives in detectron2 (layers/mask_ops.py: paste_masks_in_image / _do_paste_mask), an unpinned git
 * dependency (README :36-39) that is absent from /root/reference, so its PUBLISHED algorithm is restated here:
 * on the CPU every mask is its own chunk and is pasted with skip_empty=True -
 *     x0_int = clamp(floor(x0) - 1, min=0), x1_int = clamp(ceil(x1) + 1, max=W) (same in y);
 *     img_x = arange(x0_int, x1_int) + 0.5;  gx = (img_x - x0) / (x1 - x0) * 2 - 1   (four separately rounded fp32 ops);
 *     v = F.grid_sample(mask[None, None], grid(gx, gy), align_corners=False)           (bilinear, zero padding);
 *     out[y0_int:y1_int, x0_int:x1_int] = v >= threshold.
 * grid_sample is ATen; its vectorised CPU kernel (AVX2 / AVX-512 builds, GridSamplerKernel.cpp) rounds as
 *     ix = fma(gx + 1, S / 2, -0.5);  w = ix - floor(ix), e = 1 - w, n = iy - floor(iy), s = 1 - n;
 *     v  = fma(se_v, n * w, fma(sw_v, n * e, fma(ne_v, s * w, nw_v * (s * e))))
 * which tests/golden/make_golden.py pins by running torch here (tests/golden/paste.npz: 0 differing bits in v).
 * Parity: the sampler arithmetic is pinned by execution; the wrapper around it is restated from the dependency's
 * published source, not executed ("parity unpinned" for that part).
 */
#include <math.h>
#include <stdint.h>

/* probs (K,S,S) f32, boxes (K,4) f32 -> masks (K,H*W) u8; values (K,H*W) f32 nullable (sampled value inside the region, 0 outside) */
void oracle_paste_masks2(const float *probs, const float *boxes, int K, int S, int H, int W, float thr, int skip_empty, uint8_t *masks, float *values);

void oracle_paste_masks(const float *probs, const float *boxes, int K, int S, int H, int W, float thr, uint8_t *masks, float *values)
{
    oracle_paste_masks2(probs, boxes, K, S, H, W, thr, 1, masks, values);
}

/* skip_empty = 1: the CPU path above (integer neighbourhood of the box).  skip_empty = 0: what _do_paste_mask does for CUDA
 * tensors - the path the reference actually takes (custom_rcnn.py:880 is called on .cuda() tensors): the whole image is sampled.
 * The two agree for probabilities <= 1 and thr >= 0.5 (outside the neighbourhood the bilinear value is < 0.5). */
void oracle_paste_masks2(const float *probs, const float *boxes, int K, int S, int H, int W, float thr, int skip_empty, uint8_t *masks, float *values)
{
    const float half_S = (float)S / 2.0f;
    for (int k = 0; k < K; ++k) {
        const float *m = probs + (long)k * S * S;
        const float x0 = boxes[4 * k], y0 = boxes[4 * k + 1], x1 = boxes[4 * k + 2], y1 = boxes[4 * k + 3];
        const float dx = x1 - x0, dy = y1 - y0;
        float f;
        f = floorf(x0) - 1.0f; int rx0 = (int)(f < 0 ? 0 : (f > W ? W : f));
        f = floorf(y0) - 1.0f; int ry0 = (int)(f < 0 ? 0 : (f > H ? H : f));
        f = ceilf(x1) + 1.0f;  int rx1 = (int)(f > W ? W : (f < 0 ? 0 : f));
        f = ceilf(y1) + 1.0f;  int ry1 = (int)(f > H ? H : (f < 0 ? 0 : f));
        if (!skip_empty) { rx0 = 0; ry0 = 0; rx1 = W; ry1 = H; }
        for (long p = 0; p < (long)H * W; ++p) { masks[(long)k * H * W + p] = 0; if (values) values[(long)k * H * W + p] = 0.0f; }
        for (int py = ry0; py < ry1; ++py) {
            float gy = ((float)py + 0.5f) - y0; gy = gy / dy; gy = gy * 2.0f; gy = gy - 1.0f;
            const float iy = fmaf(gy + 1.0f, half_S, -0.5f);
            const float yn = floorf(iy), ys = yn + 1.0f, n = iy - yn, s = 1.0f - n;
            const int in_n = yn > -1.0f && yn < (float)S, in_s = ys > -1.0f && ys < (float)S;
            for (int px = rx0; px < rx1; ++px) {
                float gx = ((float)px + 0.5f) - x0; gx = gx / dx; gx = gx * 2.0f; gx = gx - 1.0f;
                const float ix = fmaf(gx + 1.0f, half_S, -0.5f);
                const float xw = floorf(ix), xe = xw + 1.0f, w = ix - xw, e = 1.0f - w;
                const int in_w = xw > -1.0f && xw < (float)S, in_e = xe > -1.0f && xe < (float)S;
                const float v_nw = (in_w && in_n) ? m[(int)yn * S + (int)xw] : 0.0f;
                const float v_ne = (in_e && in_n) ? m[(int)yn * S + (int)xe] : 0.0f;
                const float v_sw = (in_w && in_s) ? m[(int)ys * S + (int)xw] : 0.0f;
                const float v_se = (in_e && in_s) ? m[(int)ys * S + (int)xe] : 0.0f;
                float acc = v_nw * (s * e);
                acc = fmaf(v_ne, s * w, acc);
                acc = fmaf(v_sw, n * e, acc);
                acc = fmaf(v_se, n * w, acc);
                masks[(long)k * H * W + (long)py * W + px] = acc >= thr;
                if (values) values[(long)k * H * W + (long)py * W + px] = acc;
            }
        }
    }
}
