// backbone_aux.cu -- the non-GEMM kernels of Resnet18_8s (lib/networks/resnet.py:200-220,
// lib/networks/model_repository.py:64-80): stem conv 7x7/2 (Cin=3), max-pool 3x3/2,
// bilinear x2 upsampling (align_corners=True), NCHW image -> NHWC slice packing, and the
// final 1x1 conv + per-pixel argmax head that writes the reference's NCHW outputs.
// All of them are HBM-bound streaming kernels except the stem (FP32 FMA bound).
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace pvnet {

// ------------------------------------------------------------------ stem
// conv1 (3->64, 7x7, stride 2, pad 3) + folded bn1 + ReLU (resnet.py:201-203).
// in: NCHW [b,3,H,W]; out: NHWC [b,H/2,W/2,out_cs] at out_co (64 channels), tf32-rounded.
// CTA: 8 x 32 output pixels x 64 channels.  Shared memory: the 21 x 69 x 3 input patch and
// all 64*147 weights ([tap][ci][co] so a thread reads 4 consecutive co as one LDS.128
// broadcast).  Thread = one output pixel, 64 accumulators.
constexpr int STEM_TY = 8, STEM_TX = 32;
constexpr int STEM_PH = STEM_TY * 2 + 5, STEM_PW = STEM_TX * 2 + 5;   // 21 x 69
constexpr int STEM_PWP = STEM_PW + 1;

__global__ void __launch_bounds__(256)
    k_stem(const float *__restrict__ in, const float *__restrict__ w /*[49][3][64]*/,
           const float *__restrict__ bias /*[64]*/, float *__restrict__ out, int H, int W, int out_cs, int out_co)
{
    extern __shared__ float sm[];
    float *sw = sm;                               // 147*64
    float *sp = sm + 147 * 64;                    // [3][21][70]
    const int Ho = H / 2, Wo = W / 2;
    const int n = blockIdx.z;
    const int oy0 = blockIdx.y * STEM_TY, ox0 = blockIdx.x * STEM_TX;
    for (int i = threadIdx.x; i < 147 * 64; i += 256) sw[i] = w[i];
    const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
    for (int i = threadIdx.x; i < 3 * STEM_PH * STEM_PW; i += 256) {
        const int c = i / (STEM_PH * STEM_PW);
        const int r = i - c * (STEM_PH * STEM_PW);
        const int py = r / STEM_PW, px = r - py * STEM_PW;
        const int iy = iy0 + py, ix = ix0 + px;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = in[(((size_t)n * 3 + c) * H + iy) * W + ix];
        sp[(c * STEM_PH + py) * STEM_PWP + px] = v;
    }
    __syncthreads();
    const int ty = threadIdx.x / STEM_TX, tx = threadIdx.x % STEM_TX;
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    for (int kh = 0; kh < 7; ++kh) {
        for (int kw = 0; kw < 7; ++kw) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = sp[(c * STEM_PH + ty * 2 + kh) * STEM_PWP + tx * 2 + kw];
                const float4 *wv = reinterpret_cast<const float4 *>(sw + ((kh * 7 + kw) * 3 + c) * 64);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 ww = wv[j];
                    acc[4 * j + 0] = fmaf(v, ww.x, acc[4 * j + 0]);
                    acc[4 * j + 1] = fmaf(v, ww.y, acc[4 * j + 1]);
                    acc[4 * j + 2] = fmaf(v, ww.z, acc[4 * j + 2]);
                    acc[4 * j + 3] = fmaf(v, ww.w, acc[4 * j + 3]);
                }
            }
        }
    }
    const int oy = oy0 + ty, ox = ox0 + tx;
    if (oy < Ho && ox < Wo) {
        float *o = out + (((size_t)n * Ho + oy) * Wo + ox) * out_cs + out_co;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float4 v;
            v.x = ptx::round_tf32(fmaxf(acc[4 * j + 0] + bias[4 * j + 0], 0.f));
            v.y = ptx::round_tf32(fmaxf(acc[4 * j + 1] + bias[4 * j + 1], 0.f));
            v.z = ptx::round_tf32(fmaxf(acc[4 * j + 2] + bias[4 * j + 2], 0.f));
            v.w = ptx::round_tf32(fmaxf(acc[4 * j + 3] + bias[4 * j + 3], 0.f));
            reinterpret_cast<float4 *>(o)[j] = v;
        }
    }
}

int launch_stem(const float *in, const float *w, const float *bias, float *out, int b, int H, int W, int out_cs,
                int out_co, cudaStream_t s)
{
    const size_t smem = (147 * 64 + 3 * STEM_PH * STEM_PWP) * sizeof(float);
    PV_CUDA(ensure_max_smem((const void *)k_stem, (int)smem));
    dim3 grid((W / 2 + STEM_TX - 1) / STEM_TX, (H / 2 + STEM_TY - 1) / STEM_TY, b);
    k_stem<<<grid, 256, smem, s>>>(in, w, bias, out, H, W, out_cs, out_co);
    PV_LAUNCHED("k_stem");
    return PVNET_OK;
}

// ------------------------------------------------------------------ image packing
// NCHW [b,3,H,W] -> NHWC slice [.., co..co+8): 3 image channels (tf32-rounded) + 5 zeros
__global__ void k_pack_image(const float *__restrict__ in, float *__restrict__ out, int npix_per_img, long long total,
                             int out_cs, int out_co)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long n = i / npix_per_img;
    const long long p = i - n * npix_per_img;
    const float *src = in + n * 3 * (long long)npix_per_img + p;
    float4 a = make_float4(ptx::round_tf32(src[0]), ptx::round_tf32(src[npix_per_img]),
                           ptx::round_tf32(src[2 * (long long)npix_per_img]), 0.f);
    float4 *o = reinterpret_cast<float4 *>(out + i * out_cs + out_co);
    o[0] = a;
    o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
}

int launch_pack_image(const float *in, float *out, int b, int H, int W, int out_cs, int out_co, cudaStream_t s)
{
    const long long total = (long long)b * H * W;
    k_pack_image<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, out, H * W, total, out_cs, out_co);
    PV_LAUNCHED("k_pack_image");
    return PVNET_OK;
}

// ------------------------------------------------------------------ space-to-depth + image packing
// One pass over the NCHW image that writes (a) S [b,H/2,W/2,16]: the 2x2 space-to-depth image,
// channel (py*2+px)*3+c, 4 zero channels, for the tensor-core stem (a 7x7 stride-2 conv is a
// 4x4 stride-1 conv on S), and (b) the image slice of the convraw.0 input buffer (3 channels +
// 5 zeros at out_co).  Values rounded to tf32.
__global__ void __launch_bounds__(128)
    k_s2d_pack(const float *__restrict__ in, float *__restrict__ s2d, float *__restrict__ out, int H, int W, int out_cs,
               int out_co)
{
    // grid.y = image * H/2 + half-resolution row; threads over half-resolution columns
    const int W2 = W >> 1, H2 = H >> 1;
    const int x2 = blockIdx.x * blockDim.x + threadIdx.x;
    if (x2 >= W2) return;
    const int n = blockIdx.y / H2, y2 = blockIdx.y - n * H2;
    const size_t plane = (size_t)H * W;
    const float *src = in + (size_t)n * 3 * plane + (size_t)(2 * y2) * W + 2 * x2;
    float v[16];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int py = 0; py < 2; ++py) {
            const float2 t = __ldg(reinterpret_cast<const float2 *>(src + c * plane + py * W));
            v[(py * 2 + 0) * 3 + c] = ptx::round_tf32(t.x);
            v[(py * 2 + 1) * 3 + c] = ptx::round_tf32(t.y);
        }
    v[12] = v[13] = v[14] = v[15] = 0.f;
    float4 *so = reinterpret_cast<float4 *>(s2d + ((size_t)blockIdx.y * W2 + x2) * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) so[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
        for (int px = 0; px < 2; ++px) {
            const size_t pix = (size_t)n * plane + (size_t)(2 * y2 + py) * W + (2 * x2 + px);
            float4 *o = reinterpret_cast<float4 *>(out + pix * out_cs + out_co);
            const int b = (py * 2 + px) * 3;
            o[0] = make_float4(v[b], v[b + 1], v[b + 2], 0.f);
            o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
}

int launch_s2d_pack(const float *in, float *s2d, float *out, int b, int H, int W, int out_cs, int out_co,
                    cudaStream_t s)
{
    dim3 grid((unsigned)((W / 2 + 127) / 128), (unsigned)(b * (H / 2)));
    k_s2d_pack<<<grid, 128, 0, s>>>(in, s2d, out, H, W, out_cs, out_co);
    PV_LAUNCHED("k_s2d_pack");
    return PVNET_OK;
}

// ------------------------------------------------------------------ max-pool 3x3/2 pad 1
// (resnet.py:142,204).  in NHWC [b,H,W,in_cs] at in_co (C channels) -> out [b,H/2,W/2,C]
__global__ void __launch_bounds__(256)
    k_maxpool(const float *__restrict__ in, float *__restrict__ out, int H, int W, int C, int in_cs, int in_co)
{
    // grid.y = image * H/2 + output row
    const int c4 = C >> 2, Ho = H >> 1, Wo = W >> 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Wo * c4) return;
    const int ox = i / c4, cg = i - ox * c4;
    const int n = blockIdx.y / Ho, oy = blockIdx.y - n * Ho;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int iy = oy * 2 + dy;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int ix = ox * 2 + dx;
            if (ix < 0 || ix >= W) continue;
            const float4 v = __ldg(reinterpret_cast<const float4 *>(in + (((size_t)n * H + iy) * W + ix) * in_cs + in_co) + cg);
            m.x = fmaxf(m.x, v.x);
            m.y = fmaxf(m.y, v.y);
            m.z = fmaxf(m.z, v.z);
            m.w = fmaxf(m.w, v.w);
        }
    }
    reinterpret_cast<float4 *>(out)[((size_t)blockIdx.y * Wo + ox) * c4 + cg] = m;
}

int launch_maxpool(const float *in, float *out, int b, int H, int W, int C, int in_cs, int in_co, cudaStream_t s)
{
    dim3 grid((unsigned)(((W / 2) * (C / 4) + 255) / 256), (unsigned)(b * (H / 2)));
    k_maxpool<<<grid, 256, 0, s>>>(in, out, H, W, C, in_cs, in_co);
    PV_LAUNCHED("k_maxpool");
    return PVNET_OK;
}

// ------------------------------------------------------------------ bilinear x2, align_corners=True
// nn.UpsamplingBilinear2d(scale_factor=2) (model_repository.py:35,43,51).  Same arithmetic
// as ATen's upsample_bilinear2d: scale=(in-1)/(out-1) in fp32, src=scale*dst, i0=(int)src,
// l1=src-i0, l0=1-l1, out = h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11).
// in NHWC [b,h,w,C] dense -> out NHWC [b,2h,2w,out_cs] at out_co; values rounded to tf32.
__global__ void __launch_bounds__(256)
    k_upsample2x(const float *__restrict__ in, float *__restrict__ out, int h, int w, int C, int out_cs, int out_co,
                 float sy, float sx)
{
    // grid.y = image * 2h + output row; threads cover (output column, 4-channel group) of that row
    const int c4 = C >> 2, Ho = 2 * h, Wo = 2 * w;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Wo * c4) return;
    const int ox = i / c4, cg = i - ox * c4;
    const int n = blockIdx.y / Ho, oy = blockIdx.y - n * Ho;
    const float fy = sy * (float)oy, fx = sx * (float)ox;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float h1 = fy - (float)y0, h0 = 1.f - h1, w1 = fx - (float)x0, w0 = 1.f - w1;
    const float4 *base = reinterpret_cast<const float4 *>(in) + (size_t)n * h * w * c4 + cg;
    const float4 v00 = __ldg(base + (y0 * w + x0) * c4), v01 = __ldg(base + (y0 * w + x1) * c4);
    const float4 v10 = __ldg(base + (y1 * w + x0) * c4), v11 = __ldg(base + (y1 * w + x1) * c4);
    float4 o;
    o.x = ptx::round_tf32(h0 * (w0 * v00.x + w1 * v01.x) + h1 * (w0 * v10.x + w1 * v11.x));
    o.y = ptx::round_tf32(h0 * (w0 * v00.y + w1 * v01.y) + h1 * (w0 * v10.y + w1 * v11.y));
    o.z = ptx::round_tf32(h0 * (w0 * v00.z + w1 * v01.z) + h1 * (w0 * v10.z + w1 * v11.z));
    o.w = ptx::round_tf32(h0 * (w0 * v00.w + w1 * v01.w) + h1 * (w0 * v10.w + w1 * v11.w));
    *reinterpret_cast<float4 *>(out + ((size_t)blockIdx.y * Wo + ox) * out_cs + out_co + cg * 4) = o;
}

int launch_upsample2x(const float *in, float *out, int b, int h, int w, int C, int out_cs, int out_co, cudaStream_t s)
{
    const float sy = (float)(h - 1) / (float)(2 * h - 1), sx = (float)(w - 1) / (float)(2 * w - 1);
    dim3 grid((unsigned)((2 * w * (C / 4) + 255) / 256), (unsigned)(b * 2 * h));
    k_upsample2x<<<grid, 256, 0, s>>>(in, out, h, w, C, out_cs, out_co, sy, sx);
    PV_LAUNCHED("k_upsample2x");
    return PVNET_OK;
}

// ------------------------------------------------------------------ head
// convraw.3: 1x1 conv raw_dim(=32) -> seg_dim+ver_dim with bias (model_repository.py:57), in
// exact fp32, fused with torch.argmax(seg_pred,1) (tools/demo.py:52; first maximum wins).
// in NHWC [b,H,W,32] -> out NCHW [b,Cout,H,W]; mask int64 [b,H,W] (or u8) optional.
constexpr int HEAD_MAX_COUT = 64;
__global__ void __launch_bounds__(256)
    k_head(const float *__restrict__ in, const float *__restrict__ w /*[Cout][32]*/, const float *__restrict__ bias,
           float *__restrict__ out, void *__restrict__ mask, int mask_esz, int seg_dim, int Cout, int npix,
           long long total)
{
    __shared__ float sw[HEAD_MAX_COUT * 32];
    __shared__ float sb[HEAD_MAX_COUT];
    for (int i = threadIdx.x; i < Cout * 32; i += 256) sw[i] = w[i];
    for (int i = threadIdx.x; i < Cout; i += 256) sb[i] = bias[i];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long n = i / npix;
    const long long p = i - n * npix;
    float v[32];
    const float4 *src = reinterpret_cast<const float4 *>(in + i * 32);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 t = __ldg(src + j);
        v[4 * j] = t.x;
        v[4 * j + 1] = t.y;
        v[4 * j + 2] = t.z;
        v[4 * j + 3] = t.w;
    }
    float best = -INFINITY;
    int best_c = 0;
    float *o = out + n * Cout * (long long)npix + p;
    for (int co = 0; co < Cout; ++co) {
        float acc = sb[co];
        const float *wr = sw + co * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc = fmaf(v[j], wr[j], acc);
        o[(long long)co * npix] = acc;
        if (co < seg_dim && acc > best) {
            best = acc;
            best_c = co;
        }
    }
    if (mask) {
        if (mask_esz == 8)
            reinterpret_cast<long long *>(mask)[i] = best_c;
        else
            reinterpret_cast<unsigned char *>(mask)[i] = (unsigned char)best_c;
    }
}

int launch_head(const float *in, const float *w, const float *bias, float *out, void *mask, int mask_esz, int seg_dim,
                int Cout, int b, int H, int W, cudaStream_t s)
{
    PV_CHECK_ARG(Cout >= 1 && Cout <= HEAD_MAX_COUT, "head: %d output channels unsupported (max %d)", Cout,
                 HEAD_MAX_COUT);
    const long long total = (long long)b * H * W;
    k_head<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, w, bias, out, mask, mask_esz, seg_dim, Cout, H * W,
                                                          total);
    PV_LAUNCHED("k_head");
    return PVNET_OK;
}

}  // namespace pvnet
