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
// U8 = true: `in` is a raw uint8 HWC image [b,H,W,3] (what an image decoder yields) and the kernel
// applies torchvision's ToTensor + Normalize in their own fp32 arithmetic (tools/demo.py:89-95,
// lib/datasets/linemod_dataset.py:191-195): (float(v) / 255 - mean[c]) / std[c], three correctly
// rounded ops -- bit-identical to feeding the float path with the torch-normalised tensor, at a
// quarter of the input bytes.
struct Norm3 {
    float mean[3], std[3];
};
template <bool U8>
__global__ void __launch_bounds__(128)
    k_s2d_pack(const void *__restrict__ in_v, Norm3 nrm, float *__restrict__ s2d, float *__restrict__ out, int H, int W,
               int out_cs, int out_co)
{
    const float *in = static_cast<const float *>(in_v);
    // grid.y = image * H/2 + half-resolution row; a block covers 128 half-resolution columns.
    // Loads are coalesced float2 reads of the six (channel, row) lines; both outputs are staged in
    // shared memory so that the stores are coalesced too (a thread-per-pixel store touches 32
    // different 64-byte / 160-byte-strided records per instruction).
    __shared__ float4 sS[128 * 4];        // [px][4 chunks], chunk j of px at j ^ ((px >> 1) & 3)
    __shared__ float4 sI[2][256];         // [row parity][full-resolution column]: (r, g, b, 0)
    const int W2 = W >> 1, H2 = H >> 1;
    const int xb = blockIdx.x * 128, t = threadIdx.x;
    const int x2 = xb + t;
    const int n = blockIdx.y / H2, y2 = blockIdx.y - n * H2;
    const size_t plane = (size_t)H * W;
    if (x2 < W2) {
        float v[16];
        if (U8) {
            // two rows x (2 pixels x 3 bytes): three 16-bit loads per row, coalesced across the warp
            const unsigned short *src8 = reinterpret_cast<const unsigned short *>(
                static_cast<const unsigned char *>(in_v) + (((size_t)n * H + 2 * y2) * W + 2 * x2) * 3);
#pragma unroll
            for (int py = 0; py < 2; ++py) {
                unsigned bytes[6];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const unsigned q = __ldg(src8 + (size_t)py * W * 3 / 2 + j);
                    bytes[2 * j] = q & 0xffu;
                    bytes[2 * j + 1] = q >> 8;
                }
#pragma unroll
                for (int px = 0; px < 2; ++px)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float t = __fdiv_rn(__fsub_rn(__fdiv_rn((float)bytes[px * 3 + c], 255.f), nrm.mean[c]), nrm.std[c]);
                        v[(py * 2 + px) * 3 + c] = ptx::round_tf32(t);
                    }
            }
        } else {
            const float *src = in + (size_t)n * 3 * plane + (size_t)(2 * y2) * W + 2 * x2;
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int py = 0; py < 2; ++py) {
                    const float2 q = __ldg(reinterpret_cast<const float2 *>(src + c * plane + py * W));
                    v[(py * 2 + 0) * 3 + c] = ptx::round_tf32(q.x);
                    v[(py * 2 + 1) * 3 + c] = ptx::round_tf32(q.y);
                }
        }
        v[12] = v[13] = v[14] = v[15] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            sS[t * 4 + (j ^ ((t >> 1) & 3))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                const int b = (py * 2 + px) * 3;
                sI[py][2 * t + px] = make_float4(v[b], v[b + 1], v[b + 2], 0.f);
            }
    }
    __syncthreads();
    const int npx = min(128, W2 - xb);                    // half-resolution pixels this block holds
    float4 *so = reinterpret_cast<float4 *>(s2d + ((size_t)blockIdx.y * W2 + xb) * 16);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int L = t + 128 * r, px = L >> 2, j = L & 3;
        if (px < npx) so[L] = sS[px * 4 + (j ^ ((px >> 1) & 3))];
    }
    // image slice: lane pairs write the 32 bytes (3 channels + 5 zeros) of one full-resolution pixel
#pragma unroll
    for (int py = 0; py < 2; ++py) {
        float *orow = out + (((size_t)n * H + 2 * y2 + py) * W + 2 * xb) * out_cs + out_co;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int L = t + 128 * r, px = L >> 1, half = L & 1;
            if (px < 2 * npx)
                *reinterpret_cast<float4 *>(orow + (size_t)px * out_cs + half * 4) =
                    half ? make_float4(0.f, 0.f, 0.f, 0.f) : sI[py][px];
        }
    }
}

int launch_s2d_pack(const void *in, int in_is_u8, const float *mean3, const float *std3, float *s2d, float *out, int b,
                    int H, int W, int out_cs, int out_co, cudaStream_t s)
{
    dim3 grid((unsigned)((W / 2 + 127) / 128), (unsigned)(b * (H / 2)));
    Norm3 nrm{};
    if (in_is_u8) {
        for (int c = 0; c < 3; ++c) {
            nrm.mean[c] = mean3[c];
            nrm.std[c] = std3[c];
        }
        k_s2d_pack<true><<<grid, 128, 0, s>>>(in, nrm, s2d, out, H, W, out_cs, out_co);
    } else {
        k_s2d_pack<false><<<grid, 128, 0, s>>>(in, nrm, s2d, out, H, W, out_cs, out_co);
    }
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
// w0*a + w1*b with lerp3's rounding sequence (fmul, then fma), four lanes
__device__ __forceinline__ float4 lerp2_4(float w0, const float4 &a, float w1, const float4 &b)
{
    return make_float4(__fmaf_rn(w1, b.x, __fmul_rn(w0, a.x)), __fmaf_rn(w1, b.y, __fmul_rn(w0, a.y)),
                       __fmaf_rn(w1, b.z, __fmul_rn(w0, a.z)), __fmaf_rn(w1, b.w, __fmul_rn(w0, a.w)));
}
// cvt.rna.tf32.f32 (nearest, ties away) as integer add + mask: identical for every finite input and infinity,
// two instructions per value instead of the four the conversion compiles to
__device__ __forceinline__ float4 round_tf32_4(const float4 &v)
{
    return make_float4(__uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xffffe000u),
                       __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xffffe000u),
                       __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xffffe000u),
                       __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xffffe000u));
}
__global__ void __launch_bounds__(256)
    k_upsample2x(const float *__restrict__ in, float *__restrict__ out, int h, int w, int C, int out_cs, int out_co,
                 float sy, float sx)
{
    // One thread = the 2x2 output block (2j..2j+1, 2k..2k+1) of 4-channel groups cg, cg+8, ...  With
    // align_corners=True and scale 2 the source rows of output rows 2j, 2j+1 all lie in
    // {j-1, j, j+1} (same for columns), so the block needs a 3x3 neighbourhood: 9 loads for 4
    // outputs instead of 16.  Each output keeps PyTorch's separable form
    // h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11); the third row/column enters with weight 0 (and a
    // clamped row/column index reproduces v10 = v00 at the border, as ATen's index clamp does).
    // The kernel was instruction-issue bound (ncu: 79 % issue-active, 31 instructions per output
    // float), so weights and offsets are computed once per thread and kept in 32-bit arithmetic.
    // blockDim = (8 channel groups, 32 pixel pairs); grid.y = image * h + j.
    const int c4 = C >> 2;
    const int k = blockIdx.x * blockDim.y + threadIdx.y;
    if (k >= w) return;
    const int n = blockIdx.y / h, j = blockIdx.y - n * h;

    float wy[2][3], wx[2][3];
    bool pat[2][2];                                     // [o][0] = uy, [o][1] = ux
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        const float fy = sy * (float)(2 * j + o), fx = sx * (float)(2 * k + o);
        const int y0 = (int)fy, x0 = (int)fx;
        const float h1 = fy - (float)y0, h0 = 1.f - h1, w1 = fx - (float)x0, w0 = 1.f - w1;
        const bool uy = y0 >= j, ux = x0 >= k;          // source pair is rows (j, j+1) rather than (j-1, j)
        pat[o][0] = uy;
        pat[o][1] = ux;
        wy[o][0] = uy ? 0.f : h0;
        wy[o][1] = uy ? h0 : h1;
        wy[o][2] = uy ? h1 : 0.f;
        wx[o][0] = ux ? 0.f : w0;
        wx[o][1] = ux ? w0 : w1;
        wx[o][2] = ux ? w1 : 0.f;
    }
    // element offsets of the 3x3 neighbourhood (clamped) and of the 2x2 outputs; 32-bit by contract
    unsigned ioff[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const unsigned y = (unsigned)min(max(j - 1 + r, 0), h - 1);
        const unsigned rowb = ((unsigned)n * h + y) * (unsigned)w;
#pragma unroll
        for (int c = 0; c < 3; ++c) ioff[r][c] = (rowb + (unsigned)min(max(k - 1 + c, 0), w - 1)) * (unsigned)C;
    }
    const unsigned Wo = 2u * w;
    const unsigned o00 = (((unsigned)n * 2u * h + 2u * j) * Wo + 2u * k) * (unsigned)out_cs + (unsigned)out_co;
    const unsigned orow = Wo * (unsigned)out_cs;

    // All blocks but those of the first row / column (and a last one whose scale*index rounds down) have the same
    // pattern: output 2j reads source rows (j-1, j), output 2j+1 rows (j, j+1), same for columns -- two-term sums (8 instead of 12
    // floating-point instructions per float4; the kernel is issue-bound).  lerp3 with its zero weight rounds
    // identically (fmul of the first product, fma of the second; adding 0*x changes nothing), so both paths and the
    // column kernel's fused loader agree to the last bit.
    const bool interior = !pat[0][0] && pat[1][0] && !pat[0][1] && pat[1][1];
    if (interior) {
        const float a0 = wx[0][0], a1 = wx[0][1], b0 = wx[1][1], b1 = wx[1][2];      // column weights of outputs 2k, 2k+1
        const float p0 = wy[0][0], p1 = wy[0][1], q0 = wy[1][1], q1 = wy[1][2];      // row weights of outputs 2j, 2j+1
        for (int cg = threadIdx.x; cg < c4; cg += 8) {
            const float *ip = in + cg * 4;
            float4 t[3][2];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(ip + ioff[r][0]));
                const float4 bq = __ldg(reinterpret_cast<const float4 *>(ip + ioff[r][1]));
                const float4 c = __ldg(reinterpret_cast<const float4 *>(ip + ioff[r][2]));
                t[r][0] = lerp2_4(a0, a, a1, bq);
                t[r][1] = lerp2_4(b0, bq, b1, c);
            }
            float *op = out + cg * 4;
#pragma unroll
            for (int ox = 0; ox < 2; ++ox) {
                *reinterpret_cast<float4 *>(op + o00 + (unsigned)ox * (unsigned)out_cs) = round_tf32_4(lerp2_4(p0, t[0][ox], p1, t[1][ox]));
                *reinterpret_cast<float4 *>(op + o00 + orow + (unsigned)ox * (unsigned)out_cs) = round_tf32_4(lerp2_4(q0, t[1][ox], q1, t[2][ox]));
            }
        }
        return;
    }
    for (int cg = threadIdx.x; cg < c4; cg += 8) {
        const float *ip = in + cg * 4;
        float4 t[3][2];                                   // per source row: the two horizontally interpolated values
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(ip + ioff[r][0]));
            const float4 bq = __ldg(reinterpret_cast<const float4 *>(ip + ioff[r][1]));
            const float4 c = __ldg(reinterpret_cast<const float4 *>(ip + ioff[r][2]));
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                t[r][o].x = lerp3(wx[o][0], a.x, wx[o][1], bq.x, wx[o][2], c.x);
                t[r][o].y = lerp3(wx[o][0], a.y, wx[o][1], bq.y, wx[o][2], c.y);
                t[r][o].z = lerp3(wx[o][0], a.z, wx[o][1], bq.z, wx[o][2], c.z);
                t[r][o].w = lerp3(wx[o][0], a.w, wx[o][1], bq.w, wx[o][2], c.w);
            }
        }
        float *op = out + cg * 4;
#pragma unroll
        for (int oy = 0; oy < 2; ++oy)
#pragma unroll
            for (int ox = 0; ox < 2; ++ox) {
                float4 v;
                v.x = lerp3(wy[oy][0], t[0][ox].x, wy[oy][1], t[1][ox].x, wy[oy][2], t[2][ox].x);
                v.y = lerp3(wy[oy][0], t[0][ox].y, wy[oy][1], t[1][ox].y, wy[oy][2], t[2][ox].y);
                v.z = lerp3(wy[oy][0], t[0][ox].z, wy[oy][1], t[1][ox].z, wy[oy][2], t[2][ox].z);
                v.w = lerp3(wy[oy][0], t[0][ox].w, wy[oy][1], t[1][ox].w, wy[oy][2], t[2][ox].w);
                *reinterpret_cast<float4 *>(op + o00 + (unsigned)oy * orow + (unsigned)ox * (unsigned)out_cs) = round_tf32_4(v);
            }
    }
}

int launch_upsample2x(const float *in, float *out, int b, int h, int w, int C, int out_cs, int out_co, cudaStream_t s)
{
    PV_CHECK_ARG((long long)b * 4 * h * w * out_cs < (1LL << 32) && (long long)b * h * w * C < (1LL << 32),
                 "upsample: tensor too large for 32-bit element offsets");
    const float sy = (float)(h - 1) / (float)(2 * h - 1), sx = (float)(w - 1) / (float)(2 * w - 1);
    dim3 grid((unsigned)((w + 31) / 32), (unsigned)(b * h));
    k_upsample2x<<<grid, dim3(8, 32), 0, s>>>(in, out, h, w, C, out_cs, out_co, sy, sx);
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
           long long total, int nhwc)
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
    float *o = nhwc ? out + i * Cout : out + n * Cout * (long long)npix + p;
    const long long ostride = nhwc ? 1 : npix;
    for (int co = 0; co < Cout; ++co) {
        float acc = sb[co];
        const float *wr = sw + co * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc = fmaf(v[j], wr[j], acc);
        o[(long long)co * ostride] = acc;
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
                int Cout, int b, int H, int W, int nhwc, cudaStream_t s)
{
    PV_CHECK_ARG(Cout >= 1 && Cout <= HEAD_MAX_COUT, "head: %d output channels unsupported (max %d)", Cout,
                 HEAD_MAX_COUT);
    const long long total = (long long)b * H * W;
    k_head<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, w, bias, out, mask, mask_esz, seg_dim, Cout, H * W,
                                                          total, nhwc);
    PV_LAUNCHED("k_head");
    return PVNET_OK;
}

}  // namespace pvnet
