// conv_tc.cuh -- internal interface of the tensor-core convolution (conv_tc.cu) and the
// auxiliary backbone kernels (backbone_aux.cu), shared with backbone.cu.
#pragma once
#include "common.cuh"

#include <cuda.h>
#ifdef __CUDACC__
#include "ptx.cuh"
#endif

namespace pvnet {

struct ConvDesc {
    const float *in;   // NHWC buffer [b,H,W,in_cs]; channels [in_co, in_co+Cin) are the conv input
    int in_cs, in_co, Cin;
    const float *w;    // packed [Cout][taps][Cin]
    const float *bias; // [Cout]
    const float *res;  // NHWC [b,Ho,Wo,res_cs] at res_co, or null
    int res_cs, res_co;
    float *out;        // NHWC [b,Ho,Wo,out_cs], written at out_co
    int out_cs, out_co, Cout;
    int b, H, W;
    int ksize, stride, dilation;
    int act, round_out;
    // optional second source (column kernel only): channels [Cin, Cin+Cin2) of the convolution come
    // from in2 [b,H,W,in2_cs] at in2_co -- a torch.cat on the READ side, so that both producers of a
    // concatenated input write dense records of their own
    const float *in2 = nullptr;
    int in2_cs = 0, in2_co = 0, Cin2 = 0;
    // optional fused upsampling (column kernel with the fused head only): channels [0, Cin) are NOT read
    // from `in` but are F.interpolate(up_src, scale 2, bilinear, align_corners=True) of the dense
    // half-resolution tensor up_src [b,H/2,W/2,Cin], interpolated inside the kernel straight into the
    // operand stages (model_repository.py:75: the upsampled tensor is never written)
    const float *up_src = nullptr;
    int up_mode = 1;   // 1: interpolated by the epilogue warps (two CTAs per SM); 2: by eight dedicated warps (one CTA per SM)
};

// Optional fused 1x1 head (convraw.3 + argmax) for the column kernel's epilogue.
struct HeadDesc {
    const float *w;      // [cout][32] fp32
    const float *bias;   // [cout]
    float *out_nchw;     // [b,cout,H,W]
    void *mask;          // [b,H,W] int64 / u8, or null
    int mask_esz, seg_dim, cout;
};

// cuTensorMapEncodeTiled wrapper (fp32 elements); swizzle_bytes in {128,64,32}
int tma_encode(CUtensorMap *m, const void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides_bytes,
               const cuuint32_t *box, int swizzle_bytes);

// conv mode override for tests: 0 auto, 1 force per-tap kernel, 2 force column kernel
extern int g_conv_mode;
bool conv_col_eligible(const ConvDesc &d);
size_t conv_col_plan_size();
int conv_col_plan_at(const ConvDesc &d, const HeadDesc *head, void *plan_storage);
int conv_col_launch_at(const void *plan_storage, cudaStream_t s);
void conv_col_set_head_ptrs(void *plan_storage, float *out, void *mask, int mask_esz, int nhwc);

// A plan = encoded tensor maps + launch geometry; opaque bytes so callers can cache it.
size_t conv_plan_size();
int conv_plan_at(const ConvDesc &d, void *plan_storage);
int conv_launch_at(const void *plan_storage, cudaStream_t s);

int launch_stem(const float *in, const float *w, const float *bias, float *out, int b, int H, int W, int out_cs,
                int out_co, cudaStream_t s);
int launch_pack_image(const float *in, float *out, int b, int H, int W, int out_cs, int out_co, cudaStream_t s);
int launch_s2d_pack(const void *in, int in_is_u8, const float *mean3, const float *std3, float *s2d, float *out, int b,
                    int H, int W, int out_cs, int out_co, cudaStream_t s);
int launch_maxpool(const float *in, float *out, int b, int H, int W, int C, int in_cs, int in_co, cudaStream_t s);
int launch_upsample2x(const float *in, float *out, int b, int h, int w, int C, int out_cs, int out_co,
                      cudaStream_t s);
int launch_head(const float *in, const float *w, const float *bias, float *out, void *mask, int mask_esz,
                int seg_dim, int Cout, int b, int H, int W, int nhwc, cudaStream_t s);

#ifdef __CUDACC__
// Three-slot interpolation sum with a FIXED rounding sequence (one weight of the window is zero): both
// upsampling implementations -- k_upsample2x and the column kernel's fused loader -- go through it, so they
// agree to the last bit whatever contraction the compiler would have picked for `a*b + c*d + e*f`.
__device__ __forceinline__ float lerp3(float w0, float a, float w1, float b, float w2, float c)
{
    return __fmaf_rn(w2, c, __fmaf_rn(w1, b, __fmul_rn(w0, a)));
}
// Coalesced residual fetch shared by the conv epilogues (128-pixel tiles TW pixels wide, one pixel
// per thread, epilogue warp q owns tile rows q*32/TW ...).  Load i of lane l reads the 16-byte chunk
// (l % 8) of warp-pixel 4*i + l/8, i.e. four full 128-byte lines per instruction instead of 32
// strided sectors.  `base` points at (tile origin, first channel of the 32-channel group);
// out-of-image pixels read nothing.
template <int TW>
__device__ __forceinline__ void res_fetch8(float4 (&dst)[8], const float *base, int q, int lane, int y0, int x0,
                                           int Ho, int Wo, int res_cs)
{
    static_assert(TW == 8 || TW == 16, "tile width");
    constexpr int ROWS = 32 / TW;                       // tile rows per warp
    const int yy = y0 + ROWS * q, xx = x0 + (lane >> 3);
    const float *p = base + ((size_t)(ROWS * q) * Wo + (lane >> 3)) * res_cs + (lane & 7) * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int dy = (4 * i) / TW, dx = (4 * i) % TW;
        dst[i] = (yy + dy < Ho && xx + dx < Wo)
                     ? __ldg(reinterpret_cast<const float4 *>(p + ((size_t)dy * Wo + dx) * res_cs))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
// Epilogue tail shared by the conv kernels: v[32] (one pixel, 32 channels, bias already added)
// -> (+ residual) -> activation -> optional tf32 rounding -> this warp's 32 rows of a 128-byte-
// swizzled staging buffer -> ONE TMA box per warp (32 channels x 32 pixels).  Each epilogue warp owns
// two 4 KB buffers and its own bulk-store groups, so no CTA-wide barrier is needed: lane 0 waits
// until the store issued two chunks ago has finished reading its buffer, the warp writes, fences the
// async proxy and lane 0 issues the next store.  `buf` is the shared-space address of the buffer to
// use now; row r of it holds pixel r of the warp, 16-byte chunk j at position j ^ (r & 7).
template <int ACT, bool ROUND>
__device__ __forceinline__ void epi_activate(float (&v)[32])
{
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (ACT == 1) v[j] = fmaxf(v[j], 0.f);
        if (ACT == 2) v[j] = fmaxf(v[j], 0.1f * v[j]);      // LeakyReLU(0.1): max(x, 0.1x)
        if (ROUND) v[j] = ptx::round_tf32(v[j]);
    }
}
__device__ __forceinline__ void epi_activate(float (&v)[32], int act, int round_out)
{
    // one warp-uniform branch per chunk instead of per-element predicates
    if (round_out) {
        if (act == 1) epi_activate<1, true>(v);
        else if (act == 2) epi_activate<2, true>(v);
        else epi_activate<0, true>(v);
    } else {
        if (act == 1) epi_activate<1, false>(v);
        else if (act == 2) epi_activate<2, false>(v);
    }
}
// residual: rpre (coalesced fetch layout, see res_fetch8) -> rows of `buf` -> added to the owning lane's v
__device__ __forceinline__ void epi_add_residual(float (&v)[32], const float4 (&rpre)[8], uint32_t buf, int lane)
{
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + (lane >> 3);
        ptx::sts128(buf + (uint32_t)(row * 128 + (((lane & 7) ^ (row & 7)) << 4)), rpre[i]);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 rv = ptx::lds128(buf + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4)));
        v[4 * j] += rv.x;
        v[4 * j + 1] += rv.y;
        v[4 * j + 2] += rv.z;
        v[4 * j + 3] += rv.w;
    }
    __syncwarp();
}
__device__ __forceinline__ void epi_stage(const float (&v)[32], uint32_t buf, int lane)
{
#pragma unroll
    for (int j = 0; j < 8; ++j)
        ptx::sts128(buf + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4)),
               make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
}
#endif

}  // namespace pvnet
