// conv_tc.cuh -- internal interface of the tensor-core convolution (conv_tc.cu) and the
// auxiliary backbone kernels (backbone_aux.cu), shared with backbone.cu.
#pragma once
#include "common.cuh"

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
};

// A plan = encoded tensor maps + launch geometry; opaque bytes so callers can cache it.
size_t conv_plan_size();
int conv_plan_at(const ConvDesc &d, void *plan_storage);
int conv_launch_at(const void *plan_storage, cudaStream_t s);

int launch_stem(const float *in, const float *w, const float *bias, float *out, int b, int H, int W, int out_cs,
                int out_co, cudaStream_t s);
int launch_pack_image(const float *in, float *out, int b, int H, int W, int out_cs, int out_co, cudaStream_t s);
int launch_maxpool(const float *in, float *out, int b, int H, int W, int C, int in_cs, int in_co, cudaStream_t s);
int launch_upsample2x(const float *in, float *out, int b, int h, int w, int C, int out_cs, int out_co,
                      cudaStream_t s);
int launch_head(const float *in, const float *w, const float *bias, float *out, void *mask, int mask_esz,
                int seg_dim, int Cout, int b, int H, int W, cudaStream_t s);

}  // namespace pvnet
