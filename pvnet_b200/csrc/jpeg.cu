// jpeg.cu -- JPEG bytes -> uint8 HWC images on the device, the front of
// pvnet_backbone_forward_u8.  The reference decodes on the host with PIL (`Image.open`,
// lib/datasets/linemod_dataset.py:180-195; tools/demo.py:89) and uploads float32 NCHW; here the compressed
// bytes go to NVIDIA's nvJPEG (plain library use, like cuBLAS for a library GEMM: Huffman decode + IDCT +
// colour conversion are not rewritten) and its interleaved-RGB output IS the [b,h,w,3] tensor the packing
// kernel normalises.  libnvjpeg is dlopen'ed on first use, so libpvnet_b200.so has no link-time dependency
// on it and every other entry point works without it.
//
// Parity: nvJPEG and libjpeg(-turbo, behind PIL) are different decoders -- IDCT rounding differs by +-1 level,
// and for chroma-subsampled files libjpeg's "fancy" triangle upsampling differs from nvJPEG's at colour edges.
// Bit parity of the hot path therefore starts at the decoded uint8 image; tests/test_gpu_jpeg.py measures the
// decoder-to-decoder difference.
#include "common.cuh"

#include <dlfcn.h>
#include <nvjpeg.h>

#include <mutex>
#include <vector>

namespace {

struct NvJpegApi {
    void *lib = nullptr;
    nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t *) = nullptr;
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
    nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t *) = nullptr;
    nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
    nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char *, size_t, int *, nvjpegChromaSubsampling_t *, int *,
                                   int *) = nullptr;
    nvjpegStatus_t (*DecodeBatchedInitialize)(nvjpegHandle_t, nvjpegJpegState_t, int, int, nvjpegOutputFormat_t) = nullptr;
    nvjpegStatus_t (*DecodeBatched)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char *const *, const size_t *,
                                    nvjpegImage_t *, cudaStream_t) = nullptr;
    bool ok = false;
};

NvJpegApi &api()
{
    static NvJpegApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so.12", "libnvjpeg.so",
                               "/usr/local/cuda/lib64/libnvjpeg.so"};
        for (const char *n : names)
            if ((a.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
        if (!a.lib) return;
#define LOAD(field, sym) *(void **)(&a.field) = dlsym(a.lib, sym)
        LOAD(CreateSimple, "nvjpegCreateSimple");
        LOAD(Destroy, "nvjpegDestroy");
        LOAD(JpegStateCreate, "nvjpegJpegStateCreate");
        LOAD(JpegStateDestroy, "nvjpegJpegStateDestroy");
        LOAD(GetImageInfo, "nvjpegGetImageInfo");
        LOAD(DecodeBatchedInitialize, "nvjpegDecodeBatchedInitialize");
        LOAD(DecodeBatched, "nvjpegDecodeBatched");
#undef LOAD
        a.ok = a.CreateSimple && a.Destroy && a.JpegStateCreate && a.JpegStateDestroy && a.GetImageInfo &&
               a.DecodeBatchedInitialize && a.DecodeBatched;
    });
    return a;
}

}  // namespace

struct pvnet_jpeg_decoder {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    int batch = 0;       // batch size nvjpegDecodeBatchedInitialize was last called with
};

extern "C" {

int pvnet_jpeg_available(void) { return api().ok ? 1 : 0; }

int pvnet_jpeg_decoder_create(pvnet_jpeg_decoder_t **out)
{
    PV_CHECK_ARG(out, "null out pointer");
    NvJpegApi &a = api();
    if (!a.ok) {
        pvnet::set_error("libnvjpeg.so.12 could not be loaded (dlopen): JPEG decoding is unavailable");
        return PVNET_E_STATE;
    }
    pvnet_jpeg_decoder *d = new pvnet_jpeg_decoder();
    nvjpegStatus_t st = a.CreateSimple(&d->handle);
    if (st == NVJPEG_STATUS_SUCCESS) st = a.JpegStateCreate(d->handle, &d->state);
    if (st != NVJPEG_STATUS_SUCCESS) {
        pvnet::set_error("nvjpeg initialisation failed (status %d)", (int)st);
        if (d->handle) a.Destroy(d->handle);
        delete d;
        return PVNET_E_CUDA;
    }
    *out = d;
    return PVNET_OK;
}

void pvnet_jpeg_decoder_destroy(pvnet_jpeg_decoder_t *d)
{
    if (!d) return;
    NvJpegApi &a = api();
    if (a.ok) {
        if (d->state) a.JpegStateDestroy(d->state);
        if (d->handle) a.Destroy(d->handle);
    }
    delete d;
}

int pvnet_jpeg_decode_batch(pvnet_jpeg_decoder_t *d, const uint8_t *const *jpeg_data, const size_t *lengths, int b, int h,
                            int w, uint8_t *out_hwc, pvnet_stream_t stream)
{
    PV_CHECK_ARG(d && jpeg_data && lengths && out_hwc, "null pointer");
    PV_CHECK_ARG(b >= 1 && h >= 1 && w >= 1, "non-positive dimension");
    NvJpegApi &a = api();
    for (int i = 0; i < b; ++i) {
        PV_CHECK_ARG(jpeg_data[i] && lengths[i] > 0, "image %d: empty JPEG buffer", i);
        int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
        nvjpegChromaSubsampling_t sub;
        const nvjpegStatus_t st = a.GetImageInfo(d->handle, jpeg_data[i], lengths[i], &ncomp, &sub, ws, hs);
        if (st != NVJPEG_STATUS_SUCCESS) {
            pvnet::set_error("image %d: not a decodable JPEG (nvjpeg status %d)", i, (int)st);
            return PVNET_E_INVALID;
        }
        PV_CHECK_ARG(ws[0] == w && hs[0] == h, "image %d is %dx%d, expected %dx%d", i, ws[0], hs[0], w, h);
    }
    if (d->batch != b) {
        const nvjpegStatus_t st = a.DecodeBatchedInitialize(d->handle, d->state, b, 1, NVJPEG_OUTPUT_RGBI);
        if (st != NVJPEG_STATUS_SUCCESS) {
            pvnet::set_error("nvjpegDecodeBatchedInitialize failed (status %d)", (int)st);
            return PVNET_E_CUDA;
        }
        d->batch = b;
    }
    std::vector<nvjpegImage_t> dst((size_t)b);
    for (int i = 0; i < b; ++i) {
        for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) {
            dst[i].channel[c] = nullptr;
            dst[i].pitch[c] = 0;
        }
        dst[i].channel[0] = out_hwc + (size_t)i * h * w * 3;      // interleaved RGB = the [h,w,3] image itself
        dst[i].pitch[0] = (size_t)w * 3;
    }
    const nvjpegStatus_t st = a.DecodeBatched(d->handle, d->state, jpeg_data, lengths, dst.data(), (cudaStream_t)stream);
    if (st != NVJPEG_STATUS_SUCCESS) {
        pvnet::set_error("nvjpegDecodeBatched failed (status %d)", (int)st);
        return PVNET_E_CUDA;
    }
    return PVNET_OK;
}

}  // extern "C"
