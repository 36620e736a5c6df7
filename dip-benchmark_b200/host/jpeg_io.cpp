#include "jpeg_io.h"

#include <cuda_runtime.h>
#include <nvjpeg.h>

struct JpegIo {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    nvjpegEncoderState_t enc_state = nullptr;
    nvjpegEncoderParams_t enc_params = nullptr;
};

static bool ok(nvjpegStatus_t s, const char* what, std::string* err) {
    if (s == NVJPEG_STATUS_SUCCESS) return true;
    if (err) *err = std::string(what) + " failed with nvjpeg status " + std::to_string((int)s);
    return false;
}

JpegIo* jpeg_io_create(std::string* err) {
    JpegIo* io = new JpegIo();
    if (!ok(nvjpegCreateSimple(&io->handle), "nvjpegCreateSimple", err) ||
        !ok(nvjpegJpegStateCreate(io->handle, &io->state), "nvjpegJpegStateCreate", err) ||
        !ok(nvjpegEncoderStateCreate(io->handle, &io->enc_state, nullptr), "nvjpegEncoderStateCreate", err) ||
        !ok(nvjpegEncoderParamsCreate(io->handle, &io->enc_params, nullptr), "nvjpegEncoderParamsCreate", err) ||
        !ok(nvjpegEncoderParamsSetQuality(io->enc_params, 95, nullptr), "nvjpegEncoderParamsSetQuality", err) ||
        !ok(nvjpegEncoderParamsSetSamplingFactors(io->enc_params, NVJPEG_CSS_420, nullptr), "nvjpegEncoderParamsSetSamplingFactors", err) ||
        !ok(nvjpegEncoderParamsSetOptimizedHuffman(io->enc_params, 0, nullptr), "nvjpegEncoderParamsSetOptimizedHuffman", err)) {
        jpeg_io_destroy(io);
        return nullptr;
    }
    return io;
}

void jpeg_io_destroy(JpegIo* io) {
    if (!io) return;
    if (io->enc_params) nvjpegEncoderParamsDestroy(io->enc_params);
    if (io->enc_state) nvjpegEncoderStateDestroy(io->enc_state);
    if (io->state) nvjpegJpegStateDestroy(io->state);
    if (io->handle) nvjpegDestroy(io->handle);
    delete io;
}

bool jpeg_io_info(JpegIo* io, const std::vector<uint8_t>& data, int* width, int* height, std::string* err) {
    int ncomp = 0;
    nvjpegChromaSubsampling_t ss;
    int w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0};
    if (!ok(nvjpegGetImageInfo(io->handle, data.data(), data.size(), &ncomp, &ss, w, h), "nvjpegGetImageInfo", err)) return false;
    *width = w[0];
    *height = h[0];
    return w[0] > 0 && h[0] > 0;
}

bool jpeg_io_decode(JpegIo* io, const std::vector<uint8_t>& data, void* device_ptr, size_t pitch, int, int, std::string* err) {
    nvjpegImage_t img{};
    img.channel[0] = static_cast<unsigned char*>(device_ptr);
    img.pitch[0] = pitch;
    if (!ok(nvjpegDecode(io->handle, io->state, data.data(), data.size(), NVJPEG_OUTPUT_BGRI, &img, nullptr), "nvjpegDecode", err)) return false;
    return cudaStreamSynchronize(nullptr) == cudaSuccess;
}

bool jpeg_io_encode(JpegIo* io, const void* device_ptr, size_t pitch, int width, int height, std::vector<uint8_t>* out, std::string* err) {
    nvjpegImage_t img{};
    img.channel[0] = static_cast<unsigned char*>(const_cast<void*>(device_ptr));
    img.pitch[0] = pitch;
    if (!ok(nvjpegEncodeImage(io->handle, io->enc_state, io->enc_params, &img, NVJPEG_INPUT_BGRI, width, height, nullptr), "nvjpegEncodeImage", err))
        return false;
    size_t len = 0;
    if (!ok(nvjpegEncodeRetrieveBitstream(io->handle, io->enc_state, nullptr, &len, nullptr), "nvjpegEncodeRetrieveBitstream", err)) return false;
    out->resize(len);
    if (!ok(nvjpegEncodeRetrieveBitstream(io->handle, io->enc_state, out->data(), &len, nullptr), "nvjpegEncodeRetrieveBitstream", err)) return false;
    out->resize(len);
    return cudaStreamSynchronize(nullptr) == cudaSuccess;
}
