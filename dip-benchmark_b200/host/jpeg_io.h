// JPEG input/output for the native harness through nvJPEG (CUDA toolkit library, not part of the hot path).
// The reference harnesses read and write the image with OpenCV (sycl/benchmark.cpp:445,:374); OpenCV's C++
// library is not in this image, nvJPEG is.  Decoding lands directly in the device image (no raw-pixel upload)
// and encoding reads the device image (only the compressed bitstream crosses PCIe) -- SURVEY.md 8(f) rank 2.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

struct JpegIo;
JpegIo* jpeg_io_create(std::string* err);
void jpeg_io_destroy(JpegIo* io);
// width / height of a JPEG byte stream (components are always delivered as 3-channel BGR, like cv::imread)
bool jpeg_io_info(JpegIo* io, const std::vector<uint8_t>& data, int* width, int* height, std::string* err);
// decode into device memory: interleaved BGR, rows `pitch` bytes apart
bool jpeg_io_decode(JpegIo* io, const std::vector<uint8_t>& data, void* device_ptr, size_t pitch, int width, int height, std::string* err);
// encode interleaved BGR device memory (quality 95, 4:2:0 chroma: OpenCV's imwrite defaults)
bool jpeg_io_encode(JpegIo* io, const void* device_ptr, size_t pitch, int width, int height, std::vector<uint8_t>* out, std::string* err);
