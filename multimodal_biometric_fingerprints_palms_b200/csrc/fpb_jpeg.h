// Internal declarations of k_jpeg.cu (JPEG entropy decoder on the host, IDCT on the device, JSON writer).
#pragma once
#include <string>
#include "../../include/fpb200.h"
#include "fpb_kernels.h"

#define FPB_JPEG_E_FORMAT      -10   /* not a JPEG / corrupt stream                       */
#define FPB_JPEG_E_UNSUPPORTED -11   /* progressive, arithmetic, 12-bit, CMYK/RGB, EXIF   */
#define FPB_JPEG_E_SHAPE       -12   /* dimensions differ from the handle's H x W         */

struct FpbJpegInfo { int width, height, components; };
int fpb_jpeg_parse(const uint8_t* buf, size_t size, FpbJpegInfo* info);
int fpb_jpeg_entropy_decode(const uint8_t* buf, size_t size, int W, int H, int16_t* coefs, uint16_t* qt);
void fpb_jpeg_idct(FpbLaunch L, const int16_t* coefs, const uint16_t* qts, int n, int W, int H, uint8_t* dst);
// skeleton hand-off of the reference: dst = cv2.imread(cv2.imwrite(src as .jpg, quality 95)) per image crop
void fpb_jpeg_roundtrip_q95(FpbLaunch L, const uint8_t* src, int n, int W, int H, const int4* roi, uint8_t* dst);
void fpb_minutiae_json_string(const fpb_minutia* m, int n, std::string& out);
