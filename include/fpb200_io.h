/*
 * fpb200_io - C ABI of the on-disk hand-offs either side of the hot path.
 *
 *   fpb_jpeg_info / fpb_decode_jpeg_batch   replace cv2.imread(path, cv2.IMREAD_GRAYSCALE) at
 *       src/preprocessing/run_preprocessing.py:38-47 (and src/features/extract_features.py:83) for baseline JPEG
 *       files: Huffman decoding on host threads, dequantisation + libjpeg's "islow" inverse DCT + range limiting on
 *       the device, straight into the handle's input plane.  Pixels are bit-identical to cv2.imread's.
 *   fpb_run_decoded                         the hot path (fpb_run_device) on the plane fpb_decode_jpeg_batch filled
 *   fpb_minutiae_json / fpb_write_minutiae_json_batch   replace json.dump(refined, f, indent=2) at
 *       src/features/extract_features.py:104-105, byte for byte (Python float repr included).
 *
 * Status codes: FPB_OK / FPB_E_* of fpb200.h plus the three below.
 */
#ifndef FPB200_IO_H
#define FPB200_IO_H

#include "fpb200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define FPB_E_JPEG_FORMAT      -10   /* not a JPEG / corrupt stream                                         */
#define FPB_E_JPEG_UNSUPPORTED -11   /* progressive, arithmetic, 12-bit, RGB/CMYK, EXIF: decode with cv2    */
#define FPB_E_JPEG_SHAPE       -12   /* dimensions differ from the handle's H x W                           */

/* width / height / number of components of a JPEG in memory (no GPU needed) */
int fpb_jpeg_info(const uint8_t* buf, size_t size, int* width, int* height, int* components);

/* host half of the decoder alone (no GPU needed): luminance DCT coefficients [ceil(H/8)][ceil(W/8)][64] int16 in
 * natural order and the luminance quantisation table [64] of a JPEG whose size is width x height */
int fpb_jpeg_coefficients(const uint8_t* buf, size_t size, int width, int height, int16_t* coefs, uint16_t* qt);

/* n JPEG files in memory -> images 0..n-1 of the handle's device input plane.  `threads` host threads do the entropy
 * decoding (<= 0: one per core).  status[i] = FPB_OK or the reason image i could not be decoded here (its plane is
 * left zero; decode it with cv2 and use fpb_run_host instead).  Returns the number of images decoded, or < 0. */
int fpb_decode_jpeg_batch(fpb_handle* h, const uint8_t* const* bufs, const size_t* sizes, int n, int threads, int32_t* status);
/* copy the decoded input plane to the host: n*H*W bytes (parity tests against cv2.imread) */
int fpb_fetch_input(fpb_handle* h, uint8_t* dst, int n);
/* device address of the handle's input plane (max_batch*H*W bytes): what fpb_decode_jpeg_batch / fpb_synth_ridge fill; pass it
 * to fpb_run_device for an ASYNCHRONOUS run (streaming loops: fpb_download_results of the previous batch meanwhile) */
const void* fpb_input_plane(const fpb_handle* h);
/* run K1..K9 on the first n images of the input plane and download roi / counts / refined minutiae */
int fpb_run_decoded(fpb_handle* h, int n);

/* Synthetic input generated ON THE DEVICE (BASELINE configs[3]): images first_index .. first_index+n-1 of
 * the stream `seed` are written to the handle's input plane - the ridge formula of synth.ridge_image with a counter-based
 * Philox4x32-10 generator, so any image can be regenerated alone.  period <= 0: drawn per image from U[7, 11].  Follow with
 * fpb_run_decoded (the hot path on the input plane); fpb_fetch_input returns the pixels.  Asynchronous. */
int fpb_synth_ridge(fpb_handle* h, uint64_t seed, uint64_t first_index, int n, double period, double noise_sigma);

/* the skeleton hand-off as a stage: out = cv2.imread(cv2.imwrite(img as JPEG, quality 95), IMREAD_GRAYSCALE) for n
 * host images of the handle's H x W (run_preprocessing.py:137-140 -> extract_features.py:83), bit-identical */
int fpb_jpeg_roundtrip(fpb_handle* h, const uint8_t* img, int n, uint8_t* out);

/* the text json.dump(list, f, indent=2) writes for n refined minutiae.  Returns the length (without the NUL);
 * writes at most cap bytes (NUL-terminated when cap > length). */
long long fpb_minutiae_json(const fpb_minutia* m, int n, char* buf, size_t cap);
/* refined minutiae of image i of the last run -> paths[i] (written to "<path>.tmp", then renamed), `threads` writers.
 * Returns the number of files written, or < 0. */
int fpb_write_minutiae_json_batch(fpb_handle* h, const char* const* paths, int n, int threads);

#ifdef __cplusplus
}
#endif
#endif /* FPB200_IO_H */
