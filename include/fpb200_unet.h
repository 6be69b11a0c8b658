/*
 * fpb200_unet - C ABI of the U-Net++ fingerprint segmenter INFERENCE: the reference's
 * `NestedUNet` (src/preprocessing/segmentation/model.py:26-83; wrapper FingerprintSegmentationModel :89-99) as used by
 * src/preprocessing/segmentation/inference.py:87-133 - an alternative source of the fingerprint mask.
 *
 * 3x3 convolutions (+ folded eval-mode BatchNorm + ReLU) run as implicit GEMMs on the 5th-generation tensor cores
 * (tcgen05.mma kind::tf32, accumulators in tensor memory) with the 3xTF32 error-compensated split, so the logits agree
 * with the fp32 PyTorch module to fp32 round-off; max-pooling, the align_corners bilinear up-sampling and the final 1x1
 * convolution are CUDA-core kernels.  Activations live in HBM as zero-bordered NHWC planes.  Inference only (eval mode).
 *
 * Same conventions as fpb200.h: 0 on success or a negative FPB_E_* code, `fpb_unet_last_error`, plain pointers and sizes.
 */
#ifndef FPB200_UNET_H
#define FPB200_UNET_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fpb_unet fpb_unet;

/* The 15 ConvBlocks of NestedUNet.__init__ (model.py:34-58), in this order; block i holds two 3x3 convolutions. */
enum {
    FPB_UNET_CONV0_0 = 0, FPB_UNET_CONV1_0, FPB_UNET_CONV2_0, FPB_UNET_CONV3_0, FPB_UNET_CONV4_0,
    FPB_UNET_UP1_0, FPB_UNET_UP2_0, FPB_UNET_UP3_0, FPB_UNET_UP1_1, FPB_UNET_UP2_1, FPB_UNET_UP1_2,
    FPB_UNET_NUM_BLOCKS
};

/* height, width: multiples of 16 (four 2x2 poolings), 16..1024; input_channels as NestedUNet(input_channels=3) */
int  fpb_unet_create(fpb_unet** out, int device, int max_batch, int height, int width, int input_channels);
void fpb_unet_destroy(fpb_unet* u);
const char* fpb_unet_last_error(const fpb_unet* u);   /* u may be NULL: error of fpb_unet_create */
/* in / out channels of convolution `conv` (0 or 1) of block `block`, as the reference's state_dict has them */
int  fpb_unet_conv_shape(const fpb_unet* u, int block, int conv, int* in_channels, int* out_channels);

/* Parameters of one convolution of a ConvBlock (`conv.0`/`conv.1` for conv 0, `conv.3`/`conv.4` for conv 1 in the
 * reference's state_dict): weight [OC][IC][3][3], bias [OC], BatchNorm weight / bias / running_mean / running_var [OC], eps. */
int  fpb_unet_set_conv(fpb_unet* u, int block, int conv, const float* weight, const float* bias, const float* bn_weight,
                       const float* bn_bias, const float* bn_mean, const float* bn_var, double bn_eps);
/* `final` 1x1 convolution (model.py:61): weight [num_labels=1][64], bias [1] */
int  fpb_unet_set_final(fpb_unet* u, const float* weight, const float* bias);

/* forward (model.py:63-83): input float32 [n][input_channels][H][W] (host) -> logits float32 [n][H][W] (host).
 * conv4_0 is loaded but not evaluated: the reference computes x4_0 and never uses it (model.py:69). */
int  fpb_unet_forward(fpb_unet* u, const float* input, int n, float* logits);
/* number of kernel launches of the last forward, and how many of them were tensor-core convolutions */
int  fpb_unet_launches(const fpb_unet* u, int* total, int* tensor_core);

#ifdef __cplusplus
}
#endif
#endif /* FPB200_UNET_H */
