/* transform360-b200: ABI types of the drop-in boundary.
 *
 * Binary-compatible with the reference's Transform360/Library/VideoFrameTransformHelper.h
 * (enums :18-54, FrameTransformContext :56-90): same enumerator names and values, same
 * 28 x 4-byte parameter block (112 bytes, natural alignment), so a caller compiled against
 * the reference header (e.g. vf_transform360.c:111-139) links against this library unchanged.
 * The reference's FACEBOOK_LAYOUT enumerator is not published and is not part of this ABI.
 */
#ifndef TRANSFORM360_B200_VIDEOFRAMETRANSFORMHELPER_H
#define TRANSFORM360_B200_VIDEOFRAMETRANSFORMHELPER_H

#include <stdint.h>

/* cube faces in the order the 3x2 layouts place them (top row R L T, bottom row Bo F Ba) */
typedef enum TransformFaceType { RIGHT = 0, LEFT = 1, TOP = 2, BOTTOM = 3, FRONT = 4, BACK = 5 } TransformFaceType;

typedef enum Layout {
  LAYOUT_CUBEMAP_32 = 0,           /* 3 x 2 cube faces                         */
  LAYOUT_CUBEMAP_23_OFFCENTER = 1, /* 2 x 3 cube faces, off-centre projection  */
  LAYOUT_FLAT_FIXED = 2,           /* fixed-FOV window of an equirect frame    */
  LAYOUT_EQUIRECT = 3,
  LAYOUT_BARREL = 4,
  LAYOUT_BARREL_SPLIT = 5,
  LAYOUT_EAC_32 = 6, /* equi-angular 3 x 2 cube */
  LAYOUT_N = 7
} Layout;

typedef enum StereoFormat {
  STEREO_FORMAT_TB = 0,
  STEREO_FORMAT_LR = 1,
  STEREO_FORMAT_MONO = 2,
  STEREO_FORMAT_GUESS = 3,
  STEREO_FORMAT_N = 4
} StereoFormat;

/* numerically equal to OpenCV's INTER_* flags, which is what the reference hands to cv::remap */
typedef enum InterpolationAlg { NEAREST = 0, LINEAR = 1, CUBIC = 2, LANCZOS4 = 4 } InterpolationAlg;

typedef struct FrameTransformContext {
  /* projection */
  Layout input_layout;
  Layout output_layout;
  StereoFormat input_stereo_format;
  StereoFormat output_stereo_format;
  int vflip;               /* flip the second eye of a TB output vertically        */
  float input_expand_coef; /* face expansion of a cubemap INPUT                    */
  float expand_coef;       /* face expansion of the output (1.01 by default)       */
  InterpolationAlg interpolation_alg;
  float width_scale_factor;  /* render at scale x output size, then area-downscale */
  float height_scale_factor;
  float fixed_yaw;   /* degrees */
  float fixed_pitch; /* degrees */
  float fixed_roll;  /* degrees */
  float fixed_hfov;  /* degrees, LAYOUT_FLAT_FIXED */
  float fixed_vfov;  /* degrees, LAYOUT_FLAT_FIXED */
  float fixed_cube_offcenter_x;
  float fixed_cube_offcenter_y;
  float fixed_cube_offcenter_z;
  int is_horizontal_offset;
  /* segmented low-pass (anti-alias) filter */
  int enable_low_pass_filter;
  float kernel_height_scale_factor;
  float min_kernel_half_height;
  float max_kernel_half_height;
  int enable_multi_threading; /* reference: one host thread per tile; here: ignored, the GPU does all tiles */
  int num_vertical_segments;
  int num_horizontal_segments;
  int adjust_kernel;
  float kernel_adjust_factor;
} FrameTransformContext;

#if defined(__cplusplus)
static_assert(sizeof(FrameTransformContext) == 112, "FrameTransformContext must stay 28 x 4 bytes");
#elif defined(__STDC_VERSION__) && __STDC_VERSION__ >= 201112L
_Static_assert(sizeof(FrameTransformContext) == 112, "FrameTransformContext must stay 28 x 4 bytes");
#endif

#endif /* TRANSFORM360_B200_VIDEOFRAMETRANSFORMHELPER_H */
