/* The reference's ffmpeg filter includes "transform360/VideoFrameTransformHandler.h" (vf_transform360.c:27; its README
 * tells users to fix the path case by hand).  This forwarding header makes both spellings work. */
#include "../Transform360/VideoFrameTransformHandler.h"
