/* see VideoFrameTransformHandler.h in this directory (vf_transform360.c:28) */
#include "../Transform360/VideoFrameTransformHelper.h"
