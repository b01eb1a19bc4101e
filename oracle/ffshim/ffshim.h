/* oracle/ffshim/ffshim.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A stand-in for the slice of libavfilter / libavutil that the reference's Transform360/vf_transform360.c touches,
 * so that the UNMODIFIED filter source compiles here (no ffmpeg tree exists in this image) and can be driven by
 * oracle/ff_driver.c: option table -> TransformContext, config_output, filter_frame.  The same filter object is
 * linked once against the reference library (oracle/_ref/libt360ref.so) and once against the product
 * (transform360_b200/lib/libTransform360.so): identical frames from both is the filter-level drop-in test.
 */
#ifndef T360_FFSHIM_H
#define T360_FFSHIM_H

#include <errno.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define av_cold
#define AVERROR(e) (-(e))
#define FFSWAP(type, a, b) do { type SWAP_tmp = b; b = a; a = SWAP_tmp; } while (0)
#define FF_CEIL_RSHIFT(a, b) (-((-(a)) >> (b)))
#define NULL_IF_CONFIG_SMALL(x) x
#define LIBAVUTIL_VERSION_INT 0
#define av_assert1(cond) ((void)0)
#define AV_LOG_ERROR 16
#define AV_LOG_INFO 32
#define AV_LOG_VERBOSE 40
#define AV_OPT_FLAG_VIDEO_PARAM 16
#define AV_OPT_FLAG_FILTERING_PARAM (1 << 16)
#define AVERROR_EXTERNAL (-0x20545845)
#define FFALIGN(x, a) (((x) + (a) - 1) & ~((a) - 1))
#define FF_FILTER_FLAG_HWFRAME_AWARE 1

enum AVOptionType { AV_OPT_TYPE_FLAGS, AV_OPT_TYPE_INT, AV_OPT_TYPE_INT64, AV_OPT_TYPE_DOUBLE, AV_OPT_TYPE_FLOAT,
                    AV_OPT_TYPE_STRING, AV_OPT_TYPE_RATIONAL, AV_OPT_TYPE_BINARY, AV_OPT_TYPE_DICT, AV_OPT_TYPE_UINT64,
                    AV_OPT_TYPE_CONST, AV_OPT_TYPE_BOOL };
enum AVMediaType { AVMEDIA_TYPE_VIDEO = 0 };
enum { AV_CLASS_CATEGORY_FILTER = 8 };
enum AVPixelFormat { AV_PIX_FMT_NONE = -1, AV_PIX_FMT_YUV420P = 0, AV_PIX_FMT_GRAY8 = 8, AV_PIX_FMT_NV12 = 23, AV_PIX_FMT_CUDA = 117 };

typedef struct AVOption {
  const char* name;
  const char* help;
  int offset;
  enum AVOptionType type;
  union { int64_t i64; double dbl; const char* str; } default_val;
  double min, max;
  int flags;
  const char* unit;
} AVOption;

typedef struct AVClass {
  const char* class_name;
  const char* (*item_name)(void* ctx);
  const AVOption* option;
  int version;
  int category;
} AVClass;
static inline const char* av_default_item_name(void* ctx) { (void)ctx; return "transform360"; }

typedef struct AVDictionary AVDictionary;
static inline void av_dict_free(AVDictionary** d) { if (d) *d = NULL; }

typedef struct AVPixFmtDescriptor { int nb_components, log2_chroma_w, log2_chroma_h; } AVPixFmtDescriptor;
static inline const AVPixFmtDescriptor* av_pix_fmt_desc_get(int fmt) {
  static const AVPixFmtDescriptor yuv420 = {3, 1, 1}, gray = {1, 0, 0};
  return fmt == AV_PIX_FMT_GRAY8 ? &gray : &yuv420;
}
static inline int av_pix_fmt_count_planes(int fmt) { return av_pix_fmt_desc_get(fmt)->nb_components; }

/* ---- the slice of libavutil/buffer.h + hwcontext(_cuda).h that a CUDA-frame filter touches ------------------- */
typedef struct AVBufferRef {
  uint8_t* data;
  int* refs; /* shared count */
} AVBufferRef;
static inline AVBufferRef* ffshim_buffer_new(size_t bytes) {
  AVBufferRef* r = (AVBufferRef*)calloc(1, sizeof(*r));
  r->data = (uint8_t*)calloc(1, bytes);
  r->refs = (int*)calloc(1, sizeof(int));
  *r->refs = 1;
  return r;
}
static inline AVBufferRef* av_buffer_ref(AVBufferRef* b) {
  if (!b) return NULL;
  AVBufferRef* r = (AVBufferRef*)calloc(1, sizeof(*r));
  *r = *b;
  ++*r->refs;
  return r;
}
static inline void av_buffer_unref(AVBufferRef** b) {
  if (!b || !*b) return;
  if (--*(*b)->refs == 0) { free((*b)->data); free((*b)->refs); }
  free(*b);
  *b = NULL;
}

typedef void* CUcontext;
typedef void* CUstream;
typedef struct CudaFunctions { /* ffnvcodec's dynlink table: the three entries the filter calls */
  int (*cuCtxPushCurrent)(CUcontext ctx);
  int (*cuCtxPopCurrent)(CUcontext* ctx);
  int (*cuStreamSynchronize)(CUstream stream);
} CudaFunctions;
typedef struct AVCUDADeviceContextInternal { CudaFunctions* cuda_dl; } AVCUDADeviceContextInternal;
typedef struct AVCUDADeviceContext {
  CUcontext cuda_ctx;
  CUstream stream;
  AVCUDADeviceContextInternal* internal;
} AVCUDADeviceContext;
typedef struct AVHWDeviceContext { void* hwctx; } AVHWDeviceContext;
typedef struct AVHWFramesContext {
  AVBufferRef* device_ref;
  AVHWDeviceContext* device_ctx;
  int format, sw_format, width, height;
  int initialised;
  /* stand-in for the frame pool: the driver lends the planes the next av_hwframe_get_buffer hands out */
  uint8_t* lend[3];
  int lend_pitch[3];
} AVHWFramesContext;
static inline AVBufferRef* av_hwframe_ctx_alloc(AVBufferRef* device_ref) {
  if (!device_ref) return NULL;
  AVBufferRef* r = ffshim_buffer_new(sizeof(AVHWFramesContext));
  AVHWFramesContext* f = (AVHWFramesContext*)r->data;
  f->device_ref = device_ref; /* borrowed: the driver keeps the device alive for the filter's lifetime */
  f->device_ctx = (AVHWDeviceContext*)device_ref->data;
  return r;
}
static inline int av_hwframe_ctx_init(AVBufferRef* ref) {
  AVHWFramesContext* f = (AVHWFramesContext*)ref->data;
  if (f->format != AV_PIX_FMT_CUDA || f->width <= 0 || f->height <= 0) return AVERROR(EINVAL);
  f->initialised = 1;
  return 0;
}

typedef struct AVFrame {
  uint8_t* data[8];
  int linesize[8];
  int width, height, format;
  uint8_t* owned[8];
  AVBufferRef* hw_frames_ctx;
} AVFrame;
static inline AVFrame* av_frame_alloc(void) { return (AVFrame*)calloc(1, sizeof(AVFrame)); }
static inline int av_hwframe_get_buffer(AVBufferRef* ref, AVFrame* frame, int flags) {
  (void)flags;
  AVHWFramesContext* f = (AVHWFramesContext*)ref->data;
  if (!f->initialised) return AVERROR(EINVAL);
  if (!f->lend[0]) return AVERROR(ENOMEM);
  for (int p = 0; p < 3; p++) {
    frame->data[p] = f->lend[p];
    frame->linesize[p] = f->lend_pitch[p];
    f->lend[p] = NULL;
  }
  frame->format = AV_PIX_FMT_CUDA;
  frame->width = f->width;
  frame->height = f->height;
  frame->hw_frames_ctx = av_buffer_ref(ref);
  return 0;
}

struct AVFilterContext;
typedef struct AVFilterLink {
  struct AVFilterContext* src;
  struct AVFilterContext* dst;
  int w, h, format;
  AVFrame* delivered; /* what ff_filter_frame received last */
  AVBufferRef* hw_frames_ctx;
} AVFilterLink;

typedef struct AVFilterPad {
  const char* name;
  enum AVMediaType type;
  int (*filter_frame)(AVFilterLink* link, AVFrame* frame);
  int (*config_props)(AVFilterLink* link);
} AVFilterPad;

typedef struct AVFilter {
  const char* name;
  const char* description;
  int (*init_dict)(struct AVFilterContext* ctx, AVDictionary** options);
  int (*init)(struct AVFilterContext* ctx);
  int (*query_formats)(struct AVFilterContext* ctx);
  int flags_internal;
  void (*uninit)(struct AVFilterContext* ctx);
  int priv_size;
  const AVClass* priv_class;
  const AVFilterPad* inputs;
  const AVFilterPad* outputs;
} AVFilter;

typedef struct AVFilterContext {
  const AVClass* av_class;
  const AVFilter* filter;
  void* priv;
  AVFilterLink** inputs;
  AVFilterLink** outputs;
  const int* common_formats; /* what query_formats announced */
} AVFilterContext;
static inline const int* ff_make_format_list(const int* fmts) { return fmts; }
static inline int ff_set_common_formats(AVFilterContext* ctx, const int* fmts) { ctx->common_formats = fmts; return fmts ? 0 : AVERROR(ENOMEM); }

static inline void av_log(void* avcl, int level, const char* fmt, ...) { (void)avcl; (void)level; (void)fmt; }

/* w / h expressions: plain numbers are all the tests use */
static inline int av_expr_parse_and_eval(double* res, const char* s, const char* const* names, const double* values,
                                         const char* const* f1n, double (*const* f1)(void*, double), const char* const* f2n,
                                         double (*const* f2)(void*, double, double), void* opaque, int log_offset, void* log_ctx) {
  (void)names; (void)values; (void)f1n; (void)f1; (void)f2n; (void)f2; (void)opaque; (void)log_offset; (void)log_ctx;
  if (!s) { *res = NAN; return AVERROR(EINVAL); }
  char* end = NULL;
  *res = strtod(s, &end);
  return end == s ? AVERROR(EINVAL) : 0;
}

static inline AVFrame* ffshim_alloc_frame(int w, int h, int format) {
  AVFrame* f = (AVFrame*)calloc(1, sizeof(AVFrame));
  const AVPixFmtDescriptor* d = av_pix_fmt_desc_get(format);
  f->width = w; f->height = h; f->format = format;
  for (int p = 0; p < d->nb_components; p++) {
    int pw = p ? FF_CEIL_RSHIFT(w, d->log2_chroma_w) : w, ph = p ? FF_CEIL_RSHIFT(h, d->log2_chroma_h) : h;
    f->linesize[p] = (pw + 31) / 32 * 32 + 32; /* ffmpeg pads lines */
    f->owned[p] = f->data[p] = (uint8_t*)calloc((size_t)f->linesize[p] * ph + 64, 1);
  }
  return f;
}
static inline AVFrame* ff_get_video_buffer(AVFilterLink* link, int w, int h) { return ffshim_alloc_frame(w, h, link->format); }
static inline void av_frame_free(AVFrame** f) {
  if (!f || !*f) return;
  for (int p = 0; p < 8; p++) free((*f)->owned[p]);
  av_buffer_unref(&(*f)->hw_frames_ctx);
  free(*f);
  *f = NULL;
}
static inline int av_frame_copy_props(AVFrame* dst, const AVFrame* src) { (void)dst; (void)src; return 0; }
static inline int ff_filter_frame(AVFilterLink* link, AVFrame* frame) {
  if (link->delivered) av_frame_free(&link->delivered);
  link->delivered = frame;
  return 0;
}

#endif
