/* oracle/ff_driver.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Drives the reference's UNMODIFIED ffmpeg filter (Transform360/vf_transform360.c, compiled against oracle/ffshim) the
 * way libavfilter would: defaults and "key=value:key=value" arguments through its AVOption table (vf:407-987),
 * init_dict (vf:306-326), config_output (vf:167-304), then filter_frame (vf:338-402) per frame.  Whatever
 * libTransform360 the object is linked against (the reference's, or the product's) does the work.
 */
#include "ffshim.h"

/* -DFF_FILTER=ff_vf_transform360_cuda builds the same driver around the product's CUDA-frame filter
 * (transform360_b200/filter/vf_transform360_cuda.c); t360f_open_cuda / t360f_filter_cuda below feed it device planes. */
#ifndef FF_FILTER
#define FF_FILTER ff_vf_transform360
#endif
extern AVFilter FF_FILTER;
#define ff_vf_transform360 FF_FILTER

typedef struct T360Filter {
  AVFilterContext ctx;
  AVFilterLink in, out;
  AVFilterLink* ins[1];
  AVFilterLink* outs[1];
  /* CUDA frames only */
  AVBufferRef* device;
  AVCUDADeviceContext cuda;
  AVCUDADeviceContextInternal cudaInternal;
  CudaFunctions cudaFns;
  int pushes, pops, syncs;
} T360Filter;

static const AVOption* find_option(const AVOption* o, const char* name) {
  for (; o && o->name; o++)
    if (o->type != AV_OPT_TYPE_CONST && !strcmp(o->name, name)) return o;
  return NULL;
}

static int set_option(void* priv, const AVOption* table, const AVOption* o, const char* value) {
  uint8_t* dst = (uint8_t*)priv + o->offset;
  if (o->type == AV_OPT_TYPE_STRING) {
    *(char**)dst = strdup(value);
    return 0;
  }
  double v;
  char* end = NULL;
  v = strtod(value, &end);
  if (end == value || *end) { /* a named constant of the option's unit */
    const AVOption* c = table;
    for (; c->name; c++)
      if (c->type == AV_OPT_TYPE_CONST && o->unit && c->unit && !strcmp(c->unit, o->unit) && !strcmp(c->name, value)) break;
    if (!c->name) return AVERROR(EINVAL);
    v = (double)c->default_val.i64;
  }
  if (v < o->min || v > o->max) return AVERROR(ERANGE);
  if (o->type == AV_OPT_TYPE_FLOAT) *(float*)dst = (float)v;
  else *(int*)dst = (int)v;
  return 0;
}

__attribute__((visibility("default"))) void t360f_close(T360Filter* f) {
  if (!f) return;
  if (f->ctx.priv) {
    if (ff_vf_transform360.uninit) ff_vf_transform360.uninit(&f->ctx);
    free(f->ctx.priv);
  }
  if (f->out.delivered) av_frame_free(&f->out.delivered);
  av_buffer_unref(&f->out.hw_frames_ctx);
  av_buffer_unref(&f->in.hw_frames_ctx);
  av_buffer_unref(&f->device);
  free(f);
}

static T360Filter* t360f_open_on(const char* args, int in_w, int in_h, int format, int sw_format, void* cuda_ctx, void* stream, void* const* driverFns, int* err);

/* returns NULL on bad arguments; err[0] receives the AVERROR code */
__attribute__((visibility("default"))) T360Filter* t360f_open(const char* args, int in_w, int in_h, int format, int* err) {
  return t360f_open_on(args, in_w, in_h, format, format, NULL, NULL, NULL, err);
}

/* A link of AV_PIX_FMT_CUDA frames (sw_format yuv420p / gray8) on the given CUcontext and stream.  driverFns =
 * {cuCtxPushCurrent, cuCtxPopCurrent, cuStreamSynchronize} of the CUDA driver, or NULL to configure without a device. */
__attribute__((visibility("default"))) T360Filter* t360f_open_cuda(const char* args, int in_w, int in_h, int sw_format, void* cuda_ctx, void* stream,
                                                                   void* const* driverFns, int* err) {
  return t360f_open_on(args, in_w, in_h, AV_PIX_FMT_CUDA, sw_format, cuda_ctx, stream, driverFns, err);
}

static int no_push(CUcontext c) { (void)c; return 0; }
static int no_pop(CUcontext* c) { if (c) *c = NULL; return 0; }
static int no_sync(CUstream s) { (void)s; return 0; }

static T360Filter* t360f_open_on(const char* args, int in_w, int in_h, int format, int sw_format, void* cuda_ctx, void* stream, void* const* driverFns, int* err) {
  T360Filter* f = (T360Filter*)calloc(1, sizeof(*f));
  if (format == AV_PIX_FMT_CUDA) {
    f->cudaFns.cuCtxPushCurrent = driverFns ? (int (*)(CUcontext))driverFns[0] : no_push;
    f->cudaFns.cuCtxPopCurrent = driverFns ? (int (*)(CUcontext*))driverFns[1] : no_pop;
    f->cudaFns.cuStreamSynchronize = driverFns ? (int (*)(CUstream))driverFns[2] : no_sync;
    f->cudaInternal.cuda_dl = &f->cudaFns;
    f->cuda.cuda_ctx = cuda_ctx;
    f->cuda.stream = stream;
    f->cuda.internal = &f->cudaInternal;
    f->device = ffshim_buffer_new(sizeof(AVHWDeviceContext));
    ((AVHWDeviceContext*)f->device->data)->hwctx = &f->cuda;
    f->in.hw_frames_ctx = av_hwframe_ctx_alloc(f->device);
    AVHWFramesContext* frames = (AVHWFramesContext*)f->in.hw_frames_ctx->data;
    frames->format = AV_PIX_FMT_CUDA;
    frames->sw_format = sw_format;
    frames->width = in_w;
    frames->height = in_h;
    av_hwframe_ctx_init(f->in.hw_frames_ctx);
  }
  const AVFilter* flt = &ff_vf_transform360;
  int rc = 0;
  f->ctx.filter = flt;
  f->ctx.av_class = flt->priv_class;
  f->ctx.priv = calloc(1, (size_t)flt->priv_size);
  *(const AVClass**)f->ctx.priv = flt->priv_class;
  f->ins[0] = &f->in; f->outs[0] = &f->out;
  f->ctx.inputs = f->ins; f->ctx.outputs = f->outs;
  f->in.dst = &f->ctx; f->out.src = &f->ctx;
  f->in.w = in_w; f->in.h = in_h; f->in.format = f->out.format = format;
  f->out.w = in_w; f->out.h = in_h; /* libavfilter presets an output link to its input's size */
  const AVOption* table = flt->priv_class->option;
  for (const AVOption* o = table; o->name; o++) { /* av_opt_set_defaults */
    uint8_t* dst = (uint8_t*)f->ctx.priv + o->offset;
    if (o->type == AV_OPT_TYPE_INT || o->type == AV_OPT_TYPE_BOOL) *(int*)dst = (int)o->default_val.i64;
    else if (o->type == AV_OPT_TYPE_FLOAT) *(float*)dst = (float)o->default_val.dbl;
    else if (o->type == AV_OPT_TYPE_STRING) *(char**)dst = o->default_val.str ? strdup(o->default_val.str) : NULL;
  }
  char* copy = strdup(args ? args : "");
  for (char* tok = strtok(copy, ":"); tok && !rc; tok = strtok(NULL, ":")) {
    char* eq = strchr(tok, '=');
    if (!eq) { rc = AVERROR(EINVAL); break; }
    *eq = 0;
    const AVOption* o = find_option(table, tok);
    rc = o ? set_option(f->ctx.priv, table, o, eq + 1) : AVERROR(ENOENT);
  }
  free(copy);
  AVDictionary* opts = NULL;
  if (!rc && flt->init_dict) rc = flt->init_dict(&f->ctx, &opts);
  if (!rc && flt->init) rc = flt->init(&f->ctx);
  if (!rc && flt->query_formats) {
    rc = flt->query_formats(&f->ctx);
    int offered = 0;
    for (const int* fmt = f->ctx.common_formats; !rc && fmt && *fmt != AV_PIX_FMT_NONE; fmt++) offered |= *fmt == format;
    if (!rc && !offered) rc = AVERROR(EINVAL); /* format negotiation would fail */
  }
  if (!rc) rc = flt->outputs[0].config_props(&f->out);
  if (err) *err = rc;
  if (rc) { t360f_close(f); return NULL; }
  return f;
}

__attribute__((visibility("default"))) void t360f_out_size(const T360Filter* f, int* w, int* h) { *w = f->out.w; *h = f->out.h; }

/* one frame: planes/pitches of the input (copied into an AVFrame the filter owns and frees), output copied out */
__attribute__((visibility("default"))) int t360f_filter(T360Filter* f, const uint8_t* const* planes, const int* pitches, uint8_t* const* outPlanes,
                                                        const int* outPitches) {
  const AVPixFmtDescriptor* d = av_pix_fmt_desc_get(f->in.format);
  AVFrame* in = ffshim_alloc_frame(f->in.w, f->in.h, f->in.format);
  for (int p = 0; p < d->nb_components; p++) {
    int pw = p ? FF_CEIL_RSHIFT(f->in.w, d->log2_chroma_w) : f->in.w, ph = p ? FF_CEIL_RSHIFT(f->in.h, d->log2_chroma_h) : f->in.h;
    for (int y = 0; y < ph; y++) memcpy(in->data[p] + (size_t)y * in->linesize[p], planes[p] + (size_t)y * pitches[p], (size_t)pw);
  }
  int rc = ff_vf_transform360.inputs[0].filter_frame(&f->in, in);
  if (rc) return rc;
  AVFrame* out = f->out.delivered;
  if (!out) return AVERROR(EINVAL);
  for (int p = 0; p < d->nb_components; p++) {
    int pw = p ? FF_CEIL_RSHIFT(f->out.w, d->log2_chroma_w) : f->out.w, ph = p ? FF_CEIL_RSHIFT(f->out.h, d->log2_chroma_h) : f->out.h;
    for (int y = 0; y < ph; y++) memcpy(outPlanes[p] + (size_t)y * outPitches[p], out->data[p] + (size_t)y * out->linesize[p], (size_t)pw);
  }
  return 0;
}

/* one frame of device planes: `planes` become the input AVFrame's data (as NVDEC would deliver them), `outPlanes` are lent
 * to the output frame pool; returns after filter_frame, the caller synchronises its stream (or relies on sync=1) */
__attribute__((visibility("default"))) int t360f_filter_cuda(T360Filter* f, uint8_t* const* planes, const int* pitches, uint8_t* const* outPlanes,
                                                             const int* outPitches) {
  if (!f->out.hw_frames_ctx) return AVERROR(EINVAL);
  AVHWFramesContext* pool = (AVHWFramesContext*)f->out.hw_frames_ctx->data;
  const int n = av_pix_fmt_count_planes(pool->sw_format);
  AVFrame* in = av_frame_alloc();
  in->format = AV_PIX_FMT_CUDA;
  in->width = f->in.w;
  in->height = f->in.h;
  in->hw_frames_ctx = av_buffer_ref(f->in.hw_frames_ctx);
  for (int p = 0; p < n; p++) {
    in->data[p] = planes[p];
    in->linesize[p] = pitches[p];
    pool->lend[p] = outPlanes[p];
    pool->lend_pitch[p] = outPitches[p];
  }
  int rc = ff_vf_transform360.inputs[0].filter_frame(&f->in, in);
  if (rc) return rc;
  AVFrame* out = f->out.delivered;
  if (!out || out->format != AV_PIX_FMT_CUDA || !out->hw_frames_ctx) return AVERROR(EINVAL);
  for (int p = 0; p < n; p++)
    if (out->data[p] != outPlanes[p]) return AVERROR(EINVAL);
  return 0;
}
