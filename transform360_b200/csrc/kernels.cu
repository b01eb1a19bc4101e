// sm_100a kernels of the projection-remap hot path.
//
//   gatherKernel<K>   replaces cv::remap as the reference calls it (VideoFrameTransform.cpp:748-754):
//                     per output pixel a K x K window of the 8-bit source is weighted with OpenCV's
//                     15-bit fixed-point table and rounded with (sum + 16384) >> 15.  Bit-exact by
//                     construction: same table (host-built, sampling.cpp), same integer arithmetic.
//   blurTileKernel    replaces cv::sepFilter2D over the reference's tiles (cpp:173-204, 579-704):
//                     separable Gaussian, float32, fused multiply-add chain in the order cv2 4.13 uses
//                     (see oracle/t360_oracle.c for the model and its pin), round-half-even, u8.
//
// This is an HBM / issue-bound gather, not a contraction: no tensor cores.  What matters here is
//   * coalescing: a warp owns 128 consecutive output pixels of one row, a thread 4 of them: one 128-bit
//     read of the sampling plan and one 32-bit store per thread, 128 B per warp;
//   * word-wide taps: the K source bytes of one window row are fetched as 2 (K<=4) or 3 (K=8) aligned
//     32-bit words through the read-only path and aligned with a funnel shift, then folded with IDP.2A
//     (two s16 x u8 MACs per instruction);
//   * the weight table lives in shared memory, transposed so that random phases spread over banks;
//   * a persistent grid (multiple of the SM count) loads that table once per CTA.
#include "kernels.cuh"

#include <atomic>

namespace t360 {

namespace {

std::atomic<unsigned long long> gLaunches{0};

constexpr int kGatherPxPerThread = 4;
constexpr int kGatherTileW = 32 * kGatherPxPerThread;  // one warp row

__device__ __forceinline__ int dp2aLo(uint32_t w, uint32_t b, int acc) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(b), "r"(acc));
  return d;
}
__device__ __forceinline__ int dp2aHi(uint32_t w, uint32_t b, int acc) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(b), "r"(acc));
  return d;
}

// the sampling plan is streamed once per frame: read-only path, do not allocate in L1
__device__ __forceinline__ int4 loadPlan(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ int wrapIndex(int p, int n) {  // cv::borderInterpolate(BORDER_WRAP)
  if ((unsigned)p < (unsigned)n) return p;
  p %= n;
  return p < 0 ? p + n : p;
}
__device__ __forceinline__ int reflect101(int p, int n) {  // what remap uses for taps under BORDER_TRANSPARENT
  if (n == 1) return 0;
  while ((unsigned)p >= (unsigned)n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

template <int K>
struct SmemTable;
template <>
struct SmemTable<2> {
  static constexpr int kBytes = 1024 * 8;
};
template <>
struct SmemTable<4> {
  static constexpr int kBytes = 1024 * 32;
};
template <>
struct SmemTable<8> {
  static constexpr int kBytes = 1024 * 128;
};

// Copies the [1024][K][K] int16 table into shared memory as [K*K/8][1024] 16-byte vectors (K >= 4) or
// [1024] 8-byte vectors (K == 2), so that lanes with unrelated phases hit different banks.
template <int K>
__device__ __forceinline__ void stageWeights(const int16_t* __restrict__ g, unsigned char* smem) {
  if (K == 2) {
    const uint2* src = reinterpret_cast<const uint2*>(g);
    uint2* dst = reinterpret_cast<uint2*>(smem);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) dst[i] = __ldg(src + i);
  } else {
    constexpr int kVec = K * K / 8;  // uint4 per phase
    const uint4* src = reinterpret_cast<const uint4*>(g);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < 1024 * kVec; i += blockDim.x) dst[(i % kVec) * 1024 + i / kVec] = __ldg(src + i);
  }
}

struct SrcView {
  const uint32_t* words;  // source plane base rounded down to 4 bytes
  const uint8_t* bytes;   // true base
  int misalign;           // bytes - words
  int w, h, pitch;
};

// One output pixel.  Returns the 8-bit value, or -1 when BORDER_TRANSPARENT leaves the pixel untouched.
template <int K, bool TRANSPARENT>
__device__ __forceinline__ int gatherPixel(const SrcView& s, const unsigned char* smem, int col0, int rowPhase) {
  const int row0 = rowPhase >> 10, phase = rowPhase & 1023;
  int acc = 0;
  // interior: no wrapping, and the aligned word reads stay inside the row (col0 + K + 3 <= w)
  const bool interior = col0 >= 0 && row0 >= 0 && col0 + K + 3 <= s.w && row0 + K <= s.h;
  if (interior) {
    int off = row0 * s.pitch + col0 + s.misalign;
    if (K == 2) {
      const uint2 wt = reinterpret_cast<const uint2*>(smem)[phase];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const uint32_t* p = s.words + (off >> 2);
        const uint32_t b = __funnelshift_r(__ldg(p), __ldg(p + 1), (off & 3) * 8);
        acc = dp2aLo(r == 0 ? wt.x : wt.y, b, acc);
        off += s.pitch;
      }
    } else if (K == 4) {
      const uint4* tab = reinterpret_cast<const uint4*>(smem);
      const uint4 wa = tab[phase], wb = tab[1024 + phase];
      const uint32_t w01[4] = {wa.x, wa.z, wb.x, wb.z}, w23[4] = {wa.y, wa.w, wb.y, wb.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t* p = s.words + (off >> 2);
        const uint32_t b = __funnelshift_r(__ldg(p), __ldg(p + 1), (off & 3) * 8);
        acc = dp2aLo(w01[r], b, acc);
        acc = dp2aHi(w23[r], b, acc);
        off += s.pitch;
      }
    } else {
      const uint4* tab = reinterpret_cast<const uint4*>(smem);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const uint4 wt = tab[r * 1024 + phase];
        const uint32_t* p = s.words + (off >> 2);
        const uint32_t q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2);
        const int sh = (off & 3) * 8;
        const uint32_t b0 = __funnelshift_r(q0, q1, sh), b1 = __funnelshift_r(q1, q2, sh);
        acc = dp2aLo(wt.x, b0, acc);
        acc = dp2aHi(wt.y, b0, acc);
        acc = dp2aLo(wt.z, b1, acc);
        acc = dp2aHi(wt.w, b1, acc);
        off += s.pitch;
      }
    }
  } else {
    // window touches an edge: per-tap addressing.  BORDER_WRAP wraps columns AND rows (reference cpp:719).
    if (TRANSPARENT) {
      const bool inlier = col0 >= 0 && row0 >= 0 && col0 + K <= s.w && row0 + K <= s.h;
      if (K == 2 && !inlier) return -1;  // remapBilinear leaves every non-inlier alone
      const int ax = col0 + (K / 2 - 1), ay = row0 + (K / 2 - 1);
      if ((unsigned)ax >= (unsigned)s.w || (unsigned)ay >= (unsigned)s.h) return -1;
    }
    const int16_t* wt;
    int strideR, strideC;  // in int16 units, matching the transposed shared layout
    if (K == 2) {
      wt = reinterpret_cast<const int16_t*>(smem) + phase * 4;
      strideR = 2; strideC = 1;
    } else {
      wt = reinterpret_cast<const int16_t*>(smem) + phase * 8;
      strideR = 0; strideC = 0;  // handled below
    }
#pragma unroll 1
    for (int r = 0; r < K; ++r) {
      const int yy = TRANSPARENT ? reflect101(row0 + r, s.h) : wrapIndex(row0 + r, s.h);
      const uint8_t* rowp = s.bytes + (size_t)yy * s.pitch;
#pragma unroll 1
      for (int c = 0; c < K; ++c) {
        const int xx = TRANSPARENT ? reflect101(col0 + c, s.w) : wrapIndex(col0 + c, s.w);
        int w;
        if (K == 2) {
          w = wt[r * strideR + c * strideC];
        } else {
          // element (r, c) of the phase lives in vector v = (r*K + c) / 8, lane (r*K + c) % 8
          const int e = r * K + c;
          w = wt[(e >> 3) * 1024 * 8 + (e & 7)];
        }
        acc += w * (int)__ldg(rowp + xx);
      }
    }
  }
  const int v = (acc + (1 << 14)) >> 15;
  return min(max(v, 0), 255);
}

template <int K, bool TRANSPARENT>
__global__ void __launch_bounds__(K == 8 ? 512 : 256) gatherKernel(GatherParams p, int tilesX, int numTiles) {
  extern __shared__ __align__(16) unsigned char smem[];
  stageWeights<K>(p.weights, smem);
  __syncthreads();

  SrcView s;
  s.bytes = p.src;
  s.misalign = (int)(reinterpret_cast<uintptr_t>(p.src) & 3);
  s.words = reinterpret_cast<const uint32_t*>(p.src - s.misalign);
  s.w = p.srcW; s.h = p.srcH; s.pitch = p.srcPitch;

  const int rowsPerTile = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool dstWordAligned = ((reinterpret_cast<uintptr_t>(p.dst) | (unsigned)p.dstPitch) & 3) == 0;

  for (int tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
    const int ty = tile / tilesX, tx = tile - ty * tilesX;
    const int y = ty * rowsPerTile + warp;
    const int x = tx * kGatherTileW + lane * kGatherPxPerThread;
    if (y >= p.dstH || x >= p.dstW) continue;
    // 4 sampling records = 32 bytes, two 128-bit loads; the plan is re-read every frame: keep it out of L1
    const int4* sp = reinterpret_cast<const int4*>(p.samples + (size_t)y * p.samplesPitch + x);
    const int4 s01 = loadPlan(sp), s23 = loadPlan(sp + 1);
    const int cols[4] = {s01.x, s01.z, s23.x, s23.z}, rps[4] = {s01.y, s01.w, s23.y, s23.w};
    int v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = gatherPixel<K, TRANSPARENT>(s, smem, cols[i], rps[i]);
    uint8_t* out = p.dst + (size_t)y * p.dstPitch + x;
    const bool full = x + 3 < p.dstW && (!TRANSPARENT || (v[0] | v[1] | v[2] | v[3]) >= 0);
    if (full && dstWordAligned) {
      *reinterpret_cast<uint32_t*>(out) = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (x + i < p.dstW && v[i] >= 0) out[i] = (uint8_t)v[i];
    }
  }
}

template <bool TRANSPARENT>
__global__ void __launch_bounds__(256) nearestKernel(GatherParams p, int tilesX, int numTiles) {
  const int rowsPerTile = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool dstWordAligned = ((reinterpret_cast<uintptr_t>(p.dst) | (unsigned)p.dstPitch) & 3) == 0;
  for (int tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
    const int ty = tile / tilesX, tx = tile - ty * tilesX;
    const int y = ty * rowsPerTile + warp;
    const int x = tx * kGatherTileW + lane * kGatherPxPerThread;
    if (y >= p.dstH || x >= p.dstW) continue;
    const int4* sp = reinterpret_cast<const int4*>(p.samples + (size_t)y * p.samplesPitch + x);
    const int4 s01 = loadPlan(sp), s23 = loadPlan(sp + 1);
    const int cols[4] = {s01.x, s01.z, s23.x, s23.z}, rows[4] = {s01.y >> 10, s01.w >> 10, s23.y >> 10, s23.w >> 10};
    int v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int sx = cols[i], sy = rows[i];
      const bool inside = (unsigned)sx < (unsigned)p.srcW && (unsigned)sy < (unsigned)p.srcH;
      if (!inside && TRANSPARENT) { v[i] = -1; continue; }
      if (!inside) { sx = wrapIndex(sx, p.srcW); sy = wrapIndex(sy, p.srcH); }
      v[i] = __ldg(p.src + (size_t)sy * p.srcPitch + sx);
    }
    uint8_t* out = p.dst + (size_t)y * p.dstPitch + x;
    const bool full = x + 3 < p.dstW && (!TRANSPARENT || (v[0] | v[1] | v[2] | v[3]) >= 0);
    if (full && dstWordAligned) {
      *reinterpret_cast<uint32_t*>(out) = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (x + i < p.dstW && v[i] >= 0) out[i] = (uint8_t)v[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Segmented low-pass.  One CTA per job (a <= 64 x 32 rectangle inside one plan segment).
//   stage 0: source bytes of the rectangle grown by the kernel half-sizes -> shared (edge-replicated
//            against the PLANE border only: tiles see their real neighbours, reference cpp:184-197)
//   stage 1: horizontal pass -> float rows in shared:  s = kx[0]*p[0]; s = fma(kx[i], p[i], s)
//   stage 2: vertical pass, symmetric pairs:  s = ky[h]*R[y]; s = fma(ky[h+i], R[y+i] + R[y-i], s)
//            -> round-half-even, saturate, store
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) blurTileKernel(BlurParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const BlurJob job = p.jobs[blockIdx.x];
  const int hx = job.kxCount >> 1, hy = job.kyCount >> 1;
  const int rows = job.h + 2 * hy;
  const int srcStride = (job.w + 2 * hx + 3) & ~3;
  uint8_t* tile = smem;
  float* R = reinterpret_cast<float*>(smem + rows * srcStride);
  const float* __restrict__ kx = p.taps + job.kxOffset;
  const float* __restrict__ ky = p.taps + job.kyOffset;

  const int cols = job.w + 2 * hx;
  for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
    const int r = i / cols, c = i - r * cols;
    const int sy = min(max(job.y0 - hy + r, 0), p.height - 1);
    const int sx = min(max(job.x0 - hx + c, 0), p.width - 1);
    tile[r * srcStride + c] = __ldg(p.src + (size_t)sy * p.srcPitch + sx);
  }
  __syncthreads();

  for (int i = threadIdx.x; i < rows * job.w; i += blockDim.x) {
    const int r = i / job.w, x = i - r * job.w;
    const uint8_t* t = tile + r * srcStride + x;
    float s = __fmul_rn(__ldg(kx), (float)t[0]);
    for (int k = 1; k < job.kxCount; ++k) s = __fmaf_rn(__ldg(kx + k), (float)t[k], s);
    R[i] = s;
  }
  __syncthreads();

  for (int i = threadIdx.x; i < job.h * job.w; i += blockDim.x) {
    const int y = i / job.w, x = i - y * job.w;
    const float* c = R + (y + hy) * job.w + x;
    float s = __fmul_rn(__ldg(ky + hy), c[0]);
    for (int k = 1; k <= hy; ++k) s = __fmaf_rn(__ldg(ky + hy + k), __fadd_rn(c[k * job.w], c[-k * job.w]), s);
    const int v = __float2int_rn(s);
    p.dst[(size_t)(job.y0 + y) * p.dstPitch + job.x0 + x] = (uint8_t)min(max(v, 0), 255);
  }
}

// Fallback for kernels too large for a shared-memory tile (sigma can reach half the plane width):
// every thread evaluates its pixel's whole separable sum straight from global memory.  Same order.
__global__ void __launch_bounds__(256) blurDirectKernel(BlurParams p) {
  const BlurJob job = p.jobs[blockIdx.x];
  const int hx = job.kxCount >> 1, hy = job.kyCount >> 1;
  const float* __restrict__ kx = p.taps + job.kxOffset;
  const float* __restrict__ ky = p.taps + job.kyOffset;
  for (int i = threadIdx.x; i < job.h * job.w; i += blockDim.x) {
    const int y = job.y0 + i / job.w, x = job.x0 + i % job.w;
    auto rowSum = [&](int yy) {
      const uint8_t* rowp = p.src + (size_t)min(max(yy, 0), p.height - 1) * p.srcPitch;
      float s = __fmul_rn(__ldg(kx), (float)__ldg(rowp + min(max(x - hx, 0), p.width - 1)));
      for (int k = 1; k < job.kxCount; ++k)
        s = __fmaf_rn(__ldg(kx + k), (float)__ldg(rowp + min(max(x - hx + k, 0), p.width - 1)), s);
      return s;
    };
    float s = __fmul_rn(__ldg(ky + hy), rowSum(y));
    for (int k = 1; k <= hy; ++k) s = __fmaf_rn(__ldg(ky + hy + k), __fadd_rn(rowSum(y + k), rowSum(y - k)), s);
    const int v = __float2int_rn(s);
    p.dst[(size_t)y * p.dstPitch + x] = (uint8_t)min(max(v, 0), 255);
  }
}

struct LaunchCfg {
  bool ready = false;
  int perSM = 0;
};

template <auto Kern>
cudaError_t launchPersistent(const GatherParams& p, int threads, int smemBytes, int numSMs, cudaStream_t stream) {
  static thread_local LaunchCfg cfg;  // one per kernel instantiation (and per host thread / device binding)
  if (!cfg.ready) {
    cudaError_t err = cudaSuccess;
    if (smemBytes > 48 * 1024) {
      err = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
      if (err != cudaSuccess) return err;
    }
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg.perSM, Kern, threads, smemBytes);
    if (err != cudaSuccess) return err;
    if (cfg.perSM < 1) return cudaErrorLaunchOutOfResources;
    cfg.ready = true;
  }
  const int rowsPerTile = threads / 32;
  const int tilesX = (p.dstW + kGatherTileW - 1) / kGatherTileW;
  const int tilesY = (p.dstH + rowsPerTile - 1) / rowsPerTile;
  const int numTiles = tilesX * tilesY;
  int grid = numSMs * cfg.perSM;  // whole waves: a multiple of the SM count
  if (grid > numTiles) grid = numTiles;
  Kern<<<grid, threads, smemBytes, stream>>>(p, tilesX, numTiles);
  gLaunches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launchGather(const GatherParams& p, int numSMs, cudaStream_t stream) {
  if (p.dstW <= 0 || p.dstH <= 0) return cudaSuccess;
  const bool t = p.transparent != 0;
  switch (p.kernelSize) {
    case 1:
      return t ? launchPersistent<nearestKernel<true>>(p, 256, 0, numSMs, stream)
               : launchPersistent<nearestKernel<false>>(p, 256, 0, numSMs, stream);
    case 2:
      return t ? launchPersistent<gatherKernel<2, true>>(p, 256, SmemTable<2>::kBytes, numSMs, stream)
               : launchPersistent<gatherKernel<2, false>>(p, 256, SmemTable<2>::kBytes, numSMs, stream);
    case 4:
      return t ? launchPersistent<gatherKernel<4, true>>(p, 256, SmemTable<4>::kBytes, numSMs, stream)
               : launchPersistent<gatherKernel<4, false>>(p, 256, SmemTable<4>::kBytes, numSMs, stream);
    case 8:
      return t ? launchPersistent<gatherKernel<8, true>>(p, 512, SmemTable<8>::kBytes, numSMs, stream)
               : launchPersistent<gatherKernel<8, false>>(p, 512, SmemTable<8>::kBytes, numSMs, stream);
    default:
      return cudaErrorInvalidValue;
  }
}

cudaError_t launchBlur(const BlurParams& p, cudaStream_t stream) {
  if (p.numJobs <= 0) return cudaSuccess;
  static thread_local int configuredSmem = 0;
  if (p.tileSmemBytes > 48 * 1024 && p.tileSmemBytes > configuredSmem) {
    cudaError_t err = cudaFuncSetAttribute(blurTileKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBlurMaxSmem);
    if (err != cudaSuccess) return err;
    configuredSmem = kBlurMaxSmem;
  }
  blurTileKernel<<<p.numJobs, 256, p.tileSmemBytes, stream>>>(p);
  gLaunches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

cudaError_t launchBlurDirect(const BlurParams& p, cudaStream_t stream) {
  if (p.numJobs <= 0) return cudaSuccess;
  blurDirectKernel<<<p.numJobs, 256, 0, stream>>>(p);
  gLaunches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

unsigned long long kernelLaunchCount() { return gLaunches.load(std::memory_order_relaxed); }

}  // namespace t360
