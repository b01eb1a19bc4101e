// sm_100a kernels of the projection-remap hot path.
//
//   gatherFrameKernel<K>   replace cv::remap as the reference calls it (VideoFrameTransform.cpp:748-754):
//   gatherKernel<K>        per output pixel a K x K window of the 8-bit source is weighted with OpenCV's
//   nearestKernel          15-bit fixed-point table and rounded with (sum + 16384) >> 15.  Bit-exact by
//                          construction: same table (host-built, sampling.cpp), same integer arithmetic.
//   blurStripKernel<HY>    replace cv::sepFilter2D over the reference's tiles (cpp:173-204, 579-704):
//   blurTileKernel         separable Gaussian, float32, fused multiply-add chain in the order cv2 4.13 uses
//   blurDirectKernel       (see oracle/t360_oracle.c for the model and its pin), round-half-even, u8.
//   areaResizeKernel       replaces cv::resize(INTER_AREA) shrinking (cpp:770-776).
//
// This is a gather, not a contraction: no tensor cores.  What the design is built around:
//   * A warp owns 32 adjacent output columns x 4 rows; a lane computes one column, so a warp-wide tap read
//     covers ~48 contiguous source bytes per source row.  Plan reads are 8 B per lane (256 contiguous bytes
//     per warp, tile-major), stores 1 B per lane (one full 32-byte sector per warp).
//   * gatherFrameKernel: ONE persistent launch takes the tiles of all planes of a frame, handed out by an
//     atomic counter.  Staged tiles (the bulk of every plane): the source window of a 32 x 32 (K=8: 32 x 64)
//     output tile is brought into shared memory by ONE cp.async.bulk.tensor.2d (TMA) box load from the
//     pitch-linear plane, double-buffered against the arithmetic through mbarriers; taps are then read as
//     aligned 32-bit shared-memory words (bank-granular, no 32-byte-sector waste: the same reads through L1
//     measured 13-26 sectors per request) and aligned with a funnel shift.
//   * Every window row is folded with IDP.2A: two s16 x u8 multiply-adds per instruction.
//   * The 1024-phase weight table sits in shared memory, transposed so that unrelated phases spread
//     over bank groups; persistent CTAs (grid = multiple of the SM count) stage it once.
//   * Tiles whose window does not fit a box or touches a plane border (BORDER_WRAP wraps rows AND columns,
//     cpp:719) read their taps through L1 inside the same launch; BORDER_TRANSPARENT plans, nearest
//     neighbour and planes TMA cannot describe go through gatherKernel / nearestKernel.
#include "kernels.cuh"

#include <cuda.h>  // CUtensorMap (type only; no libcuda symbol is referenced)

#include <algorithm>
#include <atomic>
#include <cstring>

namespace t360 {

namespace {

std::atomic<unsigned long long> gLaunches{0};

constexpr int kRowsPerThread = 4;

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
__device__ __forceinline__ int2 loadPlan(const int2* p) {
  int2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ int recordColumn(int word0) { return (int)((unsigned)word0 >> kRecordColumnShift); }
__device__ __forceinline__ int recordCol0(int word0) { return (word0 << (32 - kRecordColumnShift)) >> (32 - kRecordColumnShift); }

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
__host__ __device__ constexpr int weightBytes() { return 1024 * K * K * 2; }

template <int K>
__device__ __forceinline__ int weightSlot(int phase) { return weightSlotOf(K, phase); }

// Copies the [1024][K][K] int16 table into shared memory as [K*K/8][1024] 16-byte vectors (K >= 4) or
// [1024] 8-byte vectors (K == 2), slot-permuted by weightSlot().
template <int K>
__device__ __forceinline__ void stageWeights(const int16_t* __restrict__ g, unsigned char* smem) {
  if constexpr (K == 2) {
    const uint2* src = reinterpret_cast<const uint2*>(g);
    uint2* dst = reinterpret_cast<uint2*>(smem);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) dst[weightSlot<K>(i)] = __ldg(src + i);
  } else {
    constexpr int kVec = K * K / 8;  // uint4 per phase
    const uint4* src = reinterpret_cast<const uint4*>(g);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < 1024 * kVec; i += blockDim.x) dst[(i % kVec) * 1024 + weightSlot<K>(i / kVec)] = __ldg(src + i);
  }
}

// K x K window in GLOBAL memory (read-only path) whose rows are `pitch` bytes apart, starting at byte offset `off`
// of a 4-byte aligned base.  No bounds handling: the caller guarantees the window (plus the tail of its last
// aligned word) is readable.
template <int K>
__device__ __forceinline__ int foldWindow(const uint32_t* __restrict__ words, int off, int pitch,
                                          const unsigned char* wsmem, int phase) {
  phase = weightSlot<K>(phase);
  auto ld = [&](int wordIndex) -> uint32_t { return __ldg(words + wordIndex); };
  int acc = 0;
  if constexpr (K == 2) {
    const uint2 wt = reinterpret_cast<const uint2*>(wsmem)[phase];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const uint32_t b = __funnelshift_r(ld(off >> 2), ld((off >> 2) + 1), (off & 3) * 8);
      acc = dp2aLo(r == 0 ? wt.x : wt.y, b, acc);
      off += pitch;
    }
  } else if constexpr (K == 4) {
    const uint4* tab = reinterpret_cast<const uint4*>(wsmem);
    const uint4 wa = tab[phase], wb = tab[1024 + phase];
    const uint32_t w01[4] = {wa.x, wa.z, wb.x, wb.z}, w23[4] = {wa.y, wa.w, wb.y, wb.w};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint32_t b = __funnelshift_r(ld(off >> 2), ld((off >> 2) + 1), (off & 3) * 8);
      acc = dp2aLo(w01[r], b, acc);
      acc = dp2aHi(w23[r], b, acc);
      off += pitch;
    }
  } else {
    const uint4* tab = reinterpret_cast<const uint4*>(wsmem);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const uint4 wt = tab[r * 1024 + phase];
      const uint32_t q0 = ld(off >> 2), q1 = ld((off >> 2) + 1), q2 = ld((off >> 2) + 2);
      const int sh = (off & 3) * 8;
      const uint32_t b0 = __funnelshift_r(q0, q1, sh), b1 = __funnelshift_r(q1, q2, sh);
      acc = dp2aLo(wt.x, b0, acc);
      acc = dp2aHi(wt.y, b0, acc);
      acc = dp2aLo(wt.z, b1, acc);
      acc = dp2aHi(wt.w, b1, acc);
      off += pitch;
    }
  }
  return acc;
}

// Shared-memory flavour used by the staged kernel.  The staging pitch is a compile-time multiple of 4, so the
// aligned word address of every window row is (base & ~3) + r * PITCH and the byte shift is the same for all rows:
// one address computation per pixel, every load uses an immediate offset.
template <int IMM>
__device__ __forceinline__ uint32_t ldsWordImm(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
  return v;
}
template <int IMM>
__device__ __forceinline__ uint4 ldsVecImm(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(IMM));
  return v;
}
template <int IMM>
__device__ __forceinline__ uint2 ldsVec2Imm(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr), "n"(IMM));
  return v;
}

template <int K, int PITCH, int R>
struct WindowRows {
  static __device__ __forceinline__ void run(uint32_t rowAddr, int sh, uint32_t wAddr, const uint4& wa, const uint4& wb,
                                             const uint2& w2, int& acc) {
    if constexpr (K == 2) {
      const uint32_t b = __funnelshift_r(ldsWordImm<R * PITCH>(rowAddr), ldsWordImm<R * PITCH + 4>(rowAddr), sh);
      acc = dp2aLo(R == 0 ? w2.x : w2.y, b, acc);
    } else if constexpr (K == 4) {
      const uint32_t b = __funnelshift_r(ldsWordImm<R * PITCH>(rowAddr), ldsWordImm<R * PITCH + 4>(rowAddr), sh);
      const uint32_t w01 = R == 0 ? wa.x : (R == 1 ? wa.z : (R == 2 ? wb.x : wb.z));
      const uint32_t w23 = R == 0 ? wa.y : (R == 1 ? wa.w : (R == 2 ? wb.y : wb.w));
      acc = dp2aLo(w01, b, acc);
      acc = dp2aHi(w23, b, acc);
    } else {
      const uint4 wt = ldsVecImm<R * 16384>(wAddr);
      const uint32_t q0 = ldsWordImm<R * PITCH>(rowAddr), q1 = ldsWordImm<R * PITCH + 4>(rowAddr), q2 = ldsWordImm<R * PITCH + 8>(rowAddr);
      const uint32_t b0 = __funnelshift_r(q0, q1, sh), b1 = __funnelshift_r(q1, q2, sh);
      acc = dp2aLo(wt.x, b0, acc);
      acc = dp2aHi(wt.y, b0, acc);
      acc = dp2aLo(wt.z, b1, acc);
      acc = dp2aHi(wt.w, b1, acc);
    }
    if constexpr (R + 1 < K) WindowRows<K, PITCH, R + 1>::run(rowAddr, sh, wAddr, wa, wb, w2, acc);
  }
};

// stageAddr / wAddr: 32-bit shared-window addresses of the staging buffer and of the weight table
template <int K, int PITCH>
__device__ __forceinline__ int foldWindowShared(uint32_t stageAddr, int off, uint32_t wAddr, int phase) {
  static_assert(PITCH % 4 == 0, "staging pitch must keep rows word-aligned");
  const uint32_t rowAddr = stageAddr + (uint32_t)(off & ~3);
  const int sh = (off & 3) * 8;
  const uint32_t slotAddr = wAddr + (uint32_t)weightSlot<K>(phase) * (K == 2 ? 8u : 16u);
  uint4 wa = make_uint4(0, 0, 0, 0), wb = wa;
  uint2 w2 = make_uint2(0, 0);
  if constexpr (K == 2) w2 = ldsVec2Imm<0>(slotAddr);
  if constexpr (K == 4) { wa = ldsVecImm<0>(slotAddr); wb = ldsVecImm<16384>(slotAddr); }
  int acc = 0;
  WindowRows<K, PITCH, 0>::run(rowAddr, sh, slotAddr, wa, wb, w2, acc);
  return acc;
}

// ---- column sharing (staged kernel) ---------------------------------------------------------------------
// A thread computes 4 vertically adjacent output pixels.  On every face whose longitude does not depend on the
// output row (4 of the 6 cube faces, and equirect->equirect) they sample the SAME source columns, and consecutive
// pixels start 1 or 2 source rows apart, so their K-row windows overlap in K-1 or K-2 rows.  The window is kept in
// registers and slid down: each further pixel fetches only its 1 or 2 new rows (2-4 shared-memory words instead of
// 8 for cubic, 3-6 instead of 24 for Lanczos4).  Selecting "shift by 1 or by 2" is a SEL per row.
template <int K>
struct RowBytes {
  uint32_t b[K / 4];  // the K source bytes of one window row, already aligned
};

template <int K>
__device__ __forceinline__ RowBytes<K> loadWindowRow(const uint32_t* __restrict__ rowWords, int sh) {
  RowBytes<K> o;
  const uint32_t q0 = rowWords[0], q1 = rowWords[1];
  o.b[0] = __funnelshift_r(q0, q1, sh);
  if constexpr (K == 8) o.b[1] = __funnelshift_r(q1, rowWords[2], sh);
  return o;
}

template <int K>
__device__ __forceinline__ int foldRows(const RowBytes<K> (&W)[K], const unsigned char* wsmem, int phase) {
  const uint4* tab = reinterpret_cast<const uint4*>(wsmem) + weightSlot<K>(phase);
  int acc = 0;
  if constexpr (K == 4) {
    const uint4 wa = tab[0], wb = tab[1024];
    acc = dp2aLo(wa.x, W[0].b[0], acc); acc = dp2aHi(wa.y, W[0].b[0], acc);
    acc = dp2aLo(wa.z, W[1].b[0], acc); acc = dp2aHi(wa.w, W[1].b[0], acc);
    acc = dp2aLo(wb.x, W[2].b[0], acc); acc = dp2aHi(wb.y, W[2].b[0], acc);
    acc = dp2aLo(wb.z, W[3].b[0], acc); acc = dp2aHi(wb.w, W[3].b[0], acc);
  } else {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const uint4 wt = tab[r * 1024];
      acc = dp2aLo(wt.x, W[r].b[0], acc); acc = dp2aHi(wt.y, W[r].b[0], acc);
      acc = dp2aLo(wt.z, W[r].b[1], acc); acc = dp2aHi(wt.w, W[r].b[1], acc);
    }
  }
  return acc;
}

// (whether the 4 records of every thread of a warp can share their window columns -- same first column, row steps
// of 1 or 2 -- is decided by the host per warp and tile: StagedTile::shareMask)
template <int K, int PITCH>
__device__ __forceinline__ void gatherColumnShared(const unsigned char* stage, int boxX, int boxY, const int2 (&rec)[4],
                                                   const unsigned char* wsmem, int (&acc)[4]) {
  static_assert(PITCH % 4 == 0 && (K == 4 || K == 8), "");
  const int off = ((rec[0].y >> 10) - boxY) * PITCH + (recordCol0(rec[0].x) - boxX);
  const uint32_t* rowWords = reinterpret_cast<const uint32_t*>(stage + (off & ~3));
  const int sh = (off & 3) * 8;
  RowBytes<K> W[K];
#pragma unroll
  for (int r = 0; r < K; ++r) W[r] = loadWindowRow<K>(rowWords + r * (PITCH / 4), sh);
  acc[0] = foldRows<K>(W, wsmem, rec[0].y & 1023);
#pragma unroll
  for (int j = 1; j < 4; ++j) {
    const int d = (rec[j].y >> 10) - (rec[j - 1].y >> 10);
    rowWords += d * (PITCH / 4);
    const RowBytes<K> last = loadWindowRow<K>(rowWords + (K - 1) * (PITCH / 4), sh);
    RowBytes<K> prev = W[K - 1];
    if (d == 2) prev = loadWindowRow<K>(rowWords + (K - 2) * (PITCH / 4), sh);
#pragma unroll
    for (int r = 0; r + 2 < K; ++r)
#pragma unroll
      for (int i = 0; i < K / 4; ++i) W[r].b[i] = d == 1 ? W[r + 1].b[i] : W[r + 2].b[i];
    W[K - 2] = prev;
    W[K - 1] = last;
    acc[j] = foldRows<K>(W, wsmem, rec[j].y & 1023);
  }
}

__device__ __forceinline__ int roundToByte(int acc) {  // FixedPtCast<int, uchar, 15>
  return min(max((acc + (1 << 14)) >> 15, 0), 255);
}

struct SrcView {
  const uint32_t* words;  // source plane base rounded down to 4 bytes
  const uint8_t* bytes;   // true base
  int misalign;           // bytes - words
  int w, h, pitch;
};

// One output pixel through L1, any border case.  Returns the 8-bit value, or -1 when BORDER_TRANSPARENT
// leaves the pixel untouched.
template <int K, bool TRANSPARENT>
__device__ __forceinline__ int gatherPixel(const SrcView& s, const unsigned char* wsmem, int col0, int rowPhase) {
  const int row0 = rowPhase >> 10, phase = rowPhase & 1023;
  // interior: no wrapping, and the aligned word reads stay inside the row (col0 + K + 3 <= w)
  const bool interior = col0 >= 0 && row0 >= 0 && col0 + K + 3 <= s.w && row0 + K <= s.h;
  if (interior)
    return roundToByte(foldWindow<K>(s.words, row0 * s.pitch + col0 + s.misalign, s.pitch, wsmem, phase));

  // window touches an edge: per-tap addressing.  BORDER_WRAP wraps columns AND rows (reference cpp:719).
  const int16_t* wt = reinterpret_cast<const int16_t*>(wsmem) + weightSlot<K>(phase) * (K == 2 ? 4 : 8);
  if (TRANSPARENT) {
    // every interpolator leaves the pixel alone when its anchor sample lies outside the source
    const int ax = col0 + (K / 2 - 1), ay = row0 + (K / 2 - 1);
    if ((unsigned)ax >= (unsigned)s.w || (unsigned)ay >= (unsigned)s.h) return -1;
    if (K == 2) {
      // bilinear, anchor inside but the 2x2 window sticks out on the last row / column: OpenCV blends the taps
      // that exist and renormalises by their weight, rounding half up (oracle/t360_oracle.c)
      int num = 0, den = 0;
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c)
          if (col0 + c < s.w && row0 + r < s.h) {
            num += wt[r * 2 + c] * (int)__ldg(s.bytes + (size_t)(row0 + r) * s.pitch + col0 + c);
            den += wt[r * 2 + c];
          }
      return den > 0 ? (2 * num + den) / (2 * den) : -1;
    }
  }
  // border columns once, then every load of a row (of the whole window for K <= 4) before the first use: a warp that
  // straddles the +-180 degree seam waits for its few border lanes, so their latency is the tile's latency
  int acc = 0;
  int xs[K];
#pragma unroll
  for (int c = 0; c < K; ++c) xs[c] = TRANSPARENT ? reflect101(col0 + c, s.w) : wrapIndex(col0 + c, s.w);
#pragma unroll(K <= 4 ? K : 1)
  for (int r = 0; r < K; ++r) {
    const int yy = TRANSPARENT ? reflect101(row0 + r, s.h) : wrapIndex(row0 + r, s.h);
    const uint8_t* rowp = s.bytes + (size_t)yy * s.pitch;
    int px[K];
#pragma unroll
    for (int c = 0; c < K; ++c) px[c] = __ldg(rowp + xs[c]);
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const int e = r * K + c;  // element (r, c) lives in vector e / 8, lane e % 8 of the transposed table
      acc += (K == 2 ? wt[e] : wt[(e >> 3) * 1024 * 8 + (e & 7)]) * px[c];
    }
  }
  return roundToByte(acc);
}

template <int K, bool TRANSPARENT>
__global__ void __launch_bounds__(gatherThreads(K), K == 8 ? 1 : 4)
gatherKernel(GatherParams p, int tilesX, int numTiles) {
  extern __shared__ __align__(16) unsigned char smem[];
  stageWeights<K>(p.weights, smem);
  __syncthreads();

  SrcView s;
  s.bytes = p.src;
  s.misalign = (int)(reinterpret_cast<uintptr_t>(p.src) & 3);
  s.words = reinterpret_cast<const uint32_t*>(p.src - s.misalign);
  s.w = p.srcW; s.h = p.srcH; s.pitch = p.srcPitch;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (int tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
    const int ty = tile / tilesX, tx = tile - ty * tilesX;
    const int y0 = ty * gatherTileH(K) + warp * kRowsPerThread;
    const int segX = tx * kGatherTileW;
    if (y0 >= p.dstH || segX + lane >= p.dstW) continue;  // records exist for every pixel of the plane, in lane order
    int2 rec[kRowsPerThread];
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j)
      rec[j] = loadPlan(p.samples + ((size_t)tile * gatherTileH(K) + warp * kRowsPerThread + j) * kGatherTileW + lane);
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) {
      if (y0 + j >= p.dstH) break;
      const int v = gatherPixel<K, TRANSPARENT>(s, smem, recordCol0(rec[j].x), rec[j].y);
      if (!TRANSPARENT || v >= 0) p.dst[(size_t)(y0 + j) * p.dstPitch + segX + recordColumn(rec[j].x)] = (uint8_t)v;
    }
  }
}

template <bool TRANSPARENT>
__global__ void __launch_bounds__(256, 4) nearestKernel(GatherParams p, int tilesX, int numTiles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
    const int ty = tile / tilesX, tx = tile - ty * tilesX;
    const int y0 = ty * gatherTileH(1) + warp * kRowsPerThread;
    const int segX = tx * kGatherTileW;
    if (y0 >= p.dstH || segX + lane >= p.dstW) continue;
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) {
      if (y0 + j >= p.dstH) break;
      const int2 rec = loadPlan(p.samples + ((size_t)tile * gatherTileH(1) + warp * kRowsPerThread + j) * kGatherTileW + lane);
      const int x = segX + recordColumn(rec.x);
      int sx = recordCol0(rec.x), sy = rec.y >> 10;
      const bool inside = (unsigned)sx < (unsigned)p.srcW && (unsigned)sy < (unsigned)p.srcH;
      if (!inside && TRANSPARENT) continue;
      if (!inside) { sx = wrapIndex(sx, p.srcW); sy = wrapIndex(sy, p.srcH); }
      p.dst[(size_t)(y0 + j) * p.dstPitch + x] = __ldg(p.src + (size_t)sy * p.srcPitch + sx);
    }
  }
}

// ---- TMA-staged tiles -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(arrivals));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmaLoadBox(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smemAddr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smemAddr(bar)) : "memory");
}

template <int K, int CLS>
__host__ __device__ constexpr int stageBytes() {  // + slack for the last word over-read; TMA destinations need 128-byte alignment
  return (stageBoxW(K, CLS) * stageBoxH(K, CLS) + 64 + 127) & ~127;
}
// One persistent kernel per plane or per frame.  The job list is sorted by kind: every CTA starts with general tiles
// (taps through L1, any border case: latency-bound, so they run while all CTAs of the SM are busy and the first TMA box
// is already on its way), then class-1 tiles (their larger box takes both stage buffers, no prefetch), then streams
// class-0 tiles through the double-buffered TMA pipeline, which leaves a short, uniform tail.  (The staging logic
// itself accepts any order.)
template <int K>
__host__ __device__ constexpr int planeSmemBytes() { return weightBytes<K>() + 2 * stageBytes<K, 0>() + 64; }

struct FrameTensorMaps {
  CUtensorMap map[kMaxFramePlanes][kNumBoxClasses];
};

template <int K, int PITCH>
__device__ __forceinline__ void computeStagedTile(const PlaneView& p, const unsigned char* stage, int outX, int outY, int boxX,
                                                  int boxY, bool shared, const int2 (&rec)[kRowsPerThread], const unsigned char* wsmem,
                                                  int lane, int warp) {
  const int y0 = outY + warp * kRowsPerThread;
  const bool active = outX + lane < p.dstW;
  // `shared` (warp-uniform, from the tile header): every active lane's 4 pixels share their columns, all 4 rows exist
  // a thread's four pixels sit in ONE output column (one lane order per 32 x 4 block): one address, four row steps
  const int dstPitch = p.dstPitch;
  uint8_t* const dst = p.dst + (size_t)y0 * dstPitch + outX + recordColumn(rec[0].x);
  if (shared) {
    if constexpr (K >= 4) {
      if (active) {
        int acc[kRowsPerThread];
        gatherColumnShared<K, PITCH>(stage, boxX, boxY, rec, wsmem, acc);
#pragma unroll
        for (int j = 0; j < kRowsPerThread; ++j)
          dst[(size_t)j * dstPitch] = (uint8_t)roundToByte(acc[j]);
      }
    }
  } else if (active) {
    const uint32_t stageAddr = smemAddr(stage), wAddr = smemAddr(wsmem);
    const int dstH = p.dstH;
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) {
      if (y0 + j >= dstH) break;
      const int row0 = rec[j].y >> 10, phase = rec[j].y & 1023;
      const int off = (row0 - boxY) * PITCH + (recordCol0(rec[j].x) - boxX);
      const int acc = foldWindowShared<K, PITCH>(stageAddr, off, wAddr, phase);
      dst[(size_t)j * dstPitch] = (uint8_t)roundToByte(acc);
    }
  }
}

template <int K>
__global__ void __launch_bounds__(gatherThreads(K), K == 8 ? 1 : 3)
gatherFrameKernel(const __grid_constant__ FrameGatherParams p, StagedParams jobs, const __grid_constant__ FrameTensorMaps maps) {
  static_assert(stageBoxW(K, 1) * stageBoxH(K, 1) + 64 <= 2 * stageBytes<K, 0>(), "a class-1 box must fit both stage buffers");
  static_assert(stageBoxW(K, 0) * stageBoxH(K, 0) % 16 == 0, "the seam merge works on 16-byte vectors");
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* wsmem = smem;
  unsigned char* stage0 = smem + weightBytes<K>();
  constexpr int kStage = stageBytes<K, 0>();
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + 2 * kStage);
  constexpr uint32_t kBox0 = stageBoxW(K, 0) * stageBoxH(K, 0), kBox1 = stageBoxW(K, 1) * stageBoxH(K, 1);

  // Programmatic dependent launch: the next launch on the stream (the next frame's gather) may place its CTAs as soon
  // as ours retire, and run its prologue -- which touches only constant data: weights, job list, sampling records --
  // under our tail.  Everything an earlier kernel may have written (the source planes, the scheduler counters) is
  // only touched after griddepcontrol.wait below.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    mbarInit(&bars[0], 1);
    mbarInit(&bars[1], 1);
    mbarInit(&bars[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  stageWeights<K>(p.weights, wsmem);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Software pipeline, two deep, so that no load is waited for in the iteration that issues it (a warp executes in
  // order: a header load followed by the record loads that need its fields would stall the whole tile on the header):
  //   iteration n:  issue header(n+2) | issue records(n+1) from header(n+1), already in registers | compute tile n
  auto loadHeader = [&](int i) { return i < jobs.numTiles ? jobs.tiles[i] : StagedTile{0, 0, 0, 0}; };
  // The header fetched two jobs ahead must not be waited for where it is issued.  The compiler keeps warp-uniform
  // values in uniform registers and converts a loaded header the moment it arrives, which parked every warp on this
  // load at the top of every job (10 % of all stall samples); so the load is opaque (asm: four ordinary registers), and
  // the header becomes uniform -- through a warp reduction whose result the compiler knows to be uniform -- only at the
  // end of the job, when it has long arrived.  An index past the list reads its last entry; validity is tracked by
  // the index itself.
  auto issueHeaderLoad = [&](int i, int (&raw)[4]) {
    const StagedTile* src = jobs.tiles + min(i, jobs.numTiles - 1);
    asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(raw[0]), "=r"(raw[1]), "=r"(raw[2]), "=r"(raw[3]) : "l"(src));
  };
  auto uniformHeader = [&](const int (&raw)[4]) {
    return StagedTile{(int)__reduce_or_sync(0xffffffffu, (unsigned)raw[0]), (int)__reduce_or_sync(0xffffffffu, (unsigned)raw[1]),
                      (int)__reduce_or_sync(0xffffffffu, (unsigned)raw[2]), (int)__reduce_or_sync(0xffffffffu, (unsigned)raw[3])};
  };
  // The plane of a tile, field by field through selects on kernel-parameter operands: an indexed load from the
  // parameter bank instead would put its latency in front of every tile's record loads (measured: +5 % on a plane).
  static_assert(kMaxFramePlanes == 3, "planeOf selects among three planes");
  auto planeOf = [&](const StagedTile& t) {
    const int pl = t.outY >> kJobPlaneShift;
    const PlaneView &a = p.plane[0], &b = p.plane[1], &c = p.plane[2];
#define T360_PICK(f) (pl == 0 ? a.f : (pl == 1 ? b.f : c.f))
    return PlaneView{T360_PICK(src), T360_PICK(dst), T360_PICK(samples), T360_PICK(srcW), T360_PICK(srcH), T360_PICK(srcPitch),
                     T360_PICK(dstW), T360_PICK(dstH), T360_PICK(dstPitch), T360_PICK(tilesPerRow), 0};
#undef T360_PICK
  };
  auto loadRecords = [&](int i, const StagedTile& t, int2 (&rec)[kRowsPerThread]) {
    const PlaneView pv = planeOf(t);
    // tile-major records: one base address per warp and tile, the four rows at immediate offsets, no bounds checks
    const int tileIndex = ((t.outY & kJobRowMask) / gatherTileH(K)) * pv.tilesPerRow + t.outX / kGatherTileW;
    const int2* base = pv.samples + ((size_t)tileIndex * gatherTileH(K) + warp * kRowsPerThread) * kGatherTileW + lane;
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) rec[j] = i < jobs.numTiles ? loadPlan(base + j * kGatherTileW) : make_int2(0, 0);
  };
  // Dynamic tile scheduling: the first four jobs of a CTA are static (blockIdx.x + k * gridDim.x), every further one is
  // claimed from a global counter by thread 0 and handed to the other threads through a double-buffered shared slot
  // across the end-of-job barrier.  The value the atomic returns is not touched in the iteration that issues it -- a
  // warp executes in order and would sit out the round trip while the rest of the CTA waits for it at the barrier --
  // but one iteration later (in an asm statement, so that the compiler cannot hoist the use).
  int* claimSlot = reinterpret_cast<int*>(bars + 3);
  const int claimBase = 4 * gridDim.x;
  int claimedRaw = (int)blockIdx.x - (int)gridDim.x;  // thread 0; claimBase + claimedRaw = the CTA's fourth static job
  int i0 = blockIdx.x, i1 = i0 + gridDim.x, i2 = i1 + gridDim.x;
  StagedTile tile = loadHeader(i0), tileNext = loadHeader(i1);
  int2 rec[kRowsPerThread];
  loadRecords(i0, tile, rec);
  // q0 / q1: class-0 / class-1 tiles this CTA has consumed; issued0: class-0 boxes it has requested.  A class-0 tile
  // with sequence number q lives in stage q & 1 and completes phase (q >> 1) & 1 of that stage's barrier.
  uint32_t q0 = 0, q1 = 0, issued0 = 0;
  asm volatile("griddepcontrol.wait;" ::: "memory");  // earlier kernels on the stream are complete and visible from here on
  auto requestClass0 = [&](const StagedTile& t) {  // thread 0 only
    const uint32_t st = issued0 & 1;
    mbarExpectTx(&bars[st], kBox0);
    tmaLoadBox(stage0 + st * kStage, &maps.map[t.outY >> kJobPlaneShift][0], t.boxXY & 0xffff, t.boxXY >> 16, &bars[st]);
  };
  for (uint32_t it = 0; i0 < jobs.numTiles; ++it) {
    const int next = i1;
    if (threadIdx.x == 0) {
      int claimed;
      asm volatile("add.s32 %0, %1, %2;" : "=r"(claimed) : "r"(claimedRaw), "r"(claimBase));
      claimSlot[it & 1] = claimed;
      claimedRaw = atomicAdd(jobs.claimCounter, 1);
    }
    int headerAfterNext[4];
    issueHeaderLoad(i2, headerAfterNext);
    int2 recNext[kRowsPerThread];
    loadRecords(next, tileNext, recNext);
    const int kind = (tile.outY >> kJobKindShift) & kJobKindMask, outY = tile.outY & kJobRowMask;
    const PlaneView pv = planeOf(tile);
    const bool nextIsClass0 = next < jobs.numTiles && ((tileNext.outY >> kJobKindShift) & kJobKindMask) == 0;
    if (kind == 0) {
      if (issued0 == q0) {  // not prefetched (first job, or it follows a class-1 tile that needed both stages)
        if (threadIdx.x == 0) requestClass0(tile);
        ++issued0;
      }
      if (nextIsClass0) {  // the other stage was released by the barrier that ended the previous job
        if (threadIdx.x == 0) requestClass0(tileNext);
        ++issued0;
      }
      const uint32_t st = q0 & 1;
      mbarWait(&bars[st], (q0 >> 1) & 1);
      computeStagedTile<K, stageBoxW(K, 0)>(pv, stage0 + st * kStage, tile.outX, outY, tile.boxXY & 0xffff, tile.boxXY >> 16,
                                            K >= 4 && ((tile.shareMask >> warp) & 1), rec, wsmem, lane, warp);
      ++q0;
    } else if (kind == 1) {
      if (threadIdx.x == 0) {  // no class-0 box is in flight here: the larger box may span both stage buffers
        mbarExpectTx(&bars[2], kBox1);
        tmaLoadBox(stage0, &maps.map[tile.outY >> kJobPlaneShift][1], tile.boxXY & 0xffff, tile.boxXY >> 16, &bars[2]);
      }
      mbarWait(&bars[2], q1 & 1);
      computeStagedTile<K, stageBoxW(K, 1)>(pv, stage0, tile.outX, outY, tile.boxXY & 0xffff, tile.boxXY >> 16,
                                            K >= 4 && ((tile.shareMask >> warp) & 1), rec, wsmem, lane, warp);
      ++q1;
    } else if (kind == kJobSeam) {
      // two complementary class-0 boxes (zero-filled outside the plane), one per stage buffer, OR-ed into the first
      const int boxX = tile.boxXY & 0xffff, boxY = tile.boxXY >> 16;
      if (threadIdx.x == 0) {
        mbarExpectTx(&bars[2], 2 * kBox0);
        const CUtensorMap* m = &maps.map[tile.outY >> kJobPlaneShift][0];
        tmaLoadBox(stage0, m, boxX, boxY, &bars[2]);
        tmaLoadBox(stage0 + kStage, m, boxX - pv.srcW, boxY, &bars[2]);
      }
      mbarWait(&bars[2], q1 & 1);
      {
        uint4* a = reinterpret_cast<uint4*>(stage0);
        const uint4* b = reinterpret_cast<const uint4*>(stage0 + kStage);
        for (int i = threadIdx.x; i < (int)(kBox0 / 16); i += gatherThreads(K)) {
          uint4 x = a[i];
          const uint4 y = b[i];
          x.x |= y.x; x.y |= y.y; x.z |= y.z; x.w |= y.w;
          a[i] = x;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // these writes precede later TMA writes to the stage
      }
      __syncthreads();
      computeStagedTile<K, stageBoxW(K, 0)>(pv, stage0, tile.outX, outY, boxX, boxY, K >= 4 && ((tile.shareMask >> warp) & 1), rec,
                                            wsmem, lane, warp);
      ++q1;
    } else {
      if (nextIsClass0 && issued0 == q0) {  // both stages are idle during a general tile: start the next box now
        if (threadIdx.x == 0) requestClass0(tileNext);
        ++issued0;
      }
      SrcView sv;
      sv.bytes = pv.src;
      sv.misalign = (int)(reinterpret_cast<uintptr_t>(pv.src) & 3);
      sv.words = reinterpret_cast<const uint32_t*>(pv.src - sv.misalign);
      sv.w = pv.srcW; sv.h = pv.srcH; sv.pitch = pv.srcPitch;
      const int y0 = outY + warp * kRowsPerThread;
      if (tile.outX + lane < pv.dstW) {
#pragma unroll
        for (int j = 0; j < kRowsPerThread; ++j) {
          if (y0 + j >= pv.dstH) break;
          const int v = gatherPixel<K, false>(sv, wsmem, recordCol0(rec[j].x), rec[j].y);
          pv.dst[(size_t)(y0 + j) * pv.dstPitch + tile.outX + recordColumn(rec[j].x)] = (uint8_t)v;
        }
      }
    }
    __syncthreads();  // everyone is done with this job's stage before it is refilled (and sees the claimed index)
    i0 = i1; i1 = i2; i2 = claimSlot[it & 1];
    tile = tileNext;
    tileNext = uniformHeader(headerAfterNext);
    // (asm: the copies stay here, ahead of the next job's loads, whose scoreboards they would otherwise share)
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) {
      asm volatile("mov.b32 %0, %1;" : "=r"(rec[j].x) : "r"(recNext[j].x));
      asm volatile("mov.b32 %0, %1;" : "=r"(rec[j].y) : "r"(recNext[j].y));
    }
  }
  // the CTA that finishes last re-arms the scheduler for the next launch (claimCounter[0] = claims, [1] = finished CTAs)
  if (threadIdx.x == 0 && atomicAdd(jobs.claimCounter + 1, 1) == (int)gridDim.x - 1) {
    jobs.claimCounter[0] = 0;
    jobs.claimCounter[1] = 0;
    __threadfence();
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

// ---------------------------------------------------------------------------------------------------
// Register-resident low-pass (the fast path; vertical half-size HY <= 3, any horizontal size).
// One WARP per job, no shared memory, no barriers.  Lane L owns columns x0 + 8L .. x0 + 8L + 7 and marches
// down the rows of the strip:
//   horizontal: the 8 running sums advance together through the taps in chunks of 4; the source bytes they
//     need form a sliding window kept as floats in a 12-register ring (3 groups of 4).  Each chunk issues
//     32 FMAs, converts one new group of 4 bytes (PRMT into the mantissa of 2^23, minus 2^23: exact) fetched
//     as one aligned 32-bit word and aligned with a funnel shift, and reads 4 taps as one 128-bit uniform load.
//     Tap arrays are zero-padded to a multiple of 4: fma(0, p, s) == s exactly, and starting the chain from
//     +0 makes the first fma equal the reference's plain multiply, so the bits match the oracle's order.
//   vertical: the last 2*HY+1 row results stay in a register ring; the symmetric-pair FMA chain of the oracle
//     produces one output row per input row; rounding is the 1.5*2^23 magic add (round-half-even), and the low
//     mantissa bytes of 8 results are packed into one 64-bit store.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float byteToFloat(uint32_t word, int k) {
  // (float)byte k of word: place it in the low mantissa byte of 8388608.0f, subtract 8388608.0f
  return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7440 | k)) - 8388608.0f;
}

// Source bytes of one strip row as seen by one lane.  Interior strips read aligned 32-bit words through the
// read-only path; edge strips replicate the plane's left/right border byte by byte (BORDER_REPLICATE against the
// parent plane, reference cpp:197).
template <bool EDGE>
struct StripRowReader {
  const uint8_t* rowBytes;
  const uint32_t* rowWords;
  int firstByte, sh, width;

  __device__ __forceinline__ StripRowReader(const StripParams& p, int y, int firstByte_) : firstByte(firstByte_), width(p.width) {
    rowBytes = p.src + (size_t)min(max(y, 0), p.height - 1) * p.srcPitch;
    const uintptr_t a = reinterpret_cast<uintptr_t>(rowBytes) + firstByte;
    rowWords = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    sh = (int)(a & 3) * 8;
  }
  // raw[0..3]: what the first three groups (window positions 0..11) are made of
  __device__ __forceinline__ void head(uint32_t (&raw)[4]) const {
    if (EDGE) {
#pragma unroll
      for (int g = 0; g < 3; ++g) raw[g] = bytes(g);
      raw[3] = 0;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) raw[i] = __ldg(rowWords + i);
    }
  }
  __device__ __forceinline__ uint32_t bytes(int group) const {
    uint32_t g = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) g |= (uint32_t)__ldg(rowBytes + min(max(firstByte + group * 4 + b, 0), width - 1)) << (8 * b);
    return g;
  }
  // group `group` (>= 3) given the previous aligned word
  __device__ __forceinline__ uint32_t next(int group, uint32_t& prevWord) const {
    if (EDGE) return bytes(group);
    const uint32_t w = __ldg(rowWords + group + 1);
    const uint32_t g = __funnelshift_r(prevWord, w, sh);
    prevWord = w;
    return g;
  }
};

// One chunk of 4 taps: 32 FMAs on the ring, then ring slots 4u..4u+3 take the next group of source bytes.
template <bool EDGE, int U, bool LAST>
__device__ __forceinline__ void stripChunk(const StripRowReader<EDGE>& rd, const float4* __restrict__ taps, int c, uint32_t& prev,
                                           float (&ring)[12], float (&s)[8]) {
  const float4 k4 = __ldg(taps + c);
  uint32_t grp = 0;
  if (!LAST) grp = rd.next(c + 3, prev);  // the group that replaces the 4 oldest window positions
  const float k[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int m = 0; m < 8; ++m) s[m] = __fmaf_rn(k[t], ring[(4 * U + m + t) % 12], s[m]);
  if (!LAST) {
#pragma unroll
    for (int b = 0; b < 4; ++b) ring[4 * U + b] = byteToFloat(grp, b);
  }
}

template <bool EDGE>
__device__ __forceinline__ void stripRow(const StripParams& p, const StripJob& job, const StripRowReader<EDGE>& rd,
                                         const uint32_t (&raw)[4], float (&s)[8]) {
  float ring[12];
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const uint32_t grp = EDGE ? raw[g] : __funnelshift_r(raw[g], raw[g + 1], rd.sh);
#pragma unroll
    for (int b = 0; b < 4; ++b) ring[g * 4 + b] = byteToFloat(grp, b);
  }
  uint32_t prev = raw[3];
#pragma unroll
  for (int m = 0; m < 8; ++m) s[m] = 0.0f;
  const float4* taps = reinterpret_cast<const float4*>(p.taps + job.kxOffset);
  // 3-, 5-, 7-, 9-tap kernels (1-3 chunks) cover most of a plane: straight-line code for them (warp-uniform switch)
  switch (job.kxChunks) {
    case 1:
      stripChunk<EDGE, 0, true>(rd, taps, 0, prev, ring, s);
      return;
    case 2:
      stripChunk<EDGE, 0, false>(rd, taps, 0, prev, ring, s);
      stripChunk<EDGE, 1, true>(rd, taps, 1, prev, ring, s);
      return;
    case 3:
      stripChunk<EDGE, 0, false>(rd, taps, 0, prev, ring, s);
      stripChunk<EDGE, 1, false>(rd, taps, 1, prev, ring, s);
      stripChunk<EDGE, 2, true>(rd, taps, 2, prev, ring, s);
      return;
    default:
      break;
  }
  for (int c = 0; c < job.kxChunks; c += 3) {  // the ring is back in phase every 3 chunks
    stripChunk<EDGE, 0, false>(rd, taps, c, prev, ring, s);
    if (c + 1 < job.kxChunks) stripChunk<EDGE, 1, false>(rd, taps, c + 1, prev, ring, s);
    if (c + 2 < job.kxChunks) stripChunk<EDGE, 2, false>(rd, taps, c + 2, prev, ring, s);
  }
}

template <int HY, bool EDGE>
__device__ __forceinline__ void stripBody(const StripParams& p, const StripJob& job, int lane) {
  constexpr int L = 2 * HY + 1;
  const int hx = job.kxCount >> 1;
  const int lx = job.x0 + lane * kStripLanePx;
  if (lane * kStripLanePx >= job.w) return;
  const int firstByte = lx - hx;  // column of window position 0
  const float* __restrict__ ky = p.taps + job.kyOffset;
  float kv[HY + 1];
#pragma unroll
  for (int i = 0; i <= HY; ++i) kv[i] = __ldg(ky + HY + i);
  const int nValid = min(kStripLanePx, job.w - lane * kStripLanePx);
  const bool wide = nValid == 8 && ((reinterpret_cast<uintptr_t>(p.dst) | (unsigned)p.dstPitch | (unsigned)lx) & 7) == 0;

  float R[L][8];
  const int rowsTotal = job.h + 2 * HY;
  // software pipeline over rows: the head of row j+1 is requested before row j is computed (rows are first
  // touches of DRAM lines; without this every row start exposes the full memory latency)
  uint32_t rawNext[4];
  StripRowReader<EDGE>(p, job.y0 - HY, firstByte).head(rawNext);
  for (int jb = 0; jb < rowsTotal; jb += L) {
#pragma unroll
    for (int u = 0; u < L; ++u) {
      const int j = jb + u;
      if (j < rowsTotal) {
        const StripRowReader<EDGE> rd(p, job.y0 - HY + j, firstByte);
        uint32_t raw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) raw[i] = rawNext[i];
        if (j + 1 < rowsTotal) StripRowReader<EDGE>(p, job.y0 - HY + j + 1, firstByte).head(rawNext);
        stripRow<EDGE>(p, job, rd, raw, R[u]);
        if (j >= 2 * HY) {
          // centre row is the one computed HY steps ago: ring slot (u - HY) mod L
          constexpr int kBig = 4 * L;
          float o[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            float acc = __fmul_rn(kv[0], R[(u - HY + kBig) % L][m]);
#pragma unroll
            for (int i = 1; i <= HY; ++i)
              acc = __fmaf_rn(kv[i], __fadd_rn(R[(u - HY + i + kBig) % L][m], R[(u - HY - i + kBig) % L][m]), acc);
            o[m] = __fadd_rn(acc, 12582912.0f);  // low mantissa byte = rint(acc), half-even
          }
          const int y = job.y0 + j - 2 * HY;
          uint8_t* out = p.dst + (size_t)y * p.dstPitch + lx;
          const uint32_t lo = __byte_perm(__byte_perm(__float_as_uint(o[0]), __float_as_uint(o[1]), 0x0040),
                                          __byte_perm(__float_as_uint(o[2]), __float_as_uint(o[3]), 0x0040), 0x5410);
          const uint32_t hi = __byte_perm(__byte_perm(__float_as_uint(o[4]), __float_as_uint(o[5]), 0x0040),
                                          __byte_perm(__float_as_uint(o[6]), __float_as_uint(o[7]), 0x0040), 0x5410);
          if (wide) {
            *reinterpret_cast<uint2*>(out) = make_uint2(lo, hi);
          } else {
#pragma unroll
            for (int m = 0; m < 8; ++m)
              if (m < nValid) out[m] = (uint8_t)((m < 4 ? lo >> (8 * m) : hi >> (8 * (m - 4))) & 0xFF);
          }
        }
      }
    }
  }
}

template <int HY>
__global__ void __launch_bounds__(128, 3) blurStripKernel(StripParams p) {
  const int job = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (job >= p.numJobs) return;
  const StripJob j = p.jobs[job];
  if (j.edge) stripBody<HY, true>(p, j, threadIdx.x & 31);
  else stripBody<HY, false>(p, j, threadIdx.x & 31);
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

// ---------------------------------------------------------------------------------------------------
// cv::resize(INTER_AREA) shrink (reference cpp:770-776, only when *_scale_factor != 1).  Not on the hot
// configurations: a straightforward one-thread-per-output-pixel kernel that follows OpenCV's order of
// operations (separate multiply and add, float32) so that the result is bit-exact.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) areaResizeKernel(AreaParams p) {
  const int dx = blockIdx.x * 32 + (threadIdx.x & 31), dy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (dx >= p.dstW || dy >= p.dstH) return;
  int v;
  if (p.cellW > 0) {
    int sum = 0;
    for (int y = 0; y < p.cellH; ++y) {
      const uint8_t* row = p.src + (size_t)(dy * p.cellH + y) * p.srcPitch + dx * p.cellW;
      for (int x = 0; x < p.cellW; ++x) sum += __ldg(row + x);
    }
    if (p.cellW == 2 && p.cellH == 2) v = (sum + 2) >> 2;
    else v = __float2int_rn(__fmul_rn((float)sum, 1.f / (float)(p.cellW * p.cellH)));
  } else {
    const int x0 = __ldg(p.xFirst + dx), x1 = __ldg(p.xFirst + dx + 1);
    const int y0 = __ldg(p.yFirst + dy), y1 = __ldg(p.yFirst + dy + 1);
    float sum = 0.f;
    for (int j = y0; j < y1; ++j) {
      const int2 ty = __ldg(p.yTaps + j);
      const uint8_t* row = p.src + (size_t)ty.x * p.srcPitch;
      float buf = 0.f;
      for (int k = x0; k < x1; ++k) {
        const int2 tx = __ldg(p.xTaps + k);
        buf = __fadd_rn(buf, __fmul_rn((float)__ldg(row + tx.x), __int_as_float(tx.y)));
      }
      const float term = __fmul_rn(__int_as_float(ty.y), buf);
      sum = j == y0 ? term : __fadd_rn(sum, term);
    }
    v = __float2int_rn(sum);
  }
  p.dst[(size_t)dy * p.dstPitch + dx] = (uint8_t)min(max(v, 0), 255);
}

struct LaunchCfg {
  bool ready = false;
  int perSM = 0;
};

template <auto Kern>
cudaError_t prepare(LaunchCfg& cfg, int threads, int smemBytes) {
  if (cfg.ready) return cudaSuccess;
  cudaError_t err = cudaSuccess;
  if (smemBytes > 48 * 1024) {
    err = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
    if (err != cudaSuccess) return err;
  }
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg.perSM, Kern, threads, smemBytes);
  if (err != cudaSuccess) return err;
  if (cfg.perSM < 1) return cudaErrorLaunchOutOfResources;
  cfg.ready = true;
  return cudaSuccess;
}

template <int K, bool T>
cudaError_t launchGatherK(const GatherParams& p, int numSMs, cudaStream_t stream) {
  static thread_local LaunchCfg cfg;  // per kernel instantiation (and per host thread / device binding)
  constexpr int threads = gatherThreads(K), smemBytes = weightBytes<K>();
  cudaError_t err = prepare<gatherKernel<K, T>>(cfg, threads, smemBytes);
  if (err != cudaSuccess) return err;
  const int tilesX = (p.dstW + kGatherTileW - 1) / kGatherTileW;
  const int tilesY = (p.dstH + gatherTileH(K) - 1) / gatherTileH(K);
  const int numTiles = tilesX * tilesY;
  if (numTiles <= 0) return cudaSuccess;
  const int grid = std::min(numSMs * cfg.perSM, numTiles);  // whole waves: a multiple of the SM count
  gatherKernel<K, T><<<grid, threads, smemBytes, stream>>>(p, tilesX, numTiles);
  gLaunches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

template <bool T>
cudaError_t launchNearest(const GatherParams& p, int numSMs, cudaStream_t stream) {
  static thread_local LaunchCfg cfg;
  cudaError_t err = prepare<nearestKernel<T>>(cfg, 256, 0);
  if (err != cudaSuccess) return err;
  const int tilesX = (p.dstW + kGatherTileW - 1) / kGatherTileW;
  const int tilesY = (p.dstH + gatherTileH(1) - 1) / gatherTileH(1);
  const int grid = std::min(numSMs * cfg.perSM, tilesX * tilesY);
  nearestKernel<T><<<grid, 256, 0, stream>>>(p, tilesX, tilesX * tilesY);
  gLaunches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

template <int K>
cudaError_t launchFrameK(const FrameGatherParams& p, const StagedParams& jobs, const FrameTensorMaps& maps, int numSMs,
                         cudaStream_t stream) {
  static thread_local LaunchCfg cfg;
  constexpr int threads = gatherThreads(K), smemBytes = planeSmemBytes<K>();
  cudaError_t err = prepare<gatherFrameKernel<K>>(cfg, threads, smemBytes);
  if (err != cudaSuccess) return err;
  const int grid = std::min(numSMs * cfg.perSM, jobs.numTiles);  // persistent: whole waves of CTAs
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(grid);
  lc.blockDim = dim3(threads);
  lc.dynamicSmemBytes = smemBytes;
  lc.stream = stream;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see griddepcontrol.* in the kernel
  attr.val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = &attr;
  lc.numAttrs = 1;
  err = cudaLaunchKernelEx(&lc, gatherFrameKernel<K>, p, jobs, maps);
  gLaunches.fetch_add(1, std::memory_order_relaxed);
  return err;
}

}  // namespace

cudaError_t launchGather(const GatherParams& p, int numSMs, cudaStream_t stream) {
  if (p.dstW <= 0 || p.dstH <= 0) return cudaSuccess;
  const bool t = p.transparent != 0;
  switch (p.kernelSize) {
    case 1: return t ? launchNearest<true>(p, numSMs, stream) : launchNearest<false>(p, numSMs, stream);
    case 2: return t ? launchGatherK<2, true>(p, numSMs, stream) : launchGatherK<2, false>(p, numSMs, stream);
    case 4: return t ? launchGatherK<4, true>(p, numSMs, stream) : launchGatherK<4, false>(p, numSMs, stream);
    case 8: return t ? launchGatherK<8, true>(p, numSMs, stream) : launchGatherK<8, false>(p, numSMs, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launchGatherFrame(const FrameGatherParams& p, const StagedParams& jobs, const void* tensorMaps, int numSMs,
                              cudaStream_t stream) {
  if (jobs.numTiles <= 0) return cudaSuccess;
  if (p.numPlanes < 1 || p.numPlanes > kMaxFramePlanes) return cudaErrorInvalidValue;
  FrameTensorMaps maps;
  std::memcpy(&maps, tensorMaps, sizeof(CUtensorMap) * kNumBoxClasses * p.numPlanes);
  for (int i = p.numPlanes; i < kMaxFramePlanes; ++i)  // unused entries: valid descriptors that no tile refers to
    for (int c = 0; c < kNumBoxClasses; ++c) maps.map[i][c] = maps.map[0][c];
  switch (p.kernelSize) {
    case 2: return launchFrameK<2>(p, jobs, maps, numSMs, stream);
    case 4: return launchFrameK<4>(p, jobs, maps, numSMs, stream);
    case 8: return launchFrameK<8>(p, jobs, maps, numSMs, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launchAreaResize(const AreaParams& p, cudaStream_t stream) {
  if (p.dstW <= 0 || p.dstH <= 0) return cudaSuccess;
  const dim3 grid((p.dstW + 31) / 32, (p.dstH + 7) / 8);
  areaResizeKernel<<<grid, 256, 0, stream>>>(p);
  gLaunches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

cudaError_t launchBlurStrips(const StripParams& p, int hy, cudaStream_t stream) {
  if (p.numJobs <= 0) return cudaSuccess;
  const int grid = (p.numJobs + 3) / 4;
  switch (hy) {
    case 0: case 1: blurStripKernel<1><<<grid, 128, 0, stream>>>(p); break;  // hy == 0: one tap, padded with two zeros
    case 2: blurStripKernel<2><<<grid, 128, 0, stream>>>(p); break;
    case 3: blurStripKernel<3><<<grid, 128, 0, stream>>>(p); break;
    default: return cudaErrorInvalidValue;
  }
  gLaunches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
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
