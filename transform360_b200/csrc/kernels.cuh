// Device-side interface of the remap hot path (sm_100a).  See kernels.cu.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace t360 {

// One gather launch: dst[y][x] = interpolate(src, samples[y][x]) for a whole plane.
struct GatherParams {
  const uint8_t* src;  // (blurred) input plane
  int srcW, srcH, srcPitch;
  uint8_t* dst;
  int dstW, dstH, dstPitch;
  const int2* samples;  // sampling records {col0, (row0 << 10) | phase}, tile-major (see "sampling records" below)
  int tilesPerRow;      // 32-column tiles per row of tiles
  const int16_t* weights;  // device copy of the [1024][k][k] table (nullptr for nearest)
  int kernelSize;          // 1, 2, 4, 8
  int transparent;         // BORDER_TRANSPARENT (barrel layouts) instead of BORDER_WRAP
};

// One tile of the TMA-staged gather: a gatherTileW x gatherTileH block of output pixels whose whole source
// window fits the fixed staging box placed at (boxX, boxY) of the source plane (boxX % 16 == 0).
struct StagedTile {
  int outX, outY;  // outY carries the job kind (kJobKindShift) and the image plane (kJobPlaneShift)
  int boxXY;       // boxX | boxY << 16
  int shareMask;   // bit w: warp w of the tile may slide its windows down the column (see gatherColumnShared); found by the host
};
// job kinds of the persistent gather kernel: 0 / 1 = staged through TMA with box class 0 / 1, 2 = general (L1) path,
// 3 = "seam": the tile's windows cross the left/right plane border (BORDER_WRAP, the +-180 degree meridian of an
// equirect source) but fit a class-0 box that wraps around it.  The box is fetched as TWO class-0 TMA loads, at
// column boxX and at column boxX - srcW: whatever lies outside the plane arrives as zeros, so the two boxes are
// complementary and their bitwise OR is the wrapped window.  The records of such a tile carry col0 relative to the
// unwrapped box (boxX <= col0 < boxX + box width, i.e. up to srcW + box width).
constexpr int kJobKindShift = 24, kJobRowMask = (1 << kJobKindShift) - 1, kJobGeneral = 2, kJobSeam = 3;
constexpr int kJobPlaneShift = 28, kJobKindMask = (1 << (kJobPlaneShift - kJobKindShift)) - 1;

// The persistent gather kernel takes the tiles of up to three image planes (Y, U, V of one frame) in ONE launch:
// one weight-table prologue and one tail per frame instead of per plane, and the dynamic scheduler balances the
// planes against each other.  Everything that differs between the planes sits in a PlaneView (+ two tensor maps).
constexpr int kMaxFramePlanes = 3;
struct PlaneView {
  const uint8_t* src;   // (blurred) input plane
  uint8_t* dst;
  const int2* samples;  // lane-ordered, tile-major records
  int srcW, srcH, srcPitch;
  int dstW, dstH, dstPitch;
  int tilesPerRow;
  int reserved;
};
struct FrameGatherParams {
  PlaneView plane[kMaxFramePlanes];
  const int16_t* weights;
  int kernelSize, numPlanes;
};

constexpr int kGatherTileW = 32;                                   // one warp = 32 adjacent columns
__host__ __device__ constexpr int gatherThreads(int k) { return k == 8 ? 512 : 256; }
__host__ __device__ constexpr int gatherTileH(int k) { return gatherThreads(k) / 32 * 4; }  // 4 rows per thread
// Staging boxes (bytes x rows), two classes per kernel size: the common one, and a larger one for tiles whose
// source window is wide (towards the poles).  Shared-memory row pitch = box width (TMA writes dense rows):
// 192 B = 48 words puts consecutive rows 16 banks apart, so a warp whose 32 adjacent pixels drift over 2-4
// source rows still reads conflict-free; 240 B = 60 words puts them 28 banks apart.
constexpr int kNumBoxClasses = 2;
__host__ __device__ constexpr int stageBoxW(int /*k*/, int cls) { return cls == 0 ? 192 : 240; }
// (three 256-thread CTAs of the common class share an SM: 32 KB of cubic weights + two 12 KB stages each.  A fourth
// fits with 63-row boxes but measured no faster -- the kernel is bound by shared-memory wavefronts, not by latency
// -- and it leaves no room for the concurrently running minority-tile kernels.)
__host__ __device__ constexpr int stageBoxH(int k, int cls) { return k == 8 ? (cls == 0 ? 112 : 144) : (cls == 0 ? 64 : 96); }

// Shared-memory slot of the weights of phase a = (fracY << 5) | fracX.  A 128-bit shared load is served 8 lanes
// (one quarter-warp) at a time out of 8 bank groups of 16 bytes -- a 64-bit one 16 lanes out of 16 groups -- and
// the group is the low bits of the slot.  The slot takes fracX's HIGH bits as its low bits: of the hashes tried in
// the offline bank simulator this one balances the phases of 32 adjacent pixels best.
__host__ __device__ constexpr int weightSlotOf(int k, int phase) {
  return k == 2 ? ((phase & ~31) | ((phase & 1) << 4) | ((phase & 31) >> 1))
                : ((phase & ~31) | ((phase & 3) << 3) | ((phase & 31) >> 2));
}
__host__ __device__ constexpr int weightBankGroups(int k) { return k == 2 ? 16 : 8; }
__host__ __device__ constexpr int weightLanesPerPass(int k) { return k == 2 ? 16 : 8; }

// Sampling records.  Tile-major: the gatherTileH(k) x 32 records of output tile (ty, tx) are contiguous,
//   records[((ty * tilesPerRow + tx) * gatherTileH(k) + rowInTile) * 32 + lane]
// (tiles that stick out of the plane are padded with zero records), so a warp fetches the records of its four rows
// from one base address with immediate offsets and needs no bounds checks.  Inside a 32-pixel row segment the records
// are in LANE order, not column order: the host deals the pixels of a segment to lanes so that the lanes served
// together by one shared-memory pass ask for different bank groups (see buildLaneOrder in video_frame_transform.cpp).
// Record word 0 therefore carries the pixel's column inside the segment in its top 5 bits:
// x = segmentX + (word0 >> 27), col0 = (word0 << 5) >> 5.
constexpr int kRecordColumnShift = 27;

struct StagedParams {
  const StagedTile* tiles;  // device list
  int numTiles;
  int* claimCounter;        // two device ints, zero before the first launch (the kernel re-arms them): tile scheduler
};

// One tile of the segmented low-pass: output rectangle and the taps to use.
struct BlurJob {
  int x0, y0, w, h;  // output rectangle (inside one plan segment)
  int kxOffset, kxCount, kyOffset, kyCount;
};

struct BlurParams {
  const uint8_t* src;
  uint8_t* dst;
  int width, height, srcPitch, dstPitch;
  const BlurJob* jobs;
  int numJobs;
  const float* taps;
  int tileSmemBytes;  // dynamic shared memory each block needs (max over jobs)
};

// One warp-job of the register-resident low-pass: a strip of up to 256 columns (8 per lane) x h rows inside one
// plan segment.  kxOffset points at the segment's horizontal taps zero-padded to kxChunks * 4 floats (16-byte
// aligned); kyOffset at the 2*hy+1 vertical taps.
struct StripJob {
  int x0, y0, w, h;
  int kxOffset, kxChunks, kxCount, kyOffset;
  int edge;  // 1: the strip's reads would cross the plane's left/right border -> clamped byte loads
};

struct StripParams {
  const uint8_t* src;
  uint8_t* dst;
  int width, height, srcPitch, dstPitch;
  const StripJob* jobs;
  int numJobs;
  const float* taps;
};

// cv::resize(INTER_AREA) shrink, one thread per destination pixel.  cellW > 0: integer ratios (sum of a cellW x cellH
// block, (sum+2)>>2 for 2x2, else rint(sum * (1.f/area))); otherwise the per-axis tap tables ({src, alpha} pairs,
// first[] offsets) with OpenCV's accumulation order: row sums over x taps, then weighted by the y taps.
struct AreaParams {
  const uint8_t* src;
  uint8_t* dst;
  int srcW, srcH, srcPitch, dstW, dstH, dstPitch;
  int cellW, cellH;
  const int2* xTaps;  // {src index, alpha bits}
  const int* xFirst;
  const int2* yTaps;
  const int* yFirst;
};
cudaError_t launchAreaResize(const AreaParams& p, cudaStream_t stream);

constexpr int kStripLanePx = 8, kStripW = 32 * kStripLanePx, kStripMaxHy = 3;

constexpr int kBlurTileW = 64, kBlurTileH = 32;
constexpr int kBlurMaxSmem = 96 * 1024;

// Launchers: enqueue on `stream`, return the CUDA status of the launch.  Each counts the kernels it launches.
// General path for a whole plane: taps through L1, every border mode (BORDER_WRAP, BORDER_TRANSPARENT), nearest.
cudaError_t launchGather(const GatherParams& p, int numSMs, cudaStream_t stream);
// Whole planes in one persistent kernel: `jobs` lists the tiles of every plane, sorted by kind (general, class 1,
// class 0).  tensorMaps: per plane two CUtensorMap (128 bytes each) describing its source with the staging boxes of
// class 0 and 1 of p.kernelSize, i.e. [numPlanes][kNumBoxClasses].  BORDER_WRAP only.
cudaError_t launchGatherFrame(const FrameGatherParams& p, const StagedParams& jobs, const void* tensorMaps, int numSMs,
                              cudaStream_t stream);
cudaError_t launchBlurStrips(const StripParams& p, int hy, cudaStream_t stream);  // register-resident, hy <= kStripMaxHy
cudaError_t launchBlur(const BlurParams& p, cudaStream_t stream);        // shared-memory tiles
cudaError_t launchBlurDirect(const BlurParams& p, cudaStream_t stream);  // any kernel size, slow
unsigned long long kernelLaunchCount();

// bytes of dynamic shared memory a blur tile of (w x h) with the given tap counts needs
inline int blurTileSmem(int w, int h, int nkx, int nky) {
  const int hx = nkx / 2, hy = nky / 2;
  const int rows = h + 2 * hy;
  const int srcStride = (w + 2 * hx + 3) & ~3;
  return rows * srcStride + rows * w * 4;
}

}  // namespace t360
