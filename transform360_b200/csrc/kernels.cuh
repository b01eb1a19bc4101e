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
  const int2* samples;  // [dstH][samplesPitch]: {col0, (row0 << 10) | phase}
  int samplesPitch;     // elements per row, multiple of 4
  const int16_t* weights;  // device copy of the [1024][k][k] table (nullptr for nearest)
  int kernelSize;          // 1, 2, 4, 8
  int transparent;         // BORDER_TRANSPARENT (barrel layouts) instead of BORDER_WRAP
};

// One tile of the TMA-staged gather: a gatherTileW x gatherTileH block of output pixels whose whole source
// window fits the fixed staging box placed at (boxX, boxY) of the source plane (boxX % 16 == 0).
struct StagedTile {
  int outX, outY, boxX, boxY;
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
__host__ __device__ constexpr int stageBoxH(int k, int cls) { return k == 8 ? (cls == 0 ? 112 : 144) : (cls == 0 ? 64 : 96); }

struct StagedParams {
  const StagedTile* tiles;  // device list
  int numTiles;
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

constexpr int kStripLanePx = 8, kStripW = 32 * kStripLanePx, kStripMaxHy = 3;

constexpr int kBlurTileW = 64, kBlurTileH = 32;
constexpr int kBlurMaxSmem = 96 * 1024;

// Launchers: enqueue on `stream`, return the CUDA status of the launch.  Each counts the kernels it launches.
// tileList == nullptr: every tile of the plane; otherwise only the listed tile indices (row-major, tilesX wide)
cudaError_t launchGather(const GatherParams& p, const int* tileList, int numListed, int numSMs, cudaStream_t stream);
// TMA-staged tiles of one box class.  tensorMap: a CUtensorMap (128 bytes, by value) describing the source plane
// with the staging box of (p.kernelSize, boxClass).  BORDER_WRAP only (tiles touching a border are never staged).
cudaError_t launchGatherStaged(const GatherParams& p, const StagedParams& sp, const void* tensorMap, int boxClass,
                               int numSMs, cudaStream_t stream);
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
