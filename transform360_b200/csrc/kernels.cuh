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

constexpr int kBlurTileW = 64, kBlurTileH = 32;
constexpr int kBlurMaxSmem = 96 * 1024;

// Launchers: enqueue on `stream`, return the CUDA status of the launch.  Each counts the kernels it launches.
cudaError_t launchGather(const GatherParams& p, int numSMs, cudaStream_t stream);
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
