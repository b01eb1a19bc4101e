// Device-independent part of the gather plan of one plane: how the output plane is cut into jobs for the persistent
// gather kernel and how the sampling records are laid out for it.  Pure host code (no CUDA call), so that the CPU
// test-suite can check it without a GPU (T360B200_hostPlanGather).  See kernels.cuh for the formats.
#pragma once

#include <cstdint>
#include <vector>

#include "host_plan.h"
#include "kernels.cuh"

namespace t360 {

struct GatherPlan {
  int tilesPerRow = 0, tileRows = 0, tileH = 0;  // grid of the FULL records: tiles of 32 x tileH pixels (general kernels)
  std::vector<int2> records;                     // full records: tile-major, lane-ordered (kernels.cuh)
  std::vector<GatherJob> jobs;                   // general, seam, class 1, share, class 0 (empty: the plan is not staged)
  std::vector<uint32_t> compact;                 // compact records of the staged jobs (GatherJob::recordOffset)
  std::vector<int> jobNeedRows;                  // per job: source rows [0, n) it reads (inH if a window wraps vertically)
  int numStaged[2] = {}, numSeam = 0, numGeneral = 0, numShare = 0;
  int totalStaged() const { return numSeam + numShare + numStaged[0] + numStaged[1]; }
};

// stageTiles: cut the plane into jobs for the persistent kernel (kernel size >= 2 and BORDER_WRAP); otherwise only the
// full records are produced (nearest neighbour, barrel layouts: whole-plane general kernels).
void buildGatherPlan(const HostPlan& h, bool stageTiles, GatherPlan& g);

// The output rectangle of a job {x0, y0, x1, y1} (clipped to the plane by the caller).
inline void jobOutputRect(const GatherJob& job, int k, int rect[4]) {
  const int kind = (job.outY >> kJobKindShift) & kJobKindMask, quad = (job.outX & kJobQuadMask) - 1;
  rect[0] = job.outX & ~kJobQuadMask;
  rect[1] = job.outY & kJobRowMask;
  if (quad >= 0) { rect[0] += 16 * (quad & 1); rect[1] += 16 * (quad >> 1); }
  const bool share = kind == kJobShare || kind == kJobShareStay;
  rect[2] = rect[0] + (quad >= 0 ? 16 : (share ? kShareW : kGatherTileW));
  rect[3] = rect[1] + (quad >= 0 ? 16 : (share ? shareH(k) : kFrameTileH));
}

// Launch order of a job list that is sorted by kind (general, class 1, share, class 0): the general jobs -- latency-bound
// reads through L1 that leave the shared-memory pipe idle -- are spread evenly over the first half of the staged jobs,
// so that they run beside them instead of all at once at the start of the launch.
void spreadGeneralJobs(std::vector<GatherJob>& jobs);


// Deals n <= 32 pixels of one warp step to lanes (and table copies) so that the lanes one shared-memory pass serves
// together ask for different bank groups of the weight table.  slot[i] = weightSlotOf(k, phase of pixel i).
// laneOf[i] = lane of pixel i, copyOf[i] = table copy it reads.  Returns the modelled wavefronts of one weight load.
int dealLanes(int k, int copies, int n, const int* slot, int* laneOf, int* copyOf);

// The shared-memory image of the interpolation weights the frame kernel copies in (layout: kernels.cuh, "Weight tables
// in shared memory"), from OpenCV's table int16 [1024][k][k].  weightImageBytes(k, weightCopies(k)) bytes.
std::vector<uint8_t> buildWeightImage(int k, const int16_t* table);

}  // namespace t360
