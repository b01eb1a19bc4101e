// Device-independent part of the gather plan of one plane: how the output plane is cut into jobs for the persistent
// gather kernel and in which order the sampling records are laid out for it.  Pure host code (no CUDA call), so that
// the CPU test-suite can check it without a GPU (T360B200_hostPlanGather).  See kernels.cuh for the formats.
#pragma once

#include <vector>

#include "host_plan.h"
#include "kernels.cuh"

namespace t360 {

struct GatherPlan {
  int tilesPerRow = 0, tileRows = 0, tileH = 0;  // tile grid of the (scaled) output plane; tile = 32 x tileH pixels
  std::vector<int2> records;                     // tile-major, lane-ordered sampling records (kernels.cuh)
  std::vector<StagedTile> jobs;                  // general, seam, class 1, class 0 (empty: the plan is not staged)
  int numStaged[kNumBoxClasses] = {}, numSeam = 0, numGeneral = 0;
  int totalStaged() const {
    int n = numSeam;
    for (int c : numStaged) n += c;
    return n;
  }
};

// stageTiles: classify the tiles for the persistent kernel (kernel size >= 2 and BORDER_WRAP); otherwise only the
// records are produced (nearest neighbour, barrel layouts: whole-plane general kernels).
void buildGatherPlan(const HostPlan& h, bool stageTiles, GatherPlan& g);

}  // namespace t360
