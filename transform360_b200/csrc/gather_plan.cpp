// Host side of the gather plan: job classification and record layout (see gather_plan.h, kernels.cuh).
#include "gather_plan.h"

#include <algorithm>
#include <climits>
#include <cstdint>
#include <thread>

namespace t360 {
namespace {

// fn(begin, end) over [0, n) on up to 32 host threads (the work items are independent and write disjoint outputs)
template <class F>
void parallelRanges(int n, size_t workPerItem, F fn) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  if (nt > 32) nt = 32;
  if (static_cast<size_t>(n) * workPerItem < (1u << 16)) nt = 1;
  const int chunk = (n + static_cast<int>(nt) - 1) / static_cast<int>(nt);
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < nt; ++t) {
    const int b = static_cast<int>(t) * chunk, e = std::min(n, b + chunk);
    if (b < e) pool.emplace_back(fn, b, e);
  }
  fn(0, std::min(n, chunk));
  for (auto& th : pool) th.join();
}

// Device order of the sampling records.  Each row is cut into segments of 32 pixels (= the width of a gather
// tile = one warp); inside a segment the pixels are dealt to LANES so that the lanes which one shared-memory pass
// serves together (8 for the 128-bit weight loads of cubic / Lanczos, 16 for the 64-bit ones of bilinear) ask for
// different bank groups of the weight table: sort the pixels by (bank group, phase), then deal them round-robin
// over the passes.  The window reads are unaffected (the warp still touches the same 32 windows) and the stores
// still fill one 32-byte sector.  The pixel's column inside the segment travels in the record's top 5 bits.
void buildLaneOrder(const HostPlan& h, std::vector<int2>& out, int tilesPerRow, int tileH, const std::vector<int>& seamBoxX) {
  const int k = h.kernelSize;
  const int groups = weightBankGroups(k), lanesPerPass = weightLanesPerPass(k), passes = 32 / lanesPerPass;
  constexpr int kRows = 4;  // rows per thread: one lane order per 32 x 4 block, so that a thread keeps ONE column
  parallelRanges((h.mapH + kRows - 1) / kRows, static_cast<size_t>(h.mapW) * kRows, [&](int blockBegin, int blockEnd) {
  for (int yb = blockBegin * kRows; yb < blockEnd * kRows; yb += kRows) {
    for (int x0 = 0; x0 < h.mapW; x0 += 32) {
      const int n = std::min(32, h.mapW - x0);
      int order[32];
      for (int i = 0; i < n; ++i) order[i] = i;
      const bool deal = k >= 2 && n == 32;
      if (deal) {
        // the bank group depends on fracX only, and fracX is the same down a column wherever the source column
        // does not depend on the output row (the four equatorial cube faces): order by the block's first row
        const SamplePoint* row = &h.samples[static_cast<size_t>(yb) * h.mapW];
        auto keyOf = [&](int c) {
          const int phase = row[x0 + c].rowPhase & 1023;
          return ((weightSlotOf(k, phase) & (groups - 1)) << 10) | phase;
        };
        std::stable_sort(order, order + n, [&](int a, int b) { return keyOf(a) < keyOf(b); });
      }
      for (int y = yb; y < std::min(h.mapH, yb + kRows); ++y) {
        const SamplePoint* row = &h.samples[static_cast<size_t>(y) * h.mapW];
        // tile-major: the records of tile (ty, tx) are contiguous, [rowInTile][lane]
        const size_t tile = static_cast<size_t>(y / tileH) * tilesPerRow + x0 / 32;
        int2* dst = &out[(tile * tileH + y % tileH) * 32];
        for (int i = 0; i < n; ++i) {
          // i-th pixel of the sorted sequence -> pass i % passes, position i / passes inside the pass
          const int lane = deal ? (i % passes) * lanesPerPass + i / passes : i;
          const int c = order[i];
          const SamplePoint& sp = row[x0 + c];
          int col0 = sp.col0;
          if (seamBoxX[tile] >= 0) {  // seam tile: first column relative to the unwrapped box (kernels.cuh, kJobSeam)
            int cw = col0 % h.inW;
            if (cw < 0) cw += h.inW;
            col0 = seamBoxX[tile] + (cw - seamBoxX[tile] + h.inW) % h.inW;
          }
          dst[lane] = int2{static_cast<int>((static_cast<unsigned>(col0) & ((1u << kRecordColumnShift) - 1)) |
                                                 (static_cast<unsigned>(c) << kRecordColumnShift)),
                                sp.rowPhase};
        }
      }
    }
  }
  });
}

// Splits the output plane into CTA tiles and finds, per tile, the bounding box of all source windows.  A tile is
// "staged" when that box lies inside the plane (no BORDER_WRAP needed) and fits the fixed TMA box; its box is
// anchored at a 16-byte aligned column.  Everything else is listed for the general (L1) kernel.
void buildGatherTiles(const HostPlan& h, GatherPlan& d, std::vector<int>& seamBoxX) {
  const int k = h.kernelSize, tw = kGatherTileW, th = gatherTileH(k);
  const int tilesX = (h.mapW + tw - 1) / tw, tilesY = (h.mapH + th - 1) / th;
  // seam tiles need whole 16-byte columns on both sides of the border and a plane much wider than the box
  const bool seamPossible = h.inW % 16 == 0 && h.inW >= 4 * stageBoxW(k, 0);
  std::vector<StagedTile> perTile(static_cast<size_t>(tilesX) * tilesY);  // classified in parallel, collected in raster order
  parallelRanges(tilesY, static_cast<size_t>(h.mapW) * th, [&](int tyBegin, int tyEnd) {
  for (int ty = tyBegin; ty < tyEnd; ++ty)
    for (int tx = 0; tx < tilesX; ++tx) {
      int minC = INT32_MAX, maxC = INT32_MIN, minR = INT32_MAX, maxR = INT32_MIN;
      const int y1 = std::min(h.mapH, (ty + 1) * th), x1 = std::min(h.mapW, (tx + 1) * tw);
      for (int y = ty * th; y < y1; ++y) {
        const SamplePoint* row = &h.samples[static_cast<size_t>(y) * h.mapW];
        for (int x = tx * tw; x < x1; ++x) {
          const int c = row[x].col0, r = row[x].rowPhase >> 10;
          minC = std::min(minC, c); maxC = std::max(maxC, c);
          minR = std::min(minR, r); maxR = std::max(maxR, r);
        }
      }
      const int boxX = minC >= 0 ? (minC & ~15) : -1;
      const bool inPlane = minC >= 0 && minR >= 0 && maxC + k <= h.inW && maxR + k <= h.inH;
      int cls = -1;
      for (int c = 0; c < kNumBoxClasses && inPlane && cls < 0; ++c)
        if (maxC + k - boxX <= stageBoxW(k, c) && maxR + k - minR <= stageBoxH(k, c)) cls = c;
      // warps (4 rows each) whose every pixel column keeps its source column down the 4 rows, 1-2 source rows apart:
      // they slide one register window down the column (gatherColumnShared) instead of fetching 4 windows
      int shareMask = 0;
      for (int w = 0; k >= 4 && w < th / 4; ++w) {
        const int ya = ty * th + 4 * w;
        bool ok = ya + 4 <= h.mapH;
        for (int x = tx * tw; ok && x < x1; ++x)
          for (int j = 1; j < 4 && ok; ++j) {
            const SamplePoint &a = h.samples[static_cast<size_t>(ya + j - 1) * h.mapW + x], &b = h.samples[static_cast<size_t>(ya + j) * h.mapW + x];
            const int d = (b.rowPhase >> 10) - (a.rowPhase >> 10);
            ok = (d == 1 || d == 2) && b.col0 == h.samples[static_cast<size_t>(ya) * h.mapW + x].col0;
          }
        if (ok) shareMask |= 1 << w;
      }
      // windows that cross the left/right border only (BORDER_WRAP): do they fit a class-0 box that wraps around it?
      int wrappedBoxX = -1;
      if (cls < 0 && seamPossible && minR >= 0 && maxR + k <= h.inH && maxR + k - minR <= stageBoxH(k, 0)) {
        const int W = h.inW, half = W / 2;  // columns rotated by half a plane: the border is in the middle of the range
        int lo = INT32_MAX, hi = INT32_MIN;
        for (int y = ty * th; y < y1; ++y)
          for (int x = tx * tw; x < x1; ++x) {
            int cw = h.samples[static_cast<size_t>(y) * h.mapW + x].col0 % W;
            if (cw < 0) cw += W;
            const int rot = cw + half >= W ? cw + half - W : cw + half;
            lo = std::min(lo, rot); hi = std::max(hi, rot);
          }
        const int first = lo - half < 0 ? lo - half + W : lo - half;  // leftmost first column, in plane coordinates
        const int bx = first & ~15;
        if (hi - lo + (first - bx) + k <= stageBoxW(k, 0) && bx + stageBoxW(k, 0) > W) wrappedBoxX = bx;
      }
      StagedTile& job = perTile[static_cast<size_t>(ty) * tilesX + tx];
      if (cls >= 0) {
        job = StagedTile{tx * tw, ty * th | (cls << kJobKindShift), boxX | (minR << 16), shareMask};
      } else if (wrappedBoxX >= 0) {
        job = StagedTile{tx * tw, ty * th | (kJobSeam << kJobKindShift), wrappedBoxX | (minR << 16), shareMask};
        seamBoxX[static_cast<size_t>(ty) * tilesX + tx] = wrappedBoxX;
      } else {
        job = StagedTile{tx * tw, ty * th | (kJobGeneral << kJobKindShift), 0, 0};
      }
    }
  });
  std::vector<StagedTile> staged[kNumBoxClasses];
  std::vector<StagedTile> fallback, seam;
  for (const StagedTile& job : perTile) {
    const int kind = (job.outY >> kJobKindShift) & kJobKindMask;
    (kind == kJobGeneral ? fallback : kind == kJobSeam ? seam : staged[kind]).push_back(job);
  }
  // order: general tiles, seam tiles, then the wide-box class, then the common class (see gatherFrameKernel)
  d.numGeneral = static_cast<int>(fallback.size());
  d.numSeam = static_cast<int>(seam.size());
  d.jobs = fallback;
  d.jobs.insert(d.jobs.end(), seam.begin(), seam.end());
  for (int c = kNumBoxClasses - 1; c >= 0; --c) {
    d.numStaged[c] = static_cast<int>(staged[c].size());
    d.jobs.insert(d.jobs.end(), staged[c].begin(), staged[c].end());
  }
}

}  // namespace

void buildGatherPlan(const HostPlan& h, bool stageTiles, GatherPlan& g) {
  g = GatherPlan{};
  if (h.kernelSize <= 0) return;
  g.tileH = gatherTileH(h.kernelSize);
  g.tilesPerRow = (h.mapW + kGatherTileW - 1) / kGatherTileW;
  g.tileRows = (h.mapH + g.tileH - 1) / g.tileH;
  std::vector<int> seamBoxX(static_cast<size_t>(g.tilesPerRow) * g.tileRows, -1);  // per tile; >= 0: seam tile
  if (stageTiles) buildGatherTiles(h, g, seamBoxX);
  g.records.assign(static_cast<size_t>(g.tilesPerRow) * g.tileRows * g.tileH * kGatherTileW, int2{0, 0});
  buildLaneOrder(h, g.records, g.tilesPerRow, g.tileH, seamBoxX);
}

}  // namespace t360
