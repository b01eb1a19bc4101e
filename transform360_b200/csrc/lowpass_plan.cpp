// Host planner for the segmented low-pass: which rectangle of the input plane gets which separable
// Gaussian.  Behavioural spec: reference VideoFrameTransform.cpp:367-501 (calcualteFilteringConfig),
// :318-364 (band walk), :210-297 (per-band tiles and view-dependent kernel scaling), :78-94 (taps),
// :126-170 (sampling-density model).  Integer truncations and float/double mixing follow the reference
// expression by expression: the tap COUNT is int(2*sigma)*2+1, so a last-bit difference in sigma can
// change a kernel from 5 to 7 taps (hard part 5).  Compile with -ffp-contract=off.
//
// Structure differs from the reference: the plan is a flat list of segments referring to a shared tap
// pool (bands that reuse the same kernel share storage), produced by a small BandWalker.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include <stdexcept>

#include "host_plan.h"

namespace t360 {
namespace {

constexpr double kTiny = 1e-9;               // cpp:33
const double kViewFov = 0.5333 * M_PI;       // cpp:35
const double kWholeSphere = 4 * M_PI;        // cpp:34

// Appends a normalised Gaussian of int(2*sigma)*2+1 taps to `pool`; returns {offset, count}.  (cpp:78-94)
std::pair<int, int> appendGaussian(std::vector<float>& pool, float sigma) {
  const int half = static_cast<int>(sigma * 2);
  const int n = half * 2 + 1;
  const int offset = static_cast<int>(pool.size());
  // (sigma can come out hugely negative: sigma_x = sigma_y / (cosf(angle) + 1e-9) and cosf is -4e-8 at the pole; the
  // reference's cv::Mat::zeros(1, negative) throws there and generateMapForPlane returns 0 -- same here)
  if (n <= 0) throw std::runtime_error("low-pass kernel with a negative number of taps (negative sigma)");
  pool.resize(pool.size() + static_cast<size_t>(n));
  float* k = pool.data() + offset;
  float total = 0;
  const float falloff = std::abs(sigma) < kTiny ? 0 : static_cast<float>(0.5 / (sigma * sigma));
  for (int u = -half; u <= half; ++u) {
    const float w = expf(-(u * u * falloff));
    k[u + half] = w;
    total += w;
  }
  // the reference normalises with cv::Mat /= float, i.e. a float multiply by (float)(1.0 / total)
  const float inv = static_cast<float>(1.0 / static_cast<double>(total));
  for (int i = 0; i < n; ++i) k[i] = k[i] * inv;
  return {offset, n};
}

// --- sampling-density model of an off-centre cube view (cpp:126-170), all double -----------------
double greatCircle(double yaw1, double pitch1, double yaw2, double pitch2) {
  return std::acos(std::sin(pitch1) * std::sin(pitch2) + std::cos(pitch1) * std::cos(pitch2) * std::cos(yaw1 - yaw2));
}
double sampledArc(double offset, double renderedArc) {
  return M_PI - 2 * std::atan2(std::cos(0.5 * renderedArc) - offset, std::sin(0.5 * renderedArc));
}
double capArea(double angle) { return (1 - std::cos(0.5 * angle)) * 2 * M_PI; }

double densityRatio(double dist, double offset) {
  const double fov = kViewFov;
  double alongView;
  if (dist - kTiny > fov / 2) {
    if (dist + fov / 2 > M_PI) {
      const double nearEdge = sampledArc(offset, (2 * M_PI - dist - fov / 2) * 2) / 2;
      const double farEdge = sampledArc(offset, (dist - fov / 2) * 2) / 2;
      alongView = (2 * M_PI - nearEdge - farEdge) / fov;
    } else {
      alongView = (sampledArc(offset, 2 * dist + fov) - sampledArc(offset, 2 * dist - fov)) / 2 / fov;
    }
  } else {
    alongView = (sampledArc(offset, 2 * dist + fov) + sampledArc(offset, fov - 2 * dist)) / 2 / fov;
  }
  const double toCoVertex = greatCircle(dist, 0.5 * fov, 0.0, 0.0);
  const double acrossView = sampledArc(offset, toCoVertex * 2) / (toCoVertex * 2);
  return std::min(alongView * acrossView * capArea(fov) / kWholeSphere, 1.0);
}

class BandWalker {
 public:
  BandWalker(HostPlan& plan, int w, int h, float sigmaY) : plan_(plan), c_(plan.ctx), w_(w), h_(h), sigmaY_(sigmaY) {
    baseKy_ = appendGaussian(plan_.taps, sigmaY_);
    centreDensity_ = densityRatio(0.0, 0.0);
  }

  // One horizontal band [top, bottom] at latitude `angle` from the equator (cpp:210-297).
  void band(int top, int bottom, float angle) {
    const float sigmaX = static_cast<float>(std::min(0.5 * w_, sigmaY_ / (std::cos(angle) + kTiny)));
    const std::pair<int, int> baseKx = appendGaussian(plan_.taps, sigmaX);
    const int columns = c_.adjust_kernel ? c_.num_horizontal_segments : 1;
    if (columns <= 0) return;  // the reference's loop (cpp:235) does not run: a band without tiles
    const int tileW = static_cast<int>(std::ceil(1.0 * w_ / columns));
    for (int i = 0; i < columns && i * tileW < w_; ++i) {
      LowPassSegment s{};
      s.left = i * tileW;
      s.top = top;
      s.width = std::min(tileW, w_ - i * tileW);
      s.height = bottom - top + 1;
      if (c_.adjust_kernel) {
        // angular distance from the tile centre to the viewing direction (or, with zero yaw/pitch and
        // a non-trivial off-centre vector, to the direction opposite that vector), cpp:256-283
        const float tileYaw = static_cast<float>(2 * M_PI * ((i * tileW + 0.5 * s.width) - 0.5 * w_) / w_);
        const float tilePitch = static_cast<float>(0.5 * M_PI * (h_ - top - bottom) / h_);
        float yaw = static_cast<float>(c_.fixed_yaw * M_PI / 180.0f);
        float pitch = static_cast<float>(c_.fixed_pitch * M_PI / 180.0f);
        float offset = std::abs(c_.fixed_cube_offcenter_z);
        const float ox = c_.fixed_cube_offcenter_x, oy = c_.fixed_cube_offcenter_y, oz = c_.fixed_cube_offcenter_z;
        if (std::abs(yaw) < kTiny && std::abs(pitch) < kTiny &&
            (std::abs(ox) > kTiny || std::abs(oy) > kTiny || oz > kTiny)) {
          offset = sqrtf(ox * ox + oy * oy + oz * oz);
          yaw = atan2f(-ox / offset, -oz / offset);
          pitch = asinf(-oy / offset);
        }
        const double dist = greatCircle(yaw, pitch, tileYaw, tilePitch);
        const double scale = c_.kernel_adjust_factor * centreDensity_ / densityRatio(dist, offset);
        const auto kx = appendGaussian(plan_.taps, static_cast<float>(scale * sigmaX));
        const auto ky = appendGaussian(plan_.taps, static_cast<float>(scale * sigmaY_));
        s.kxOffset = kx.first; s.kxCount = kx.second;
        s.kyOffset = ky.first; s.kyCount = ky.second;
      } else {
        s.kxOffset = baseKx.first; s.kxCount = baseKx.second;
        s.kyOffset = baseKy_.first; s.kyCount = baseKy_.second;
      }
      plan_.segments.push_back(s);
    }
  }

  // Bands from the equator outwards: first up to the north pole, then down to the south pole (cpp:318-364).
  // The two halves measure the band centre with different integer expressions on purpose.
  void halves(int firstTopBelow, int firstBottomAbove, int bandH) {
    for (int bottom = firstBottomAbove; bottom >= 0; bottom -= bandH) {
      const int top = std::max(bottom - bandH + 1, 0);
      band(top, bottom, static_cast<float>(0.5 * M_PI * (h_ - top - bottom) / h_));
    }
    for (int top = firstTopBelow; top < h_; top += bandH) {
      const int bottom = std::min(top + bandH - 1, h_ - 1);
      band(top, bottom, static_cast<float>(0.5 * M_PI * (top + bottom - h_) / h_));
    }
  }

 private:
  HostPlan& plan_;
  const FrameTransformContext& c_;
  int w_, h_;
  float sigmaY_;
  std::pair<int, int> baseKy_;
  double centreDensity_;
};

}  // namespace

bool buildLowPassPlan(HostPlan& plan) {
  const FrameTransformContext& c = plan.ctx;
  plan.segments.clear();
  plan.taps.clear();
  // one eye only: the same tiles are applied to both halves of a stereo frame (cpp:377-401)
  int inW = plan.inW, inH = plan.inH, outW = plan.mapW, outH = plan.mapH;
  if (c.input_stereo_format == STEREO_FORMAT_LR) inW = static_cast<int>(inW * 0.5);
  else if (c.input_stereo_format == STEREO_FORMAT_TB) inH = static_cast<int>(inH * 0.5);
  if (c.output_stereo_format == STEREO_FORMAT_LR) outW = static_cast<int>(outW * 0.5);
  else if (c.output_stereo_format == STEREO_FORMAT_TB) outH = static_cast<int>(outH * 0.5);

  float hFov, vFov;  // degrees covered by the output layout (cpp:404-446)
  switch (c.output_layout) {
    case LAYOUT_CUBEMAP_32: case LAYOUT_EAC_32: hFov = 270.0f; vFov = 180.0f; break;
    case LAYOUT_CUBEMAP_23_OFFCENTER: hFov = 180.0f; vFov = 270.0f; break;
    case LAYOUT_FLAT_FIXED: hFov = c.fixed_hfov; vFov = c.fixed_vfov; break;
    case LAYOUT_EQUIRECT: hFov = 360.0f; vFov = 180.0f; break;
    case LAYOUT_BARREL: case LAYOUT_BARREL_SPLIT: hFov = 450.0f; vFov = 90.0f; break;
    default: std::printf("Invalid layout type %d.\n", static_cast<int>(c.output_layout)); return false;
  }
  // vertical sigma = half the input-pixels-per-output-pixel ratio, clamped (cpp:448-454)
  const float sigmaY =
      0.5f * std::min(c.max_kernel_half_height,
                      std::max(c.min_kernel_half_height,
                               c.kernel_height_scale_factor * std::min(inW / 360.0f, inH / 180.0f) /
                                   std::max(outW / hFov, outH / vFov)));
  if (c.num_vertical_segments < 1) {
    std::printf("num_vertical_segments must be positive.\n");
    return false;
  }
  BandWalker walker(plan, inW, inH, sigmaY);
  const int bandH = static_cast<int>(std::ceil(1.0 * inH / c.num_vertical_segments));
  if (bandH < 1) return false;
  if (c.num_vertical_segments % 2 == 0) {
    walker.halves(static_cast<int>(0.5 * inH), static_cast<int>(0.5 * inH - 1), bandH);
  } else {
    const int top = static_cast<int>(0.5 * (inH - bandH));  // equator band first (cpp:474-487)
    const int bottom = top + bandH - 1;
    walker.band(top, bottom, 0);
    walker.halves(bottom + 1, top - 1, bandH);
  }
  return true;
}

}  // namespace t360
