// ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Builds the UNMODIFIED reference translation units of the hot path (pointcloud.cpp, transformation.cpp,
// segmentation.cpp, quadrilateralTest.cpp, stairs.cpp, window.cpp) by path from /root/reference against the
// shim headers in oracle/shim/, and exposes the stage classes of pointcloud.cpp's anonymous namespace through
// a small C API (ssd_ref_*). Output goes to oracle/_ref/ only (git-ignored). Nothing here is copied from the
// reference: the sources are #included / compiled where they lie. Build recipe: oracle/Makefile.
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load the resulting library.
#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <numeric>
#include <optional>
#include <span>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../include/ssd_gpu.h"
#include "opencv2/imgproc.hpp"

#ifndef SSD_REF_DIR
#error "SSD_REF_DIR must point at the reference checkout"
#endif
#define SSD_STR2(x) #x
#define SSD_STR(x) SSD_STR2(x)
#define SSD_REF_FILE(name) SSD_STR(SSD_REF_DIR/name)

// reach the private stage helpers (getPointsInQuadrilateral, _camera, ...) without editing the reference
#define private public
#define protected public
#include SSD_REF_FILE(pointcloud.cpp)
#undef private
#undef protected
#include SSD_REF_FILE(geometricCalibration.h) // GeometricCalibration::load (the TU itself is compiled by path, see Makefile)
#include SSD_REF_FILE(calibrationMark.h)

// assert() in the reference aborts the process; turn it into an exception the harness can report.
struct ssd_ref_assert_failure : std::runtime_error
{
  using std::runtime_error::runtime_error;
};
extern "C" __attribute__((visibility("hidden"))) void __assert_fail(const char *expr, const char *file, unsigned line, const char *)
{
  throw ssd_ref_assert_failure(std::string("assert(") + expr + ") " + file + ":" + std::to_string(line));
}

namespace stairs
{
// GL overlay (drawing.h:57): the GL calls are stubbed; the arguments drawStairStep (pointcloud.cpp:588-597) hands
// over -- the projected quadrilateral, the labelling corners and the z label -- are recorded per thread.
struct ssd_ref_overlay_call
{
  float px[4][2];
  double label[2][2];
  double z_label;
};
thread_local std::vector<ssd_ref_overlay_call> g_overlay_calls;
void drawQuadrilateral(const Quadrilateralf_t &q, const Quadrilateral_t &labeling, Coordinate_t zLabel)
{
  ssd_ref_overlay_call c{};
  for(int i = 0; i < 4; i++)
  {
    c.px[i][0] = q[i].x;
    c.px[i][1] = q[i].y;
  }
  for(int i = 0; i < 2; i++)
  {
    c.label[i][0] = labeling[i].x;
    c.label[i][1] = labeling[i].y;
  }
  c.z_label = zLabel;
  g_overlay_calls.push_back(c);
}
// The mark-detection half of geometricCalibration.cpp (calibration tool, never called here) refers to these:
void setDrawOffset(int, int) {}
void resetDrawOffset() {}
void drawSubframeRect(const Rect2i &, Rgb) {}
void drawMarker(const Point2 &, const Contour_t &) {}
void drawMarker(const Point2 &, const Point3f &) {}
CalibrationMark::Detections CalibrationMark::detect(const Image &) { throw std::logic_error("CalibrationMark::detect is outside the oracle"); }
}

namespace
{
using namespace stairs;

thread_local std::string g_err;

const ProcessingConfiguration &cfg()
{
  return __processingConfiguration;
}

struct RefWorld
{
  Window window{ "ssd_ref" };
  GeometricTransformation trans; // identity; overwritten per call
};

void setTransform(GeometricTransformation &t, const ssd_gpu_transform &xf)
{
  auto &cam = const_cast<Transformation &>(t._camera);
  for(int i = 0; i < 3; i++)
    for(int j = 0; j < 3; j++)
      cam._a.a[i][j] = xf.a[i * 3 + j];
  cam._b = Point3(xf.b[0], xf.b[1], xf.b[2]);
  cam._aInv = boost::qvm::inverse(cam._a);
  auto &ext = const_cast<Transformation2D &>(t._toExternalWorld._world);
  for(int i = 0; i < 2; i++)
    for(int j = 0; j < 2; j++)
      ext._a.a[i][j] = xf.ext_a[i * 2 + j];
  ext._b = Point2(xf.ext_b[0], xf.ext_b[1]);
  const_cast<Coordinate_t &>(t._toExternalWorld._worldZ) = xf.ext_z;
}

void getTransform(const GeometricTransformation &t, ssd_gpu_transform &xf)
{
  for(int i = 0; i < 3; i++)
    for(int j = 0; j < 3; j++)
      xf.a[i * 3 + j] = t._camera._a.a[i][j];
  xf.b[0] = t._camera._b.x;
  xf.b[1] = t._camera._b.y;
  xf.b[2] = t._camera._b.z;
  for(int i = 0; i < 2; i++)
    for(int j = 0; j < 2; j++)
      xf.ext_a[i * 2 + j] = t._toExternalWorld._world._a.a[i][j];
  xf.ext_b[0] = t._toExternalWorld._world._b.x;
  xf.ext_b[1] = t._toExternalWorld._world._b.y;
  xf.ext_z = t._toExternalWorld._worldZ;
}

// intrinsics of the stub depth frame (only DepthFrame::project reads them): default fx = fy = width, centre principal point
thread_local bool g_intr_set = false;
thread_local float g_intr[4];

Camera::DepthFrame makeFrame(const float *xyz)
{
  auto d = std::make_shared<rs2::frame_data>();
  d->width = cfg().streams.depth.width;
  d->height = cfg().streams.depth.height;
  d->vertices = reinterpret_cast<const rs2::vertex *>(xyz);
  d->intr.width = d->width;
  d->intr.height = d->height;
  if(g_intr_set)
  {
    d->intr.fx = g_intr[0];
    d->intr.fy = g_intr[1];
    d->intr.ppx = g_intr[2];
    d->intr.ppy = g_intr[3];
  }
  else
  {
    d->intr.fx = d->intr.fy = float(d->width);
    d->intr.ppx = d->width * 0.5f;
    d->intr.ppy = d->height * 0.5f;
  }
  return Camera::DepthFrame(rs2::depth_frame(rs2::frame(d)));
}

template<class F>
int guarded(F &&f)
{
  try
  {
    return f();
  }
  catch(const std::exception &e)
  {
    g_err = e.what();
    return SSD_E_STATE;
  }
}

} // namespace

extern "C"
{
#define SSD_API __attribute__((visibility("default")))

SSD_API const char *ssd_ref_last_error(void)
{
  return g_err.c_str();
}

SSD_API int ssd_ref_config(ssd_gpu_config *out)
{
  const auto &c = cfg();
  out->width = c.streams.depth.width;
  out->height = c.streams.depth.height;
  out->x_min = c.measuringRange.x.min;
  out->x_max = c.measuringRange.x.max;
  out->y_min = c.measuringRange.y.min;
  out->y_max = c.measuringRange.y.max;
  out->z_min = c.measuringRange.z.min;
  out->z_max = c.measuringRange.z.max;
  out->height_interval = c.heightInterval;
  out->min_height_above_ground = c.minHeightAboveGround;
  out->min_step_depth = c.minStepDepth;
  out->min_peak_points = 2000;
  out->reserved = 0;
  return SSD_OK;
}

// derived constants of ProcessingConfiguration / Projection2D (pointcloud.cpp:60-106) for oracle pinning
SSD_API int ssd_ref_derived(int *min_height, int *min_img_y_extent, double *xy_ratio, double *height_interval_reciprocal, int *n_bins)
{
  const auto &c = cfg();
  *min_height = c.minHeight;
  *min_img_y_extent = c.minImgYExtent;
  *xy_ratio = c.projection.getXYRatio();
  *height_interval_reciprocal = c.heightIntervalReciprocal;
  *n_bins = int(static_cast<size_t>((c.measuringRange.z.max - c.measuringRange.z.min) * c.heightIntervalReciprocal) + 1);
  return SSD_OK;
}

// Projection2D::worldToImage / imageToWorld (pointcloud.cpp:79-88) on n points
SSD_API int ssd_ref_world_to_image(const double *xy, int n, int *ixy)
{
  for(int i = 0; i < n; i++)
  {
    const Point2i p = cfg().projection.worldToImage(Point3(xy[i * 2], xy[i * 2 + 1], 0));
    ixy[i * 2] = p.x;
    ixy[i * 2 + 1] = p.y;
  }
  return SSD_OK;
}

SSD_API int ssd_ref_image_to_world(const double *pxy, int n, double *xy)
{
  for(int i = 0; i < n; i++)
  {
    const Point2 p = cfg().projection.imageToWorld(Point2(pxy[i * 2], pxy[i * 2 + 1]));
    xy[i * 2] = p.x;
    xy[i * 2 + 1] = p.y;
  }
  return SSD_OK;
}

SSD_API int ssd_ref_make_transform(const double world_pts[9], const double camera_pts[9], ssd_gpu_transform *out)
{
  return guarded([&]
                 {
                   GeometricTransformation::RefPoints w, c;
                   for(int i = 0; i < 3; i++)
                   {
                     w[i] = Point3(world_pts[i * 3], world_pts[i * 3 + 1], world_pts[i * 3 + 2]);
                     c[i] = Point3(camera_pts[i * 3], camera_pts[i * 3 + 1], camera_pts[i * 3 + 2]);
                   }
                   const GeometricTransformation t(w, c);
                   getTransform(t, *out);
                   return SSD_OK;
                 });
}

// GeometricCalibration::load() (geometricCalibration.cpp:185-203) on the files "calibration-triangle" and
// "calibration-points" of the CURRENT DIRECTORY (the reference resolves them there; the caller chdir()s).
SSD_API int ssd_ref_load_calibration(ssd_gpu_transform *out)
{
  return guarded([&]
                 {
                   const GeometricTransformation t = GeometricCalibration::load();
                   getTransform(t, *out);
                   return SSD_OK;
                 });
}

// One frame through the stage classes exactly as Pointcloud::process (pointcloud.cpp:608-626) chains them,
// keeping every intermediate.
SSD_API int ssd_ref_process(const ssd_gpu_transform *xf, const float *xyz, uint8_t *labels, uint32_t *hist_out, int hist_cap,
                            ssd_gpu_frame_info *info, ssd_gpu_plateau *plats, ssd_gpu_step *steps, char *line, size_t line_cap)
{
  return guarded([&]
                 {
                   RefWorld w;
                   setTransform(w.trans, *xf);
                   const Camera::DepthFrame frame = makeFrame(xyz);
                   const rs2::pointcloud pc;
                   stairs::g_overlay_calls.clear();
                   const int W = cfg().streams.depth.width, H = cfg().streams.depth.height;
                   const size_t N = size_t(W) * H;
                   std::memset(info, 0, sizeof(*info));
                   info->ground_index = info->first_valid_index = -1;

                   Points3_t points;
                   PointsHt_t pointsHt;
                   const PointsExtraction pointsEx(pc, frame, w.trans.cameraToWorld());
                   pointsEx.extract(points, pointsHt);

                   // pixel index of every in-range point: same functor, same compares, same order
                   std::vector<uint32_t> pixelOf;
                   pixelOf.reserve(points.size());
                   const auto *v = reinterpret_cast<const rs2::vertex *>(xyz);
                   const auto &r = cfg().measuringRange;
                   uint32_t nNonZero = 0;
                   for(size_t i = 0; i < N; i++)
                   {
                     if(!(v[i].z > 0))
                     {
                       if(labels)
                         labels[i] = SSD_LABEL_INVALID;
                       continue;
                     }
                     nNonZero++;
                     const Point3 p = w.trans.cameraToWorld()(v[i]);
                     const bool in = p.x > r.x.min && p.x < r.x.max && p.y > r.y.min && p.y < r.y.max && p.z > r.z.min && p.z < r.z.max;
                     if(labels)
                       labels[i] = in ? SSD_LABEL_REMAINDER : SSD_LABEL_OUT_OF_RANGE;
                     if(in)
                       pixelOf.push_back(uint32_t(i));
                   }
                   if(pixelOf.size() != points.size())
                     throw std::runtime_error("harness: in-range count mismatch");
                   info->n_nonzero = nNonZero;
                   info->n_in_range = uint32_t(points.size());

                   Histogram_t heightsHistogram;
                   Peaks_t histogramPeaks;
                   HeightsHistogram::calcHistogram(pointsHt, heightsHistogram, histogramPeaks);
                   info->n_bins = int(heightsHistogram.size());
                   if(hist_out)
                     for(int i = 0; i < hist_cap && i < int(heightsHistogram.size()); i++)
                       hist_out[i] = heightsHistogram[i];

                   const PlateausExtraction plateausEx(heightsHistogram, histogramPeaks);
                   Plateaus_t plateaus = plateausEx.extractPlateaus(pointsHt);
                   info->n_plateaus = int(plateaus.size());
                   if(plateaus.size() > SSD_GPU_MAX_PLATEAUS)
                     throw std::runtime_error("harness: more plateaus than SSD_GPU_MAX_PLATEAUS");

                   if(labels)
                     for(size_t k = 0; k < plateaus.size(); k++)
                       for(const PointHt &ph : plateaus[k].plateauPoints)
                         labels[pixelOf[ph.pointIdx]] = uint8_t(k);

                   const StairsDetector detector(w.window, w.trans, frame, points);
                   std::vector<uint32_t> sizes;
                   for(const Plateau &p : plateaus)
                     sizes.push_back(uint32_t(p.plateauPoints.size()));

                   Stairs stairs;
                   bool degenerate = false;
                   try
                   {
                     stairs = detector.detectStairs(plateaus);
                   }
                   catch(const std::invalid_argument &e)
                   {
                     degenerate = true; // the reference would terminate here (uncaught)
                     g_err = e.what();
                   }

                   // replay the bookkeeping of detectStairSteps (pointcloud.cpp:402-429) for the debug record
                   size_t i = 0, maxGround = 0;
                   for(; i < plateaus.size(); i++)
                   {
                     if(plateaus[i].height >= cfg().minHeight)
                       break;
                     if(maxGround < sizes[i])
                     {
                       maxGround = sizes[i];
                       info->ground_index = int(i);
                     }
                   }
                   const size_t firstOutlined = i;
                   for(; i < plateaus.size(); i++)
                     if(plateaus[i].valid && info->first_valid_index < 0)
                       info->first_valid_index = int(i);
                   if(info->first_valid_index < 0)
                     info->ground_index = info->ground_index; // ground exists but is never made valid

                   for(size_t k = 0; k < plateaus.size(); k++)
                   {
                     const Plateau &p = plateaus[k];
                     ssd_gpu_plateau &o = plats[k];
                     std::memset(&o, 0, sizeof(o));
                     o.height = p.height;
                     const Height_t pred = p.height - 1, succ = p.height + 1;
                     if(heightsHistogram[pred] > heightsHistogram[succ])
                     {
                       o.hmin = pred;
                       o.hmax = p.height;
                     }
                     else
                     {
                       o.hmin = p.height;
                       o.hmax = succ;
                     }
                     o.n_points = sizes[k];
                     o.valid = p.valid;
                     o.outlined = k >= firstOutlined;
                     o.quad_status = -1;
                     for(int c = 0; c < 4; c++)
                     {
                       o.quad_world[c][0] = p.quadriWorld2D[c].x;
                       o.quad_world[c][1] = p.quadriWorld2D[c].y;
                     }
                     o.mean_z = 0;
                     if(p.valid)
                     {
                       try
                       {
                         const PointsHt_t in = detector.getPointsInQuadrilateral(p.plateauPoints, p.quadriWorld2D);
                         o.n_in_quad = uint32_t(in.size());
                         o.mean_z = detector.calcAverageZ(in);
                         o.quad_status = 0;
                       }
                       catch(const std::invalid_argument &)
                       {
                         o.quad_status = 1;
                       }
                     }
                   }

                   info->n_steps = int(stairs.stairSteps.size());
                   for(int s = 0; s < info->n_steps && s < SSD_GPU_MAX_STEPS; s++)
                   {
                     steps[s].height = stairs.stairSteps[s].height;
                     for(int c = 0; c < 4; c++)
                     {
                       steps[s].quad[c][0] = stairs.stairSteps[s].quadrilateral[c].x;
                       steps[s].quad[c][1] = stairs.stairSteps[s].quadrilateral[c].y;
                     }
                   }
                   if(info->n_steps == 0)
                     info->status |= SSD_STATUS_NO_STEPS;
                   if(degenerate)
                     info->status |= SSD_STATUS_DEGENERATE_QUAD;
                   if(line && line_cap)
                   {
                     const std::string s = stairs.serialize();
                     std::snprintf(line, line_cap, "%s", s.c_str());
                   }
                   return SSD_OK;
                 });
}

// The drawQuadrilateral calls of the last ssd_ref_process on this thread (detectStairs draws every step twice:
// depth viewport with {corner 0, corner 1, z} of the world quadrilateral as the label, then infrared viewport with the
// external-world corners and height, pointcloud.cpp:367-368,388-392). px: n x 8 floats, label: n x 4 doubles, z: n.
SSD_API int ssd_ref_last_overlay(float *px, double *label, double *z_label, int cap, int *n)
{
  const auto &v = stairs::g_overlay_calls;
  *n = int(v.size());
  for(int i = 0; i < *n && i < cap; i++)
  {
    if(px)
      std::memcpy(px + 8 * i, v[i].px, sizeof(v[i].px));
    if(label)
      std::memcpy(label + 4 * i, v[i].label, sizeof(v[i].label));
    if(z_label)
      z_label[i] = v[i].z_label;
  }
  return SSD_OK;
}

// Intrinsics of the stub depth frame for the following calls on this thread (NULL: back to the default).
SSD_API int ssd_ref_set_intrinsics(const ssd_gpu_intrinsics *intr)
{
  g_intr_set = intr != nullptr;
  if(intr)
  {
    g_intr[0] = intr->fx;
    g_intr[1] = intr->fy;
    g_intr[2] = intr->ppx;
    g_intr[3] = intr->ppy;
  }
  return SSD_OK;
}

// Transformation_<3>::_aInv as the reference holds it for this transform (setTransform: boost::qvm::inverse(_a))
SSD_API int ssd_ref_a_inv(const ssd_gpu_transform *xf, double a_inv[9])
{
  return guarded([&]
                 {
                   RefWorld w;
                   setTransform(w.trans, *xf);
                   for(int i = 0; i < 3; i++)
                     for(int j = 0; j < 3; j++)
                       a_inv[i * 3 + j] = w.trans._camera._aInv.a[i][j];
                   return SSD_OK;
                 });
}

// The reference's own entry point, timed: Pointcloud::process per frame (prints suppressed), frames spread
// over n_threads host threads (threading lives in this harness only; the reference is single-threaded).
SSD_API int ssd_ref_process_timed(const ssd_gpu_transform *xf, const float *xyz, int n_frames, int n_threads, int repeat,
                                  double *seconds, int *n_failed)
{
  return guarded([&]
                 {
                   const size_t N = size_t(cfg().streams.depth.width) * cfg().streams.depth.height;
                   struct NullBuf : std::streambuf
                   {
                     int overflow(int c) override { return c; }
                     std::streamsize xsputn(const char *, std::streamsize n) override { return n; }
                   } nullBuf;
                   std::streambuf *old = std::cout.rdbuf(&nullBuf);
                   std::atomic<int> next{ 0 }, failed{ 0 };
                   if(n_threads < 1)
                     n_threads = 1;
                   const int total = n_frames * (repeat < 1 ? 1 : repeat);
                   auto worker = [&]
                   {
                     RefWorld w;
                     setTransform(w.trans, *xf);
                     const Pointcloud pointcloud(w.window, w.trans);
                     for(;;)
                     {
                       const int j = next.fetch_add(1);
                       if(j >= total)
                         break;
                       const Camera::DepthFrame frame = makeFrame(xyz + size_t(j % n_frames) * N * 3);
                       try
                       {
                         pointcloud.process(frame);
                       }
                       catch(const std::exception &)
                       {
                         failed++;
                       }
                     }
                   };
                   const auto t0 = std::chrono::steady_clock::now();
                   std::vector<std::thread> th;
                   for(int t = 1; t < n_threads; t++)
                     th.emplace_back(worker);
                   worker();
                   for(auto &t : th)
                     t.join();
                   const auto t1 = std::chrono::steady_clock::now();
                   std::cout.rdbuf(old);
                   *seconds = std::chrono::duration<double>(t1 - t0).count();
                   if(n_failed)
                     *n_failed = failed.load();
                   return SSD_OK;
                 });
}

SSD_API int ssd_ref_detect_outline(const uint8_t *image, int width, int height, int min_img_y_extent, double xy_ratio, double quad_px[8],
                                   int *valid)
{
  return guarded([&]
                 {
                   std::vector<uint8_t> copy(image, image + size_t(width) * height);
                   const Image img(width, height, copy.data(), width);
                   const Segmentation::Outline o = Segmentation::detectOutline(img, min_img_y_extent, xy_ratio, "ref");
                   for(int c = 0; c < 4; c++)
                   {
                     quad_px[c * 2] = o.quadrilateral[c].x;
                     quad_px[c * 2 + 1] = o.quadrilateral[c].y;
                   }
                   *valid = o.valid;
                   return SSD_OK;
                 });
}

SSD_API int ssd_ref_detect_front_edge(const uint8_t *image, int width, int height, double left_px[2], double right_px[2], int *valid)
{
  return guarded([&]
                 {
                   std::vector<uint8_t> copy(image, image + size_t(width) * height);
                   const Image img(width, height, copy.data(), width);
                   const Segmentation::FrontEdge e = Segmentation::detectFrontEdge(img, "ref");
                   left_px[0] = e.pointLeft.x;
                   left_px[1] = e.pointLeft.y;
                   right_px[0] = e.pointRight.x;
                   right_px[1] = e.pointRight.y;
                   *valid = e.valid;
                   return SSD_OK;
                 });
}

SSD_API int ssd_ref_close(uint8_t *image, int width, int height)
{
  return guarded([&]
                 {
                   const Image img(width, height, image, width);
                   cv::morphologyEx(img, img, cv::MORPH_CLOSE, cv::Mat());
                   return SSD_OK;
                 });
}

SSD_API int ssd_ref_points_in_quad(const double quad[8], const double *xy, int n, uint8_t *inside, int *ctor_status)
{
  return guarded([&]
                 {
                   const Quadrilateral_t q{ Point2(quad[0], quad[1]), Point2(quad[2], quad[3]), Point2(quad[4], quad[5]),
                                            Point2(quad[6], quad[7]) };
                   std::memset(inside, 0, size_t(n));
                   try
                   {
                     const QuadrilateralTest test(q);
                     for(int i = 0; i < n; i++)
                       inside[i] = test.isPointWithin(Point2(xy[i * 2], xy[i * 2 + 1]));
                     *ctor_status = 0;
                   }
                   catch(const std::invalid_argument &e)
                   {
                     g_err = e.what();
                     *ctor_status = 1;
                   }
                   return SSD_OK;
                 });
}

SSD_API int ssd_ref_camera_to_world(const ssd_gpu_transform *xf, const float *xyz, int n, double *world)
{
  return guarded([&]
                 {
                   GeometricTransformation t;
                   setTransform(t, *xf);
                   const auto *v = reinterpret_cast<const rs2::vertex *>(xyz);
                   for(int i = 0; i < n; i++)
                   {
                     const Point3 p = t.cameraToWorld()(v[i]);
                     world[i * 3] = p.x;
                     world[i * 3 + 1] = p.y;
                     world[i * 3 + 2] = p.z;
                   }
                   return SSD_OK;
                 });
}

SSD_API int ssd_ref_serialize(const ssd_gpu_step *steps, int n, char *buf, size_t cap)
{
  return guarded([&]
                 {
                   Stairs s;
                   for(int i = 0; i < n; i++)
                   {
                     Stairs::StairStep st;
                     st.height = steps[i].height;
                     for(int c = 0; c < 4; c++)
                       st.quadrilateral[c] = Point2(steps[i].quad[c][0], steps[i].quad[c][1]);
                     s.stairSteps.push_back(st);
                   }
                   const std::string str = s.serialize();
                   if(buf && cap)
                     std::snprintf(buf, cap, "%s", str.c_str());
                   return int(str.size());
                 });
}

} // extern "C"
