// geometricCalibration.h -- the loading half of the reference's GeometricCalibration (geometricCalibration.h,
// geometricCalibration.cpp:73-98,127-141,185-203): "calibration-triangle" + "calibration-points" -> the
// GeometricTransformation whose 12 + 7 doubles the GPU path binds. The detection half (IR marks, OpenCV contours,
// RealSense deprojection: geometricCalibration.cpp:157-183,205-235) is the calibration tool and out of scope.
#pragma once
#include <array>
#include <string>
#include <vector>
#include "transformation.h"
#include "types.h"

namespace stairs
{

class GeometricCalibration
{
public:
  static const int numMarkers = 3;
  using MarkerPoints3_t = std::array<Point3f, numMarkers>;
  using PointSets_t = std::vector<MarkerPoints3_t>;

  // Like the reference: files in the current directory; on any failure the identity transformation
  // (geometricCalibration.cpp:199-202, transformation.h:51-55,107).
  static GeometricTransformation load();
  static GeometricTransformation load(const std::string &directory);

  enum LoadStatus
  {
    ok = 0,
    triangleMissing = 1,  // file absent / wrong header / value missing
    triangleInvalid = 2,  // CalibrationTriangle::isValid() false
    pointsMissing = 3     // "calibration-points" absent, wrong header or fewer than 10 rows
  };
  // the same with the reason, and the six reference points (external world / camera) it found
  static LoadStatus loadPoints(const std::string &directory, GeometricTransformation::RefPoints &world, GeometricTransformation::RefPoints &camera);
};

} // namespace stairs
