// calibrationTriangle.h -- the reference's CalibrationTriangle surface (calibrationTriangle.h:27-51), loading part:
// the three marks' external-world coordinates from the text file "calibration-triangle"
// (calibrationTriangle.cpp:97-125). Host only; the calibration *tool* (mark detection, save) is out of scope.
#pragma once
#include <array>
#include <string>
#include "types.h"

namespace stairs
{

class CalibrationTriangle
{
public:
  // reads "calibration-triangle" in the current directory, like the reference; 0 ok, -1 no file / wrong header,
  // -2 a value is missing (calibrationTriangle.cpp:97-125)
  int load();
  int load(const std::string &path);
  bool isValid() const; // calibrationTriangle.cpp:148-168

  typedef Point3 TriangleCorner;
  typedef std::array<TriangleCorner, 3> TriangleCorners;
  enum class Side
  {
    undefined,
    left,
    right
  };

  const TriangleCorners &getTriangleCorners() { return triangleCorners; }
  const Side &getLowerQuadrant() { return lowerQuadrant; }

private:
  TriangleCorners triangleCorners{};
  Side lowerQuadrant{};
};

} // namespace stairs
