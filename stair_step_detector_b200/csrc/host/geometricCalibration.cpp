// geometricCalibration.cpp -- see geometricCalibration.h.
#include "geometricCalibration.h"
#include <fstream>
#include "calibrationTriangle.h"

namespace stairs
{

namespace
{

const int numIterations = 10; // geometricCalibration.cpp:38

// "x, y, z" (geometricCalibration.cpp:52-57)
std::istream &operator>>(std::istream &is, Point3f &p)
{
  char ch;
  is >> p.x >> ch >> p.y >> ch >> p.z;
  return is;
}

// geometricCalibration.cpp:73-98: header line, then rows "p0; p1; p2"; exactly ten rows are needed
bool load_point_sets(const std::string &path, GeometricCalibration::PointSets_t &sets)
{
  sets.reserve(numIterations);
  std::ifstream file(path);
  std::string id;
  std::getline(file, id);
  if(id != "calibration points")
    return false;
  while(true)
  {
    char ch;
    GeometricCalibration::MarkerPoints3_t mp;
    file >> mp[0] >> ch >> mp[1] >> ch >> mp[2];
    if(!file)
      break;
    sets.push_back(mp);
    if((int)sets.size() == numIterations)
      return true;
  }
  return false;
}

std::string join(const std::string &dir, const char *name)
{
  if(dir.empty())
    return name;
  return dir.back() == '/' ? dir + name : dir + "/" + name;
}

} // namespace

GeometricCalibration::LoadStatus GeometricCalibration::loadPoints(const std::string &directory, GeometricTransformation::RefPoints &world,
                                                                  GeometricTransformation::RefPoints &camera)
{
  CalibrationTriangle triangle;
  if(triangle.load(join(directory, "calibration-triangle")))
    return triangleMissing;
  if(!triangle.isValid())
    return triangleInvalid;
  PointSets_t sets;
  if(!load_point_sets(join(directory, "calibration-points"), sets))
    return pointsMissing;
  // calcAverageRefPointSet (geometricCalibration.cpp:127-141): double sums of the float samples in file order, / count
  for(int m = 0; m < numMarkers; m++)
  {
    Point3 avg;
    for(const MarkerPoints3_t &mp : sets)
    {
      avg.x += (double)mp[m].x;
      avg.y += (double)mp[m].y;
      avg.z += (double)mp[m].z;
    }
    const double n = (double)sets.size();
    avg.x /= n;
    avg.y /= n;
    avg.z /= n;
    camera[m] = avg;
    world[m] = triangle.getTriangleCorners()[m];
  }
  return ok;
}

GeometricTransformation GeometricCalibration::load()
{
  return load("");
}

GeometricTransformation GeometricCalibration::load(const std::string &directory)
{
  GeometricTransformation::RefPoints world, camera;
  if(loadPoints(directory, world, camera) == ok)
    return { world, camera };
  return {};
}

} // namespace stairs
