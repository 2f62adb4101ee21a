// calibrationTriangle.cpp -- see calibrationTriangle.h. File grammar (calibrationTriangle.cpp:32-35,49-71,97-125 of
// the reference): first line "calibration triangle", then whitespace-separated tokens; a value is the token after
// "<name> =", parsed by operator>> (so "x1 = -1.121, y1 = ..." leaves the comma for the next token).
#include "calibrationTriangle.h"
#include <fstream>

namespace stairs
{

namespace
{

template<typename ValueType>
bool read_value(std::ifstream &file, const std::string &name, ValueType &value)
{
  while(file)
  {
    std::string token;
    file >> token;
    if(token == name)
    {
      std::string sign;
      file >> sign;
      if(sign == "=")
      {
        file >> value;
        return static_cast<bool>(file);
      }
    }
  }
  return false;
}

} // namespace

int CalibrationTriangle::load()
{
  return load("calibration-triangle");
}

int CalibrationTriangle::load(const std::string &path)
{
  *this = {};
  std::ifstream file(path);
  std::string id;
  std::getline(file, id);
  if(id != "calibration triangle")
    return -1;
  bool ok = true;
  for(int n = 1; n <= 3; n++)
  {
    TriangleCorner &c = triangleCorners[n - 1];
    const std::string ns = std::to_string(n);
    ok = ok && read_value(file, "x" + ns, c.x);
    ok = ok && read_value(file, "y" + ns, c.y);
    ok = ok && read_value(file, "z" + ns, c.z);
  }
  std::string side;
  ok = ok && read_value(file, "lowerQuadrant", side);
  lowerQuadrant = side == "left" ? Side::left : (side == "right" ? Side::right : Side::undefined);
  return ok ? 0 : -2;
}

bool CalibrationTriangle::isValid() const
{
  const double minDistQu = 0.01 * 0.01; // minSideLength squared (calibrationTriangle.cpp:47,150)
  auto distQu = [](const TriangleCorner &a, const TriangleCorner &b) {
    const double dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
    return dx * dx + dy * dy + dz * dz;
  };
  if(distQu(triangleCorners[0], triangleCorners[1]) < minDistQu || distQu(triangleCorners[1], triangleCorners[2]) < minDistQu ||
     distQu(triangleCorners[2], triangleCorners[0]) < minDistQu)
    return false;
  return lowerQuadrant == Side::left || lowerQuadrant == Side::right;
}

} // namespace stairs
