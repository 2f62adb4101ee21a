// Stairs::serialize -- the wire format consumed by print-stairs.py:54-72 and the ROS wrapper
// (reference stairs.cpp:34-70). printf("%.3f") formats exactly like ostream << fixed << setprecision(3).
#include "stairs.h"
#include "../../../include/ssd_gpu.h"
#include <cstdio>

namespace
{

void appendNumber(std::string &s, double v)
{
  char tmp[352];
  const int n = std::snprintf(tmp, sizeof(tmp), "%.3f", v);
  s.append(tmp, size_t(n));
}

void appendStep(std::string &s, double height, const double quad[4][2])
{
  s += "[[\"height\",";
  appendNumber(s, height);
  s += "],[\"quadrilateral\"";
  for(int c = 0; c < 4; c++)
  {
    s += ",[";
    appendNumber(s, quad[c][0]);
    s += ',';
    appendNumber(s, quad[c][1]);
    s += ']';
  }
  s += "]]";
}

std::string serializeSteps(const ssd_gpu_step *steps, int n)
{
  std::string s = "[\"stairs\",[\"stairSteps\"," + std::to_string(n) + "]";
  if(n > 0)
  {
    s += ",[";
    for(int i = 0; i < n; i++)
    {
      if(i)
        s += ',';
      appendStep(s, steps[i].height, steps[i].quad);
    }
    s += ']';
  }
  s += ']';
  return s;
}

} // namespace

namespace stairs
{

std::string Stairs::serialize() const
{
  std::vector<ssd_gpu_step> tmp(stairSteps.size());
  for(size_t i = 0; i < stairSteps.size(); i++)
  {
    tmp[i].height = stairSteps[i].height;
    for(int c = 0; c < 4; c++)
    {
      tmp[i].quad[c][0] = stairSteps[i].quadrilateral[c].x;
      tmp[i].quad[c][1] = stairSteps[i].quadrilateral[c].y;
    }
  }
  return serializeSteps(tmp.data(), int(tmp.size()));
}

} // namespace stairs

extern "C" int ssd_stairs_serialize(const ssd_gpu_step *steps, int n, char *buf, size_t cap)
{
  if(n < 0 || (n > 0 && !steps))
    return SSD_E_INVALID_ARG;
  const std::string s = serializeSteps(steps, n);
  if(buf && cap)
  {
    const size_t m = s.size() < cap - 1 ? s.size() : cap - 1;
    s.copy(buf, m);
    buf[m] = 0;
  }
  return int(s.size());
}
