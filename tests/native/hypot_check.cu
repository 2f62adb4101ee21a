// Host-side evaluation of hypot_cr (ssd_device.cuh: the same function the kernels compile for the device) on operand pairs read
// from a file of doubles {x0, y0, x1, y1, ...}; writes hypot_cr and libm's hypot of every pair. Built and run by
// tests/test_hypot.py (nvcc, host code only; no GPU needed), which checks the results against exact arithmetic.
#include <cmath>
#include <cstdio>
#include <vector>
#include "../../stair_step_detector_b200/csrc/ssd_device.cuh"

int main(int argc, char **argv)
{
  if(argc != 3)
    return 2;
  FILE *f = std::fopen(argv[1], "rb");
  if(!f)
    return 3;
  std::vector<double> in;
  double buf[1024];
  size_t n;
  while((n = std::fread(buf, sizeof(double), 1024, f)) > 0)
    in.insert(in.end(), buf, buf + n);
  std::fclose(f);
  std::vector<double> out(in.size());
  for(size_t i = 0; i + 1 < in.size(); i += 2)
  {
    out[i] = hypot_cr(in[i], in[i + 1]);
    out[i + 1] = hypot(in[i], in[i + 1]);
  }
  f = std::fopen(argv[2], "wb");
  if(!f)
    return 4;
  std::fwrite(out.data(), sizeof(double), out.size(), f);
  std::fclose(f);
  return 0;
}
