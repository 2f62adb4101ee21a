// ssd_host.cpp -- host-only entry points of the C ABI (no GPU needed): default configuration,
// transformation builder, synthetic scene generator (host leg).
#include <cstring>
#include "../../include/ssd_gpu.h"
#include "host/transformation.h"
#include "host/geometricCalibration.h"
#include <exception>

extern "C"
{

int ssd_gpu_abi_version(void)
{
  return SSD_GPU_ABI_VERSION;
}

// Configuration (reference configuration.h:27-52) with the stream size as a parameter
void ssd_gpu_default_config(ssd_gpu_config *c, int32_t width, int32_t height)
{
  c->width = width;
  c->height = height;
  c->x_min = -0.6;
  c->x_max = 0.6;
  c->y_min = 0.1;
  c->y_max = 1.3;
  c->z_min = -0.1;
  c->z_max = 1.1;
  c->height_interval = 0.01;
  c->min_height_above_ground = 0.05;
  c->min_step_depth = 0.1;
  c->min_peak_points = 2000;
  c->reserved = 0;
}

int ssd_make_transform(const double world_pts[9], const double camera_pts[9], ssd_gpu_transform *out)
{
  if(!world_pts || !camera_pts || !out)
    return SSD_E_INVALID_ARG;
  try
  {
    stairs::GeometricTransformation::RefPoints w, c;
    for(int i = 0; i < 3; i++)
    {
      w[i] = stairs::Point3(world_pts[i * 3], world_pts[i * 3 + 1], world_pts[i * 3 + 2]);
      c[i] = stairs::Point3(camera_pts[i * 3], camera_pts[i * 3 + 1], camera_pts[i * 3 + 2]);
    }
    const stairs::GeometricTransformation t(w, c);
    *out = t.abi();
    return SSD_OK;
  }
  catch(const std::exception &)
  {
    return SSD_E_INVALID_ARG;
  }
}

int ssd_make_transform_ex(const double world_pts[9], const double camera_pts[9], ssd_gpu_transform *out, double a_inv[9])
{
  if(!world_pts || !camera_pts || !out || !a_inv)
    return SSD_E_INVALID_ARG;
  try
  {
    stairs::GeometricTransformation::RefPoints w, c;
    for(int i = 0; i < 3; i++)
    {
      w[i] = stairs::Point3(world_pts[i * 3], world_pts[i * 3 + 1], world_pts[i * 3 + 2]);
      c[i] = stairs::Point3(camera_pts[i * 3], camera_pts[i * 3 + 1], camera_pts[i * 3 + 2]);
    }
    const stairs::GeometricTransformation t(w, c);
    *out = t.abi();
    t.abiInverse(a_inv);
    return SSD_OK;
  }
  catch(const std::exception &)
  {
    return SSD_E_INVALID_ARG;
  }
}

int ssd_inverse3(const double a[9], double a_inv[9])
{
  if(!a || !a_inv)
    return SSD_E_INVALID_ARG;
  try
  {
    stairs::Matrix_<3> m;
    for(int i = 0; i < 3; i++)
      for(int j = 0; j < 3; j++)
        m.a[i][j] = a[i * 3 + j];
    const stairs::Matrix_<3> r = stairs::inverseMatrix3(m);
    for(int i = 0; i < 3; i++)
      for(int j = 0; j < 3; j++)
        a_inv[i * 3 + j] = r.a[i][j];
    return SSD_OK;
  }
  catch(const std::exception &)
  {
    return SSD_E_INVALID_ARG;
  }
}

int ssd_load_calibration(const char *directory, ssd_gpu_transform *out, double world_pts[9], double camera_pts[9])
{
  if(!out)
    return SSD_E_INVALID_ARG;
  try
  {
    stairs::GeometricTransformation::RefPoints w, c;
    const int st = (int)stairs::GeometricCalibration::loadPoints(directory ? directory : "", w, c);
    if(st != 0)
    {
      const stairs::GeometricTransformation identity;
      *out = identity.abi();
      return st;
    }
    for(int i = 0; i < 3; i++)
      for(int j = 0; j < 3; j++)
      {
        if(world_pts)
          world_pts[i * 3 + j] = w[i][j];
        if(camera_pts)
          camera_pts[i * 3 + j] = c[i][j];
      }
    const stairs::GeometricTransformation t(w, c);
    *out = t.abi();
    return SSD_OK;
  }
  catch(const std::exception &)
  {
    return SSD_E_INVALID_ARG;
  }
}

} // extern "C"
