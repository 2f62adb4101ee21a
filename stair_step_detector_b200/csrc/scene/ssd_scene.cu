// ssd_scene.cu -- libssd_scene.so: the synthetic input source (include/ssd_scene.h), host and device versions.
// Test / benchmark infrastructure, not part of the product library.
#include "scene_model.h"
#include <algorithm>
#include <cstring>
#include <cuda_runtime.h>

__global__ void k_synth_frames(ssd_scene base, uint64_t base_seed, long long first_index, int n_frames, int min_steps, int max_steps,
                               float *__restrict__ xyz, uint16_t *__restrict__ depth)
{
  __shared__ ssd_scene s;
  __shared__ ssd_scene_rt rt;
  const int f = blockIdx.y;
  if(threadIdx.x == 0)
  {
    if(min_steps > 0)
      ssd_scene_randomize_hd(&s, &base, base_seed, first_index + f, min_steps, max_steps);
    else
      s = base;
    ssd_scene_prepare(&s, &rt);
  }
  __syncthreads();
  const int N = s.width * s.height;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
  {
    const int v = i / s.width, u = i - v * s.width;
    const uint16_t d = ssd_scene_depth(&s, &rt, u, v);
    if(depth)
      depth[(size_t)f * N + i] = d;
    if(xyz)
    {
      float o[3];
      ssd_deproject_pixel(&s, u, v, d, o);
      float *dst = xyz + ((size_t)f * N + i) * 3;
      dst[0] = o[0];
      dst[1] = o[1];
      dst[2] = o[2];
    }
  }
}


extern "C"
{

void ssd_scene_default(ssd_scene *s, int32_t width, int32_t height)
{
  ssd_scene_default_hd(s, width, height);
}

void ssd_scene_randomize(ssd_scene *s, const ssd_scene *base, uint64_t base_seed, int64_t index, int min_steps, int max_steps)
{
  ssd_scene_randomize_hd(s, base, base_seed, index, min_steps, max_steps);
}

// Three marks on the calibration plane (z = 0 in scene coordinates), laid out like the reference's
// calibration-triangle file (top-left, top-right, bottom), and where the scene's camera sees them.
void ssd_scene_calibration_points(const ssd_scene *s, double world_pts[9], double camera_pts[9])
{
  ssd_scene_rt rt;
  ssd_scene_prepare(s, &rt);
  const double marks[3][3] = { { s->cam_x - 0.45, s->cam_y + 1.25, 0.0 }, { s->cam_x + 0.45, s->cam_y + 1.25, 0.0 },
                               { s->cam_x + 0.30, s->cam_y + 0.35, 0.0 } };
  for(int i = 0; i < 3; i++)
  {
    for(int j = 0; j < 3; j++)
      world_pts[i * 3 + j] = marks[i][j];
    ssd_scene_to_camera(&rt, marks[i], camera_pts + i * 3);
  }
}

void ssd_scene_intrinsics(const ssd_scene *s, ssd_gpu_intrinsics *out)
{
  memset(out, 0, sizeof(*out));
  out->fx = s->fx;
  out->fy = s->fy;
  out->ppx = s->ppx;
  out->ppy = s->ppy;
  out->depth_unit = s->depth_unit;
}

int ssd_synth_depth_host(const ssd_scene *s, uint16_t *depth_out)
{
  if(!s || !depth_out || s->width <= 0 || s->height <= 0)
    return SSD_E_INVALID_ARG;
  ssd_scene_rt rt;
  ssd_scene_prepare(s, &rt);
  for(int v = 0; v < s->height; v++)
    for(int u = 0; u < s->width; u++)
      depth_out[size_t(v) * s->width + u] = ssd_scene_depth(s, &rt, u, v);
  return SSD_OK;
}

int ssd_deproject_host(const ssd_scene *s, const uint16_t *depth, float *xyz_out)
{
  if(!s || !depth || !xyz_out)
    return SSD_E_INVALID_ARG;
  for(int v = 0; v < s->height; v++)
    for(int u = 0; u < s->width; u++)
    {
      const size_t i = size_t(v) * s->width + u;
      ssd_deproject_pixel(s, u, v, depth[i], xyz_out + i * 3);
    }
  return SSD_OK;
}


int ssd_scene_synth_frames_device(int device, const ssd_scene *base, uint64_t base_seed, int64_t first_index, int n_frames, int min_steps,
                                  int max_steps, float *xyz_dev, uint16_t *depth_dev)
{
  if(!base || (!xyz_dev && !depth_dev) || n_frames <= 0 || base->width <= 0 || base->height <= 0)
    return SSD_E_INVALID_ARG;
  if(cudaSetDevice(device) != cudaSuccess)
    return SSD_E_CUDA;
  const int N = base->width * base->height;
  const int bx = std::min((N + 255) / 256, 512);
  for(int f0 = 0; f0 < n_frames; f0 += 32768)
  {
    const int nf = std::min(32768, n_frames - f0);
    k_synth_frames<<<dim3(bx, nf), 256>>>(*base, base_seed, first_index + f0, nf, min_steps, max_steps, xyz_dev ? xyz_dev + (size_t)f0 * N * 3 : nullptr,
                                           depth_dev ? depth_dev + (size_t)f0 * N : nullptr);
  }
  if(cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
    return SSD_E_CUDA;
  return SSD_OK;
}

} // extern "C"
