// Pointcloud: thin host driver over ssd_gpu_* (replaces reference pointcloud.cpp:602-626; the stage classes of
// its anonymous namespace are the CUDA kernels in ssd_kernels_points.cuh / ssd_kernels_outline.cuh).
#include "pointcloud.h"
#include "transformation.h"
#include "window.h"
#include <iostream>
#include <stdexcept>

namespace stairs
{

Pointcloud::Pointcloud(const Window &window, const GeometricTransformation &trans) : _window(window), _transformation(trans) {}

Pointcloud::Pointcloud(const Window &window, const GeometricTransformation &trans, const Configuration &config, int device, int maxFrames)
: _window(window), _transformation(trans), _config(config), _device(device), _maxFrames(maxFrames), _explicitConfig(true)
{
}

Pointcloud::~Pointcloud()
{
  ssd_gpu_destroy(_ctx);
}

void Pointcloud::ensureContext(int width, int height) const
{
  if(_ctx)
  {
    if(_config.streams.depth.width != width || _config.streams.depth.height != height)
      throw std::invalid_argument("Pointcloud: frame size differs from the configured depth stream");
    return;
  }
  if(!_explicitConfig)
    _config = Configuration(width, height); // the reference takes the size from its compile-time Configuration
  else if(_config.streams.depth.width != width || _config.streams.depth.height != height)
    throw std::invalid_argument("Pointcloud: frame size differs from the configured depth stream");
  const ssd_gpu_config cfg = _config.abi();
  const ssd_gpu_transform xf = _transformation.abi();
  if(ssd_gpu_create(&cfg, &xf, _device, _maxFrames, &_ctx) != SSD_OK)
    throw std::runtime_error(std::string("ssd_gpu_create: ") + ssd_gpu_last_error(nullptr)); // no CPU fallback
}

void Pointcloud::enableOverlay(const ssd_gpu_intrinsics &intrinsics) const
{
  _overlayIntrinsics = intrinsics;
  _overlay = true;
  _overlayApplied = false;
  _risersApplied = false;
}

std::vector<Quadrilateralf_t> Pointcloud::overlay(int frame) const
{
  if(!_ctx)
    throw std::logic_error("Pointcloud::overlay before any frame was processed");
  ssd_gpu_overlay q[SSD_GPU_MAX_STEPS];
  int n = 0;
  if(ssd_gpu_get_overlay(_ctx, frame, q, SSD_GPU_MAX_STEPS, &n) != SSD_OK)
    throw std::runtime_error(std::string("ssd_gpu_get_overlay: ") + ssd_gpu_last_error(_ctx));
  std::vector<Quadrilateralf_t> out(static_cast<size_t>(n));
  for(int i = 0; i < n; i++)
    for(int c = 0; c < 4; c++)
      out[static_cast<size_t>(i)][static_cast<size_t>(c)] = Point2f{ q[i].px[c][0], q[i].px[c][1] };
  return out;
}

std::vector<Stairs> Pointcloud::processBatch(const Camera::DepthFrame &frames, int nFrames, std::vector<unsigned> *status) const
{
  ensureContext(frames.width(), frames.height());
  if(_overlay && !_overlayApplied)
  {
    double aInv[9];
    _transformation.abiInverse(aInv);
    if(ssd_gpu_set_overlay(_ctx, aInv, &_overlayIntrinsics) != SSD_OK)
      throw std::runtime_error(std::string("ssd_gpu_set_overlay: ") + ssd_gpu_last_error(_ctx));
    _overlayApplied = true;
  }
  if(_risers != _risersApplied)
  {
    if(ssd_gpu_set_vertical_faces(_ctx, _risers ? 1 : 0) != SSD_OK)
      throw std::runtime_error(std::string("ssd_gpu_set_vertical_faces: ") + ssd_gpu_last_error(_ctx));
    _risersApplied = _risers;
  }
  int rc;
  if(frames.z16)
    rc = frames.onDevice ? ssd_gpu_process_depth_device(_ctx, frames.z16, &frames.intrinsics, nFrames)
                         : ssd_gpu_process_depth_host(_ctx, frames.z16, &frames.intrinsics, nFrames);
  else
    rc = frames.onDevice ? ssd_gpu_process_device(_ctx, frames.vertices, nFrames) : ssd_gpu_process_host(_ctx, frames.vertices, nFrames);
  if(rc != SSD_OK)
    throw std::runtime_error(std::string("ssd_gpu_process: ") + ssd_gpu_last_error(_ctx));
  std::vector<Stairs> out(static_cast<size_t>(nFrames));
  if(status)
    status->assign(static_cast<size_t>(nFrames), 0u);
  ssd_gpu_step steps[SSD_GPU_MAX_STEPS];
  for(int f = 0; f < nFrames; f++)
  {
    int n = 0;
    uint32_t st = 0;
    ssd_gpu_get_steps(_ctx, f, steps, SSD_GPU_MAX_STEPS, &n, &st);
    if(status)
      (*status)[static_cast<size_t>(f)] = st;
    Stairs &s = out[static_cast<size_t>(f)];
    s.stairSteps.resize(static_cast<size_t>(n));
    for(int i = 0; i < n; i++)
    {
      s.stairSteps[static_cast<size_t>(i)].height = steps[i].height;
      for(int c = 0; c < 4; c++)
        s.stairSteps[static_cast<size_t>(i)].quadrilateral[static_cast<size_t>(c)] = Point2(steps[i].quad[c][0], steps[i].quad[c][1]);
    }
  }
  return out;
}

Stairs Pointcloud::detect(const Camera::DepthFrame &frame) const
{
  return processBatch(frame, 1).front();
}

void Pointcloud::process(const Camera::DepthFrame &frame) const
{
  _window.setViewport(viewportId::depth);
  const Stairs stairs = detect(frame);
  _window.setViewport(viewportId::infrared);
  std::cout << stairs.serialize() << std::endl;
}

void Pointcloud::enableVerticalFaces(bool enable) const
{
  _risers = enable;
}

std::vector<ssd_gpu_riser> Pointcloud::verticalFaces(int frame) const
{
  if(!_ctx)
    throw std::logic_error("Pointcloud::verticalFaces before any frame was processed");
  ssd_gpu_riser r[SSD_GPU_MAX_PLATEAUS];
  int n = 0;
  if(ssd_gpu_get_vertical_faces(_ctx, frame, r, SSD_GPU_MAX_PLATEAUS, &n) != SSD_OK)
    throw std::runtime_error(std::string("ssd_gpu_get_vertical_faces: ") + ssd_gpu_last_error(_ctx));
  return std::vector<ssd_gpu_riser>(r, r + n);
}

std::vector<uint8_t> Pointcloud::labels(int frame) const
{
  if(!_ctx)
    throw std::logic_error("Pointcloud::labels before any frame was processed");
  std::vector<uint8_t> l(static_cast<size_t>(_config.streams.depth.width) * _config.streams.depth.height);
  if(ssd_gpu_get_labels(_ctx, frame, l.data()) != SSD_OK)
    throw std::runtime_error(std::string("ssd_gpu_get_labels: ") + ssd_gpu_last_error(_ctx));
  return l;
}

} // namespace stairs
