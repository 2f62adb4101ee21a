// detect_stairs_synthetic -- the reference's main loop (detect-stairs.cpp:26-45) on the kept class surface,
// with the RealSense capture replaced by the synthetic scene source and the GL window stubbed:
//   Window app; GeometricTransformation trans(world, camera); Pointcloud pointcloud(app, trans);
//   for each frame: pointcloud.process(frame)   -> one result line per frame on stdout
// Usage: detect_stairs_synthetic [width height n_frames [overlay]]   (overlay: the step quadrilaterals projected into the
//        camera image, what the reference draws over the depth view, one line per step on stderr; with it also the vertical
//        faces from the remainder points, one "riser" line each)
#include "../stair_step_detector_b200/csrc/host/pointcloud.h"
#include "../stair_step_detector_b200/csrc/host/transformation.h"
#include "../stair_step_detector_b200/csrc/host/window.h"
#include "../include/ssd_scene.h" // synthetic input source (libssd_scene.so), in place of Camera::waitForFrames
#include <cstdlib>
#include <iostream>
#include <vector>
using namespace stairs;

int main(int argc, char **argv)
{
  const int w = argc > 2 ? std::atoi(argv[1]) : 640, h = argc > 2 ? std::atoi(argv[2]) : 480;
  const int nFrames = argc > 3 ? std::atoi(argv[3]) : 3;
  Window app("stair-step-detector");

  ssd_scene base;
  ssd_scene_default(&base, w, h);
  base.noise_sigma = 0.0025f;
  base.dropout = 0.03f;
  base.n_holes = 3;
  double world[9], cam[9];
  ssd_scene_calibration_points(&base, world, cam);
  GeometricTransformation::RefPoints wp, cp;
  for(int i = 0; i < 3; i++)
  {
    wp[size_t(i)] = Point3(world[i * 3], world[i * 3 + 1], world[i * 3 + 2]);
    cp[size_t(i)] = Point3(cam[i * 3], cam[i * 3 + 1], cam[i * 3 + 2]);
  }
  const GeometricTransformation trans(wp, cp);
  const Pointcloud pointcloud(app, trans);

  // like the reference's loop (frames.depthFrame() -> process): the frame handed over is the z16 depth image;
  // the deprojection the reference leaves to rs2::pointcloud::calculate runs on the GPU
  std::vector<uint16_t> depth(size_t(w) * h);
  ssd_gpu_intrinsics intr;
  ssd_scene_intrinsics(&base, &intr);
  if(argc > 4)
  {
    pointcloud.enableOverlay(intr);
    pointcloud.enableVerticalFaces();
  }
  for(int f = 0; f < nFrames && app; f++)
  {
    ssd_scene sc;
    ssd_scene_randomize(&sc, &base, 2026, f, 3, 8);
    ssd_synth_depth_host(&sc, depth.data());
    pointcloud.process(Camera::DepthFrame(depth.data(), intr, w, h));
    if(argc > 4)
      for(const Quadrilateralf_t &q : pointcloud.overlay())
      {
        std::cerr << "overlay " << f;
        for(const Point2f &p : q)
          std::cerr << ' ' << p.x << ' ' << p.y;
        std::cerr << std::endl;
      }
    if(argc > 4)
      for(const ssd_gpu_riser &r : pointcloud.verticalFaces())
        std::cerr << "riser " << f << ' ' << r.lower_plateau << ' ' << r.n_points << ' ' << r.y_mean << ' ' << r.z_bottom << ' ' << r.z_top << std::endl;
  }
  return 0;
}
