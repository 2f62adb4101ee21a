// Test-infrastructure shim (NOT product code): stands in for the RealSense SDK example window.
// The GL window is stubbed out per BASELINE.json north_star.
#ifndef SSD_SHIM_EXAMPLE_HPP
#define SSD_SHIM_EXAMPLE_HPP
#include "librealsense2/rs.hpp"

struct GLFWwindow;
struct rect
{
  float x, y, w, h;
};
inline void set_viewport(const rect &) {}

class window
{
public:
  window(int, int, const char *) {}
  operator bool() { return true; }
  operator GLFWwindow *() { return nullptr; }
  void show(const rs2::frame &, const rect &) {}
};

#define GLFW_RESIZABLE 0
#define GL_FALSE 0
inline void glfwSetWindowAttrib(GLFWwindow *, int, int) {}
inline void glfwSwapBuffers(GLFWwindow *) {}
inline int glfwWindowShouldClose(GLFWwindow *) { return 1; }
inline void glfwWaitEvents() {}
#endif
