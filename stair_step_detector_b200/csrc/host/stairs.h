// Host mirror of the reference's result type (stairs.h:30-39) and its line serializer.
#pragma once
#include "types.h"
#include <string>

namespace stairs
{

struct Stairs
{
  struct StairStep
  {
    Coordinate_t height;
    Quadrilateral_t quadrilateral;
  };
  std::vector<StairStep> stairSteps;

  // ["stairs",["stairSteps",n],[ [["height",h],["quadrilateral",[x,y],[x,y],[x,y],[x,y]]], ... ]]
  // three decimals, fixed notation; the step list is omitted when n == 0 (reference stairs.cpp:55-70)
  std::string serialize() const;
};

} // namespace stairs
