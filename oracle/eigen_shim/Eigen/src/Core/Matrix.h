// Stand-in for the internal Eigen header the reference includes directly
// (Scene.hpp:10, Scene.cpp:6, Renderer.cpp:6).  See ../../Dense.
#pragma once
#include "../../Dense"
