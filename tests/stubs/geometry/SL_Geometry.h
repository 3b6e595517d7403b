// Test infrastructure: empty stand-in for LibVisualSLAM geometry/SL_Geometry.h (see math/SL_Matrix.h).
#pragma once
