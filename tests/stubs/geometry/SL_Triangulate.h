// Test infrastructure: empty stand-in for LibVisualSLAM geometry/SL_Triangulate.h (see math/SL_Matrix.h).
#pragma once
