// Test infrastructure: empty stand-in for LibVisualSLAM math/SL_LinAlg.h (see math/SL_Matrix.h).
#pragma once
