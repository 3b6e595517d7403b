/* Stand-in for LibVisualSLAM geometry/SL_Triangulate.h: nothing of it is used by SL_IntraCamPose.cpp. */
#pragma once
