// Test infrastructure: stand-in for LibVisualSLAM's geometry/SL_Point.h (see math/SL_Matrix.h).
#pragma once
class Point2d {
 public:
  union {
    struct {
      double x, y;
    };
    double m[2];
  };
  Point2d() : x(0), y(0) {}
  Point2d(double x_, double y_) : x(x_), y(y_) {}
};
class Point3d {
 public:
  union {
    struct {
      double x, y, z;
    };
    double M[3];
  };
  Point3d() : x(0), y(0), z(0) {}
  Point3d(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
};
class Point3dId : public Point3d {
 public:
  int id;
  Point3dId() : Point3d(), id(-1) {}
  Point3dId(double x_, double y_, double z_, int id_) : Point3d(x_, y_, z_), id(id_) {}
};
