// Mesh primitives of the public API: Point, Face, Tet (reference: src/primitives.h:13-103).
// The objects are a facade over the flat tables the device uses; Mesh owns them.
#pragma once
#include <array>
#include <cmath>
#include <iostream>
#include <string>
#include <vector>

#include "typedefs.h"

namespace VlasovTucker {

class Point {
public:
    Point() : coords{{0, 0, 0}}, index(-1) {}
    Point(Vector3d c) : coords(c), index(-1) {}

    Vector3d coords;
    int index;   // position in Mesh::points

    double operator[](int i) const { return coords[i]; }
    double& operator[](int i) { return coords[i]; }
    Point operator+(const Point& p) const { return Point({coords[0] + p[0], coords[1] + p[1], coords[2] + p[2]}); }
    Point operator-(const Point& p) const { return Point({coords[0] - p[0], coords[1] - p[1], coords[2] - p[2]}); }
    Point operator*(double d) const { return Point({coords[0] * d, coords[1] * d, coords[2] * d}); }
    Point operator/(double d) const { return Point({coords[0] / d, coords[1] / d, coords[2] / d}); }
    friend Point operator*(double d, const Point& p) { return p * d; }
    friend Point operator-(const Point& p) { return p * (-1); }
    bool operator==(const Point& p) const { return coords == p.coords; }
    double Abs() const { return std::sqrt(DotProduct(*this)); }
    double DotProduct(const Point& p) const { return coords[0] * p[0] + coords[1] * p[1] + coords[2] * p[2]; }
    Point CrossProduct(const Point& p) const
    {
        return Point({coords[1] * p[2] - coords[2] * p[1], coords[2] * p[0] - coords[0] * p[2],
                      coords[0] * p[1] - coords[1] * p[0]});
    }
    friend std::ostream& operator<<(std::ostream& os, const Point& p)
    {
        return os << "Point: {" << p[0] << ", " << p[1] << ", " << p[2] << "}";
    }
};

class Tet;

enum class FaceType { Internal, Boundary };

class Face {
public:
    Face() = default;
    Face(Point* p0, Point* p1, Point* p2);

    std::array<Point*, 3> points{{nullptr, nullptr, nullptr}};
    Tet* adjTet = nullptr;   // owning tetrahedron
    int adjTetInd = -1;      // face number inside the owner (0..3)
    double area = 0;
    Point centroid;
    Point normal;            // unit, outward for the owner
    FaceType type = FaceType::Internal;
    std::vector<std::string> bcTypes;   // physical group names of the boundary entity
    int index = -1;          // 4*tet + adjTetInd
    int entity = -1;         // Gmsh elementary tag of the boundary surface, -1 inside

    friend std::ostream& operator<<(std::ostream& os, const Face& f);
};

class Tet {
public:
    Tet() = default;
    Tet(Point* p0, Point* p1, Point* p2, Point* p3);
    double Orientation() const;   // det [[p,1]]; the loader requires <= 0

    std::array<Point*, 4> points{{nullptr, nullptr, nullptr, nullptr}};
    std::array<Face*, 4> faces{{nullptr, nullptr, nullptr, nullptr}};
    std::array<Tet*, 4> adjTets{{nullptr, nullptr, nullptr, nullptr}};
    Point centroid;
    double volume = 0;
    int index = -1;

    friend std::ostream& operator<<(std::ostream& os, const Tet& t);
};

}  // namespace VlasovTucker
