// Mesh of the public API (reference: src/mesh.h:13-53).  Reads Gmsh MSH 2.2 ASCII directly (the
// reference goes through MshIO) and rebuilds the reference's index contract: points[i] = i-th
// node line, tets[t] = t-th tetrahedron of the file, faces[4t+j] with the per-tet vertex order of
// src/mesh.cpp:157-160, neighbours by reversed-triangle lookup, periodic pairing by the sorted
// centroid lists of src/mesh.cpp:224-303.  Besides the Point/Face/Tet facade the mesh keeps the
// flat tables the device layer uploads.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "primitives.h"

namespace VlasovTucker {

// Flattened mesh in the reference's tet order: what vt_mesh_upload / vt_poisson_setup take.
struct FlatMesh {
    std::vector<int32_t> nbr;          // 4 per tet, -1 = none
    std::vector<double> area;          // 4 per tet
    std::vector<double> volume;        // 1 per tet
    std::vector<double> normal;        // 12 per tet
    std::vector<int32_t> entity;       // 4 per tet
    std::vector<double> tetCentroid;   // 3 per tet
    std::vector<double> faceCentroid;  // 12 per tet
    std::vector<int32_t> order;        // locality permutation for the device layout
};

class Mesh {
public:
    explicit Mesh(std::string mshFile);
    ~Mesh();
    Mesh(const Mesh&) = delete;
    Mesh& operator=(const Mesh&) = delete;

    void Reconstruct(double scaleFactor = 1);

    std::unordered_map<int, std::vector<std::string>> BoundaryLabels() const;
    void PrintBoundaryLabels() const;

    void SetPeriodicBounaries(const std::vector<std::array<int, 2>>& periodicPairs);   // [sic], as the reference
    std::vector<std::array<int, 2>> PeriodicBoundaries() const;

    const std::unordered_map<int, std::vector<Face*>>& EntityToFaces() const;
    double AverageCellSize() const;

    // device-layer view (not in the reference API)
    const FlatMesh& Flat() const { return flat_; }

public:
    std::vector<Point*> points;
    std::vector<Face*> faces;
    std::vector<Tet*> tets;

private:
    struct Raw;
    Raw* raw_;
    std::vector<std::array<int, 2>> periodic_;
    std::unordered_map<int, std::vector<std::string>> labels_;
    std::unordered_map<int, std::vector<Face*>> entityFaces_;
    FlatMesh flat_;
};

}  // namespace VlasovTucker
