// Host-side Mesh: MSH 2.2 ASCII ingestion and the reference's topology contract, built on flat
// arrays first and wrapped in the Point/Face/Tet facade afterwards.
#include "mesh.h"

#include <algorithm>
#include <cassert>
#include <cstdint>
#include <fstream>
#include <map>
#include <numeric>
#include <set>
#include <sstream>
#include <stdexcept>

namespace VlasovTucker {

// ---------------------------------------------------------------- primitives
Face::Face(Point* p0, Point* p1, Point* p2) : points{{p0, p1, p2}}
{
    centroid = (*p0 + *p1 + *p2) / 3.0;
    const Point n = (*p1 - *p0).CrossProduct(*p2 - *p0);
    const double len = n.Abs();
    if (len == 0) throw std::runtime_error("");   // degenerate triangle, as src/primitives.cpp:105-106
    normal = n / len;
    area = len / 2.0;
}

Tet::Tet(Point* p0, Point* p1, Point* p2, Point* p3) : points{{p0, p1, p2, p3}}
{
    centroid = (*p0 + *p1 + *p2 + *p3) / 4.0;
    volume = std::fabs(Orientation()) / 6.0;
}

double Tet::Orientation() const
{
    // det of the 4x4 matrix [x y z 1] by rows = -(p1-p0).((p2-p0)x(p3-p0))
    const Point a = *points[1] - *points[0], b = *points[2] - *points[0], c = *points[3] - *points[0];
    return -a.DotProduct(b.CrossProduct(c));
}

std::ostream& operator<<(std::ostream& os, const Face& f)
{
    return os << "Face: {" << *f.points[0] << ",\n       " << *f.points[1] << ",\n       " << *f.points[2];
}
std::ostream& operator<<(std::ostream& os, const Tet& t)
{
    return os << "Tet: {" << *t.points[0] << ",\n      " << *t.points[1] << ",\n      " << *t.points[2] << ",\n      "
              << *t.points[3] << "}";
}

// ---------------------------------------------------------------- MSH 2.2 ASCII
struct Mesh::Raw {
    std::vector<std::array<double, 3>> nodes;          // in file order
    std::vector<std::array<int, 4>> tets;              // 0-based node positions, file order
    struct Tri { std::array<int, 3> v; int entity; };
    std::vector<Tri> tris;                             // file order
    std::map<int, std::string> physName;               // physical tag -> name
    std::map<int, std::set<int>> surfacePhys;          // surface entity -> physical tags
};

namespace {
int nodesOfType(int type)
{
    static const std::map<int, int> n = {{15, 1}, {1, 2}, {2, 3}, {3, 4}, {4, 4}, {5, 8}, {6, 6}, {7, 5}, {8, 3}, {9, 6}, {11, 10}};
    auto it = n.find(type);
    if (it == n.end()) throw std::runtime_error("Unsupported MSH element type " + std::to_string(type));
    return it->second;
}
// canonical rotation of an oriented triangle (cyclic permutations are the same face)
std::array<int, 3> rotated(int a, int b, int c)
{
    if (a < b && a < c) return {{a, b, c}};
    if (b < a && b < c) return {{b, c, a}};
    return {{c, a, b}};
}
}  // namespace

Mesh::Mesh(std::string mshFile) : raw_(new Raw())
{
    std::ifstream in(mshFile);
    if (!in) throw std::runtime_error("Input file does not exist!");
    std::string tok;
    while (in >> tok) {
        if (tok == "$MeshFormat") {
            std::string version;
            int fileType, dataSize;
            in >> version >> fileType >> dataSize;
            if (version != "2.2" || fileType != 0) throw std::runtime_error("Only MSH 2.2 ASCII files are supported");
        } else if (tok == "$PhysicalNames") {
            int n;
            in >> n;
            for (int i = 0; i < n; i++) {
                int dim, tag;
                in >> dim >> tag;
                std::string rest;
                std::getline(in, rest);
                const size_t a = rest.find('"'), b = rest.rfind('"');
                raw_->physName[tag] = (a != std::string::npos && b > a) ? rest.substr(a + 1, b - a - 1) : rest;
            }
        } else if (tok == "$Nodes") {
            size_t n;
            in >> n;
            raw_->nodes.resize(n);
            for (size_t i = 0; i < n; i++) {
                long tag;
                in >> tag >> raw_->nodes[i][0] >> raw_->nodes[i][1] >> raw_->nodes[i][2];
            }
        } else if (tok == "$Elements") {
            size_t n;
            in >> n;
            for (size_t i = 0; i < n; i++) {
                int id, type, ntags;
                in >> id >> type >> ntags;
                std::vector<int> tags(ntags);
                for (auto& t : tags) in >> t;
                const int nn = nodesOfType(type);
                std::vector<int> v(nn);
                for (auto& x : v) in >> x;
                const int entity = ntags > 1 ? tags[1] : (ntags > 0 ? tags[0] : 1);
                if (type == 4) raw_->tets.push_back({{v[0] - 1, v[1] - 1, v[2] - 1, v[3] - 1}});
                if (type == 2) {
                    raw_->tris.push_back({{{v[0] - 1, v[1] - 1, v[2] - 1}}, entity});
                    if (ntags > 1) raw_->surfacePhys[entity].insert(tags[0]);
                }
            }
        }
    }
    for (auto& kv : raw_->surfacePhys)
        for (int tag : kv.second) labels_[kv.first].push_back(raw_->physName[tag]);
}

Mesh::~Mesh()
{
    for (auto* p : points) delete p;
    for (auto* f : faces) delete f;
    for (auto* t : tets) delete t;
    delete raw_;
}

std::unordered_map<int, std::vector<std::string>> Mesh::BoundaryLabels() const { return labels_; }

void Mesh::PrintBoundaryLabels() const
{
    for (const auto& kv : labels_) {
        std::cout << kv.first << ": {";
        for (size_t i = 0; i < kv.second.size(); i++) std::cout << (i ? ", " : "") << kv.second[i];
        std::cout << "}\n";
    }
}

void Mesh::SetPeriodicBounaries(const std::vector<std::array<int, 2>>& pairs) { periodic_ = pairs; }
std::vector<std::array<int, 2>> Mesh::PeriodicBoundaries() const { return periodic_; }
const std::unordered_map<int, std::vector<Face*>>& Mesh::EntityToFaces() const { return entityFaces_; }

double Mesh::AverageCellSize() const
{
    double s = 0;
    for (auto* t : tets) s += std::pow(t->volume * 6 * std::sqrt(2), 1 / 3.);
    return s / (double)tets.size();
}

void Mesh::Reconstruct(double scale)
{
    const size_t nT = raw_->tets.size();
    // points
    for (size_t i = 0; i < raw_->nodes.size(); i++) {
        Point* p = new Point({raw_->nodes[i][0], raw_->nodes[i][1], raw_->nodes[i][2]});
        *p = *p * scale;
        p->index = (int)i;
        points.push_back(p);
    }
    // tets and their four outward faces
    static const int fv[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
    std::map<std::array<int, 3>, int> faceOf;   // oriented triangle -> face index
    for (size_t t = 0; t < nT; t++) {
        const auto& v = raw_->tets[t];
        Tet* tet = new Tet(points[v[0]], points[v[1]], points[v[2]], points[v[3]]);
        tet->index = (int)t;
        assert(tet->Orientation() <= 0);
        tets.push_back(tet);
        for (int j = 0; j < 4; j++) {
            Face* f = new Face(points[v[fv[j][0]]], points[v[fv[j][1]]], points[v[fv[j][2]]]);
            f->adjTet = tet;
            f->adjTetInd = j;
            f->index = (int)faces.size();
            tet->faces[j] = f;
            faces.push_back(f);
            faceOf[rotated(v[fv[j][0]], v[fv[j][1]], v[fv[j][2]])] = f->index;
        }
    }
    // boundary triangles: entity tag and physical names
    for (const auto& tri : raw_->tris) {
        auto it = faceOf.find(rotated(tri.v[0], tri.v[1], tri.v[2]));
        if (it == faceOf.end()) throw std::runtime_error("Boundary triangle does not match any tetrahedron face");
        Face* f = faces[it->second];
        f->type = FaceType::Boundary;
        f->entity = tri.entity;
        auto lab = labels_.find(tri.entity);
        if (lab != labels_.end()) f->bcTypes = lab->second;
        entityFaces_[tri.entity].push_back(f);
    }
    // interior adjacency: the neighbour owns the reversed triangle
    for (Face* f : faces) {
        auto it = faceOf.find(rotated(f->points[0]->index, f->points[2]->index, f->points[1]->index));
        if (it != faceOf.end()) f->adjTet->adjTets[f->adjTetInd] = faces[it->second]->adjTet;
    }
    // periodic planes: both planes sorted with the reference's tolerance comparator, paired by rank
    for (const auto& pr : periodic_) {
        std::array<std::vector<Face*>, 2> plane;
        for (int s = 0; s < 2; s++)
            for (Face* f : faces)
                if (f->type == FaceType::Boundary && f->entity == pr[s]) plane[s].push_back(f);
        if (plane[0].size() != plane[1].size())
            throw std::runtime_error("Mismatch between the sizes of the periodic planes " + std::to_string(pr[0]) +
                                     " and  " + std::to_string(pr[1]));
        for (auto& pl : plane)
            std::sort(pl.begin(), pl.end(), [](const Face* a, const Face* b) {
                for (int k = 0; k < 3; k++)
                    if (!(std::abs(a->centroid[k] - b->centroid[k]) < 1e-8)) return a->centroid[k] < b->centroid[k];
                return false;
            });
        for (size_t i = 0; i < plane[0].size(); i++) {
            Face *a = plane[0][i], *b = plane[1][i];
            a->adjTet->adjTets[a->adjTetInd] = b->adjTet;
            b->adjTet->adjTets[b->adjTetInd] = a->adjTet;
        }
    }
    // flat tables for the device layer
    flat_.nbr.resize(4 * nT);
    flat_.area.resize(4 * nT);
    flat_.volume.resize(nT);
    flat_.normal.resize(12 * nT);
    flat_.entity.resize(4 * nT);
    flat_.tetCentroid.resize(3 * nT);
    flat_.faceCentroid.resize(12 * nT);
    for (size_t t = 0; t < nT; t++) {
        const Tet* tet = tets[t];
        flat_.volume[t] = tet->volume;
        for (int k = 0; k < 3; k++) flat_.tetCentroid[3 * t + k] = tet->centroid[k];
        for (int j = 0; j < 4; j++) {
            const Face* f = tet->faces[j];
            flat_.nbr[4 * t + j] = tet->adjTets[j] ? tet->adjTets[j]->index : -1;
            flat_.area[4 * t + j] = f->area;
            flat_.entity[4 * t + j] = f->entity;
            for (int k = 0; k < 3; k++) {
                flat_.normal[12 * t + 3 * j + k] = f->normal[k];
                flat_.faceCentroid[12 * t + 3 * j + k] = f->centroid[k];
            }
        }
    }
    // locality order for the device: Morton code of the centroid (21 bits per axis), ties by index
    {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (size_t t = 0; t < nT; t++)
            for (int k = 0; k < 3; k++) {
                lo[k] = std::min(lo[k], flat_.tetCentroid[3 * t + k]);
                hi[k] = std::max(hi[k], flat_.tetCentroid[3 * t + k]);
            }
        const double ext = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], 1e-300});
        auto spread = [](uint64_t x) {
            x &= 0x1fffff;
            x = (x | x << 32) & 0x1f00000000ffffULL;
            x = (x | x << 16) & 0x1f0000ff0000ffULL;
            x = (x | x << 8) & 0x100f00f00f00f00fULL;
            x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
            x = (x | x << 2) & 0x1249249249249249ULL;
            return x;
        };
        std::vector<uint64_t> code(nT);
        for (size_t t = 0; t < nT; t++) {
            uint64_t c = 0;
            for (int k = 0; k < 3; k++) {
                const double u = (flat_.tetCentroid[3 * t + k] - lo[k]) / ext;
                c |= spread((uint64_t)(u * 2097151.0)) << k;
            }
            code[t] = c;
        }
        flat_.order.resize(nT);
        std::iota(flat_.order.begin(), flat_.order.end(), 0);
        std::stable_sort(flat_.order.begin(), flat_.order.end(), [&](int32_t a, int32_t b) { return code[a] < code[b]; });
    }
}

}  // namespace VlasovTucker
