// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h).
// Restates src/mesh.cpp:13-303, src/primitives.cpp:94-164 and the MSH-2.2 ASCII path of
// libs/MshIO-main (load_msh_elements.cpp:97-170, load_msh_post_process.cpp:12-101).
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace oracle {

namespace {

Vec3 sub(const Vec3& a, const Vec3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
Vec3 add(const Vec3& a, const Vec3& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
Vec3 divs(const Vec3& a, double d) { return {a[0] / d, a[1] / d, a[2] / d}; }
Vec3 cross(const Vec3& a, const Vec3& b)
{
    // primitives.cpp:81-86
    return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
double norm(const Vec3& a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// Key for a triangle up to cyclic permutation (primitives.cpp:152-157): rotate so the
// smallest vertex id comes first, keeping orientation.
std::array<int, 3> CyclicKey(int a, int b, int c)
{
    if (a <= b && a <= c) return {a, b, c};
    if (b <= a && b <= c) return {b, c, a};
    return {c, a, b};
}

struct RawMsh {
    std::vector<Vec3> nodes;
    struct Elem { int type; int physTag; int entityTag; std::vector<int> nodes; };
    std::vector<Elem> elems;
    std::map<int, std::string> physNames;  // tag -> name (all dims, as mesh.cpp:19-21)
};

int NodesPerElement(int type)
{
    switch (type) {
        case 15: return 1;
        case 1: return 2;
        case 2: return 3;
        case 3: return 4;
        case 4: return 4;
        case 5: return 8;
        case 6: return 6;
        case 7: return 5;
        case 8: return 3;
        case 9: return 6;
        case 11: return 10;
        default: throw std::runtime_error("oracle: unsupported element type " + std::to_string(type));
    }
}

RawMsh ReadMsh22(const std::string& file)
{
    std::ifstream in(file);
    if (!in) throw std::runtime_error("Input file does not exist!");  // load_msh.cpp:74-76
    RawMsh raw;
    std::string line;
    while (std::getline(in, line)) {
        if (line.rfind("$MeshFormat", 0) == 0) {
            std::string ver; int ftype, dsize;
            in >> ver >> ftype >> dsize;
            if (ver != "2.2" || ftype != 0)
                throw std::runtime_error("oracle: only MSH 2.2 ASCII is restated");
        } else if (line.rfind("$PhysicalNames", 0) == 0) {
            int n; in >> n;
            for (int i = 0; i < n; i++) {
                int dim, tag; in >> dim >> tag;
                std::string rest; std::getline(in, rest);
                size_t a = rest.find('"'), b = rest.rfind('"');
                raw.physNames[tag] = (a != std::string::npos && b > a) ? rest.substr(a + 1, b - a - 1) : rest;
            }
        } else if (line.rfind("$Nodes", 0) == 0) {
            size_t n; in >> n;
            raw.nodes.resize(n);
            for (size_t i = 0; i < n; i++) {
                long tag; double x, y, z;
                in >> tag >> x >> y >> z;
                // The reference indexes entity_blocks[0].data by position (mesh.cpp:115-119):
                // points[i] is the i-th node line whatever its tag.
                raw.nodes[i] = {x, y, z};
            }
        } else if (line.rfind("$Elements", 0) == 0) {
            size_t n; in >> n;
            raw.elems.resize(n);
            for (size_t i = 0; i < n; i++) {
                int num, type, ntags; in >> num >> type >> ntags;
                std::vector<int> tags(ntags);
                for (int j = 0; j < ntags; j++) in >> tags[j];
                auto& e = raw.elems[i];
                e.type = type;
                // load_msh_elements.cpp:139-158: entity tag is the 2nd tag, else the 1st, else 1
                e.physTag = ntags > 1 ? tags[0] : -1;
                e.entityTag = ntags > 1 ? tags[1] : (ntags > 0 ? tags[0] : 1);
                int npe = NodesPerElement(type);
                e.nodes.resize(npe);
                for (int j = 0; j < npe; j++) in >> e.nodes[j];
            }
        }
    }
    return raw;
}

}  // namespace

static Mesh BuildFromRaw(RawMsh& raw, const std::vector<std::array<int, 2>>& periodicPairs,
                         double scale);

Mesh LoadMesh(const std::string& file, const std::vector<std::array<int, 2>>& periodicPairs,
              double scale)
{
    RawMsh raw = ReadMsh22(file);
    return BuildFromRaw(raw, periodicPairs, scale);
}

// Same pipeline fed from arrays (what a .msh with "triangles first, then tets" would hold);
// used for synthetic meshes so tests need not write multi-hundred-MB ASCII files.
Mesh MeshFromArrays(const double* nodes, int nNodes, const int* tets, int nTets, const int* tris,
                    const int* triEntity, int nTris,
                    const std::vector<std::array<int, 2>>& periodicPairs, double scale)
{
    RawMsh raw;
    raw.nodes.resize(nNodes);
    for (int i = 0; i < nNodes; i++) raw.nodes[i] = {nodes[3 * i], nodes[3 * i + 1], nodes[3 * i + 2]};
    raw.elems.reserve((size_t)nTris + nTets);
    for (int i = 0; i < nTris; i++)
        raw.elems.push_back({2, -1, triEntity[i], {tris[3 * i] + 1, tris[3 * i + 1] + 1, tris[3 * i + 2] + 1}});
    for (int i = 0; i < nTets; i++)
        raw.elems.push_back({4, -1, 1, {tets[4 * i] + 1, tets[4 * i + 1] + 1, tets[4 * i + 2] + 1, tets[4 * i + 3] + 1}});
    return BuildFromRaw(raw, periodicPairs, scale);
}

static Mesh BuildFromRaw(RawMsh& raw, const std::vector<std::array<int, 2>>& periodicPairs,
                         double scale)
{
    Mesh m;
    m.periodicPairs = periodicPairs;

    // mesh.cpp:23-28: surface entity -> physical group names.  MshIO keeps a std::set of
    // physical tags per entity (load_msh_elements.cpp:122,147-152) so names come out in
    // increasing physical-tag order without duplicates.
    {
        std::map<int, std::vector<int>> surfPhys;
        for (auto& e : raw.elems) {
            int dim = (e.type == 2 || e.type == 3 || e.type == 9) ? 2 : -1;
            if (dim == 2 && e.physTag >= 0) surfPhys[e.entityTag].push_back(e.physTag);
        }
        for (auto& kv : surfPhys) {
            std::sort(kv.second.begin(), kv.second.end());
            kv.second.erase(std::unique(kv.second.begin(), kv.second.end()), kv.second.end());
            for (int tag : kv.second) m.entityToPhysGroups[kv.first].push_back(raw.physNames[tag]);
        }
    }

    // _ExtractPoints, mesh.cpp:113-127
    m.points.resize(raw.nodes.size());
    for (size_t i = 0; i < raw.nodes.size(); i++)
        m.points[i] = {raw.nodes[i][0] * scale, raw.nodes[i][1] * scale, raw.nodes[i][2] * scale};

    // _ExtractTetsAndFaces, mesh.cpp:129-181.  Blocks are runs of equal (dim, entity, type)
    // in file order, so walking the elements in file order visits tets in the same order.
    std::map<std::array<int, 3>, int> pointsToFace;
    for (auto& e : raw.elems) {
        if (e.type != 4) continue;
        int p0 = e.nodes[0] - 1, p1 = e.nodes[1] - 1, p2 = e.nodes[2] - 1, p3 = e.nodes[3] - 1;
        m.tets.push_back({p0, p1, p2, p3});
        const Vec3 &P0 = m.points[p0], &P1 = m.points[p1], &P2 = m.points[p2], &P3 = m.points[p3];
        // primitives.cpp:122-127
        m.tetCentroid.push_back(divs(add(add(add(P0, P1), P2), P3), 4.0));
        // primitives.cpp:130-139: det [[P,1]] = -(P1-P0).((P2-P0)x(P3-P0))
        Vec3 a = sub(P1, P0), b = sub(P2, P0), c = sub(P3, P0);
        Vec3 bc = cross(b, c);
        double orientation = -(a[0] * bc[0] + a[1] * bc[1] + a[2] * bc[2]);
        m.tetVolume.push_back(std::fabs(orientation) / 6.0);

        const int fp[4][3] = {{p1, p2, p3}, {p0, p3, p2}, {p0, p1, p3}, {p0, p2, p1}};  // mesh.cpp:157-160
        for (int j = 0; j < 4; j++) {
            const Vec3 &A = m.points[fp[j][0]], &B = m.points[fp[j][1]], &C = m.points[fp[j][2]];
            // primitives.cpp:94-111
            Vec3 cen = divs(add(add(A, B), C), 3.0);
            Vec3 u = sub(B, A), w = sub(C, A);
            Vec3 nrm = cross(u, w);
            double len = norm(nrm);
            if (len == 0) throw std::runtime_error("");
            int fi = (int)m.facePoints.size();
            m.facePoints.push_back({fp[j][0], fp[j][1], fp[j][2]});
            m.faceCentroid.push_back(cen);
            m.faceNormal.push_back(divs(nrm, len));
            m.faceArea.push_back(len / 2.0);
            m.faceEntity.push_back(-1);
            m.faceBoundary.push_back(0);
            pointsToFace[CyclicKey(fp[j][0], fp[j][1], fp[j][2])] = fi;
        }
    }
    m.adj.assign(m.facePoints.size(), -1);

    // _LabelBoundaryFaces, mesh.cpp:183-206
    for (auto& e : raw.elems) {
        if (e.type != 2) continue;
        auto it = pointsToFace.find(CyclicKey(e.nodes[0] - 1, e.nodes[1] - 1, e.nodes[2] - 1));
        if (it == pointsToFace.end())
            throw std::runtime_error("oracle: boundary triangle matches no tet face");
        int fi = it->second;
        m.faceBoundary[fi] = 1;
        m.faceEntity[fi] = e.entityTag;
        m.entityToFaces[e.entityTag].push_back(fi);
    }

    // _FillAdjacencyInfo, mesh.cpp:208-222: neighbour = owner of the reversed triangle
    for (size_t fi = 0; fi < m.facePoints.size(); fi++) {
        auto& p = m.facePoints[fi];
        auto it = pointsToFace.find(CyclicKey(p[0], p[2], p[1]));
        if (it != pointsToFace.end()) m.adj[fi] = it->second / 4;
    }

    // _ConfigurePeriodicity, mesh.cpp:255-303 with SortFacesInPlane, mesh.cpp:224-253
    for (auto& pr : periodicPairs) {
        std::array<std::vector<int>, 2> plane;
        for (int i : {0, 1})
            for (size_t fi = 0; fi < m.facePoints.size(); fi++)
                if (m.faceBoundary[fi] && m.faceEntity[fi] == pr[i]) plane[i].push_back((int)fi);
        if (plane[0].size() != plane[1].size())
            throw std::runtime_error("Mismatch between the sizes of the periodic planes " +
                                     std::to_string(pr[0]) + " and  " + std::to_string(pr[1]));
        auto approxEq = [](double a, double b) { return std::fabs(a - b) < 1e-8; };
        for (int i : {0, 1}) {
            std::sort(plane[i].begin(), plane[i].end(), [&](int l, int r) {
                const Vec3& lc = m.faceCentroid[l];
                const Vec3& rc = m.faceCentroid[r];
                if (!approxEq(lc[0], rc[0])) return lc[0] < rc[0];
                if (!approxEq(lc[1], rc[1])) return lc[1] < rc[1];
                if (!approxEq(lc[2], rc[2])) return lc[2] < rc[2];
                return false;
            });
        }
        // NOTE: the reference sorts all planes first and pairs afterwards (mesh.cpp:286-302);
        // pairing never feeds back into the sort keys, so doing both per pair is equivalent.
        for (size_t i = 0; i < plane[0].size(); i++) {
            int f = plane[0][i], g = plane[1][i];
            m.adj[f] = g / 4;
            m.adj[g] = f / 4;
        }
    }
    return m;
}

double AverageCellSize(const Mesh& m)
{
    double s = 0;
    for (double v : m.tetVolume) s += std::pow(v * 6 * std::sqrt(2), 1 / 3.);
    return s / (double)m.tetVolume.size();
}

}  // namespace oracle
