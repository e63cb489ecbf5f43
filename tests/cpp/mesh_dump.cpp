// Dumps the host Mesh (vlasovtucker_b200/host/mesh.cpp) as flat binary tables so the CPU test can
// compare index for index with the oracle's restatement of Mesh::Reconstruct.
#include <fstream>
#include <iostream>
#include <string>

#include "mesh.h"

using namespace VlasovTucker;

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    Mesh mesh(argv[1]);
    std::vector<std::array<int, 2>> pairs;
    for (int i = 3; i + 1 < argc; i += 2) pairs.push_back({{atoi(argv[i]), atoi(argv[i + 1])}});
    mesh.SetPeriodicBounaries(pairs);
    try {
        mesh.Reconstruct();
    } catch (std::exception& e) {
        std::cerr << "EXCEPTION: " << e.what() << "\n";
        return 3;
    }
    const FlatMesh& f = mesh.Flat();
    std::ofstream out(argv[2], std::ios::binary);
    const int nT = (int)mesh.tets.size(), nP = (int)mesh.points.size();
    out.write((const char*)&nT, 4);
    out.write((const char*)&nP, 4);
    for (auto* t : mesh.tets)
        for (auto* p : t->points) out.write((const char*)&p->index, 4);
    out.write((const char*)f.nbr.data(), f.nbr.size() * 4);
    out.write((const char*)f.entity.data(), f.entity.size() * 4);
    out.write((const char*)f.volume.data(), f.volume.size() * 8);
    out.write((const char*)f.area.data(), f.area.size() * 8);
    out.write((const char*)f.normal.data(), f.normal.size() * 8);
    out.write((const char*)f.tetCentroid.data(), f.tetCentroid.size() * 8);
    out.write((const char*)f.faceCentroid.data(), f.faceCentroid.size() * 8);
    out.write((const char*)f.order.data(), f.order.size() * 4);
    const double acs = mesh.AverageCellSize();
    out.write((const char*)&acs, 8);
    // pointer-graph facade must agree with the flat tables
    for (auto* t : mesh.tets)
        for (int j = 0; j < 4; j++) {
            const int a = t->adjTets[j] ? t->adjTets[j]->index : -1;
            if (a != f.nbr[4 * t->index + j] || t->faces[j]->index != 4 * t->index + j || t->faces[j]->adjTet != t) return 4;
        }
    int dup = 0;
    for (auto& kv : mesh.EntityToFaces()) dup += (int)kv.second.size();
    std::cout << "tets " << nT << " points " << nP << " entity-face entries " << dup << "\n";
    return 0;
}
