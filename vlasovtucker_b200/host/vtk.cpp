#include "vtk.h"

#include <fstream>

namespace VlasovTucker {

namespace {
void Header(std::ofstream& out, const char* dataset)
{
    out << "# vtk DataFile Version 2.0\nVlasov-T\nASCII\nDATASET " << dataset << "\n";
}
void TetGrid(std::ofstream& out, const Mesh& mesh)
{
    Header(out, "UNSTRUCTURED_GRID");
    out << "POINTS " << mesh.points.size() << " float\n";
    for (const Point* p : mesh.points) out << (*p)[0] << " " << (*p)[1] << " " << (*p)[2] << "\n";
    out << "CELLS " << mesh.tets.size() << " " << mesh.tets.size() * 5 << "\n";
    for (const Tet* t : mesh.tets) {
        out << 4;
        for (const Point* p : t->points) out << " " << p->index;
        out << "\n";
    }
    out << "\nCELL_TYPES " << mesh.tets.size() << "\n";
    for (size_t i = 0; i < mesh.tets.size(); i++) out << 10 << "\n";
}
}  // namespace

void WriteCellScalarDataVTK(std::string fileName, const Mesh& mesh, const std::vector<double>& data)
{
    std::ofstream out(fileName + ".vtk");
    TetGrid(out, mesh);
    if (!data.empty()) {
        out << "CELL_DATA " << mesh.tets.size() << "\nSCALARS data double 1\nLOOKUP_TABLE default\n";
        for (size_t i = 0; i < mesh.tets.size(); i++) out << data[i] << "\n";
    }
}

void WriteCellVectorDataVTK(std::string fileName, const Mesh& mesh, const std::vector<Vector3d>& data)
{
    std::ofstream out(fileName + ".vtk");
    TetGrid(out, mesh);
    if (!data.empty()) {
        out << "CELL_DATA " << mesh.tets.size() << "\nVECTORS data double\n";
        for (size_t i = 0; i < mesh.tets.size(); i++) out << data[i][0] << " " << data[i][1] << " " << data[i][2] << "\n";
    }
}

void WriteMeshVTK(std::string fileName, const Mesh& mesh)
{
    // tetrahedra with their index, and boundary faces with their entity tag
    {
        std::ofstream out(fileName + "_tets.vtk");
        TetGrid(out, mesh);
        out << "CELL_DATA " << mesh.tets.size() << "\nSCALARS index int 1\nLOOKUP_TABLE default\n";
        for (const Tet* t : mesh.tets) out << t->index << "\n";
    }
    std::ofstream out(fileName + "_faces.vtk");
    Header(out, "UNSTRUCTURED_GRID");
    out << "POINTS " << mesh.points.size() << " float\n";
    for (const Point* p : mesh.points) out << (*p)[0] << " " << (*p)[1] << " " << (*p)[2] << "\n";
    std::vector<const Face*> boundary;
    for (const Face* f : mesh.faces)
        if (f->type == FaceType::Boundary) boundary.push_back(f);
    out << "CELLS " << boundary.size() << " " << boundary.size() * 4 << "\n";
    for (const Face* f : boundary) out << 3 << " " << f->points[0]->index << " " << f->points[1]->index << " " << f->points[2]->index << "\n";
    out << "\nCELL_TYPES " << boundary.size() << "\n";
    for (size_t i = 0; i < boundary.size(); i++) out << 5 << "\n";
    out << "CELL_DATA " << boundary.size() << "\nSCALARS entity int 1\nLOOKUP_TABLE default\n";
    for (const Face* f : boundary) out << f->entity << "\n";
}

void WriteDistributionVTK(std::string fileName, const VelocityGrid& g, const Tensor3d& f)
{
    std::ofstream out(fileName + ".vtk");
    Header(out, "STRUCTURED_POINTS");
    out << "DIMENSIONS " << g.nCells[0] << " " << g.nCells[1] << " " << g.nCells[2] << "\n";
    out << "ORIGIN " << g.minV[0] << " " << g.minV[1] << " " << g.minV[2] << "\n";
    out << "SPACING " << g.step[0] << " " << g.step[1] << " " << g.step[2] << "\n";
    out << "POINT_DATA " << g.nCellsTotal << "\nSCALARS distribution double 1\nLOOKUP_TABLE default\n";
    for (int i2 = 0; i2 < g.nCells[2]; i2++)
        for (int i1 = 0; i1 < g.nCells[1]; i1++)
            for (int i0 = 0; i0 < g.nCells[0]; i0++) out << f(i0, i1, i2) << "\n";
}

}  // namespace VlasovTucker
