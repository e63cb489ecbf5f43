#include "particle_data.h"

#include <algorithm>
#include <cmath>
#include <iostream>
#include <stdexcept>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <type_traits>

#include "device.h"
#include "full.h"
#include "tucker.h"

namespace VlasovTucker {

template <typename T>
ParticleData<T>::ParticleData(const Mesh* mesh, const VelocityGrid* vGrid) : mass(1), charge(1), _mesh(mesh), _vGrid(vGrid)
{
    _maxRank = *std::max_element(vGrid->nCells.begin(), vGrid->nCells.end());
    std::cout << "The maximum rank is " << _maxRank << "\n";
    _dev = device::ContextOf(mesh);
    const int32_t n[3] = {vGrid->nCells[0], vGrid->nCells[1], vGrid->nCells[2]};
    device::Check(vt_species_create(_dev->ctx, n, vGrid->minV.data(), vGrid->maxV.data(), mass, charge, &_species));
    if (std::is_same<T, Tucker>::value) device::Check(vt_tucker_enable(_dev->ctx, _species, _comprErr, _maxRank));
}

template <typename T>
void ParticleData<T>::PushParams() const
{
    if (_dev) device::Check(vt_species_set_params(_dev->ctx, _species, mass, charge));
}

template <>
void ParticleData<Full>::RebindRows()
{
    pdf.clear();
    pdf.reserve(_mesh->tets.size());
    for (size_t t = 0; t < _mesh->tets.size(); t++)
        pdf.push_back(Full::DeviceRow(_dev, _species, (int)t, _vGrid->nCells));
}

template <>
void ParticleData<Full>::SetMaxwellPDF(const MaxwellPDF& params)
{
    PushParams();   // the Maxwellian uses `mass`
    device::Check(vt_species_set_maxwell(_dev->ctx, _species, params.physDensity.data(), params.temperature,
                                         params.mostProbableV.data()));
    RebindRows();
    std::cout << "PDF size reduction: " << 1.0 << " times on average\n";
}

template <>
void ParticleData<Tucker>::SyncFromDevice()
{
    const int n0 = _vGrid->nCells[0], n1 = _vGrid->nCells[1], n2 = _vGrid->nCells[2];
    std::vector<double> core((size_t)n0 * n1 * n2), u0((size_t)n0 * n0), u1((size_t)n1 * n1), u2((size_t)n2 * n2);
    pdf.clear();
    pdf.reserve(_mesh->tets.size());
    for (size_t t = 0; t < _mesh->tets.size(); t++) {
        int32_t r[3];
        device::Check(vt_tucker_get_factors(_dev->ctx, _species, (int)t, r, core.data(), u0.data(), u1.data(), u2.data()));
        Tensor3d c(r[0], r[1], r[2]);
        std::copy(core.begin(), core.begin() + (size_t)r[0] * r[1] * r[2], c.data());
        std::array<Eigen::MatrixXd, 3> u = {Eigen::MatrixXd(n0, r[0]), Eigen::MatrixXd(n1, r[1]), Eigen::MatrixXd(n2, r[2])};
        std::copy(u0.begin(), u0.begin() + (size_t)n0 * r[0], u[0].data());
        std::copy(u1.begin(), u1.begin() + (size_t)n1 * r[1], u[1].data());
        std::copy(u2.begin(), u2.begin() + (size_t)n2 * r[2], u[2].data());
        pdf.push_back(Tucker(c, u));
    }
}
template <>
void ParticleData<Full>::SyncFromDevice() {}
template <>
void ParticleData<Tucker>::RebindRows()
{
    SyncFromDevice();
}

template <>
void ParticleData<Tucker>::SetMaxwellPDF(const MaxwellPDF& params)
{
    // the Maxwellian is tabulated and scaled on the device as for Full, then stored as exact Tucker
    // tensors (precision 0, as particle_data.cpp:64-69 builds them)
    PushParams();
    device::Check(vt_species_set_maxwell(_dev->ctx, _species, params.physDensity.data(), params.temperature,
                                         params.mostProbableV.data()));
    SyncFromDevice();
    double reduction = 0;
    for (const auto& t : pdf) reduction += _vGrid->nCellsTotal / (double)t.Size();
    std::cout << "PDF size reduction: " << reduction / (double)_mesh->tets.size() << " times on average\n";
}

// Density()/Velocity() (particle_data.cpp:93-128) are device reductions for both formats; a
// Tucker species is reconstructed on the device first
template <typename T>
std::vector<double> ParticleData<T>::Density() const
{
    std::vector<double> r(_mesh->tets.size());
    device::Check(vt_species_density(_dev->ctx, _species, r.data()));
    return r;
}
template <typename T>
std::vector<Vector3d> ParticleData<T>::Velocity() const
{
    std::vector<Vector3d> r(_mesh->tets.size());
    device::Check(vt_species_velocity(_dev->ctx, _species, &r[0][0]));
    return r;
}
template <typename T>
void ParticleData<T>::SetCompressionError(double e)
{
    _comprErr = e;
    if (std::is_same<T, Tucker>::value) device::Check(vt_tucker_enable(_dev->ctx, _species, _comprErr, _maxRank));
}
template <typename T>
double ParticleData<T>::CompressionError() const { return _comprErr; }
template <typename T>
int ParticleData<T>::MaxRank() const { return _maxRank; }
template <typename T>
void ParticleData<T>::SetMaxRank(int r)
{
    _maxRank = r;
    if (std::is_same<T, Tucker>::value) device::Check(vt_tucker_enable(_dev->ctx, _species, _comprErr, _maxRank));
}

namespace {
struct SnapshotHeader {
    char magic[8];
    int32_t nTets, n[3], format;
};
const char kSnapMagic[8] = {'V', 'T', 'S', 'N', 'A', 'P', '1', 0};
}  // namespace

template <typename T>
void ParticleData<T>::WriteSnapshot(const std::string& path) const
{
    std::ofstream out(path, std::ios::binary);
    if (!out) throw std::runtime_error("Cannot open snapshot file for writing: " + path);
    SnapshotHeader h;
    std::memcpy(h.magic, kSnapMagic, 8);
    h.nTets = (int32_t)_mesh->tets.size();
    for (int k = 0; k < 3; k++) h.n[k] = _vGrid->nCells[k];
    h.format = std::is_same<T, Tucker>::value ? 1 : 0;
    out.write((const char*)&h, sizeof(h));
    const size_t N = (size_t)_vGrid->nCellsTotal;
    const int batch = (int)std::max<size_t>(1, (64u << 20) / (N * sizeof(double)));
    std::vector<double> buf((size_t)batch * N);
    for (int first = 0; first < h.nTets; first += batch) {
        const int count = std::min(batch, h.nTets - first);
        device::Check(vt_species_get_pdf(_dev->ctx, _species, first, count, buf.data()));
        out.write((const char*)buf.data(), (std::streamsize)((size_t)count * N * sizeof(double)));
    }
    if (!out) throw std::runtime_error("Error while writing snapshot: " + path);
}

template <typename T>
void ParticleData<T>::ReadSnapshot(const std::string& path)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("Cannot open snapshot file: " + path);
    SnapshotHeader h;
    in.read((char*)&h, sizeof(h));
    if (!in || std::memcmp(h.magic, kSnapMagic, 8) != 0) throw std::runtime_error("Not a snapshot file: " + path);
    if (h.nTets != (int32_t)_mesh->tets.size() || h.n[0] != _vGrid->nCells[0] || h.n[1] != _vGrid->nCells[1] ||
        h.n[2] != _vGrid->nCells[2])
        throw std::invalid_argument("Snapshot does not match the mesh or the velocity grid");
    const size_t N = (size_t)_vGrid->nCellsTotal;
    const int batch = (int)std::max<size_t>(1, (64u << 20) / (N * sizeof(double)));
    std::vector<double> buf((size_t)batch * N);
    for (int first = 0; first < h.nTets; first += batch) {
        const int count = std::min(batch, h.nTets - first);
        in.read((char*)buf.data(), (std::streamsize)((size_t)count * N * sizeof(double)));
        if (!in) throw std::runtime_error("Snapshot file is truncated: " + path);
        device::Check(vt_species_set_pdf(_dev->ctx, _species, first, count, buf.data()));
    }
    RebindRows();   // row handles (Full) / host mirror (Tucker), as SetMaxwellPDF leaves them
}

template class ParticleData<Full>;
template class ParticleData<Tucker>;

std::vector<double> ScalarField(const Mesh* mesh, std::function<double(const Point&)> f)
{
    std::vector<double> v(mesh->tets.size());
    for (size_t i = 0; i < v.size(); i++) v[i] = f(mesh->tets[i]->centroid);
    return v;
}
double DebyeLength(double temperature, double density, double charge)
{
    return std::sqrt(epsilon0 * boltzConst * temperature / density) / charge;
}
double PlasmaFrequency(double density, double charge, double mass)
{
    return charge * std::sqrt(density / (mass * epsilon0));
}

}  // namespace VlasovTucker
