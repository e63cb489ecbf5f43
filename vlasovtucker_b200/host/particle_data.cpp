#include "particle_data.h"

#include <algorithm>
#include <cmath>
#include <iostream>

#include "device.h"
#include "full.h"
#include "tucker.h"

namespace VlasovTucker {

template <typename T>
ParticleData<T>::ParticleData(const Mesh* mesh, const VelocityGrid* vGrid) : mass(1), charge(1), _mesh(mesh), _vGrid(vGrid)
{
    _maxRank = *std::max_element(vGrid->nCells.begin(), vGrid->nCells.end());
    std::cout << "The maximum rank is " << _maxRank << "\n";
}

template <>
ParticleData<Full>::ParticleData(const Mesh* mesh, const VelocityGrid* vGrid) : mass(1), charge(1), _mesh(mesh), _vGrid(vGrid)
{
    _maxRank = *std::max_element(vGrid->nCells.begin(), vGrid->nCells.end());
    std::cout << "The maximum rank is " << _maxRank << "\n";
    _dev = device::ContextOf(mesh);
    const int32_t n[3] = {vGrid->nCells[0], vGrid->nCells[1], vGrid->nCells[2]};
    device::Check(vt_species_create(_dev->ctx, n, vGrid->minV.data(), vGrid->maxV.data(), mass, charge, &_species));
}

template <typename T>
void ParticleData<T>::PushParams() const
{
    if (_dev) device::Check(vt_species_set_params(_dev->ctx, _species, mass, charge));
}

template <>
void ParticleData<Full>::SetMaxwellPDF(const MaxwellPDF& params)
{
    PushParams();   // the Maxwellian uses `mass`
    device::Check(vt_species_set_maxwell(_dev->ctx, _species, params.physDensity.data(), params.temperature,
                                         params.mostProbableV.data()));
    pdf.clear();
    pdf.reserve(_mesh->tets.size());
    for (size_t t = 0; t < _mesh->tets.size(); t++)
        pdf.push_back(Full::DeviceRow(_dev, _species, (int)t, _vGrid->nCells));
    std::cout << "PDF size reduction: " << 1.0 << " times on average\n";
}

template <>
void ParticleData<Tucker>::SetMaxwellPDF(const MaxwellPDF& params)
{
    // host construction of the initial Tucker tensors (particle_data.cpp:23-90, uncompressed ranks)
    const int n0 = _vGrid->nCells[0], n1 = _vGrid->nCells[1], n2 = _vGrid->nCells[2];
    Tensor3d t3d(n0, n1, n2);
    double reduction = 0;
    pdf.clear();
    for (auto* tet : _mesh->tets) {
        if (params.temperature != 0.0) {
            double normConst = 0;
            for (int i0 = 0; i0 < n0; i0++)
                for (int i1 = 0; i1 < n1; i1++)
                    for (int i2 = 0; i2 < n2; i2++) {
                        const Vector3d v = _vGrid->At(i0, i1, i2);
                        double v2 = 0;
                        for (int j = 0; j < 3; j++) v2 += (v[j] - params.mostProbableV[j]) * (v[j] - params.mostProbableV[j]);
                        t3d(i0, i1, i2) = std::exp(-mass * v2 / (2 * boltzConst * params.temperature));
                        normConst += t3d(i0, i1, i2);
                    }
            const double scale = params.physDensity[tet->index] / (_vGrid->cellVolume * normConst);
            for (long i = 0; i < t3d.size(); i++) t3d.data()[i] *= scale;
        } else {
            t3d.setZero();
            const int i0 = (int)((params.mostProbableV[0] - _vGrid->minV[0]) / _vGrid->step[0]);
            const int i1 = (int)((params.mostProbableV[1] - _vGrid->minV[1]) / _vGrid->step[1]);
            const int i2 = (int)((params.mostProbableV[2] - _vGrid->minV[2]) / _vGrid->step[2]);
            t3d(i0, i1, i2) = params.physDensity[tet->index] / _vGrid->cellVolume;
        }
        Tucker tensor(t3d);
        reduction += t3d.size() / (double)tensor.Size();
        pdf.push_back(tensor);
    }
    std::cout << "PDF size reduction: " << reduction / (double)_mesh->tets.size() << " times on average\n";
}

template <>
std::vector<double> ParticleData<Full>::Density() const
{
    std::vector<double> r(_mesh->tets.size());
    device::Check(vt_species_density(_dev->ctx, _species, r.data()));
    return r;
}
template <>
std::vector<Vector3d> ParticleData<Full>::Velocity() const
{
    std::vector<Vector3d> r(_mesh->tets.size());
    device::Check(vt_species_velocity(_dev->ctx, _species, &r[0][0]));
    return r;
}
template <>
std::vector<double> ParticleData<Tucker>::Density() const
{
    std::vector<double> r(_mesh->tets.size());
    for (size_t i = 0; i < r.size(); i++) r[i] = pdf[i].Sum() * _vGrid->cellVolume;
    return r;
}
template <>
std::vector<Vector3d> ParticleData<Tucker>::Velocity() const
{
    std::vector<Vector3d> r(_mesh->tets.size());
    const std::vector<double> density = Density();
    for (size_t i = 0; i < r.size(); i++)
        for (int k = 0; k < 3; k++) {
            const Tucker vPDF = Tucker(_vGrid->v[k]) * pdf[i];
            r[i][k] = density[i] != 0 ? vPDF.Sum() * _vGrid->cellVolume / density[i] : 0.0;
        }
    return r;
}

template <typename T>
void ParticleData<T>::SetCompressionError(double e) { _comprErr = e; }
template <typename T>
double ParticleData<T>::CompressionError() const { return _comprErr; }
template <typename T>
int ParticleData<T>::MaxRank() const { return _maxRank; }
template <typename T>
void ParticleData<T>::SetMaxRank(int r) { _maxRank = r; }

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
