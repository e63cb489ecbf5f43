// Per-species state of the public API (reference: src/particle_data.h:21-59).  For
// TensorType = Full the distribution function lives on the GPU and `pdf` holds row handles;
// for TensorType = Tucker the compressed tensors live on the GPU as well (vt_tucker_*) and `pdf`
// is a host mirror (core + factors per tet) refreshed by SyncFromDevice().
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "constants.h"
#include "mesh.h"
#include "typedefs.h"
#include "velocity_grid.h"

namespace VlasovTucker {
namespace device {
struct MeshContext;
}

// Parameters of SetMaxwellPDF: density per tet, temperature (0 = all particles at mostProbableV)
struct MaxwellPDF {
    std::vector<double> physDensity;
    double temperature;
    Vector3d mostProbableV;
};

template <typename TensorType>
class ParticleData {
public:
    // public data, as the drivers use it
    std::string species = "";
    double mass;
    double charge;
    std::vector<TensorType> pdf;   // one entry per tet

    ParticleData(const Mesh* mesh, const VelocityGrid* vGrid);

    // initial condition: Maxwellian normalised discretely to physDensity (particle_data.cpp:23-90)
    void SetMaxwellPDF(const MaxwellPDF& paramsPDF);

    // moments: sum_v f * cellVolume and sum_v v f * cellVolume / density (device reductions)
    std::vector<double> Density() const;
    std::vector<Vector3d> Velocity() const;

    // Tucker rounding parameters (ignored by Full)
    void SetCompressionError(double error);
    double CompressionError() const;
    int MaxRank() const;
    void SetMaxRank(int maxRank);   // addition: fixed-rank configurations

    // ---- device binding (not in the reference API)
    std::shared_ptr<device::MeshContext> DeviceContext() const { return _dev; }
    int DeviceSpecies() const { return _species; }
    void PushParams() const;        // mass/charge are public members assigned after construction
    void SyncFromDevice();          // Tucker: refresh the host mirror `pdf` from the device state

    // ---- lossless binary snapshot of the distribution function (addition): every
    // tet's tensor as doubles in the reference's tet order.  Full: restores the state bit for bit.
    // Tucker: stores the reconstruction; reading re-rounds it with precision 0.
    void WriteSnapshot(const std::string& path) const;
    void ReadSnapshot(const std::string& path);

private:
    void RebindRows();   // pdf = row handles (Full) or the host mirror of the device state (Tucker)

    const Mesh* _mesh;
    const VelocityGrid* _vGrid;
    std::shared_ptr<device::MeshContext> _dev;
    int _species = -1;
    int _maxRank;
    double _comprErr = 1e-10;
};

// helpers of the drivers (particle_data.cpp:145-162)
double DebyeLength(double temperature, double density, double charge);
double PlasmaFrequency(double density, double charge, double mass);
std::vector<double> ScalarField(const Mesh* mesh, std::function<double(const Point&)> densityFunc);
}  // namespace VlasovTucker
