// Parity driver for the C++ host API (tests/test_host_api_gpu.py): the reference's own drivers
// (examples/oscillations.cpp, examples/sheath.cpp) with a short run and a binary dump of the state.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>

#include "header.h"

using namespace VlasovTucker;
using namespace std;

static void Dump(ofstream& out, const vector<double>& v) { out.write((const char*)v.data(), v.size() * sizeof(double)); }

template <typename PD>
static void DumpState(ofstream& out, const Mesh& mesh, PD& pd)
{
    for (size_t i = 0; i < mesh.tets.size(); i++) {
        Tensor3d t = pd.pdf[i].Reconstructed();
        out.write((const char*)t.data(), t.size() * sizeof(double));
    }
    Dump(out, pd.Density());
    auto vel = pd.Velocity();
    out.write((const char*)vel.data(), vel.size() * sizeof(Vector3d));
}

// examples/sheath.cpp, shortened; TensorType is the alias the example flips (sheath.cpp:10)
template <typename TensorType>
static void RunSheath(const string& meshFile, int iterations, ofstream& out, int n0)
{
    double elTemperature = 1 * electronvolt, ionTemperature = 400, density = 1e17, ionMass = atomicMass;
    double debyeLength = DebyeLength(elTemperature, density, elCharge);
    double plasmaT = 2 * pi / PlasmaFrequency(density, elCharge, elMass);
    Mesh mesh(meshFile);
    mesh.SetPeriodicBounaries({{3, 4}, {5, 6}});
    mesh.Reconstruct(22 * debyeLength);
    double maxVE = sqrt(-log(1e-6) * 2 * boltzConst * elTemperature / elMass);
    VelocityGrid vGridE({n0, 5, 5}, {-4 * maxVE, -maxVE, -maxVE}, {4 * maxVE, maxVE, maxVE});
    double maxVI = sqrt(-log(1e-6) * 2 * boltzConst * ionTemperature / ionMass);
    VelocityGrid vGridI({n0, 5, 5}, {-4 * maxVI, -maxVI, -maxVI}, {4 * maxVI, maxVI, maxVI});
    auto rhoFunc = [density](const Point&) { return density; };
    ParticleData<TensorType> particleDataE(&mesh, &vGridE);
    particleDataE.species = "electron";
    particleDataE.mass = elMass;
    particleDataE.charge = -elCharge;
    particleDataE.SetCompressionError(1e-6);
    MaxwellPDF maxwellE;
    maxwellE.physDensity = ScalarField(&mesh, rhoFunc);
    maxwellE.temperature = elTemperature;
    maxwellE.mostProbableV = {0, 0, 0};
    particleDataE.SetMaxwellPDF(maxwellE);
    ParticleData<TensorType> particleDataI(&mesh, &vGridI);
    particleDataI.species = "ion";
    particleDataI.mass = ionMass;
    particleDataI.charge = elCharge;
    particleDataI.SetCompressionError(1e-6);
    MaxwellPDF maxwellI;
    maxwellI.physDensity = ScalarField(&mesh, rhoFunc);
    maxwellI.temperature = ionTemperature;
    maxwellI.mostProbableV = {0, 0, 0};
    particleDataI.SetMaxwellPDF(maxwellI);
    Solver<TensorType> solverE(&mesh, &vGridE, &particleDataE);
    Solver<TensorType> solverI(&mesh, &vGridI, &particleDataI);
    ParticleBC<TensorType> particleBC1;
    particleBC1.type = ParticleBCType::Absorbing;
    particleBC1.collectCharge = true;
    solverE.SetParticleBC(1, particleBC1);
    solverI.SetParticleBC(1, particleBC1);
    ParticleBC<TensorType> particleBC2;
    particleBC2.type = ParticleBCType::Free;
    solverE.SetParticleBC(2, particleBC2);
    solverI.SetParticleBC(2, particleBC2);
    FieldBC fieldBC1;
    fieldBC1.type = FieldBCType::ChargedPlane;
    fieldBC1.chargeDensity = 0;
    solverE.SetFieldBC(1, fieldBC1);
    FieldBC fieldBC2;
    fieldBC2.type = FieldBCType::ConstantPotential;
    fieldBC2.potential = 0;
    solverE.SetFieldBC(2, fieldBC2);
    MulticomponentSolver<TensorType> multiSolver(&solverE);
    multiSolver.AddSolver(&solverI);
    multiSolver.timeStep = 1e-4 * plasmaT;
    multiSolver.stepMultipliers[&solverI] = 10;
    multiSolver.stepMultipliers[&solverE] = 1;
    multiSolver.nIterations = iterations;
    multiSolver.Solve();
    DumpState(out, mesh, particleDataE);
    DumpState(out, mesh, particleDataI);
}

int main(int argc, char** argv)
{
    if (argc < 5) {
        cerr << "usage: host_parity <oscillations|sheath> <mesh.msh> <iterations> <out.bin>\n";
        return 2;
    }
    const string which = argv[1], meshFile = argv[2], outFile = argv[4];
    const int iterations = atoi(argv[3]);
    ofstream out(outFile, ios::binary);
    if (which == "snapshot" || which == "snapshot_tucker") {
        // WriteSnapshot / ReadSnapshot round trip: run a few iterations, write, read into a second
        // ParticleData on the same mesh; both states are dumped
        Mesh mesh(meshFile);
        mesh.SetPeriodicBounaries({{1, 2}, {3, 4}, {5, 6}});
        mesh.Reconstruct();
        VelocityGrid vGrid({11, 9, 7}, {-3, -1, -1}, {3, 1, 1});
        double L = 0;
        for (auto* p : mesh.points) L = max(L, (*p)[0]);
        MaxwellPDF paramsPDF;
        paramsPDF.physDensity = ScalarField(&mesh, [L](const Point& p) { return 10 + 0.2 * sin(p[0] / L * (2 * pi)); });
        paramsPDF.temperature = 0.3 / boltzConst;
        paramsPDF.mostProbableV = {0.4, 0, 0};
        const string snap = outFile + ".snap";
        auto run = [&](auto& first, auto& second, auto& solver) {
            first.species = "custom";
            first.mass = 1;
            first.charge = 2.975e-5;
            first.SetCompressionError(1e-6);
            first.SetMaxwellPDF(paramsPDF);
            solver.backgroundChargeDensity = vector<double>(mesh.tets.size(), -first.charge * 10);
            solver.timeStep = 1e-3;
            solver.nIterations = iterations;
            solver.Solve();
            first.WriteSnapshot(snap);
            second.mass = first.mass;
            second.charge = first.charge;
            second.SetCompressionError(1e-6);
            second.ReadSnapshot(snap);
            DumpState(out, mesh, first);
            DumpState(out, mesh, second);
        };
        if (which == "snapshot") {
            ParticleData<Full> a(&mesh, &vGrid), b(&mesh, &vGrid);
            Solver<Full> solver(&mesh, &vGrid, &a);
            run(a, b, solver);
        } else {
            ParticleData<Tucker> a(&mesh, &vGrid), b(&mesh, &vGrid);
            Solver<Tucker> solver(&mesh, &vGrid, &a);
            run(a, b, solver);
        }
        remove(snap.c_str());
    } else if (which == "oscillations_tucker") {
        // the same driver with TensorType = Tucker (examples/oscillations.cpp switches by one typedef);
        // a warm drifting Maxwellian so that the ranks are not trivially 1
        Mesh mesh(meshFile);
        mesh.SetPeriodicBounaries({{1, 2}, {3, 4}, {5, 6}});
        mesh.Reconstruct();
        VelocityGrid vGrid({11, 9, 7}, {-3, -1, -1}, {3, 1, 1});
        ParticleData<Tucker> particleData(&mesh, &vGrid);
        particleData.species = "custom";
        particleData.mass = 1;
        particleData.charge = 2.975e-5;
        MaxwellPDF paramsPDF;
        double L = 0;
        for (auto* p : mesh.points) L = max(L, (*p)[0]);
        auto rhoFunc = [L](const Point& p) { return 10 + 0.2 * sin(p[0] / L * (2 * pi)); };
        paramsPDF.physDensity = ScalarField(&mesh, rhoFunc);
        paramsPDF.temperature = 0.3 / boltzConst;
        paramsPDF.mostProbableV = {0.4, 0, 0};
        particleData.SetCompressionError(1e-6);
        particleData.SetMaxwellPDF(paramsPDF);
        Solver<Tucker> solver(&mesh, &vGrid, &particleData);
        solver.backgroundChargeDensity = vector<double>(mesh.tets.size(), -particleData.charge * 10);
        solver.timeStep = 1e-3;
        solver.nIterations = iterations;
        solver.Solve();
        DumpState(out, mesh, particleData);
    } else if (which == "oscillations") {
        // examples/oscillations.cpp with the stable parameters of SURVEY.md §8d (C1s)
        Mesh mesh(meshFile);
        mesh.SetPeriodicBounaries({{1, 2}, {3, 4}, {5, 6}});
        mesh.Reconstruct();
        VelocityGrid vGrid({11, 11, 11}, {-3, -0.1, -0.1}, {3, 0.1, 0.1});
        ParticleData<Full> particleData(&mesh, &vGrid);
        particleData.species = "custom";
        particleData.mass = 1;
        particleData.charge = 2.975e-5;
        MaxwellPDF paramsPDF;
        double L = 0;
        for (auto* p : mesh.points) L = max(L, (*p)[0]);
        auto rhoFunc = [L](const Point& p) { return 10 + 0.2 * sin(p[0] / L * (2 * pi)); };
        paramsPDF.physDensity = ScalarField(&mesh, rhoFunc);
        paramsPDF.temperature = 0;
        paramsPDF.mostProbableV = {0, 0, 0};
        particleData.SetCompressionError(1e-6);
        particleData.SetMaxwellPDF(paramsPDF);
        Solver<Full> solver(&mesh, &vGrid, &particleData);
        solver.backgroundChargeDensity = vector<double>(mesh.tets.size(), -particleData.charge * 10);
        solver.timeStep = 1e-4;
        solver.nIterations = iterations;
        solver.Solve();
        DumpState(out, mesh, particleData);
    } else if (which == "sheath") {
        RunSheath<Full>(meshFile, iterations, out, 50);
    } else if (which == "sheath_tucker") {
        // the same driver with `using TensorType = Tucker;` (examples/sheath.cpp:10), smaller grid
        RunSheath<Tucker>(meshFile, iterations, out, 20);
    } else {
        cerr << "unknown case\n";
        return 2;
    }
    return 0;
}
