// host/dugks_run.cpp — stand-alone driver that plays the role of dugksFoam.C's time loop
// (reference src/dugksFoam.C:63-109, setDeltaTvar.H:34-47) on a case dumped by
// `python -m dugksfoam_b200.dump_case`.  Usage:
//     dugks_run <case.bin> <nSteps> <maxCo or 0 for the case's fixed deltaT> <out.bin> [maxDeltaT]
// Writes rho[nc], U[nc][3], T[nc], q[nc][3] as raw doubles to <out.bin>.
// Exit code 2 + the library's message when no CUDA device is present (there is no CPU fallback).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "fvDVM.hpp"

namespace {
struct Blob { std::string dtype; std::vector<char> bytes; };

std::map<std::string, Blob> read_case_file(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    std::map<std::string, Blob> out;
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        std::istringstream ss(line);
        std::string name, dtype; size_t nbytes = 0;
        ss >> name >> dtype >> nbytes;
        Blob b; b.dtype = dtype; b.bytes.resize(nbytes);
        f.read(b.bytes.data(), (std::streamsize)nbytes);
        if (!f) throw std::runtime_error("truncated case file at " + name);
        out[name] = std::move(b);
    }
    return out;
}
template <class T>
std::vector<T> take(const std::map<std::string, Blob>& m, const char* name) {
    auto it = m.find(name);
    if (it == m.end()) throw std::runtime_error(std::string("case file lacks ") + name);
    std::vector<T> v(it->second.bytes.size() / sizeof(T));
    std::memcpy(v.data(), it->second.bytes.data(), v.size() * sizeof(T));
    return v;
}
}  // namespace

int main(int argc, char** argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: %s case.bin nSteps maxCo out.bin\n", argv[0]); return 1; }
    try {
        auto m = read_case_file(argv[1]);
        dugks::CaseArrays c;
        auto sizes = take<int32_t>(m, "sizes");        // nCells nInternalFaces nBoundaryFaces nSolutionD KInner
        c.nCells = sizes[0]; c.nInternalFaces = sizes[1]; c.nBoundaryFaces = sizes[2]; c.nSolutionD = sizes[3];
        c.owner = take<int32_t>(m, "owner"); c.neighbour = take<int32_t>(m, "neighbour");
        c.C = take<double>(m, "C"); c.V = take<double>(m, "V"); c.Cf = take<double>(m, "Cf"); c.Sf = take<double>(m, "Sf");
        c.ownLs = take<double>(m, "ownLs"); c.neiLs = take<double>(m, "neiLs"); c.patchLs = take<double>(m, "patchLs");
        c.deltaCoeffs = take<double>(m, "deltaCoeffs");
        c.Xis = take<double>(m, "Xis"); c.weights = take<double>(m, "weights");
        auto sc = take<double>(m, "scalars");          // xiMax xiMin R omega Tref muRef Pr deltaT
        c.xiMax = sc[0]; c.xiMin = sc[1];
        c.gas = dugks_gas_t{sc[2], sc[3], sc[4], sc[5], sc[6], sizes[4], 0};
        auto pk = take<int32_t>(m, "patches");         // kind start size U_bc T_bc per patch
        auto pp = take<double>(m, "patch_pressure");
        for (size_t p = 0; p * 5 < pk.size(); p++)
            c.patches.push_back(dugks_patch_t{pk[p * 5], pk[p * 5 + 1], pk[p * 5 + 2], pk[p * 5 + 3], pk[p * 5 + 4], 0, pp[p]});
        c.rho = take<double>(m, "rho"); c.U = take<double>(m, "U"); c.T = take<double>(m, "T");
        c.rho_b = take<double>(m, "rho_b"); c.U_b = take<double>(m, "U_b"); c.T_b = take<double>(m, "T_b");
        const int nSteps = std::atoi(argv[2]);
        const double maxCo = std::atof(argv[3]);
        double dt = sc[7];
        const double maxDeltaT = argc > 5 ? std::atof(argv[5]) : 1e300;     // controlDict maxDeltaT (readTimeControlsExplicit.H:39-52)

        dugks::fvDVM dvm(c);
        std::printf("dugks_run: %d cells, %d faces, %d discrete velocities\n", c.nCells,
                    c.nInternalFaces + c.nBoundaryFaces, dvm.nXi());
        for (int step = 0; step < nSteps; step++) {
            double maxCoNum = 0, meanCoNum = 0;
            dvm.getCoNum(dt, maxCoNum, meanCoNum);                          // CourantNo.H:35
            if (maxCo > 0) {                                                // setDeltaTvar.H:34-47: cuts are immediate, growth is damped
                const double maxDeltaTFact = maxCo / (maxCoNum + 1e-15);   // SMALL
                const double deltaTFact = std::min(std::min(maxDeltaTFact, 1.0 + 0.1 * maxDeltaTFact), 1.2);
                dt = std::min(deltaTFact * dt, maxDeltaT);
            }
            dvm.evolution(dt);                                              // dugksFoam.C:78
            if (step == nSteps - 1 || step % 10 == 0)
                std::printf("step %d  deltaT = %.9e  Courant max %.6f mean %.6f\n", step + 1, dt, maxCoNum, meanCoNum);
        }
        {   // dugksFoam.C:88-107 (one check over the whole run)
            double dT = 0, dRho = 0, dU = 0;
            dvm.convergence(dT, dRho, dU);
            std::printf("Temperature changes = %.6e\nDensity     changes = %.6e\nVelocity    changes = %.6e\n", dT, dRho, dU);
        }
        std::ofstream o(argv[4], std::ios::binary);
        auto put = [&](const std::vector<double>& v) { o.write((const char*)v.data(), (std::streamsize)(v.size() * 8)); };
        put(dvm.rhoVol()); put(dvm.Uvol()); put(dvm.Tvol()); put(dvm.qVol());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "dugks_run: FATAL ERROR: %s\n", e.what());
        return 2;
    }
    return 0;
}
