// host/fvDVM.hpp — C++17 host-side mirror of Foam::fvDVM (reference src/fvDVM/fvDVM/fvDVM.H:269-376)
// above the C-ABI of include/dugks.h, for hosts without OpenFOAM.  Same member names, argument meaning
// and error behaviour as the reference class: construction fails loudly (the reference's FatalError
// becomes a std::runtime_error carrying dugks_last_error), evolution() is one time step, the macro
// accessors return the fields OpenFOAM's volScalarField/volVectorField would hold (vectors = 3
// contiguous doubles).  The OpenFOAM-side adapter in INTEGRATION.md marshals exactly these arrays.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/dugks.h"

namespace dugks {

// What the fvDVM constructor reads: mesh + LS vectors, patch table, constant/{Xis,weights},
// constant/DVMProperties, the 0/{rho,U,T} fields (fvDVM.C:886-1075).
struct CaseArrays {
    int nCells = 0, nInternalFaces = 0, nBoundaryFaces = 0, nSolutionD = 3;
    std::vector<int32_t> owner, neighbour;
    std::vector<double> C, V, Cf, Sf, ownLs, neiLs, patchLs, deltaCoeffs;
    std::vector<dugks_patch_t> patches;
    std::vector<double> Xis, weights;
    double xiMax = 0, xiMin = 0;
    dugks_gas_t gas{};
    std::vector<double> rho, U, T, rho_b, U_b, T_b;
};

class fvDVM {
  public:
    // rank/nRanks: velocity-space decomposition (-dvParallel, fvDVM.C:228-260); nccl_id: 128 bytes
    // from dugks_nccl_unique_id on rank 0, broadcast by the caller (MPI_Bcast in a real host)
    explicit fvDVM(const CaseArrays& c, int rank = 0, int nRanks = 1, int device = -1,
                   const void* nccl_id = nullptr, int store_h = 0)
        : nCells_(c.nCells), nFaces_(c.nInternalFaces + c.nBoundaryFaces), nBnd_(c.nBoundaryFaces),
          nXiPerDim_((int)c.Xis.size()), xiMax_(c.xiMax), xiMin_(c.xiMin), gas_(c.gas) {
        dugks_mesh_t m{};
        m.nCells = c.nCells; m.nInternalFaces = c.nInternalFaces; m.nBoundaryFaces = c.nBoundaryFaces;
        m.nSolutionD = c.nSolutionD;
        m.owner = c.owner.data(); m.neighbour = c.neighbour.data();
        m.C = c.C.data(); m.V = c.V.data(); m.Cf = c.Cf.data(); m.Sf = c.Sf.data();
        m.ownLs = c.ownLs.data(); m.neiLs = c.neiLs.data(); m.patchLs = c.patchLs.data();
        m.deltaCoeffs = c.deltaCoeffs.data();
        dugks_dvset_t dv{(int32_t)c.Xis.size(), 0, c.Xis.data(), c.weights.data(), c.xiMax, c.xiMin};
        dugks_par_t par{};
        par.rank = rank; par.nRanks = nRanks; par.device = device; par.nccl_unique_id = nccl_id;
        par.store_h = store_h;
        int rc = dugks_create(&m, c.patches.data(), (int32_t)c.patches.size(), &dv, &c.gas, &par, c.rho.data(),
                              c.U.data(), c.T.data(), c.rho_b.data(), c.U_b.data(), c.T_b.data(), &h_);
        if (rc != DUGKS_OK) throw std::runtime_error(std::string("fvDVM::fvDVM: ") + dugks_last_error(nullptr));
        dugks_sizes(h_, &nXi_, &nXiLocal_, nullptr, nullptr);
    }
    fvDVM(const fvDVM&) = delete;              // fvDVM.H:254-258
    fvDVM& operator=(const fvDVM&) = delete;
    ~fvDVM() {
        for (std::vector<double>* v : {&rho_, &U_, &T_, &q_, &tau_})
            if (pinned_ && !v->empty()) dugks_host_unregister(v->data());
        dugks_destroy(h_);
    }

    // fvDVM.H:288 — one time step; dt = runTime.deltaTValue()
    void evolution(double dt) { check(dugks_step(h_, dt), "fvDVM::evolution"); dirty_ = true; }
    // fvDVM.H:293
    void getCoNum(double dt, double& maxCoNum, double& meanCoNum) {
        check(dugks_courant(h_, dt, &maxCoNum, &meanCoNum), "fvDVM::getCoNum");
    }
    // convergence monitor of the time loop (dugksFoam.C:88-107) since the previous call, evaluated on the device
    void convergence(double& TemperatureChange, double& rhoChange, double& Uchange) {
        double c[3];
        check(dugks_convergence(h_, c), "fvDVM::convergence");
        TemperatureChange = c[0]; rhoChange = c[1]; Uchange = c[2];
    }
    // fvDVM.H:309-322 (cells) and :324-338 (faces: internal then boundary)
    const std::vector<double>& rhoVol() { sync(); return rho_; }
    const std::vector<double>& Uvol() { sync(); return U_; }
    const std::vector<double>& Tvol() { sync(); return T_; }
    const std::vector<double>& qVol() { sync(); return q_; }
    const std::vector<double>& tauVol() { sync(); return tau_; }
    const std::vector<double>& rhoSurf() { syncSurf(); return rhoS_; }
    const std::vector<double>& Usurf() { syncSurf(); return US_; }
    const std::vector<double>& Tsurf() { syncSurf(); return TS_; }
    const std::vector<double>& qSurf() { syncSurf(); return qS_; }
    const std::vector<double>& tauSurf() { syncSurf(); return tauS_; }
    // fvDVM.C:539-581
    void wallDiagnostics(std::vector<double>& qWall, std::vector<double>& stressWall) {
        qWall.assign((size_t)nBnd_ * 3, 0.0); stressWall.assign((size_t)nBnd_ * 9, 0.0);
        check(dugks_get_wall_diag(h_, qWall.data(), stressWall.data()), "fvDVM::wallDiagnostics");
    }
    // fvDVM.H:341-364
    int nXi() const { return nXi_; }
    int nXiPerDim() const { return nXiPerDim_; }
    double xiMax() const { return xiMax_; }
    double xiMin() const { return xiMin_; }
    double R() const { return gas_.R; }
    double omega() const { return gas_.omega; }
    double Tref() const { return gas_.Tref; }
    double muRef() const { return gas_.muRef; }
    double Pr() const { return gas_.Pr; }
    int KInner() const { return gas_.KInner; }
    // fvDVM.H:375 — gTilde (and hTilde) of one cell for all global DVs
    void writeDFonCell(int cell, std::vector<double>& g, std::vector<double>& h) {
        g.assign(nXi_, 0.0); h.assign(nXi_, 0.0);
        check(dugks_get_df(h_, cell, g.data(), h.data()), "fvDVM::writeDFonCell");
    }
    // exact restart (the reference's own restart drops the distribution functions, discreteVelocity.C:220-249):
    // everything the next evolution() reads, as one rank-local blob
    std::vector<char> checkpoint() {
        uint64_t n = 0;
        check(dugks_checkpoint_size(h_, &n), "fvDVM::checkpoint");
        std::vector<char> blob(n);
        check(dugks_checkpoint_save(h_, blob.data(), n), "fvDVM::checkpoint");
        return blob;
    }
    void restore(const std::vector<char>& blob) {
        check(dugks_checkpoint_load(h_, blob.data(), blob.size()), "fvDVM::restore");
        dirty_ = surfDirty_ = true;
    }
    dugks_handle_t* handle() { return h_; }

  private:
    void check(int rc, const char* where) {
        if (rc != DUGKS_OK) throw std::runtime_error(std::string(where) + ": " + dugks_last_error(h_));
    }
    void sync() {
        if (!dirty_ && !rho_.empty()) return;
        if (rho_.empty()) {
            rho_.resize(nCells_); T_.resize(nCells_); tau_.resize(nCells_);
            U_.resize((size_t)nCells_ * 3); q_.resize((size_t)nCells_ * 3);
            // page-lock the field storage once (best effort): the accessor then copies straight into it
            pinned_ = true;
            for (std::vector<double>* v : {&rho_, &U_, &T_, &q_, &tau_})
                if (dugks_host_register(v->data(), v->size() * sizeof(double)) != DUGKS_OK) pinned_ = false;
        }
        check(dugks_get_cell_macros(h_, rho_.data(), U_.data(), T_.data(), q_.data(), tau_.data()), "fvDVM::sync");
        surfDirty_ = true; dirty_ = false;
    }
    void syncSurf() {
        sync();
        if (!surfDirty_ && !rhoS_.empty()) return;
        rhoS_.resize(nFaces_); TS_.resize(nFaces_); tauS_.resize(nFaces_);
        US_.resize((size_t)nFaces_ * 3); qS_.resize((size_t)nFaces_ * 3);
        check(dugks_get_face_macros(h_, rhoS_.data(), US_.data(), TS_.data(), qS_.data(), tauS_.data()), "fvDVM::syncSurf");
        surfDirty_ = false;
    }
    dugks_handle_t* h_ = nullptr;
    int nCells_, nFaces_, nBnd_, nXiPerDim_, nXi_ = 0, nXiLocal_ = 0;
    double xiMax_, xiMin_;
    dugks_gas_t gas_;
    bool dirty_ = true, surfDirty_ = true, pinned_ = false;
    std::vector<double> rho_, U_, T_, q_, tau_, rhoS_, US_, TS_, qS_, tauS_;
};

}  // namespace dugks
