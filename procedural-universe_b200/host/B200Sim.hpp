// B200Sim -- the reference's INBodySim implemented on libnbody_b200.so.
//
// Drop-in for the three sims CreateNBodySim can return (reference src/Sim/INBodySim.cpp:7-27):
// same interface (INBodySim.hpp:19-25), same ownership contract (Init stores a pointer to the
// caller's std::vector<Particle>, BruteForceCPU.cpp:20-23; Update leaves Position / Velocity /
// Forces of that vector up to date when it returns, SimulationState.cpp:52-60), same theta plumbing
// (BHThetaChanged -> Octree::Theta, BarnesHut.cpp:29-31), same constructor log line
// (BruteForceCPU.cpp:17, BarnesHut.cpp:12).
//
// This header includes the reference's own headers by name; it is compiled inside the reference's
// build (or, for the tests in this repository, against oracle/ref_shim).  Nothing of the reference
// is copied here.
#pragma once

#include <memory>
#include <vector>

#include "Sim/INBodySim.hpp"
#include "Render/Model/Cube.hpp"

#include "nbody_b200.h"

class B200Sim : public INBodySim
{
    public:
        enum class EMode { AllPairs, BarnesHut };

        // `context` is accepted (and ignored) so that the factory signature stays the reference's.
        B200Sim(ID3D11DeviceContext* context, EMode mode, int device = 0);
        ~B200Sim();

        void Init(std::vector<Particle>& particles) override;
        void Update(float dt) override;
        // BarnesHut::RenderDebug (BarnesHut.cpp:98-101): one cube per occupied octree leaf, drawn with the
        // reference's own Cube; like the reference, only when a D3D context was given.
        void RenderDebug(DirectX::SimpleMath::Matrix view, DirectX::SimpleMath::Matrix proj) override;

        // INBodySim has no virtual destructor, so deleting through the base pointer never runs
        // ~B200Sim; owners that care about the device memory call Shutdown() first.
        void Shutdown();

        void SetTheta(float theta);
        bool IsValid() const { return Handle != nullptr; }

    private:
        nb_handle Handle = nullptr;
        EMode Mode;
        std::vector<Particle>* Particles = nullptr;
        std::unique_ptr<Cube> DebugCube;
        std::vector<float> DebugCells;
        void* Pinned = nullptr;
        size_t PinnedBytes = 0;
        bool ThetaRegistered = false;

        void Pin();
        void Unpin();
};

// What a maintainer adds to CreateNBodySim: BruteForceGPU and BarnesHut map to
// the B200 engine, BruteForceCPU stays the reference's own CPU path.
std::unique_ptr<INBodySim> CreateB200NBodySim(ID3D11DeviceContext* context, ENBodySim type);
