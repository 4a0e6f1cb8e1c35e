// Headless stand-in for <d3d11.h>.  TEST INFRASTRUCTURE ONLY (oracle build).
// The reference's src/Sim headers name a handful of D3D11 COM types but the CPU
// simulation path never dereferences them (the sims are constructed with a null
// context, exactly as SimulationState::RunSimulation does,
// src/States/Simulation/SimulationState.cpp:286).  Opaque declarations suffice.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cmath>
#include <mutex>   // src/Core/ThreadPool.hpp uses std::mutex without including <mutex>

struct ID3D11Device;
struct ID3D11DeviceContext;
struct ID3D11Buffer;
struct ID3D11ComputeShader;
struct ID3D11ShaderResourceView;
struct ID3D11UnorderedAccessView;
