// Shadow of the reference's Render/DX/ConstantBuffer.hpp (headless oracle build).
// src/Sim/BruteForceCPU.hpp includes it but the CPU sim never uses a constant buffer.
#pragma once
#include <d3d11.h>
template <class T> class ConstantBuffer;
