// Shadow of the reference's Render/DX/Shader.hpp (headless oracle build).
// src/Sim/BruteForceCPU.cpp includes it but calls nothing from it.
#pragma once
#include <string>
