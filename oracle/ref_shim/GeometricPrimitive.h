// Headless stand-in for DirectXTK's <GeometricPrimitive.h> (oracle build only).
// BarnesHut only creates the debug sphere when it is given a D3D context
// (src/Sim/BarnesHut.cpp:23-27); the oracle always passes nullptr.
#pragma once
#include <memory>
struct ID3D11DeviceContext;
namespace DirectX
{
    class GeometricPrimitive
    {
    public:
        static std::unique_ptr<GeometricPrimitive> CreateSphere(ID3D11DeviceContext*) { return nullptr; }
    };
}
