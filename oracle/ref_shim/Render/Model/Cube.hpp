// Shadow of the reference's Render/Model/Cube.hpp for the headless oracle build.
// The real header pulls in the D3D render pipeline; the simulation path only needs
// the type to exist (debug drawing of octree cells, src/Sim/Octree.cpp:147-175,
// is never reached without a D3D context).
#pragma once
#include <d3d11.h>
#include <SimpleMath.h>
class Cube
{
public:
    explicit Cube(ID3D11DeviceContext*) {}
    void Render(DirectX::SimpleMath::Vector3, float, DirectX::SimpleMath::Matrix, bool = true) {}
};
