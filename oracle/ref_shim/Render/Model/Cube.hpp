// Shadow of the reference's Render/Model/Cube.hpp for the headless oracle build.
// The real header pulls in the D3D render pipeline; the simulation path only needs the type to
// exist.  Render() records what it is asked to draw (position, scale), so the tests can compare the
// octree cells the reference's debug view draws (src/Sim/Octree.cpp:147-175) with the adapter's.
#pragma once
#include <vector>
#include <d3d11.h>
#include <SimpleMath.h>
class Cube
{
public:
    explicit Cube(ID3D11DeviceContext*) {}
    void Render(DirectX::SimpleMath::Vector3 position, float scale, DirectX::SimpleMath::Matrix, bool = true)
    {
        std::vector<float>& d = Drawn();
        d.push_back(position.x); d.push_back(position.y); d.push_back(position.z); d.push_back(scale);
    }
    static std::vector<float>& Drawn() { static std::vector<float> drawn; return drawn; }
};
